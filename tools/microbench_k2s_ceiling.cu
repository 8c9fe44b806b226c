// Measured ceiling of the K2s (dtw_pairs_stream4_kernel) consumer mapping on B200: the kernel's own step body (block_step:
// 2 template rows x 8 resident window columns = 16 DP cells, 128 FFMA2 + the interleaved DP chain) run with NOTHING else —
// rows already unit-length in shared memory, no producers, no barriers, no exchange, no block switches, no masks — at one
// and two consumer warps per scheduler (the occupancy its 184-register state allows) and, for reference, the same body at
// 3-4 warps per scheduler, which the register file only admits for a HALF-width block (4 columns = 64 FFMA2 per step).
// What it prints is the time per 1 M (120x16, 100x16) pairs that the mapping would need if feeding and synchronisation were
// free: 4656 issued cells per pair = 291 steps per group of 32 pairs (3810 useful cells).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rustpotter_b200/csrc -I include -o tools/microbench_k2s_ceiling tools/microbench_k2s_ceiling.cu
#include <cstdio>
#include <cstdlib>

#include "../rustpotter_b200/csrc/dtw_stream4_kernel.cu"   // the product kernel's device functions (anonymous namespace)

using namespace rp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int STEPS = 2048;

// WARPS consumer warps per CTA reading one 16-slot ring of row pairs (LDS traffic as in the kernel: 8 LDS.128 per step)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_full_block(long long* cyc, float* sink, const float* __restrict__ in) {
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;   // one 16-slot ring of row pairs per group of 32 pairs, read by all the CTA's warps (as in the kernel)
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * RING_PAIR_F; i += blockDim.x) ring[i] = 0.25f * __sinf(0.37f * (float)(i + 7 * blockIdx.x));
    __syncthreads();
    const float* ring_p = ring + lane * RING_PAIR_F;
    f2 bcol[CB][8];
#pragma unroll
    for (int j = 0; j < CB; j++)
#pragma unroll
        for (int q = 0; q < 8; q++) bcol[j][q] = pk(in[(j * 8 + q) * 64 + lane], in[(j * 8 + q) * 64 + 32 + lane]);
    float D1[CB], D2[CB], cost2[CB];
#pragma unroll
    for (int j = 0; j < CB; j++) { D1[j] = 0.f; D2[j] = 0.f; cost2[j] = 1.f; }
    f2 ar1[8];
    load_row(ring_p, ar1);
    float out1 = 0.f, out2 = 0.f, li1_prev = 1.f;
    const f2 one = pk(1.f, 0.f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int st = 0; st < STEPS; st++) {
        const float* rp0 = ring_p + (st & 15) * SLOT_F;
        const float* rp1 = ring_p + ((st + 1) & 15) * SLOT_F;
        block_step(rp0, rp1, true, 0x3ffu, out2, out1, li1_prev, bcol, ar1, D1, D2, cost2, out1, out2, one);
        li1_prev = out1;
    }
    const long long t1 = clock64();
    float s = out1 + out2;
#pragma unroll
    for (int j = 0; j < CB; j++) s += D1[j] + D2[j];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Half-width block (4 columns, 64 block registers): what 3-4 consumer warps per scheduler would have to run. Same arithmetic
// per cell; the row loads now serve half as many cells (8 LDS.128 per 8 cells).
__device__ __forceinline__ void half_step4(const f2 (&ar)[8], const f2 (&bcol)[4][8], f2 (&acc)[4], const float (&cprev)[4],
                                           const float (&Dsrc)[4], float (&Ddst)[4], float left, float diag, f2 one) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j] = fma2(ar[q], bcol[j][q], q == 0 ? one : acc[j]);
        if (q < 4) {
            const float up = Dsrc[q];
            const float v = cprev[q] + min3(up, diag, left);
            diag = up;
            left = v;
            Ddst[q] = v;
        }
    }
}
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_half_block(long long* cyc, float* sink, const float* __restrict__ in) {
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;   // one 16-slot ring of row pairs per group of 32 pairs, read by all the CTA's warps (as in the kernel)
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * RING_PAIR_F; i += blockDim.x) ring[i] = 0.25f * __sinf(0.37f * (float)(i + 7 * blockIdx.x));
    __syncthreads();
    const float* ring_p = ring + lane * RING_PAIR_F;
    f2 bcol[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < 8; q++) bcol[j][q] = pk(in[(j * 8 + q) * 64 + lane], in[(j * 8 + q) * 64 + 32 + lane]);
    float D1[4], D2[4], cost1[4], cost2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { D1[j] = 0.f; D2[j] = 0.f; cost2[j] = 1.f; }
    f2 ar1[8], ar2[8], acc[4];
    load_row(ring_p, ar1);
    float left = 0.f, diag = 0.f;
    const f2 one = pk(1.f, 0.f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int st = 0; st < STEPS; st++) {
        const float* rp0 = ring_p + (st & 15) * SLOT_F;
        const float* rp1 = ring_p + ((st + 1) & 15) * SLOT_F;
        load_row(rp0 + kD, ar2);
        half_step4(ar1, bcol, acc, cost2, D1, D2, left, diag, one);
#pragma unroll
        for (int j = 0; j < 4; j++) cost1[j] = hsum(acc[j]);
        load_row(rp1, ar1);
        half_step4(ar2, bcol, acc, cost1, D2, D1, diag, left, one);
#pragma unroll
        for (int j = 0; j < 4; j++) cost2[j] = hsum(acc[j]);
        left = D1[3];
        diag = D2[3];
    }
    const long long t1 = clock64();
    float s = left + diag;
#pragma unroll
    for (int j = 0; j < 4; j++) s += D1[j] + D2[j];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
double run(K kernel, int warps, int ctas_per_sm, int nsm, long long* cyc, float* sink, const float* in) {
    const size_t smem = (size_t)32 * RING_PAIR_F * sizeof(float);
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; rep++) {
        kernel<<<nsm * ctas_per_sm, warps * 32, smem>>>(cyc, sink, in);
        CK(cudaDeviceSynchronize());
    }
    static long long h[4096];
    CK(cudaMemcpy(h, cyc, nsm * ctas_per_sm * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < nsm * ctas_per_sm; i++) avg += (double)h[i];
    return avg / (nsm * ctas_per_sm) / STEPS;   // cycles per step per warp
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    long long* cyc;
    float* sink;
    CK(cudaMalloc(&cyc, 4096 * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    float* in;
    {
        static float h[64 * 64];
        for (int i = 0; i < 64 * 64; i++) h[i] = -0.25f + 0.5f * (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
        CK(cudaMalloc(&in, sizeof(h)));
        CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    printf("device %s, %d SMs, %.3f GHz; one step = 2 rows x W columns; per 1 M pairs: 4656 issued cells per pair (8-column blocks)\n", p.name, nsm, ghz);
    struct Cfg { const char* name; int warps, ctas, cols; double cyc; };
    Cfg cfg[] = {
        {"8-col block, 1 warp/scheduler (4 warps/SM)", 4, 1, 8, 0}, {"8-col block, 2 warps/scheduler (8 warps/SM, the kernel's occupancy)", 4, 2, 8, 0},
        {"4-col block, 2 warps/scheduler", 4, 2, 4, 0}, {"4-col block, 3 warps/scheduler", 4, 3, 4, 0}, {"4-col block, 4 warps/scheduler", 8, 2, 4, 0},
    };
    for (auto& c : cfg) {
        if (c.cols == 8) c.cyc = c.warps == 4 ? run(k_full_block<4>, 4, c.ctas, nsm, cyc, sink, in) : 0;
        else c.cyc = c.warps == 4 ? run(k_half_block<4>, 4, c.ctas, nsm, cyc, sink, in) : run(k_half_block<8>, 8, c.ctas, nsm, cyc, sink, in);
        const double cells_per_step = 2.0 * c.cols;
        const double warp_steps = 1e6 / 32.0 * 4656.0 / cells_per_step;                       // per 1 M pairs
        const double ms = warp_steps * c.cyc / ((double)nsm * c.warps * c.ctas) / (ghz * 1e9) * 1e3;
        const double fma_pct = (c.cols * 16.0 * 2.0) / c.cyc * (c.warps * c.ctas / 4.0) * 100.0;   // FFMA2 pipe cycles (2 per FFMA2) per scheduler
        printf("%-70s %7.1f cycles/step/warp  -> %6.3f ms per 1 M pairs (%.2f of the 6547 GB/s roofline), FFMA2 pipe %.0f %%\n", c.name, c.cyc, ms,
               14.084 / ms / 6547.2 * 1e3, fma_pct);
    }
    return 0;
}
