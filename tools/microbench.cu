// Throwaway B200 micro-benchmarks that steer the DTW / MFCC kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Not product code; results are recorded in profiles/r01_microbench.txt.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3 distinct register sources per FMA (acc = x*y + acc) like a dot product
__global__ void k_ffma_dot(float* out, const float* in) {
    float x[8], y[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = in[i] + threadIdx.x; y[i] = in[8 + i]; acc[i] = 0.f; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(x[i], y[(i + it) & 7], acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[8];
    unsigned long long A = pack2(a, a), B = pack2(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = ffma2(acc[i], A, B);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2_dot(float* out, const float* in) {
    unsigned long long x[8], y[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = pack2(in[i] + threadIdx.x, in[i]); y[i] = pack2(in[8 + i], in[9 + i]); acc[i] = 0ull; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = ffma2(x[i], y[(i + it) & 7], acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; unpack2(acc[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// DP-like mix: per "cell" 8 ffma2 + 1 add + min3 + add  (what the fused DTW inner loop looks like)
__global__ void k_mix2(float* out, const float* in) {
    unsigned long long x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = pack2(in[i] + threadIdx.x, in[i]); y[i] = pack2(in[8 + i], in[9 + i]); }
    float d0 = in[0], d1 = in[1], d2 = in[2], d3 = in[3];
    float p0 = in[4], p1 = in[5], p2 = in[6], p3 = in[7];
    for (int it = 0; it < ITERS; it++) {
        // four independent cells
        unsigned long long a0 = pack2(1.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a0 = ffma2(x[i], y[i], a0);
            a1 = ffma2(x[i], y[(i + 1) & 7], a1);
            a2 = ffma2(x[i], y[(i + 2) & 7], a2);
            a3 = ffma2(x[i], y[(i + 3) & 7], a3);
        }
        float l, h;
        unpack2(a0, l, h); float c0 = l + h;
        unpack2(a1, l, h); float c1 = l + h;
        unpack2(a2, l, h); float c2 = l + h;
        unpack2(a3, l, h); float c3 = l + h;
        float n0 = c0 + min3(p0, p1, d3);
        float n1 = c1 + min3(p1, p2, n0);
        float n2 = c2 + min3(p2, p3, n1);
        float n3 = c3 + min3(p3, d0, n2);
        p0 = d0; p1 = d1; p2 = d2; p3 = d3;
        d0 = n0; d1 = n1; d2 = n2; d3 = n3;
        x[it & 7] = pack2(n3, n0);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
}
// same with scalar FFMA
__global__ void k_mix1(float* out, const float* in) {
    float x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = in[i] + threadIdx.x; y[i] = in[8 + i]; }
    float d0 = in[0], d1 = in[1], d2 = in[2], d3 = in[3];
    float p0 = in[4], p1 = in[5], p2 = in[6], p3 = in[7];
    for (int it = 0; it < ITERS; it++) {
        float c0 = 1.f, c1 = 1.f, c2 = 1.f, c3 = 1.f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            c0 = fmaf(x[i], y[i], c0);
            c1 = fmaf(x[i], y[(i + 1) & 15], c1);
            c2 = fmaf(x[i], y[(i + 2) & 15], c2);
            c3 = fmaf(x[i], y[(i + 3) & 15], c3);
        }
        float n0 = c0 + min3(p0, p1, d3);
        float n1 = c1 + min3(p1, p2, n0);
        float n2 = c2 + min3(p2, p3, n1);
        float n3 = c3 + min3(p3, d0, n2);
        p0 = d0; p1 = d1; p2 = d2; p3 = d3;
        d0 = n0; d1 = n1; d2 = n2; d3 = n3;
        x[it & 15] = n3;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
}

__global__ void k_min3(float* out, const float* in) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = in[i] + threadIdx.x;
    float a = in[8], b = in[9];
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = min3(v[i], a, b) + 1.0f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out, const float* in) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = in[i] + threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128: mode 0 broadcast (all lanes same address), 1 = consecutive 16B per lane, 2 = stride 64B per lane
__global__ void k_lds(float* out, int mode) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int base = mode == 0 ? 0 : (mode == 1 ? lane : lane * 4);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v = sm[(base + i * 128 + it) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}
__global__ void k_read(const float4* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 a = make_float4(0, 0, 0, 0);
    for (; i < n; i += stride) { float4 v = in[i]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    if (a.x + a.y + a.z + a.w == 12345.678f) out[0] = a.x;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d smem/block optin=%zu clock=%d kHz\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.clockRate);
    int nsm = p.multiProcessorCount;
    int blocks = nsm * 8, threads = 256;  // 2048 thr/SM
    float *out, *in;
    CK(cudaMalloc(&out, blocks * threads * sizeof(float)));
    CK(cudaMalloc(&in, 64 * sizeof(float)));
    float hin[64]; for (int i = 0; i < 64; i++) hin[i] = 0.001f * i;
    CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
    double nthr = (double)blocks * threads;
    float ms;
    ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 0.999f, 0.001f); });
    printf("FFMA (2-src same)   : %.2f TFMA/s  (%.3f ms)\n", nthr * ITERS * 8 / ms / 1e9, ms);
    ms = time_ms([&] { k_ffma_dot<<<blocks, threads>>>(out, in); });
    printf("FFMA dot (3 regs)   : %.2f TFMA/s  (%.3f ms)\n", nthr * ITERS * 8 / ms / 1e9, ms);
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.999f, 0.001f); });
    printf("FFMA2               : %.2f TFMA/s  (%.3f ms)\n", nthr * ITERS * 16 / ms / 1e9, ms);
    ms = time_ms([&] { k_ffma2_dot<<<blocks, threads>>>(out, in); });
    printf("FFMA2 dot (3 regs)  : %.2f TFMA/s  (%.3f ms)\n", nthr * ITERS * 16 / ms / 1e9, ms);
    ms = time_ms([&] { k_mix1<<<blocks, threads>>>(out, in); });
    printf("mix scalar (4 cells x16 fma + dp): %.2f Gcell/s  (%.3f ms)\n", nthr * ITERS * 4 / ms / 1e6, ms);
    ms = time_ms([&] { k_mix2<<<blocks, threads>>>(out, in); });
    printf("mix ffma2  (4 cells x8 fma2 + dp): %.2f Gcell/s  (%.3f ms)\n", nthr * ITERS * 4 / ms / 1e6, ms);
    ms = time_ms([&] { k_min3<<<blocks, threads>>>(out, in); });
    printf("min3+add pairs      : %.2f T(pairs)/s  (%.3f ms)\n", nthr * ITERS * 8 / ms / 1e9, ms);
    ms = time_ms([&] { k_shfl<<<blocks, threads>>>(out, in); });
    printf("SHFL.UP             : %.2f T lane-shfl/s = %.1f warp-shfl/clk/SM@1.9GHz (%.3f ms)\n", nthr * ITERS * 8 / ms / 1e9,
           nthr * ITERS * 8 / 32 / (ms * 1e-3) / nsm / 1.9e9, ms);
    for (int mode = 0; mode < 3; mode++) {
        ms = time_ms([&] { k_lds<<<blocks, threads, 32768>>>(out, mode); });
        printf("LDS.128 mode %d      : %.2f TB/s lane-bytes, %.2f warp-LDS/clk/SM@1.9GHz (%.3f ms)\n", mode,
               nthr * ITERS * 8 * 16 / ms / 1e9, nthr * ITERS * 8 / 32 / (ms * 1e-3) / nsm / 1.9e9, ms);
    }
    size_t n = (size_t)1 << 30;  // bytes
    float4 *a, *b;
    CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
    CK(cudaMemset(a, 1, n));
    ms = time_ms([&] { k_copy<<<nsm * 16, 512>>>(a, b, n / 16); });
    printf("copy 1GiB           : %.1f GB/s (r+w)\n", 2.0 * n / ms / 1e6);
    ms = time_ms([&] { k_read<<<nsm * 16, 512>>>(a, out, n / 16); });
    printf("read 1GiB           : %.1f GB/s\n", 1.0 * n / ms / 1e6);
    // pinned H2D bandwidth
    void* h; CK(cudaMallocHost(&h, n));
    ms = time_ms([&] { CK(cudaMemcpyAsync(a, h, n, cudaMemcpyHostToDevice)); }, 3);
    printf("H2D pinned 1GiB     : %.1f GB/s\n", 1.0 * n / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(h, a, n, cudaMemcpyDeviceToHost)); }, 3);
    printf("D2H pinned 1GiB     : %.1f GB/s\n", 1.0 * n / ms / 1e6);
    return 0;
}
