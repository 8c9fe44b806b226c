// Tensor-memory microbenchmark (B200): can TMEM hold a thread's resident operands? Measures tcgen05.ld 32x32b.x16 (16 registers
// per lane from the warp's own 32 TMEM lanes) throughput per SM for 4 and 8 warps, alone and interleaved with an FFMA2 stream
// that consumes the loaded registers. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tmem tools/microbench_tmem.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2000;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 pk(uint32_t lo, uint32_t hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

// mode 0: loads only; mode 1: loads + 64 FFMA2 per 8 loads (the DTW step's ratio: 8 x16-loads feed 128 FFMA2 per step -> 16 per load)
template <int MODE>
__global__ void __launch_bounds__(256) k_tmem(long long* cycles, float* sink, int check) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((uint64_t)__cvta_generic_to_shared(&tbase)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w: TMEM lanes 32*(w%4) .., columns 256*(w/4) .. +128 (its "resident block": 128 columns = 8 x 16)
    const uint32_t my = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 8; c++) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = __float_as_uint(0.001f * (float)(lane + 1) + (float)(c * 16 + i));
        tmem_st16(my + c * 16, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    f2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0ull;
    float s = 0.f;
    const long long t0 = clock64();
    if (MODE == 2) {   // eight loads in flight, one wait: bandwidth
        for (int it = 0; it < ITERS; it++) {
            uint32_t r[8][16];
#pragma unroll
            for (int c = 0; c < 8; c++) tmem_ld16(my + c * 16, r[c]);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int i = 0; i < 16; i += 4) s += __uint_as_float(r[c][i]);
        }
    } else if (MODE == 3) {   // double-buffered: slice c+1 is loading while slice c feeds 16 FFMA2
        uint32_t ra[16], rb[16];
        tmem_ld16(my, ra);
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                tmem_ld16(my + (c + 1) * 16, rb);
                const f2 a0 = pk(__float_as_uint(1.0f + it), __float_as_uint(0.5f)), a1 = pk(__float_as_uint(0.25f), __float_as_uint(2.0f + c));
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] = fma2(a0, pk(ra[2 * j], ra[2 * j + 1]), acc[j]);
                    acc[j] = fma2(a1, pk(ra[2 * j], ra[2 * j + 1]), acc[j]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                tmem_ld16(my + ((c + 2) & 7) * 16, ra);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] = fma2(a0, pk(rb[2 * j], rb[2 * j + 1]), acc[j]);
                    acc[j] = fma2(a1, pk(rb[2 * j], rb[2 * j + 1]), acc[j]);
                }
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        s += __uint_as_float(ra[0]);
    } else
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t r[16];
            tmem_ld16(my + c * 16, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) s += __uint_as_float(r[i]);
            } else {
                // 16 FFMA2: 8 column pairs x 2 rows, as one q-slice of the DTW step
                const f2 a0 = pk(__float_as_uint(1.0f + it), __float_as_uint(0.5f)), a1 = pk(__float_as_uint(0.25f), __float_as_uint(2.0f + c));
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] = fma2(a0, pk(r[2 * j], r[2 * j + 1]), acc[j]);
                    acc[j] = fma2(a1, pk(r[2 * j], r[2 * j + 1]), acc[j]);
                }
            }
        }
    }
    const long long t1 = clock64();
    if (check && warp == 5) {   // (warp-uniform: tcgen05.ld is .sync.aligned)
        uint32_t r[16];
        tmem_ld16(my + 2 * 16, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (lane == 3) sink[1] = __uint_as_float(r[5]);   // expect 0.001*4 + 37 = 37.004
    }
#pragma unroll
    for (int j = 0; j < 8; j++) s += __uint_as_float((uint32_t)(acc[j] & 0xffffffffu)) + __uint_as_float((uint32_t)(acc[j] >> 32));
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    long long* cyc;
    float* sink;
    CK(cudaMalloc(&cyc, nsm * sizeof(long long)));
    CK(cudaMalloc(&sink, 2 * sizeof(float)));
    CK(cudaMemset(sink, 0, 2 * sizeof(float)));
    const char* names[4] = {"loads only, waited one by one", "load, wait, 16 FFMA2", "eight loads in flight, one wait", "double-buffered load + 16 FFMA2"};
    for (int mode = 0; mode < 4; mode++)
        for (int threads : {128, 256}) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k_tmem<0><<<nsm, threads>>>(cyc, sink, threads == 256);
                else if (mode == 1) k_tmem<1><<<nsm, threads>>>(cyc, sink, 0);
                else if (mode == 2) k_tmem<2><<<nsm, threads>>>(cyc, sink, 0);
                else k_tmem<3><<<nsm, threads>>>(cyc, sink, 0);
                CK(cudaDeviceSynchronize());
            }
            long long h[256];
            CK(cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < nsm; i++) avg += (double)h[i];
            avg /= nsm;
            const double loads = (double)ITERS * 8;          // per warp
            const double bytes = loads * 32 * 64 * (threads / 32);
            printf("mode %d (%s), %d warps/SM: %.1f cycles per x16 load per warp, %.1f B/cycle/SM%s\n", mode, names[mode], threads / 32,
                   avg / loads, bytes / avg, (mode & 1) ? "  (16 FFMA2 alone: 32 cycles per warp)" : "");
        }
    float hs[2];
    CK(cudaMemcpy(hs, sink, sizeof(hs), cudaMemcpyDeviceToHost));
    printf("readback check (expect 37.004): %.3f\n", hs[1]);
    return 0;
}
