// Phase accounting of the SHIPPED K2s v4 kernel on BASELINE config 4 (1 M pairs 120x16 vs 100x16): where a consumer warp's
// and a producer warp's cycles go (prologue, super-step barrier, active steps, epilogue). The kernel source is included with
// RP_K2S_PROFILE defined, which turns its PROF_* macros into clock() reads + one shared-memory add per event (a few per step
// of ~1000 cycles; units of 16 cycles); the library build never defines it. A number from this tool is an explanation, not a bench value.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rustpotter_b200/csrc -I include -o tools/profile_k2s_phases tools/profile_k2s_phases.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef K2S_SRC
#define RP_K2S_PROFILE 1
#include "../rustpotter_b200/csrc/dtw_stream4_kernel.cu"
#else   // plain timing of another copy of the kernel source (same-box A/B of source variants)
#include K2S_SRC
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int64_t n_pairs = argc > 1 ? atoll(argv[1]) : 1000000;
    const int m = 120, n = 100, d = 16;
    float *tmpl, *win, *out;
    CK(cudaMalloc(&tmpl, n_pairs * m * d * sizeof(float)));
    CK(cudaMalloc(&win, n_pairs * n * d * sizeof(float)));
    CK(cudaMalloc(&out, n_pairs * sizeof(float)));
    {   // any non-degenerate data: the kernel's time does not depend on the values
        const size_t chunk = 1 << 24;
        std::vector<float> h(chunk);
        unsigned s = 12345u;
        for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) / 16777216.f - 0.5f; }
        for (size_t o = 0; o < (size_t)n_pairs * m * d; o += chunk) CK(cudaMemcpy(tmpl + o, h.data(), std::min(chunk, (size_t)n_pairs * m * d - o) * 4, cudaMemcpyHostToDevice));
        for (size_t o = 0; o < (size_t)n_pairs * n * d; o += chunk) CK(cudaMemcpy(win + o, h.data() + 77, std::min(chunk - 77, (size_t)n_pairs * n * d - o) * 4, cudaMemcpyHostToDevice));
    }
    rp::DtwPairsArgs a;
    a.tmpl = tmpl; a.tmpl_len_max = m; a.win = win; a.win_len_max = n; a.n_pairs = n_pairs; a.d = d; a.band = 5; a.out = out;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < (argc > 2 ? atoi(argv[2]) : 3); rep++) {
#ifdef RP_K2S_PROFILE
        unsigned long long zero[8][8] = {};
        CK(cudaMemcpyToSymbol(rp::g_prof, zero, sizeof(zero)));
#endif
        CK(cudaEventRecord(e0));
        CK(rp::launch_dtw_pairs_stream4(a, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
#ifndef RP_K2S_PROFILE
        printf("%.4f ms\n", ms);
        continue;
#else
        unsigned long long h[8][8];
        CK(cudaMemcpyFromSymbol(h, rp::g_prof, sizeof(h)));
        if (rep < 2) continue;
        for (int w = 0; w < 8; w++)
            for (int k = 0; k < 8; k++)
                if (k != 3) h[w][k] *= 16;   // counters are in units of 16 cycles
        const double groups = (double)((n_pairs + 31) / 32);
        printf("K2s v4 with phase accounting: %.3f ms for %lld pairs (the uninstrumented kernel: see bench.py's roofline)\n", ms, (long long)n_pairs);
        printf("consumer warps, cycles per group (and share of the group):\n");
        printf("  warp  prologue        barrier wait    active steps    epilogue        group     active steps/group  cycles/active step\n");
        for (int w = 0; w < 4; w++) {
            const double tot = h[w][5] / groups;
            printf("  %d   %8.0f (%4.1f%%) %8.0f (%4.1f%%) %8.0f (%4.1f%%) %8.0f (%4.1f%%) %9.0f   %6.1f            %7.1f\n", w, h[w][0] / groups,
                   100.0 * h[w][0] / h[w][5], h[w][1] / groups, 100.0 * h[w][1] / h[w][5], h[w][2] / groups, 100.0 * h[w][2] / h[w][5],
                   h[w][4] / groups, 100.0 * h[w][4] / h[w][5], tot, h[w][3] / groups, (double)h[w][2] / (double)h[w][3]);
        }
        printf("producer warps, cycles per group: waiting for free slots / loads + scale + store\n");
        for (int w = 4; w < 8; w++) printf("  %d   %8.0f   %8.0f\n", w, h[w][1] / groups, h[w][2] / groups);
#endif
    }
    return 0;
}
