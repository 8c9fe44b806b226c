// K2 (generic) — banded DTW score of (template, window) pairs, and K3 — window judgement.
//
// Replaces MfccNormalizer::normalize (reference src/mfcc/normalizer.rs:3-31), cosine distance and
// MfccComparator::compare (src/mfcc/comparator.rs:15-48), Dtw::compute_optimal_path_with_window
// (src/mfcc/dtw.rs:56-105) and the scoring half of WakewordComparator::run_detection
// (src/wakewords/comp/wakeword_comp.rs:22-37,77-152).
//
// This file holds the GENERIC kernel: any m, n, mfcc width and band; one warp per pair walking the
// cost matrix along anti-diagonals (cells (r, c) with r + c = k are independent given diagonals
// k-1 and k-2). It evaluates every floating-point operation in the reference's order with explicit
// round-to-nearest mul/add (no FMA contraction), so its costs are bit-identical to the reference's
// f32 arithmetic; only the final expf may differ in the last ulp. The tuned kernels
// (dtw_stream4_kernel.cu, dtw_window_kernel.cu, dtw_cadence_kernel.cu) are validated against it and the oracle.
//
// Reference quirks reproduced:
//   - a = template (m rows), b = window (n cols); window = max(band, |m-n|)
//   - row r touches columns [max(1, r-window), min(n, r+window-1)]  (asymmetric band)
//   - D[0][0] = 0, everything else +inf; cell = cost + min(D[r-1][c], D[r][c-1], D[r-1][c-1])
//   - the returned cell is D[m-1][n] (NOT D[m][n]); +inf there gives score 0
//   - cost = 1 - dot/sqrt(|a|^2 |b|^2), similarity 0 when the product of norms is 0
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kGenericWarps = 4;

struct PairView {
    const float* a;  // template rows (global)
    const float* b;  // window rows (global)
    int m, n;
};

// One warp scores one pair. smem layout per warp (floats):
//   A[m_max*d] B[n_max*d] na[m_max] nb[n_max] d0[m_max+1] d1[m_max+1] d2[m_max+1]
__device__ float dtw_pair_faithful(const PairView pv, int d, int band, float score_ref, int cmn, float* sm, int m_max,
                                   int n_max, int lane) {
    const int m = pv.m, n = pv.n;
    float* A = sm;
    float* B = A + (size_t)m_max * d;
    float* na = B + (size_t)n_max * d;
    float* nb = na + m_max;
    float* dbuf0 = nb + n_max;
    float* dbuf1 = dbuf0 + (m_max + 1);
    float* dbuf2 = dbuf1 + (m_max + 1);
    if (m < 1 || n < 1) return 0.f;

    for (int i = lane; i < m * d; i += 32) A[i] = __ldg(pv.a + i);
    for (int i = lane; i < n * d; i += 32) B[i] = __ldg(pv.b + i);
    __syncwarp();

    if (cmn) {  // normalizer.rs:3-31: per-coefficient sum over frames in ascending order, then x -= sum / n
        for (int j = lane; j < d; j += 32) {
            float s = 0.f;
            for (int r = 0; r < n; r++) s = __fadd_rn(s, B[r * d + j]);
            const float mean = __fdiv_rn(s, (float)n);
            for (int r = 0; r < n; r++) B[r * d + j] = __fsub_rn(B[r * d + j], mean);
        }
        __syncwarp();
    }
    // |a_r|^2, |b_c|^2 in dimension order (comparator.rs:35-41 accumulates them per call in this order)
    for (int r = lane; r < m; r += 32) {
        float s = 0.f;
        for (int j = 0; j < d; j++) s = __fadd_rn(s, __fmul_rn(A[r * d + j], A[r * d + j]));
        na[r] = s;
    }
    for (int c = lane; c < n; c += 32) {
        float s = 0.f;
        for (int j = 0; j < d; j++) s = __fadd_rn(s, __fmul_rn(B[c * d + j], B[c * d + j]));
        nb[c] = s;
    }
    // diagonal 0 holds D[0][0] = 0; diagonal 1 holds only border cells (+inf)
    for (int i = lane; i <= m; i += 32) {
        dbuf2[i] = i == 0 ? 0.f : INFINITY;
        dbuf1[i] = INFINITY;
    }
    __syncwarp();

    const int diff = m > n ? m - n : n - m;
    const int window = band > diff ? band : diff;
    float* d0 = dbuf0;
    float* d1 = dbuf1;
    float* d2 = dbuf2;
    const int k_res = m - 1 + n;  // diagonal of the returned cell (m-1, n)
    float result = INFINITY;
    for (int k = 2; k <= m + n; k++) {
        for (int i = lane; i <= m; i += 32) d0[i] = INFINITY;
        __syncwarp();
        int r_lo = (k - window + 2) >> 1;  // ceil((k - window + 1) / 2)  from c <= r + window - 1
        int r_hi = (k + window) >> 1;      // floor((k + window) / 2)     from c >= r - window
        r_lo = max(r_lo, max(1, k - n));
        r_hi = min(r_hi, min(m, k - 1));
        for (int r = r_lo + lane; r <= r_hi; r += 32) {
            const int c = k - r;
            const float* av = A + (size_t)(r - 1) * d;
            const float* bv = B + (size_t)(c - 1) * d;
            float dot = 0.f;
            for (int j = 0; j < d; j++) dot = __fadd_rn(dot, __fmul_rn(av[j], bv[j]));
            const float mag = __fsqrt_rn(__fmul_rn(na[r - 1], nb[c - 1]));
            const float sim = mag == 0.f ? 0.f : __fdiv_rn(dot, mag);
            const float cost = __fsub_rn(1.f, sim);
            const float best = fminf(fminf(fminf(INFINITY, d1[r - 1]), d1[r]), d2[r - 1]);
            d0[r] = __fadd_rn(cost, best);
        }
        __syncwarp();
        if (k == k_res) result = (m - 1 >= 1) ? d0[m - 1] : INFINITY;
        float* t = d2;
        d2 = d1;
        d1 = d0;
        d0 = t;
    }
    // comparator.rs:21-26
    const float normalized = __fdiv_rn(result, (float)(m + n));
    return __fdiv_rn(1.f, __fadd_rn(1.f, expf(__fdiv_rn(__fsub_rn(normalized, score_ref), score_ref))));
}

__global__ void __launch_bounds__(kGenericWarps * 32) dtw_pairs_generic_kernel(DtwPairsArgs a, int per_warp_floats) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = smem + (size_t)warp * per_warp_floats;
    for (int64_t p = (int64_t)blockIdx.x * kGenericWarps + warp; p < a.n_pairs; p += (int64_t)gridDim.x * kGenericWarps) {
        PairView pv;
        pv.m = a.tmpl_len ? a.tmpl_len[p] : a.tmpl_len_max;
        pv.n = a.win_len ? a.win_len[p] : a.win_len_max;
        pv.a = a.tmpl + (a.tmpl_off ? a.tmpl_off[p] : p * (int64_t)a.tmpl_len_max * a.d);
        pv.b = a.win + (a.win_off ? a.win_off[p] : p * (int64_t)a.win_len_max * a.d);
        if (pv.m > a.tmpl_len_max) pv.m = a.tmpl_len_max;  // never overrun shared memory
        if (pv.n > a.win_len_max) pv.n = a.win_len_max;
        const float s = dtw_pair_faithful(pv, a.d, a.band, a.score_ref, a.cmn, sm, a.tmpl_len_max, a.win_len_max, lane);
        if (lane == 0) a.out[p] = s;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kGenericWarps * 32) dtw_windows_generic_kernel(DtwWindowsArgs a, int per_warp_floats) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sm = smem + (size_t)warp * per_warp_floats;
    const int64_t n_pairs = a.n_streams * (int64_t)a.n_new * a.n_slots;
    for (int64_t p = (int64_t)blockIdx.x * kGenericWarps + warp; p < n_pairs; p += (int64_t)gridDim.x * kGenericWarps) {
        const int s = (int)(p % a.n_slots);
        const int64_t w = p / a.n_slots;
        const int j = (int)(w % a.n_new);
        const int64_t b = w / a.n_new;
        if (j < a.first_window) continue;  // warp-uniform
        PairView pv;
        pv.m = a.slot_len[s];
        // cut_and_normalize_frame keeps the first m frames of the window (wakeword_comp.rs:22-27); the window holds
        // max_mfcc_frames rows, so an avg_features matrix longer than every template meets n < m
        pv.n = a.window_len > 0 && a.window_len < pv.m ? a.window_len : pv.m;
        pv.a = a.tmpl + a.slot_off[s];
        pv.b = a.frames + ((b * a.frame_rows) + a.first_window_row + j) * (int64_t)a.d;
        const float sc = dtw_pair_faithful(pv, a.d, a.band, a.score_ref, 1, sm, a.max_len, a.max_len, lane);
        if (lane == 0) a.scores[p] = sc;
        __syncwarp();
    }
}

// K3: one thread per window.
__global__ void judge_windows_kernel(JudgeArgs a) {
    const int64_t total = a.n_streams * (int64_t)a.n_new;
    const int stride = 5 + a.max_templates;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        if ((int)(w % a.n_new) < a.first_window) continue;
        const float* row = a.scores + w * a.n_slots;
        const Judgement jd = judge_window(row, a.metas, a.n_wakewords, a.score_mode);
        if (jd.wakeword < 0) continue;
        const int idx = atomicAdd(a.hit_count, 1);
        if (idx >= a.capacity) continue;
        float* rec = a.hits + (int64_t)idx * stride;
        const WakewordMeta m = a.metas[jd.wakeword];
        rec[0] = __int_as_float(a.stream_base + (int)(w / a.n_new));
        rec[1] = __int_as_float((int)(w % a.n_new));
        rec[2] = __int_as_float(jd.wakeword);
        rec[3] = jd.avg_score;
        rec[4] = jd.score;
        const int s0 = m.slot_begin + (m.has_avg ? 1 : 0);
        for (int t = 0; t < m.n_templates; t++) rec[5 + t] = row[s0 + t];
    }
}

size_t generic_per_warp_floats(int m_max, int n_max, int d) {
    return (size_t)(m_max + n_max) * (d + 1) + 3 * (size_t)(m_max + 1);
}

}  // namespace

cudaError_t launch_dtw_pairs_generic(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    const size_t pw = generic_per_warp_floats(a.tmpl_len_max, a.win_len_max, a.d);
    const size_t bytes = pw * kGenericWarps * sizeof(float);
    if (bytes > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(dtw_pairs_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    int64_t blocks = (a.n_pairs + kGenericWarps - 1) / kGenericWarps;
    const int64_t cap = 148 * 64;
    if (blocks > cap) blocks = cap;
    dtw_pairs_generic_kernel<<<(unsigned)blocks, kGenericWarps * 32, bytes, stream>>>(a, (int)pw);
    return cudaGetLastError();
}

cudaError_t launch_dtw_windows_generic(const DtwWindowsArgs& a, cudaStream_t stream) {
    const int64_t n_pairs = a.n_streams * (int64_t)a.n_new * a.n_slots;
    if (n_pairs <= 0) return cudaSuccess;
    const size_t pw = generic_per_warp_floats(a.max_len, a.max_len, a.d);
    const size_t bytes = pw * kGenericWarps * sizeof(float);
    if (bytes > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(dtw_windows_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    int64_t blocks = (n_pairs + kGenericWarps - 1) / kGenericWarps;
    const int64_t cap = 148 * 64;
    if (blocks > cap) blocks = cap;
    dtw_windows_generic_kernel<<<(unsigned)blocks, kGenericWarps * 32, bytes, stream>>>(a, (int)pw);
    return cudaGetLastError();
}

cudaError_t launch_judge_windows(const JudgeArgs& a, cudaStream_t stream) {
    const int64_t total = a.n_streams * (int64_t)a.n_new;
    if (total <= 0) return cudaSuccess;
    int64_t blocks = (total + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    judge_windows_kernel<<<(unsigned)blocks, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace rp
