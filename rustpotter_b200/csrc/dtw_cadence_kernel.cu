// K2c — window scorer for SHORT calls (the reference's own cadence: one 30 ms chunk = 3 new windows per stream and
// call, src/detector.rs:347-376), mfcc_size <= 16, band_size <= 5, sm_100a.
//
// Same arithmetic contract as K2p (dtw_window_kernel.cu; reference src/wakewords/comp/wakeword_comp.rs:22-37 over
// src/mfcc/{normalizer,comparator,dtw}.rs) and the same algebra — cost(r, c) = 1 - (G_r[u] - A_r) * inv_c with
// G_r[u] = a^_r . x_u shared by the windows that overlap in frame u — but a different mapping. K2p gives every window a
// thread and needs 128 consecutive windows of one stream to fill a CTA; a 30 ms call has three, so its 36 864 CTAs (4096
// streams x 9 templates) ran with three live threads each: 7.2 ms per call, whatever the call length (profiles/r02_*).
// Here a HALF-WARP scores up to three consecutive windows of one stream against one template, the two halves of a warp
// two streams against the same template (so the template row is fetched once for both):
//   lanes 0 .. 2W+1 of the half   one frame each of the row's shared range: G_r[u] (a 16-dim dot, 8 FFMA2)
//   lanes 12 .. 14                A_r of window 0 .. 2  (the same dot against the window's mean, which they keep in registers)
//   lanes 0 .. 2 again            the thread-serial in-place band DP of window 0 .. 2 (K2p's; G and A arrive by shuffle)
// A CTA = all templates ("slots") of a stream pair's window triple, one warp each (each stream's ~105 frames are staged
// in shared memory once); the template row, identical for the warp, is a broadcast load from L1 / L2. ~80 warp instructions
// per template row for six windows: well below K2p's efficiency on long calls, several times faster on short ones
// (4096 streams x 36 templates, 30 ms calls: see DESIGN.md section 6). The engine takes this kernel when a call brings at
// most 24 new windows per stream.
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int W = 5;               // band half-width of the kernel's cell layout (band_size <= 5; smaller bands are masked)
constexpr int NB = 2 * W;          // band cells per row
constexpr int NWIN = 3;            // windows per warp
constexpr int kMaxWarps = 12;      // warps (slots in flight) per CTA
constexpr int kXS = 20;            // shared-memory row stride of the frame tile in floats (80 B: conflict-free LDS.128)

typedef unsigned long long f2;

__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float hsum(f2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
struct Row16 {
    f2 p[8];
};
__device__ __forceinline__ Row16 ld_row(const float* s) {   // 16 floats, 16-byte aligned (shared or global)
    Row16 r;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 v = *reinterpret_cast<const float4*>(s + 4 * q);
        r.p[2 * q] = pk(v.x, v.y);
        r.p[2 * q + 1] = pk(v.z, v.w);
    }
    return r;
}
__device__ __forceinline__ float dot16(const Row16& a, const Row16& b) {
    f2 acc = fma2(a.p[0], b.p[0], 0ull);
#pragma unroll
    for (int q = 1; q < 8; q++) acc = fma2(a.p[q], b.p[q], acc);
    return hsum(acc);
}

// Floats per stream tile: the second stream's tile starts 16 banks after the first's, so the two halves of a warp reading the
// same (row, coefficient) of their streams (the window means) do not meet in a bank.
__host__ __device__ inline int x_tile_floats(int x_rows) {
    const int n = x_rows * kXS;
    return n + ((16 - n % 32) + 32) % 32;
}

// grid: ceil(n_streams / 2) * triples CTAs; block: n_warps * 32 threads; dynamic shared memory:
//   Xs[2][x_rows][kXS] frames of the two streams | per warp and half: Mu[NWIN][16] (negated means) | Inv[NWIN][inv_cols]
__global__ void __launch_bounds__(kMaxWarps * 32, 2) dtw_windows_cadence_kernel(DtwWindowsArgs a, const float* __restrict__ tmpl_unit,
                                                                            const int64_t* __restrict__ unit_off, int triples, int x_rows,
                                                                            int inv_cols) {
    extern __shared__ __align__(16) float sm[];
    const int n_warps = blockDim.x >> 5;
    float* XsAll = sm;
    const int x_tile = x_tile_floats(x_rows);
    float* MuAll = XsAll + (size_t)2 * x_tile;
    float* InvAll = MuAll + (size_t)n_warps * 2 * NWIN * kD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, hl = lane & 15, hbase = lane & 16;   // half-warp, lane inside it, its first lane
    const int64_t pair = blockIdx.x / triples;
    const int q = (int)(blockIdx.x - pair * triples);
    const int j0 = a.first_window + q * NWIN;                 // first window (new frame index) of this triple
    const int n_win = min(NWIN, a.n_new - j0);

    // ---- stage the frames of the triple for both streams: window jj covers tile rows jj .. jj + m - 1
    for (int h = 0; h < 2; h++) {
        const int64_t b = 2 * pair + h;
        float* Xh = XsAll + (size_t)h * x_tile;
        const int64_t row0 = (int64_t)a.first_window_row + j0;
        const int64_t avail = b < a.n_streams ? a.frame_rows - row0 : 0;   // (an odd stream count: the last pair's second half is zeros)
        if (a.d == kD) {
            const float* src = a.frames + (b * a.frame_rows + row0) * kD;
            for (int i = tid; i < x_rows * 4; i += blockDim.x) {
                const int u = i >> 2, qq = i & 3;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < avail) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)u * kD) + qq);
                *reinterpret_cast<float4*>(Xh + u * kXS + 4 * qq) = v;
            }
        } else {   // mfcc_size < 16: zero-padded to 16
            const float* src = a.frames + (b * a.frame_rows + row0) * a.d;
            for (int i = tid; i < x_rows * kD; i += blockDim.x) {
                const int u = i >> 4, qq = i & 15;
                Xh[u * kXS + qq] = (u < avail && qq < a.d) ? __ldg(src + (size_t)u * a.d + qq) : 0.f;
            }
        }
    }
    __syncthreads();

    const int64_t b = 2 * pair + half;                          // this half-warp's stream
    const float* Xs = XsAll + (size_t)half * x_tile;
    float* Mu = MuAll + (size_t)(warp * 2 + half) * (NWIN * kD);
    float* Inv = InvAll + (size_t)(warp * 2 + half) * NWIN * inv_cols;
    const unsigned band_mask = ((1u << (2 * a.band)) - 1u) << (W - a.band);   // cells inside [r-band, r+band-1]
    const bool masked = a.band != W;

    for (int s = warp; s < a.n_slots; s += n_warps) {
        const int m = a.slot_len[s];
        const float* trow = tmpl_unit + unit_off[s];           // unit template rows, 16 floats each
        // ---- window means (normalizer.rs:3-31): lane hl = coefficient, frames summed in ascending order as the reference
        // does; windows 1 and 2 follow by sliding the sum; stored NEGATED
        {
            float sacc = 0.f;
            for (int f = 0; f < m; f++) sacc += Xs[f * kXS + hl];
            const float fm = (float)m;
#pragma unroll
            for (int jj = 0; jj < NWIN; jj++) {
                Mu[jj * kD + hl] = -__fdiv_rn(sacc, fm);
                sacc = (sacc - Xs[jj * kXS + hl]) + Xs[(jj + m) * kXS + hl];
            }
        }
        __syncwarp();
        // ---- 1 / |x_u - mu_jj| for every column of every window (0 for the zero vector: similarity 0)
        for (int jj = 0; jj < NWIN; jj++) {
            const Row16 nmu = ld_row(Mu + jj * kD);
            for (int c = 1 + hl; c < inv_cols; c += 16) {       // column c <-> tile row jj + c - 1
                const Row16 x = ld_row(Xs + (jj + c - 1) * kXS);
                f2 nn = 0ull;
#pragma unroll
                for (int qq = 0; qq < 8; qq++) {
                    const f2 y = add2(x.p[qq], nmu.p[qq]);
                    nn = fma2(y, y, nn);
                }
                const float n2 = hsum(nn);
                Inv[jj * inv_cols + c] = n2 > 0.f ? rsqrtf(n2) : 0.f;
            }
        }
        __syncwarp();

        // ---- lane roles of the row loop
        const int wj = min(hl, NWIN - 1);                        // DP lanes (hl < NWIN): their window
        const bool is_a = hl >= 12 && hl < 12 + NWIN;            // A lanes: dot with the (negated) mean of window hl - 12
        const float* Invw = Inv + wj * inv_cols;
        float D[NB], inv[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) {
            D[i] = INFINITY;
            inv[i] = 0.f;
        }
        D[W] = 0.f;   // D[0][0] seen from row 1 as the (r-1, c-1) neighbour of column 1
#pragma unroll
        for (int c = 1; c < W; c++) inv[c % NB] = Invw[c];       // columns 1 .. W-1 enter the band before row 1

        const int last_row = m - 1;
        Row16 v = ld_row(is_a ? Mu + (hl - 12) * kD : Xs);       // A lanes: the window's negated mean, for every row
        for (int r0 = 0; r0 < last_row; r0 += NB) {
#pragma unroll
            for (int k = 0; k < NB; k++) {
                const int r = r0 + k + 1;
                if (r <= last_row) {                             // warp-uniform
                    // the template row, the same for the whole warp: an L1 / L2 broadcast load. (Fetching it one row ahead
                    // costs 16 registers: with the A lanes' means resident that spills at two CTAs per SM -- measured 2.12 ms
                    // per call against 1.53 without the look-ahead and 1.78 for the round's first version.)
                    const Row16 ar = ld_row(trow + (size_t)(r - 1) * kD);
                    // one dot per lane: frame u = r - W - 1 + hl of the stream's tile (lanes 0 .. NB+1 of the half), or the
                    // window's negated mean
                    const int u = r - W - 1 + hl;
                    if (!is_a) v = ld_row(Xs + min(max(u, 0), x_rows - 1) * kXS);   // (the A lanes keep their mean: a mean row
                    float g = dot16(ar, v);                                          // read beside seven frame rows conflicts)
                    if (!is_a && u < 0) g = 0.f;
                    // the DP lanes fetch A (negated mean: A = -dot) and their ten G values from their own half
                    const float A = -__shfl_sync(0xffffffffu, g, hbase + 12 + wj);
                    inv[(k + W) % NB] = Invw[min(r + W - 1, inv_cols - 1)];   // column r+W-1 enters the band
                    float gv[NB];
#pragma unroll
                    for (int i = 0; i < NB; i++) gv[i] = __shfl_sync(0xffffffffu, g, hbase + wj + i);
                    if (r == 1) {
                        // row 1: columns c < 1 must stay +inf (they would otherwise inherit D[0][0])
#pragma unroll
                        for (int i = W; i < NB; i++) {
                            const float sim = (gv[i] - A) * inv[(k + i + 1 + NB - W) % NB];
                            const float best = min3(i + 1 < NB ? D[i + 1] : INFINITY, D[i], D[i - 1]);
                            const float v = (1.f - sim) + best;
                            D[i] = (!masked || ((band_mask >> i) & 1u)) ? v : INFINITY;
                        }
                        D[W - 1] = INFINITY;
                    } else {
#pragma unroll
                        for (int i = 0; i < NB; i++) {
                            const float sim = (gv[i] - A) * inv[(k + i + 1 + NB - W) % NB];
                            const float best = min3(i + 1 < NB ? D[i + 1] : INFINITY, D[i], i > 0 ? D[i - 1] : INFINITY);
                            const float v = (1.f - sim) + best;
                            D[i] = (!masked || ((band_mask >> i) & 1u)) ? v : INFINITY;
                        }
                    }
                }
            }
        }
        if (hl < n_win && b < a.n_streams) {
            // D[m-1][m] = band offset W+1 of row m-1 (dtw.rs:101); m == 1 has no such cell -> +inf
            const float cost = m >= 2 ? D[W + 1] : INFINITY;
            const float normalized = __fdiv_rn(cost, (float)(2 * m));
            const float score = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
            a.scores[(b * a.n_new + (j0 + hl)) * a.n_slots + s] = score;
        }
        __syncwarp();   // Mu / Inv are rewritten for the warp's next slot
    }
}

}  // namespace

int dtw_windows_cadence_max_new() { return 24; }

bool dtw_windows_cadence_supported(int d, int band, int max_slot_len, int window_len) {
    return d >= 1 && d <= kD && band >= 1 && band <= W && max_slot_len >= 1 && max_slot_len <= window_len;
}

cudaError_t launch_dtw_windows_cadence(const DtwWindowsArgs& a, const float* tmpl_unit, const int64_t* unit_off, cudaStream_t stream) {
    if (!dtw_windows_cadence_supported(a.d, a.band, a.max_len, a.window_len > 0 ? a.window_len : a.max_len)) return cudaErrorInvalidValue;
    const int n = a.n_new - a.first_window;
    if (n <= 0 || a.n_streams <= 0 || a.n_slots <= 0) return cudaSuccess;
    const int triples = (n + NWIN - 1) / NWIN;
    const int64_t ctas = ((a.n_streams + 1) / 2) * (int64_t)triples;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int n_warps = a.n_slots < kMaxWarps ? a.n_slots : kMaxWarps;
    const int x_rows = NWIN + a.max_len + W + 1;
    const int inv_cols = a.max_len + W + 1;
    const size_t bytes = ((size_t)2 * x_tile_floats(x_rows) + (size_t)n_warps * 2 * NWIN * kD + (size_t)n_warps * 2 * NWIN * inv_cols) * sizeof(float);
    if (bytes > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(dtw_windows_cadence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    dtw_windows_cadence_kernel<<<(unsigned)ctas, n_warps * 32, bytes, stream>>>(a, tmpl_unit, unit_off, triples, x_rows, inv_cols);
    return cudaGetLastError();
}

}  // namespace rp
