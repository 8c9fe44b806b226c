// K2s — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Same contract as the generic kernel (reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105);
// this is the variant BASELINE.json's "DTW micro-bench" streams from HBM: every pair has its own
// template and window, nothing is shared, 14 084 algorithmic bytes per 120x16 / 100x16 pair.
//
// Mapping ("column-block systolic array"): a pair is scored by L = 6 lanes, 5 pairs per warp. The
// window's columns are cut into blocks of CB = 8; lane l owns blocks l, l+6, l+12 ... one at a time and
// keeps the block's 8 unit-normalised columns in registers (loaded straight from HBM, 512 contiguous
// bytes per block). Template rows stream through a small shared-memory ring (cp.async, 32 rows of 64 B
// per pair, prefetched 8 rows ahead). At step tau a lane handles row r = tau - B of its block B: one
// broadcast-free LDS of the 64-byte template row feeds 8 cells x 8 FFMA2, followed by the 8-cell DP
// chain in registers (min3 + add per cell). Block B needs D(r, last column of block B-1), which the
// previous lane produced one step earlier: one warp shuffle per step. The skew of one step per block
// makes all six lanes of a pair busy on six consecutive rows, so only ~6 template rows per pair are
// live at any time — which is what lets ten warps per SM stay resident.
//
// Band (dtw.rs:64-78): window = max(band, |m-n|); row r touches columns [max(1,r-window),
// min(n,r+window-1)]; a block is active for the 2*window+7 rows where any of its columns is in the
// band, which must fit the (CB+1)*L = 54-step period of a lane: window <= 23. The returned cell is
// D[m-1][n] (dtw.rs:101). Cosine distance with similarity 0 for zero vectors (comparator.rs:42-47).
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int CB = 8;      // columns per block
constexpr int L = 6;       // lanes per pair
constexpr int PPW = 5;     // pairs per warp (30 of 32 lanes)
constexpr int RING = 32;   // template rows in the ring
constexpr int LA = 8;      // prefetch distance (rows)
constexpr int RS = 20;     // ring row stride in floats (80 B: the six rows a pair reads hit disjoint banks)
constexpr int PAIR_FLOATS = RING * RS + 8;  // +8 floats rotates the bank phase of consecutive pairs by two 16 B groups

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
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Loads block B (columns 8B+1 .. 8B+8 of the window, 1-based) as NEGATED unit vectors.
__device__ __forceinline__ void load_block(const float* __restrict__ win, int n, int B, f2 (&bcol)[CB][8]) {
#pragma unroll
    for (int j = 0; j < CB; j++) {
        const int c0 = B * CB + j;  // 0-based column
        f2 x[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = c0 < n ? __ldg(reinterpret_cast<const ulonglong2*>(win + (size_t)c0 * kD) + q) : make_ulonglong2(0ull, 0ull);
            x[2 * q] = v.x;       // little endian: .x = floats (0,1) of the 16-byte chunk
            x[2 * q + 1] = v.y;
        }
        f2 n2 = mul2(x[0], x[0]);
#pragma unroll
        for (int q = 1; q < 8; q++) n2 = fma2(x[q], x[q], n2);
        const float nn = hsum(n2);
        const float s = nn > 0.f ? -rsqrtf(nn) : 0.f;
        const f2 s2 = pk(s, s);
#pragma unroll
        for (int q = 0; q < 8; q++) bcol[j][q] = mul2(x[q], s2);
    }
}

// Per-block constants of a lane (recomputed only when the lane moves to its next block).
struct BlockInfo {
    int c0;     // first column of the block, 1-based
    int rlo, rhi;  // active rows
    int jmax;   // last valid column offset inside the block (n may end inside it)
    bool left_possible;  // c0 - 1 >= 1
};
__device__ __forceinline__ BlockInfo block_info(int B, int n, int n_blocks, int w, int last_row, bool valid) {
    BlockInfo bi;
    bi.c0 = B * CB + 1;
    bi.rlo = max(1, bi.c0 + 1 - w);
    bi.rhi = min(last_row, bi.c0 + CB - 1 + w);
    if (!valid || B >= n_blocks) {  // never active
        bi.rlo = 1 << 29;
        bi.rhi = -1;
    }
    bi.jmax = min(CB - 1, n - bi.c0);
    bi.left_possible = bi.c0 - 1 >= 1;
    return bi;
}

__global__ void __launch_bounds__(32) dtw_pairs_stream_kernel(DtwPairsArgs a, int64_t n_groups, int window) {
    __shared__ __align__(16) float ring_all[PPW * PAIR_FLOATS];
    const int lane = threadIdx.x;
    const int g = lane / L, l = lane - g * L;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int w = window;
    const int n_blocks = (n + CB - 1) / CB;
    const int last_row = m - 1;                      // rows 1 .. m-1 (result cell D[m-1][n])
    const int steps = last_row + (n_blocks - 1);
    const int left_lane = l == 0 ? lane + (L - 1) : lane - 1;
    float* ring = ring_all + (g < PPW ? g : 0) * PAIR_FLOATS;

    // costs of (row r, the 8 columns of the lane's block): 1 - a^_r . b^_c with the template row taken from
    // the ring; evaluated unconditionally (inactive lanes read a clamped slot and discard the result)
    auto costs = [&](int r, const f2 (&bcol)[CB][8], float (&cost)[CB]) {
        const float* arow = ring + (r & (RING - 1)) * RS;
        f2 ar[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(arow + 4 * q);
            ar[2 * q] = v.x;
            ar[2 * q + 1] = v.y;
        }
        f2 na2 = mul2(ar[0], ar[0]);
#pragma unroll
        for (int q = 1; q < 8; q++) na2 = fma2(ar[q], ar[q], na2);
        const float na = hsum(na2);
        const float inva = na > 0.f ? rsqrtf(na) : 0.f;
#pragma unroll
        for (int j = 0; j < CB; j++) {
            f2 acc = mul2(ar[0], bcol[j][0]);
#pragma unroll
            for (int q = 1; q < 8; q++) acc = fma2(ar[q], bcol[j][q], acc);
            cost[j] = fmaf(hsum(acc), inva, 1.f);   // 1 - a^.b^  (bcol is negated)
        }
    };

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPW + g;
        const bool valid = g < PPW && p < a.n_pairs;
        const float* tmpl = a.tmpl + (a.tmpl_off && valid ? a.tmpl_off[p] : (valid ? p : 0) * (int64_t)m * kD);
        const float* win = a.win + (a.win_off && valid ? a.win_off[p] : (valid ? p : 0) * (int64_t)n * kD);

        // ---- prologue: template rows 1..LA into the ring (lanes 0..3 of each pair move 16 B each)
        __syncwarp();  // every lane is done with the previous group's ring
#pragma unroll
        for (int row = 1; row <= LA; row++) {
            if (valid && l < 4 && row <= last_row) cp_async16(ring + (row & (RING - 1)) * RS + 4 * l, tmpl + (size_t)(row - 1) * kD + 4 * l);
            cp_async_commit();
        }

        int B = l;                     // current block of this lane
        BlockInfo bi = block_info(B, n, n_blocks, w, last_row, valid);
        f2 bcol[CB][8];
        if (valid && B < n_blocks) {
            load_block(win, n, B, bcol);
            if (B + L < n_blocks) {    // warm L2 with this lane's next block (512 contiguous bytes)
#pragma unroll
                for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L) * CB * kD + 32 * k);
            }
        } else {
#pragma unroll
            for (int j = 0; j < CB; j++)
#pragma unroll
                for (int q = 0; q < 8; q++) bcol[j][q] = 0ull;
        }
        float Dcol[CB];
#pragma unroll
        for (int j = 0; j < CB; j++) Dcol[j] = INFINITY;
        float diag_in = (l == 0) ? 0.f : INFINITY;   // D[0][0] for the very first cell of block 0
        float out_last = INFINITY;
        float result = INFINITY;

        cp_async_wait<LA - 1>();       // row 1 has landed
        __syncwarp();
        float cost[CB];
        costs(max(1 - B, 1), bcol, cost);  // costs of step tau = 1

        for (int tau = 1; tau <= steps; tau++) {
            cp_async_wait<LA - 2>();   // row tau+1 has landed for its issuing lane ...
            __syncwarp();              // ... and is visible to the pair's other lanes
            const float shf = __shfl_sync(0xffffffffu, out_last, left_lane);
            const int r = tau - B;
            const bool active = r >= bi.rlo && r <= bi.rhi;
            // left neighbour D(r, c0-1), produced by the previous lane one step ago: in band iff r >= 1 and
            // max(1, r-w) <= c0-1 <= r+w-1. It is also the diagonal input of row r+1, so it is tracked on
            // every step (a block becomes active one row after its left neighbour column does).
            const bool left_ok = bi.left_possible && r >= 1 && (bi.c0 - 1 >= r - w) && (bi.c0 - 1 <= r + w - 1);
            const float left_in = left_ok ? shf : INFINITY;
            // in-band cells of this row inside the block: offsets jlo..jhi, as a bit mask (0 when inactive)
            const int jlo = max(0, r - w - bi.c0);
            const int jhi = min(bi.jmax, r + w - 1 - bi.c0);
            const unsigned mask = active ? ((2u << jhi) - (1u << jlo)) : 0u;

            // ---- next step's costs (independent of the DP chain below: lets the FFMA2s fill its latency)
            float cost_next[CB];
            costs(max(r + 1, 1), bcol, cost_next);

            // ---- DP over the 8 cells of (row r, block B)
            float left = left_in, diag = diag_in;
#pragma unroll
            for (int j = 0; j < CB; j++) {
                const float up = Dcol[j];
                float v = cost[j] + min3(up, diag, left);
                v = (mask >> j) & 1u ? v : INFINITY;
                diag = up;
                left = v;
                Dcol[j] = v;
            }
            out_last = Dcol[CB - 1];
            if (r == last_row) {
                const int jn = n - bi.c0;   // column n inside this block?
#pragma unroll
                for (int j = 0; j < CB; j++)
                    if (j == jn) result = Dcol[j];
            }
            diag_in = left_in;
#pragma unroll
            for (int j = 0; j < CB; j++) cost[j] = cost_next[j];

            if (active && r == bi.rhi) {   // block finished: move to this lane's next block
                B += L;
                bi = block_info(B, n, n_blocks, w, last_row, valid);
                if (B < n_blocks) {
                    load_block(win, n, B, bcol);
                    if (B + L < n_blocks) {
#pragma unroll
                        for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L) * CB * kD + 32 * k);
                    }
                }
#pragma unroll
                for (int j = 0; j < CB; j++) Dcol[j] = INFINITY;
                diag_in = INFINITY;
                // out_last is NOT reset: the right-hand neighbour reads D(rhi, last column) on the next step.
                costs(max(tau + 1 - B, 1), bcol, cost);  // the pipelined costs were made with the old block
            }
            // prefetch template row tau + LA into the ring
            const int row = tau + LA;
            if (valid && l < 4 && row <= last_row) cp_async16(ring + (row & (RING - 1)) * RS + 4 * l, tmpl + (size_t)(row - 1) * kD + 4 * l);
            cp_async_commit();
            if (tau == steps / 2 && grp + gridDim.x < n_groups) {
                // warm L2 with the next group's first blocks and template head
                const int64_t pn = p + (int64_t)gridDim.x * PPW;
                if (g < PPW && pn < a.n_pairs && !a.win_off) {
                    const float* wn = a.win + pn * (int64_t)n * kD + (size_t)l * CB * kD;
#pragma unroll
                    for (int k = 0; k < 4; k++) prefetch_l2(wn + 32 * k);
                    if (l < 4) prefetch_l2(a.tmpl + pn * (int64_t)m * kD + 32 * l);
                }
            }
        }
        cp_async_wait<0>();
        float mine = INFINITY;  // exactly one lane of the pair holds the result; the others hold +inf
#pragma unroll
        for (int k = 0; k < L; k++) mine = fminf(mine, __shfl_sync(0xffffffffu, result, min(g * L + k, 31)));
        if (valid && l == 0) {
            const float cst = m >= 2 ? mine : INFINITY;
            const float normalized = __fdiv_rn(cst, (float)(m + n));
            a.out[p] = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Two-rows-per-step variant: a step handles template rows (2u-1, 2u) of the lane's block, i.e. a
// 2 x 8 tile of cells = 128 FFMA2 for 8 LDS.128 + 2 shuffles + 2 band masks. Five lanes per pair
// (a block's 2*window+7 rows take <= 25 steps, the lane period is 5*L = 25 steps: window <= 21),
// six pairs per warp.
constexpr int L2 = 5;
constexpr int PPW2 = 6;
constexpr int LAS = 3;     // prefetch distance in steps (two rows each)
constexpr int PAIR_FLOATS2 = RING * RS + 4;  // +4 floats: best bank phase found for 5 lanes reading rows two apart

__global__ void __launch_bounds__(32) dtw_pairs_stream2_kernel(DtwPairsArgs a, int64_t n_groups, int window) {
    __shared__ __align__(16) float ring_all[PPW2 * PAIR_FLOATS2];
    const int lane = threadIdx.x;
    const int g = lane / L2, l = lane - g * L2;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int w = window;
    const int n_blocks = (n + CB - 1) / CB;
    const int last_row = m - 1;
    const int steps = (last_row + 1) / 2 + (n_blocks - 1);
    const int left_lane = l == 0 ? lane + (L2 - 1) : lane - 1;
    float* ring = ring_all + (g < PPW2 ? g : 0) * PAIR_FLOATS2;

    auto costs = [&](int r, const f2 (&bcol)[CB][8], float (&cost)[CB]) {
        const float* arow = ring + (r & (RING - 1)) * RS;
        f2 ar[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(arow + 4 * q);
            ar[2 * q] = v.x;
            ar[2 * q + 1] = v.y;
        }
        f2 na2 = mul2(ar[0], ar[0]);
#pragma unroll
        for (int q = 1; q < 8; q++) na2 = fma2(ar[q], ar[q], na2);
        const float na = hsum(na2);
        const float inva = na > 0.f ? rsqrtf(na) : 0.f;
#pragma unroll
        for (int j = 0; j < CB; j++) {
            f2 acc = mul2(ar[0], bcol[j][0]);
#pragma unroll
            for (int q = 1; q < 8; q++) acc = fma2(ar[q], bcol[j][q], acc);
            cost[j] = fmaf(hsum(acc), inva, 1.f);
        }
    };
    auto issue_rows = [&](const float* tmpl, bool valid, int step) {  // rows 2*step-1 and 2*step
#pragma unroll
        for (int k = 1; k >= 0; k--) {
            const int row = 2 * step - k;
            if (valid && l < 4 && row <= last_row)
                cp_async16(ring + (row & (RING - 1)) * RS + 4 * l, tmpl + (size_t)(row - 1) * kD + 4 * l);
        }
        cp_async_commit();
    };

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPW2 + g;
        const bool valid = g < PPW2 && p < a.n_pairs;
        const float* tmpl = a.tmpl + (a.tmpl_off && valid ? a.tmpl_off[p] : (valid ? p : 0) * (int64_t)m * kD);
        const float* win = a.win + (a.win_off && valid ? a.win_off[p] : (valid ? p : 0) * (int64_t)n * kD);

        __syncwarp();
#pragma unroll
        for (int st = 1; st <= LAS; st++) issue_rows(tmpl, valid, st);

        int B = l;
        BlockInfo bi = block_info(B, n, n_blocks, w, last_row, valid);
        f2 bcol[CB][8];
        if (valid && B < n_blocks) {
            load_block(win, n, B, bcol);
            if (B + L2 < n_blocks) {
#pragma unroll
                for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L2) * CB * kD + 32 * k);
            }
        } else {
#pragma unroll
            for (int j = 0; j < CB; j++)
#pragma unroll
                for (int q = 0; q < 8; q++) bcol[j][q] = 0ull;
        }
        float Dcol[CB];
#pragma unroll
        for (int j = 0; j < CB; j++) Dcol[j] = INFINITY;
        float diag_in = (l == 0) ? 0.f : INFINITY;   // D[0][0]
        float out1 = INFINITY, out2 = INFINITY;
        float result = INFINITY;

        for (int st = 1; st <= steps; st++) {
            cp_async_wait<LAS - 1>();   // rows 2*st-1, 2*st have landed
            __syncwarp();
            const float shf1 = __shfl_sync(0xffffffffu, out1, left_lane);
            const float shf2 = __shfl_sync(0xffffffffu, out2, left_lane);
            const int u = st - B;
            const int r1 = 2 * u - 1, r2 = 2 * u;
            const bool act1 = r1 >= bi.rlo && r1 <= bi.rhi;
            const bool act2 = r2 >= bi.rlo && r2 <= bi.rhi;
            const int cl = bi.c0 - 1;   // left neighbour column
            const bool ok1 = bi.left_possible && r1 >= 1 && cl >= r1 - w && cl <= r1 + w - 1;
            const bool ok2 = bi.left_possible && r2 >= 1 && cl >= r2 - w && cl <= r2 + w - 1;
            const float li1 = ok1 ? shf1 : INFINITY;
            const float li2 = ok2 ? shf2 : INFINITY;
            const int jlo1 = max(0, r1 - w - bi.c0), jhi1 = min(bi.jmax, r1 + w - 1 - bi.c0);
            const int jlo2 = max(0, r2 - w - bi.c0), jhi2 = min(bi.jmax, r2 + w - 1 - bi.c0);
            const unsigned mask1 = act1 ? ((2u << jhi1) - (1u << jlo1)) : 0u;
            const unsigned mask2 = act2 ? ((2u << jhi2) - (1u << jlo2)) : 0u;

            float c1[CB], c2[CB];
            costs(max(r1, 1), bcol, c1);
            costs(max(r2, 1), bcol, c2);

            float D1[CB];
            {
                float left = li1, diag = diag_in;
#pragma unroll
                for (int j = 0; j < CB; j++) {
                    const float up = Dcol[j];
                    float v = c1[j] + min3(up, diag, left);
                    v = (mask1 >> j) & 1u ? v : INFINITY;
                    diag = up;
                    left = v;
                    D1[j] = v;
                }
            }
            {
                float left = li2, diag = li1;
#pragma unroll
                for (int j = 0; j < CB; j++) {
                    const float up = D1[j];
                    float v = c2[j] + min3(up, diag, left);
                    v = (mask2 >> j) & 1u ? v : INFINITY;
                    diag = up;
                    left = v;
                    Dcol[j] = v;
                }
            }
            out1 = D1[CB - 1];
            out2 = Dcol[CB - 1];
            if (r1 == last_row || r2 == last_row) {
                const int jn = n - bi.c0;
#pragma unroll
                for (int j = 0; j < CB; j++)
                    if (j == jn) result = (r1 == last_row) ? D1[j] : Dcol[j];
            }
            diag_in = li2;

            if ((act1 || act2) && r2 >= bi.rhi) {   // block finished: next block of this lane
                B += L2;
                bi = block_info(B, n, n_blocks, w, last_row, valid);
                if (B < n_blocks) {
                    load_block(win, n, B, bcol);
                    if (B + L2 < n_blocks) {
#pragma unroll
                        for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L2) * CB * kD + 32 * k);
                    }
                }
#pragma unroll
                for (int j = 0; j < CB; j++) Dcol[j] = INFINITY;
                diag_in = INFINITY;
                // out1/out2 are kept: the right-hand neighbour reads them on the next step
            }
            issue_rows(tmpl, valid, st + LAS);
            if (st == steps / 2 && grp + gridDim.x < n_groups) {   // warm L2 for the next group
                const int64_t pn = p + (int64_t)gridDim.x * PPW2;
                if (g < PPW2 && pn < a.n_pairs && !a.win_off) {
                    const float* wn = a.win + pn * (int64_t)n * kD + (size_t)l * CB * kD;
#pragma unroll
                    for (int k = 0; k < 4; k++) prefetch_l2(wn + 32 * k);
                    if (l < 4) prefetch_l2(a.tmpl + pn * (int64_t)m * kD + 32 * l);
                }
            }
        }
        cp_async_wait<0>();
        float mine = INFINITY;
#pragma unroll
        for (int k = 0; k < L2; k++) mine = fminf(mine, __shfl_sync(0xffffffffu, result, min(g * L2 + k, 31)));
        if (valid && l == 0) {
            const float cst = m >= 2 ? mine : INFINITY;
            const float normalized = __fdiv_rn(cst, (float)(m + n));
            a.out[p] = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
        }
    }
}

}  // namespace

// 0 = automatic (v3 half-step kernel where it applies), 1 = one-row-per-step kernel, 2 = round-1 two-rows-per-step
// kernel (both kept for A/B measurements and for windows 21..23)
int g_stream_rows = 0;
void set_dtw_stream_rows(int rows) { g_stream_rows = rows; }

bool dtw_pairs_stream3_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream3(const DtwPairsArgs& a, cudaStream_t stream);
bool dtw_pairs_stream4_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream4(const DtwPairsArgs& a, cudaStream_t stream);

bool dtw_pairs_stream_supported(const DtwPairsArgs& a) {
    if (a.d != kD || a.cmn || a.tmpl_len || a.win_len) return false;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    if (m < 2 || n < 1) return false;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    if (2 * window + CB - 1 > (CB + 1) * L) return false;          // a block's active rows must fit the lane period
    const int n_blocks = (n + CB - 1) / CB;
    if (n_blocks - 1 + LA + 1 > RING) return false;                 // rows in flight must fit the ring
    return true;
}

cudaError_t launch_dtw_pairs_stream(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    // 0: newest kernel that supports the shape; 3: v3 (lanes of a warp as the systolic array); 1, 2: round-1 kernels
    if (g_stream_rows == 0 && dtw_pairs_stream4_supported(a)) return launch_dtw_pairs_stream4(a, stream);
    if ((g_stream_rows == 0 || g_stream_rows == 3) && dtw_pairs_stream3_supported(a)) return launch_dtw_pairs_stream3(a, stream);
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    const int64_t n_groups = (a.n_pairs + PPW - 1) / PPW;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_blocks = (n + CB - 1) / CB;
    // two-rows-per-step variant: window <= 21 and rows in flight (2*(n_blocks-1)+2 behind, 2*LAS ahead) fit the ring
    const bool two_rows = g_stream_rows != 1 && window <= 21 && 2 * (n_blocks - 1) + 2 + 2 * LAS <= RING;
    int per_sm = 0;
    cudaError_t e = two_rows ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream2_kernel, 32, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream_kernel, 32, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t groups = two_rows ? (a.n_pairs + PPW2 - 1) / PPW2 : n_groups;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > groups) blocks = groups;
    if (two_rows) dtw_pairs_stream2_kernel<<<(unsigned)blocks, 32, 0, stream>>>(a, groups, window);
    else dtw_pairs_stream_kernel<<<(unsigned)blocks, 32, 0, stream>>>(a, groups, window);
    return cudaGetLastError();
}

}  // namespace rp
