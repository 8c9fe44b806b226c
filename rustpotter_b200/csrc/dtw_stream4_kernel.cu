// K2s v4 — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Contract: reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105 (banded DTW with the
// asymmetric band [r-w, r+w-1], result cell D[m-1][n], cosine distance with similarity 0 for zero
// vectors, cost/(m+n) -> logistic score).
//
// Same arithmetic as dtw_stream3_kernel.cu (a block of 8 window columns lives in registers as negated
// unit vectors, the template streams past it two rows per step, FFMA2 dots with the DP chain of the
// previous row interleaved), but the systolic array is turned by 90 degrees: what v3 mapped to the five
// LANES of a pair is mapped to the four WARPS of a CTA, and a lane is one pair of a group of 32.
//
//  * Everything that differed between the lanes of a v3 warp (block index, band mask, block switches,
//    prefetch front) is now warp-uniform: it runs on the uniform datapath, costs no divergence, and steps
//    whose 8x2 cells lie fully inside the band take a code path without any mask (15 of a block's 24 steps).
//  * All 32 lanes work (v3: 30), a warp whose block is not in the band skips the step instead of issuing
//    masked work (v3 lanes could not), and the pipeline fill/drain of a group costs idle WARPS, which the
//    second resident CTA fills, instead of idle issue slots. Cells issued per pair: 4656 (3810 useful) against
//    6144 in v3.
//  * Block b waits sigma steps (2 for windows 17..20, else 1) behind block b-1, so one warp owns blocks
//    b, b+4, b+8, .. back to back with no idle step in between. Boundary columns travel through a 4-deep
//    exchange array in shared memory; the CTA meets at ONE barrier per step.
//  * HBM -> registers -> shared: every step the 128 threads fetch one template row pair of all 32 pairs (and,
//    four steps out of P, a quarter of the next window block) with 32-byte loads, scale it to unit length one
//    step later and store it to the ring. No cp.async raw copy, no in-ring rewrite: shared-memory traffic
//    drops from 81 KB to 53 KB per pair. Four threads cover 128 contiguous bytes of one pair.
//  * A block switch is 32 conflict-free LDS.128 (the staged block is already negated and normalised).
//
// Shapes: d == 16, uniform m >= 2, n >= 1, no CMN, 3 <= window = max(band, |m-n|) <= 20. Everything else
// takes the older kernels. The static schedule is checked on the CPU by tools/sim_stream4_schedule.py.
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int CB = 8;                                 // window columns per block
constexpr int NW = 4;                                 // warps per CTA = blocks of one pair in flight
constexpr int PPG = 32;                               // pairs per group (one per lane)
constexpr int SLOTS = 16;                             // ring slots (template row pairs) per pair
constexpr int SLOT_F = 2 * kD;                        // floats per slot: rows 2k-1, 2k
constexpr int RING_PAIR_F = SLOTS * SLOT_F + 4;       // +16 bytes: consecutive pairs rotate one bank group
constexpr int STAGE_PAIR_F = 2 * CB * kD + 4;         // two staged window blocks per pair (+16 bytes)
constexpr int XS = 4;                                 // exchange slots (row pairs) per warp
constexpr int RING_F = PPG * RING_PAIR_F;
constexpr int STAGE_F = PPG * STAGE_PAIR_F;
constexpr int XCH_F = NW * XS * 2 * 32;
constexpr int XDRAIN_F = NW * 32;                     // one more value per warp and pair: the last row of a finished block
constexpr int SMEM_FLOATS = RING_F + STAGE_F + XCH_F + XDRAIN_F;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;           // 103,936 bytes: two CTAs per SM
constexpr int KPRO = 4;                               // row pairs fetched by the prologue
constexpr int MIN_WINDOW = 3, MAX_WINDOW = 20;

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
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 32 bytes of a stream that is read exactly once
__device__ __forceinline__ void ldg32(const float* p, ulonglong2& v0, ulonglong2& v1) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v0.x), "=l"(v0.y) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2+16];" : "=l"(v1.x), "=l"(v1.y) : "l"(p));
}

// One window column (16 floats as 8 pairs) -> NEGATED unit vector (zero columns stay zero).
__device__ __forceinline__ void unit_column(const f2 (&x)[8], f2 (&col)[8]) {
    f2 n2 = mul2(x[0], x[0]);
    f2 n3 = mul2(x[1], x[1]);
#pragma unroll
    for (int q = 2; q < 8; q += 2) {
        n2 = fma2(x[q], x[q], n2);
        n3 = fma2(x[q + 1], x[q + 1], n3);
    }
    const float nn = hsum(n2) + hsum(n3);
    const float s = nn > 0.f ? -rsqrtf(nn) : 0.f;
    const f2 s2 = pk(s, s);
#pragma unroll
    for (int q = 0; q < 8; q++) col[q] = mul2(x[q], s2);
}

// Block B (0-based columns 8B .. 8B+7 of the window) straight from global memory. Columns >= n repeat column
// n-1: their cells are computed but nothing that reaches D[m-1][n] reads them (dependencies only go left/up).
__device__ __forceinline__ void load_block_global(const float* __restrict__ win, int n, int B, f2 (&bcol)[CB][8]) {
#pragma unroll
    for (int j = 0; j < CB; j++) {
        const int c = min(B * CB + j, n - 1);
        f2 x[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = *(reinterpret_cast<const ulonglong2*>(win + (size_t)c * kD) + q);
            x[2 * q] = v.x;
            x[2 * q + 1] = v.y;
        }
        unit_column(x, bcol[j]);
    }
}

// The staged copy of a block is already negated and of unit length.
__device__ __forceinline__ void load_block_staged(const float* __restrict__ stage, f2 (&bcol)[CB][8]) {
#pragma unroll
    for (int j = 0; j < CB; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(stage + j * kD + 4 * q);
            bcol[j][2 * q] = v.x;
            bcol[j][2 * q + 1] = v.y;
        }
}

// Reads one 64-byte (already unit-length) template row from the ring.
__device__ __forceinline__ void load_row(const float* __restrict__ p, f2 (&ar)[8]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + 4 * q);
        ar[2 * q] = v.x;
        ar[2 * q + 1] = v.y;
    }
}

// Half a 64-byte vector (v0, v1) -> scaled by sign / |vector| (the partner lane holds the other half; a zero
// vector stays zero: similarity 0, distance 1) -> shared memory.
__device__ __forceinline__ void norm_store(const ulonglong2& v0, const ulonglong2& v1, float* __restrict__ dst, float sign) {
    const f2 s = fma2(v1.y, v1.y, fma2(v1.x, v1.x, fma2(v0.y, v0.y, mul2(v0.x, v0.x))));
    const float part = hsum(s);
    const float nn = part + __shfl_xor_sync(0xffffffffu, part, 1);
    const float sc = nn > 0.f ? sign * rsqrtf(nn) : 0.f;
    const f2 s2 = pk(sc, sc);
    ulonglong2 o0, o1;
    o0.x = mul2(v0.x, s2);
    o0.y = mul2(v0.y, s2);
    o1.x = mul2(v1.x, s2);
    o1.y = mul2(v1.y, s2);
    *reinterpret_cast<ulonglong2*>(dst) = o0;
    *reinterpret_cast<ulonglong2*>(dst + 4) = o1;
}

// One half-step: the dots of template row `ar` with the block's eight columns (acc), interleaved with the
// DP chain of the previous row: Ddst[j] = cprev[j] + min(Dsrc[j], diag, left).
__device__ __forceinline__ void half_step(const f2 (&ar)[8], const f2 (&bcol)[CB][8], f2 (&acc)[CB], const float (&cprev)[CB],
                                          const float (&Dsrc)[CB], float (&Ddst)[CB], float left, float diag, f2 one) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int j = 0; j < CB; j++) acc[j] = fma2(ar[q], bcol[j][q], q == 0 ? one : acc[j]);   // acc = (1, 0) - a^.b^
        const float up = Dsrc[q];
        const float v = cprev[q] + min3(up, diag, left);
        diag = up;
        left = v;
        Ddst[q] = v;
    }
}

// Band mask of a step: bit i <-> d = r - c = (2u-1) - c0 + (i - 7), valid iff -w+1 <= d <= w. Bits 7-j (row 2u-1) and
// 8-j (row 2u) are the cells of column c0+j, bits 8 and 9 the two cells of column c0-1 (the left neighbour's last).
__device__ __forceinline__ unsigned band_mask(int u, int c0, int w) {
    const int t = 2 * u - c0 + w - 2;
    const int lo = min(max(7 - t, 0), 10), hi = min(max(7 - t + 2 * w, 0), 10);
    return (1u << hi) - (1u << lo);
}

// One step of a block: template rows 2u-1 and 2u against the eight columns. FULL: every cell is inside the band.
template <bool FULL>
__device__ __forceinline__ void block_step(const float* __restrict__ ring_p, int u, unsigned M, float li1, float li2p, float li1_prev,
                                           const f2 (&bcol)[CB][8], f2 (&ar1)[8], float (&D1)[CB], float (&D2)[CB], float (&cost2)[CB],
                                           float& out1, float& out2, f2 one) {
    f2 ar2[8], acc[CB];
    float cost1[CB];
    // ---- H1: dots of row 2u-1, DP of row 2u-2
    load_row(ring_p + (u & (SLOTS - 1)) * SLOT_F + kD, ar2);
    half_step(ar1, bcol, acc, cost2, D1, D2, li2p, li1_prev, one);
    out2 = D2[CB - 1];
#pragma unroll
    for (int j = 0; j < CB; j++) cost1[j] = (FULL || ((M >> (7 - j)) & 1u)) ? hsum(acc[j]) : INFINITY;
    // ---- H2: dots of row 2u, DP of row 2u-1
    load_row(ring_p + ((u + 1) & (SLOTS - 1)) * SLOT_F, ar1);
    half_step(ar2, bcol, acc, cost1, D2, D1, li1, li2p, one);
    out1 = D1[CB - 1];
#pragma unroll
    for (int j = 0; j < CB; j++) cost2[j] = (FULL || ((M >> (8 - j)) & 1u)) ? hsum(acc[j]) : INFINITY;
}

__global__ void __launch_bounds__(NW * 32, 2) dtw_pairs_stream4_kernel(DtwPairsArgs a, int64_t n_groups, int window, int sigma) {
    extern __shared__ __align__(16) float smem[];
    float* const ring_all = smem;
    float* const stage_all = smem + RING_F;
    float* const xch_all = smem + RING_F + STAGE_F;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform
    const int fp = tid >> 2, part = tid & 3;                 // fetch role: 32-byte part `part` of pair `fp`'s 128-byte unit
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int w = window;
    const int n_blocks = (n + CB - 1) / CB;
    const int last_row = m - 1;                 // rows 1 .. m-1 (the result cell is D[m-1][n])
    const int half = (last_row + 1) / 2;        // row pairs that contain a needed row
    const int P = 4 + sigma;                    // steps between the starts of consecutive blocks
    const int steps = half + sigma * (n_blocks - 1);
    const int fin0 = 4 + (w + 1) / 2;           // block B is in the band for row pairs 4B + 1 - w/2 .. 4B + fin0
    const int kmax = (m + 1) / 2;               // row pairs that contain a template row
    const int owner_w = (n_blocks - 1) & (NW - 1);
    const int left_w = (wid + NW - 1) & (NW - 1);
    const f2 one = pk(1.f, 0.f);

    float* const ring_p = ring_all + lane * RING_PAIR_F;       // this lane's pair
    float* const stage_p = stage_all + lane * STAGE_PAIR_F;
    float* const ring_f = ring_all + fp * RING_PAIR_F + part * 8;    // where this thread stores what it fetches
    float* const stage_f = stage_all + fp * STAGE_PAIR_F + (part >> 1) * kD + (part & 1) * 8;
    float* const xch_w = xch_all + wid * (XS * 2 * 32) + lane;
    const float* const xch_r = xch_all + left_w * (XS * 2 * 32) + lane;
    float* const xdrain_w = xch_all + XCH_F + wid * 32 + lane;
    const float* const xdrain_r = xch_all + XCH_F + left_w * 32 + lane;

    for (int i = tid; i < SMEM_FLOATS; i += NW * 32) smem[i] = 0.f;
    __syncthreads();

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPG + lane;
        const bool valid = p < a.n_pairs;
        const int64_t pc = valid ? p : a.n_pairs - 1;                       // clamped: every lane computes on real data
        const int64_t pf = min(grp * PPG + fp, a.n_pairs - 1);
        const float* const win = a.win + (a.win_off ? a.win_off[pc] : pc * (int64_t)n * kD);
        const float* const winf = a.win + (a.win_off ? a.win_off[pf] : pf * (int64_t)n * kD);
        const float* const tmplf = a.tmpl + (a.tmpl_off ? a.tmpl_off[pf] : pf * (int64_t)m * kD);
        const int my_row = part >> 1;   // this thread fetches (half of) row 2k-1+my_row of row pair k

        // ---- prologue: row pairs 1 .. KPRO through the fetch path, the warp's first block straight from global
        {
            ulonglong2 pr[KPRO][2];
#pragma unroll
            for (int k = 1; k <= KPRO; k++) {
                pr[k - 1][0] = make_ulonglong2(0ull, 0ull);
                pr[k - 1][1] = make_ulonglong2(0ull, 0ull);
                if (2 * k - 1 + my_row <= m) ldg32(tmplf + (size_t)(2 * k - 2) * kD + part * 8, pr[k - 1][0], pr[k - 1][1]);
            }
            // L2: the row pairs the loop fetches first, window block NW of this group, and the head of this CTA's next group
            if (2 * (KPRO + 1 + part) - 1 <= m) prefetch_l2(tmplf + (size_t)(2 * (KPRO + 1 + part) - 2) * kD);
            if (2 * (KPRO + 5 + part) - 1 <= m) prefetch_l2(tmplf + (size_t)(2 * (KPRO + 5 + part) - 2) * kD);
            if (NW < n_blocks) prefetch_l2(winf + (size_t)min(NW * CB + 2 * part, n - 1) * kD);
            {
                const int64_t gn = grp + gridDim.x;
                if (gn < n_groups && !a.win_off && !a.tmpl_off) {
                    const int64_t pn = min(gn * PPG + fp, a.n_pairs - 1);
                    const float* wn = a.win + pn * (int64_t)n * kD;
                    const float* tn = a.tmpl + pn * (int64_t)m * kD;
                    if (2 * part + 1 <= m) prefetch_l2(tn + (size_t)(2 * part) * kD);
#pragma unroll
                    for (int b = 0; b < NW; b++) prefetch_l2(wn + (size_t)min(b * CB + 2 * part, n - 1) * kD);
                }
            }
#pragma unroll
            for (int k = 1; k <= KPRO; k++) norm_store(pr[k - 1][0], pr[k - 1][1], ring_f + (k & (SLOTS - 1)) * SLOT_F, 1.f);
        }
        int B = wid;                      // warp-uniform: the block this warp works on
        f2 bcol[CB][8];
        if (B < n_blocks) {
            load_block_global(win, n, B, bcol);
        } else {
#pragma unroll
            for (int j = 0; j < CB; j++)
#pragma unroll
                for (int q = 0; q < 8; q++) bcol[j][q] = 0ull;
        }
        float D1[CB], D2[CB], cost2[CB];
#pragma unroll
        for (int j = 0; j < CB; j++) {
            D1[j] = INFINITY;
            D2[j] = INFINITY;
            cost2[j] = INFINITY;
        }
        float out1 = INFINITY, out2 = INFINITY;
        float li1_prev = INFINITY;
        float dseed = (wid == 0) ? 0.f : INFINITY;   // D[0][0], the diagonal input of cell (1,1)
        bool ok2_prev = false;
        f2 ar1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) ar1[q] = 0ull;

        // fetch state (CTA-uniform)
        int kf = KPRO + 1;                 // next template row pair to fetch
        int qB = NW, qj = 0;               // next window block quarter to fetch
        bool rows_pending = false, cols_pending = false;
        ulonglong2 pend_r0 = make_ulonglong2(0ull, 0ull), pend_r1 = pend_r0, pend_c0 = pend_r0, pend_c1 = pend_r0;
        float* dst_r = ring_f;
        float* dst_c = stage_f;
        __syncthreads();                   // prologue rows visible; previous group's ring reads are over

        for (int st = 1; st <= steps; st++) {
            // ---- fetch duties of all 128 threads: finish last step's loads, issue this step's
            if (rows_pending) norm_store(pend_r0, pend_r1, dst_r, 1.f);
            if (cols_pending) norm_store(pend_c0, pend_c1, dst_c, -1.f);
            rows_pending = false;
            cols_pending = false;
            if (kf <= kmax) {
                const int bmin = kf > fin0 ? (kf - fin0 + 3) >> 2 : 0;   // first block that reads row pair kf
                if (st >= kf + sigma * bmin - 4) {
                    const float* src = tmplf + (size_t)(2 * kf - 2) * kD + part * 8;
                    pend_r0 = make_ulonglong2(0ull, 0ull);
                    pend_r1 = pend_r0;
                    if (2 * kf - 1 + my_row <= m) ldg32(src, pend_r0, pend_r1);
                    if (part == 0 && 2 * (kf + 8) - 1 <= m) prefetch_l2(src + 8 * SLOT_F);
                    dst_r = ring_f + (kf & (SLOTS - 1)) * SLOT_F;
                    rows_pending = true;
                    kf++;
                }
            }
            if (qB < n_blocks && st >= P * (qB - NW) + fin0 - 6 + qj) {
                const int c = min(qB * CB + 2 * qj + my_row, n - 1);
                ldg32(winf + (size_t)c * kD + (part & 1) * 8, pend_c0, pend_c1);
                if (part == 0 && qB + 1 < n_blocks) prefetch_l2(winf + (size_t)min((qB + 1) * CB + 2 * qj, n - 1) * kD);
                dst_c = stage_f + (qB & 1) * (CB * kD) + 2 * qj * kD;
                cols_pending = true;
                if (++qj == 4) {
                    qj = 0;
                    qB++;
                }
            }

            // ---- this warp's block
            const int u = st - sigma * B;
            const int ufirst = max(1, 4 * B + 1 - (w >> 1));
            const int ulast = 4 * B + fin0;
            if (B < n_blocks && u >= ufirst && u <= ulast) {
                const int c0 = B * CB + 1;
                if (u == ufirst) {   // fresh block: no look-ahead happened, and the step before it was skipped
                    load_row(ring_p + (u & (SLOTS - 1)) * SLOT_F, ar1);
                    ok2_prev = ((band_mask(u - 1, c0, w) >> 9) & 1u) && B > 0;
                }
                const unsigned M = band_mask(u, c0, w);
                const bool ok1 = ((M >> 8) & 1u) && B > 0;
                const bool ok2 = ((M >> 9) & 1u) && B > 0;
                float shf1 = INFINITY, shf2 = INFINITY;
                if (B > 0) {
                    // D[2u-2][c0-1]: the left block's H1 of row pair u, or its drain if row pair u-1 was its last
                    shf2 = (u == 4 * (B - 1) + fin0 + 1) ? xdrain_r[0] : xch_r[(u & (XS - 1)) * 64];
                    shf1 = xch_r[(u & (XS - 1)) * 64 + 32];
                }
                const float li2p = ok2_prev ? shf2 : dseed;   // left input of row 2u-2 = diagonal input of row 2u-1
                const float li1 = ok1 ? shf1 : INFINITY;      // left input of row 2u-1 = diagonal input of row 2u
                dseed = INFINITY;
                if ((M & 0x1ffu) == 0x1ffu)
                    block_step<true>(ring_p, u, M, li1, li2p, li1_prev, bcol, ar1, D1, D2, cost2, out1, out2, one);
                else
                    block_step<false>(ring_p, u, M, li1, li2p, li1_prev, bcol, ar1, D1, D2, cost2, out1, out2, one);
                li1_prev = li1;
                ok2_prev = ok2;
                xch_w[(u & (XS - 1)) * 64] = out2;
                xch_w[(u & (XS - 1)) * 64 + 32] = out1;
                if (u == ulast && st < steps) {   // block finished
                    {   // DP of its last row (2*ulast): nothing to its left is in the band any more; the right neighbour reads D2[7]
                        float left = INFINITY, diag = li1_prev;
#pragma unroll
                        for (int j = 0; j < CB; j++) {
                            const float up = D1[j];
                            const float v = cost2[j] + min3(up, diag, left);
                            diag = up;
                            left = v;
                        }
                        xdrain_w[0] = left;
                    }
                    // the warp's next block, already staged (negated, unit length)
                    B += NW;
                    if (B < n_blocks) load_block_staged(stage_p + (B & 1) * (CB * kD), bcol);
#pragma unroll
                    for (int j = 0; j < CB; j++) {
                        D1[j] = INFINITY;
                        D2[j] = INFINITY;
                        cost2[j] = INFINITY;
                    }
                    li1_prev = INFINITY;
                    ok2_prev = false;
                }
            }
            __syncthreads();
        }

        // ---- result: the warp that owns the last block
        if (wid == owner_w) {
            const int Bl = n_blocks - 1;
            float res = INFINITY;
            if (B == Bl) {   // still on the last block: the result cell is inside its band
                if (!(last_row & 1)) {
                    // drain: DP of the last step's second row (row 2*half == m-1)
                    const int u = steps + 1 - sigma * Bl;
                    float left = dseed;
                    if (ok2_prev) left = (u == 4 * (Bl - 1) + fin0 + 1) ? xdrain_r[0] : xch_r[(u & (XS - 1)) * 64];
                    float diag = li1_prev;
#pragma unroll
                    for (int j = 0; j < CB; j++) {
                        const float up = D1[j];
                        const float v = cost2[j] + min3(up, diag, left);
                        diag = up;
                        left = v;
                        D2[j] = v;
                    }
                }
                const int jn = n - (Bl * CB + 1);
#pragma unroll
                for (int j = 0; j < CB; j++)
                    if (j == jn) res = (last_row & 1) ? D1[j] : D2[j];
            }
            if (valid) {
                const float normalized = __fdiv_rn(res, (float)(m + n));
                a.out[p] = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
            }
        }
        // the next group's prologue writes the ring: every warp is past its last read (barrier of the last step)
    }
}

}  // namespace

bool dtw_pairs_stream4_supported(const DtwPairsArgs& a) {
    if (a.d != kD || a.cmn || a.tmpl_len || a.win_len) return false;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    if (m < 2 || n < 1) return false;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    return window >= MIN_WINDOW && window <= MAX_WINDOW;
}

cudaError_t launch_dtw_pairs_stream4(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    const int sigma = (4 + window + 3) / 4 - 4 > 1 ? (4 + window + 3) / 4 - 4 : 1;
    const int64_t n_groups = (a.n_pairs + PPG - 1) / PPG;
    cudaError_t e = cudaFuncSetAttribute(dtw_pairs_stream4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream4_kernel, NW * 32, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > n_groups) blocks = n_groups;
    dtw_pairs_stream4_kernel<<<(unsigned)blocks, NW * 32, SMEM_BYTES, stream>>>(a, n_groups, window, sigma);
    return cudaGetLastError();
}

}  // namespace rp
