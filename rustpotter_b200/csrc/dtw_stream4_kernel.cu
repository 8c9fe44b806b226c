// K2s v4 — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Contract: reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105 (banded DTW with the
// asymmetric band [r-w, r+w-1], result cell D[m-1][n], cosine distance with similarity 0 for zero
// vectors, cost/(m+n) -> logistic score).
//
// Same arithmetic as dtw_stream3_kernel.cu (a block of 8 window columns lives in registers as negated
// unit vectors, the template streams past it two rows per step, FFMA2 dots with the DP chain of the
// previous row interleaved), but the systolic array is turned by 90 degrees and the CTA is split in roles:
//
//  * CONSUMER warps 0..3: what v3 mapped to the five LANES of a pair is mapped to four WARPS, and a lane is
//    one pair of a group of 32. Block index, band mask, block switches are warp-uniform: no divergence, and
//    steps whose 8x2 cells lie fully inside the band take a code path without any mask (15 of a block's 24
//    steps at window 20). All 32 lanes work (v3: 30); a warp whose block is outside the band skips the step
//    instead of issuing masked work; the pipeline fill/drain of a group costs idle WARPS, which the second
//    resident CTA fills, instead of idle issue slots. Cells issued per pair: 4656 (3810 useful), v3: 6144.
//    Block b runs two steps behind block b-1, so one warp owns blocks b, b+4, b+8, .. back to back, boundary
//    columns travel through a 4-deep exchange array in shared memory, and the consumers meet only every
//    SECOND step ("super-step").
//  * PRODUCER warps 4..7 (setmaxnreg hands their registers to the consumers: 32 vs 224): HBM -> registers ->
//    unit length -> shared memory. Per super-step they fetch the template row pairs and the quarters of the
//    next window block that a host-built static schedule (Stream4Sched, checked on the CPU by
//    tests/test_host_logic.py) assigns to it, 32 bytes per thread, four threads per 128 contiguous bytes of
//    one pair. No cp.async raw copy and no in-ring rewrite: shared-memory traffic is 53 KB per pair (v3: 81).
//    Producers run one super-step ahead of the consumers: two "full" and two "empty" named barriers
//    (bar.arrive / bar.sync, 256 threads) order the ring slots, so a late HBM line stalls nobody.
//  * A block switch is 32 conflict-free LDS.128 (the staged block is already negated and normalised).
//
// Shapes: d == 16, uniform m >= 2, n >= 1, no CMN, 3 <= window = max(band, |m-n|) <= 20, at most 238 steps.
// Everything else takes the older kernels.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int CB = 8;                                 // window columns per block
constexpr int NW = 4;                                 // consumer warps per CTA = blocks of one pair in flight
constexpr int NTHREADS = 2 * NW * 32;                 // + as many producer threads
constexpr int PPG = 32;                               // pairs per group (one per lane)
constexpr int SIGMA = 2;                              // block b runs SIGMA steps behind block b-1
constexpr int PITCH = 4 + SIGMA;                      // steps between the starts of consecutive blocks
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
constexpr int MIN_WINDOW = 3, MAX_WINDOW = 20;
constexpr int CONSUMER_REGS = 224, PRODUCER_REGS = 32;
constexpr int BAR_FULL = 1, BAR_EMPTY = 3;            // named barriers 1,2 (full) and 3,4 (empty)

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
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory"); }

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
__device__ __forceinline__ void norm_store(const ulonglong2& v0, const ulonglong2& v1, float* __restrict__ dst, float sign, bool store) {
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
    if (store) {
        *reinterpret_cast<ulonglong2*>(dst) = o0;
        *reinterpret_cast<ulonglong2*>(dst + 4) = o1;
    }
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

struct Geometry {
    int m, n, w, n_blocks, last_row, half, steps, n_super, fin0, kmax;
};

__device__ __forceinline__ Geometry make_geometry(const DtwPairsArgs& a, int window) {
    Geometry g;
    g.m = a.tmpl_len_max;
    g.n = a.win_len_max;
    g.w = window;
    g.n_blocks = (g.n + CB - 1) / CB;
    g.last_row = g.m - 1;                         // rows 1 .. m-1 (the result cell is D[m-1][n])
    g.half = (g.last_row + 1) / 2;                // row pairs that contain a needed row
    g.steps = g.half + SIGMA * (g.n_blocks - 1);
    g.n_super = (g.steps + 1) / 2;
    g.fin0 = 4 + (g.w + 1) / 2;                   // block B is in the band for row pairs 4B + 1 - w/2 .. 4B + fin0
    g.kmax = (g.m + 1) / 2;                       // row pairs that contain a template row
    return g;
}

// ------------------------------------------------------------------------------------------------ consumers
__device__ __forceinline__ void consumer_loop(const DtwPairsArgs& a, int64_t n_groups, const Geometry& g, float* smem) {
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform, 0..3
    const int m = g.m, n = g.n, w = g.w, n_blocks = g.n_blocks, fin0 = g.fin0, steps = g.steps;
    const int owner_w = (n_blocks - 1) & (NW - 1);
    const int left_w = (wid + NW - 1) & (NW - 1);
    const f2 one = pk(1.f, 0.f);
    const float* const ring_p = smem + lane * RING_PAIR_F;
    const float* const stage_p = smem + RING_F + lane * STAGE_PAIR_F;
    float* const xch_all = smem + RING_F + STAGE_F;
    float* const xch_w = xch_all + wid * (XS * 2 * 32) + lane;
    const float* const xch_r = xch_all + left_w * (XS * 2 * 32) + lane;
    float* const xdrain_w = xch_all + XCH_F + wid * 32 + lane;
    const float* const xdrain_r = xch_all + XCH_F + left_w * 32 + lane;
    unsigned gc = 0;   // super-steps done by this CTA's consumers (prologues included)

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPG + lane;
        const bool valid = p < a.n_pairs;
        const int64_t pc = valid ? p : a.n_pairs - 1;                       // clamped: every lane computes on real data
        const float* const win = a.win + (a.win_off ? a.win_off[pc] : pc * (int64_t)n * kD);

        // ---- super-step 0: the warp's first block straight from global memory (the producers fill the ring meanwhile)
        int B = wid;
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
        if (gc > 0) bar_sync(BAR_FULL + ((gc - 1) & 1));   // (keeps the barrier phases of the two roles in step)
        bar_arrive(BAR_EMPTY + (gc & 1));
        gc++;

        for (int S = 1; S <= g.n_super; S++) {
            bar_sync(BAR_FULL + ((gc - 1) & 1));   // what the producers stored for this super-step, and the neighbours' boundary values
#pragma unroll 1
            for (int st = 2 * S - 1; st <= min(2 * S, steps); st++) {
                const int u = st - SIGMA * B;
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
            }
            // this super-step's ring and stage reads are over (nobody waits for the CTA's very last one)
            if (S < g.n_super || grp + gridDim.x < n_groups) bar_arrive(BAR_EMPTY + (gc & 1));
            gc++;
        }

        // ---- result: the warp that owns the last block
        // the drain of the last row needs the left neighbour's last values: one more consumer-only rendezvous
        asm volatile("bar.sync %0, %1;" ::"n"(5), "n"(NW * 32) : "memory");
        if (wid == owner_w) {
            const int Bl = n_blocks - 1;
            float res = INFINITY;
            if (B == Bl) {
                if (!(g.last_row & 1)) {
                    const int u = steps + 1 - SIGMA * Bl;
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
                    if (j == jn) res = (g.last_row & 1) ? D1[j] : D2[j];
            }
            if (valid) {
                const float normalized = __fdiv_rn(res, (float)(m + n));
                a.out[p] = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ producers
// Unit codes of the schedule: 0 nothing; 1 .. kmax: template row pair k; 0x8000 | (B << 2) | j: quarter j of window block B.
__device__ __forceinline__ void producer_loop(const DtwPairsArgs& a, int64_t n_groups, const Geometry& g, float* smem, const Stream4Sched& sched) {
    const int tid = threadIdx.x - NW * 32;               // 0..127
    const int fp = tid >> 2, part = tid & 3;              // 32-byte part `part` of pair `fp`'s 128-byte unit
    const int my_row = part >> 1;                         // rows: (half of) row 2k-1+my_row; quarters: column 2j+my_row
    const int m = g.m, n = g.n, n_blocks = g.n_blocks;
    float* const ring_f = smem + fp * RING_PAIR_F + part * 8;
    float* const stage_f = smem + RING_F + fp * STAGE_PAIR_F + my_row * kD + (part & 1) * 8;
    unsigned gi = 0;   // batches done by this CTA's producers

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t pf = min(grp * PPG + fp, a.n_pairs - 1);
        const float* const winf = a.win + (a.win_off ? a.win_off[pf] : pf * (int64_t)n * kD);
        const float* const tmplf = a.tmpl + (a.tmpl_off ? a.tmpl_off[pf] : pf * (int64_t)m * kD);
        {   // L2: rows and the window block the first batches fetch, and the head of this CTA's next group
            if (2 * (part + 1) - 1 <= m) prefetch_l2(tmplf + (size_t)(2 * part) * kD);
            if (2 * (part + 5) - 1 <= m) prefetch_l2(tmplf + (size_t)(2 * (part + 4)) * kD);
            if (NW < n_blocks) prefetch_l2(winf + (size_t)min(NW * CB + 2 * part, n - 1) * kD);
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
        for (int c = 0; c <= g.n_super; c++) {
            // batch c: what the consumers first read in super-step c+1 (the last batch of a group is empty)
#pragma unroll
            for (int sb = 0; sb < 2; sb++) {
                const unsigned ua = sched.unit[c][2 * sb], ub = sched.unit[c][2 * sb + 1];
                ulonglong2 va0 = make_ulonglong2(0ull, 0ull), va1 = va0, vb0 = va0, vb1 = va0;
                float* da = ring_f;
                float* db = ring_f;
                float sa = 1.f, sb_sign = 1.f;
                if (ua) {
                    if (ua & 0x8000u) {
                        const int Bq = (ua & 0x7fffu) >> 2, j = ua & 3u;
                        const int col = min(Bq * CB + 2 * j + my_row, n - 1);
                        ldg32(winf + (size_t)col * kD + (part & 1) * 8, va0, va1);
                        if (part == 0 && Bq + 1 < n_blocks) prefetch_l2(winf + (size_t)min((Bq + 1) * CB + 2 * j, n - 1) * kD);
                        da = stage_f + (Bq & 1) * (CB * kD) + 2 * j * kD;
                        sa = -1.f;
                    } else {
                        const int k = (int)ua;
                        const float* src = tmplf + (size_t)(2 * k - 2) * kD + part * 8;
                        if (2 * k - 1 + my_row <= m) ldg32(src, va0, va1);
                        if (part == 0 && 2 * (k + 8) - 1 <= m) prefetch_l2(src + 8 * SLOT_F);
                        da = ring_f + (k & (SLOTS - 1)) * SLOT_F;
                    }
                }
                if (ub) {
                    if (ub & 0x8000u) {
                        const int Bq = (ub & 0x7fffu) >> 2, j = ub & 3u;
                        const int col = min(Bq * CB + 2 * j + my_row, n - 1);
                        ldg32(winf + (size_t)col * kD + (part & 1) * 8, vb0, vb1);
                        if (part == 0 && Bq + 1 < n_blocks) prefetch_l2(winf + (size_t)min((Bq + 1) * CB + 2 * j, n - 1) * kD);
                        db = stage_f + (Bq & 1) * (CB * kD) + 2 * j * kD;
                        sb_sign = -1.f;
                    } else {
                        const int k = (int)ub;
                        const float* src = tmplf + (size_t)(2 * k - 2) * kD + part * 8;
                        if (2 * k - 1 + my_row <= m) ldg32(src, vb0, vb1);
                        if (part == 0 && 2 * (k + 8) - 1 <= m) prefetch_l2(src + 8 * SLOT_F);
                        db = ring_f + (k & (SLOTS - 1)) * SLOT_F;
                    }
                }
                // the slots this batch overwrites were last read in super-step c-1 at the latest (Stream4Sched invariant)
                if (sb == 0 && gi > 0) bar_sync(BAR_EMPTY + ((gi - 1) & 1));
                if (ua | ub) {   // uniform; the shuffle inside needs every lane
                    norm_store(va0, va1, da, sa, ua != 0);
                    norm_store(vb0, vb1, db, sb_sign, ub != 0);
                }
            }
            __threadfence_block();
            if (c < g.n_super || grp + gridDim.x < n_groups) bar_arrive(BAR_FULL + (gi & 1));   // (the CTA's very last batch is empty)
            gi++;
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) dtw_pairs_stream4_kernel(DtwPairsArgs a, int64_t n_groups, int window, Stream4Sched sched) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < SMEM_FLOATS; i += NTHREADS) smem[i] = 0.f;
    __syncthreads();
    const Geometry g = make_geometry(a, window);
    if (threadIdx.x >= NW * 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        producer_loop(a, n_groups, g, smem, sched);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
        consumer_loop(a, n_groups, g, smem);
    }
}

}  // namespace

// Static schedule of the producer warps (see the kernel header). Batch c (0 .. n_super-1) is stored while the consumers
// run super-step c (steps 2c-1, 2c; c = 0 is their prologue) and is first read in super-step c+1. Invariants, checked
// here and again by tests/test_host_logic.py through rp_debug_stream4_schedule:
//  (1) row pair k is stored before the step that reads it first, look-ahead included: batch <= (st_first(k) - 2) / 2;
//  (2) ring slot k % 16 is free: the last read of row pair k-16 lies in a super-step before the batch's;
//  (3) the four quarters of window block B >= 4 are stored before the super-step of the step after which its warp
//      switches to it, and after the super-step in which block B-2 (same staging slot) was read.
bool build_stream4_schedule(int m, int n, int band, Stream4Sched* out, int* n_super_out) {
    Stream4Sched s;
    std::memset(&s, 0, sizeof(s));
    const int diff = m > n ? m - n : n - m;
    const int w = band > diff ? band : diff;
    if (m < 2 || n < 1 || w < MIN_WINDOW || w > MAX_WINDOW) return false;
    const int n_blocks = (n + CB - 1) / CB;
    const int half = m / 2;
    const int steps = half + SIGMA * (n_blocks - 1);
    const int n_super = (steps + 1) / 2;
    if (n_super + 1 > STREAM4_MAX_BATCHES) return false;
    const int fin0 = 4 + (w + 1) / 2;
    const int kmax = (m + 1) / 2;
    auto b_min = [&](int k) { return k > fin0 ? (k - fin0 + 3) / 4 : 0; };
    auto b_max = [&](int k) { const int b = (k - 1 + w / 2) / 4; return b < n_blocks - 1 ? b : n_blocks - 1; };
    auto st_first = [&](int k) { return k + SIGMA * b_min(k); };
    auto st_last = [&](int k) { return k + SIGMA * b_max(k); };
    auto s_sw = [&](int B) { return PITCH * (B - NW) + fin0; };      // step after which block B replaces block B-4
    auto c_due = [&](int B) { return (s_sw(B) + 1) / 2 - 1; };         // last batch that may store block B
    int kf = 1, qB = NW, qj = 0;
    for (int c = 0; c < n_super; c++) {
        int slot = 0;
        while (kf <= kmax && st_first(kf) <= 2 * c + 3) {
            if (slot == 4) return false;
            if (kf > SLOTS && (st_last(kf - SLOTS) + 1) / 2 > c - 1) return false;   // (2)
            s.unit[c][slot++] = (unsigned short)kf++;
        }
        while (qB < n_blocks && slot < 4 && c >= c_due(qB) - 2) {
            if (c > c_due(qB)) return false;                                            // (3) too late
            if (qB - 2 >= NW && (s_sw(qB - 2) + 1) / 2 > c - 1) return false;           // (3) slot still in use
            s.unit[c][slot++] = (unsigned short)(0x8000u | (unsigned)(qB << 2) | (unsigned)qj);
            if (++qj == 4) {
                qj = 0;
                qB++;
            }
        }
    }
    // every row pair a step reads and every block a warp switches to inside the loop must have been scheduled
    for (int k = 1; k <= kmax && k <= half + 1; k++) {
        const bool read = b_min(k) <= b_max(k) && st_first(k) <= steps + 1;
        if (read && k >= kf) return false;
    }
    for (int B = NW; B < n_blocks; B++)
        if (s_sw(B) < steps && B >= qB) return false;
    if (out) *out = s;
    if (n_super_out) *n_super_out = n_super;
    return true;
}

bool dtw_pairs_stream4_supported(const DtwPairsArgs& a) {
    if (a.d != kD || a.cmn || a.tmpl_len || a.win_len) return false;
    return build_stream4_schedule(a.tmpl_len_max, a.win_len_max, a.band, nullptr, nullptr);
}

cudaError_t launch_dtw_pairs_stream4(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    Stream4Sched sched;
    if (!build_stream4_schedule(m, n, a.band, &sched, nullptr)) return cudaErrorInvalidValue;
    const int64_t n_groups = (a.n_pairs + PPG - 1) / PPG;
    cudaError_t e = cudaFuncSetAttribute(dtw_pairs_stream4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream4_kernel, NTHREADS, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > n_groups) blocks = n_groups;
    dtw_pairs_stream4_kernel<<<(unsigned)blocks, NTHREADS, SMEM_BYTES, stream>>>(a, n_groups, window, sched);
    return cudaGetLastError();
}

}  // namespace rp
