// K2s v4 — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Contract: reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105 (banded DTW with the
// asymmetric band [r-w, r+w-1], result cell D[m-1][n], cosine distance with similarity 0 for zero
// vectors, cost/(m+n) -> logistic score).
//
// Same arithmetic as the retired v3 kernel (a block of 8 window columns lives in registers as negated
// unit vectors, the template streams past it two rows per step, FFMA2 dots with the DP chain of the
// previous row interleaved), but the systolic array is turned by 90 degrees and the CTA is split in roles:
//
//  * CONSUMER warps 0..3: what v3 mapped to the five LANES of a pair is mapped to four WARPS, and a lane is
//    one pair of a group of 32. Block index, band mask, block switches are warp-uniform and come from a
//    host-built table of control words: no divergence, one straight-line step body of 277 instructions (128
//    FFMA2; the band mask is applied with R2P + FSEL, measured faster than a second, mask-free body for the 15
//    of a block's 24 steps that lie fully inside the band). All 32 lanes work (v3: 30); a warp whose block is outside the band skips the step
//    instead of issuing masked work; the pipeline fill/drain of a group costs idle WARPS, which the second
//    resident CTA fills, instead of idle issue slots. Cells issued per pair: 4656 (3810 useful), v3: 6144.
//    Block b runs two steps behind block b-1, so one warp owns blocks b, b+4, b+8, .. back to back, boundary
//    columns travel through a 4-deep exchange array in shared memory, and the consumers meet only every
//    SECOND step ("super-step").
//  * PRODUCER warps 4..7 (setmaxnreg hands their registers to the consumers: 56 vs 200): HBM -> registers ->
//    unit length -> shared memory. Per super-step they fetch the template row pairs and the quarters of the
//    next window block that a host-built static schedule (Stream4Sched, checked on the CPU by
//    tests/test_host_logic.py) assigns to it, 32 bytes per thread, four threads per 128 contiguous bytes of
//    one pair. No cp.async raw copy and no in-ring rewrite: shared-memory traffic is 53 KB per pair (v3: 81).
//    Producers run up to two super-steps ahead of the consumers: four "full" and four "empty" named barriers
//    (bar.arrive / bar.sync, 256 threads) order the ring slots, so a late HBM line stalls nobody.
//  * A block switch is 32 conflict-free LDS.128 (the staged block is already negated and normalised).
//
// Shapes: d == 16, uniform m >= 2, n >= 1, no CMN, 3 <= window = max(band, |m-n|) <= 20, at most 238 steps.
// Everything else takes the older kernels. Measured (B200, 1 M pairs 120x16 / 100x16): 6.46 ms = 2.18 TB/s algorithmic;
// what bounds it (FP32 issue, two consumer warps per scheduler) and the variants that were measured and dropped are in
// DESIGN.md section 6 and profiles/r01_SUMMARY.md.
#include <cfloat>
#include <cmath>
#include <algorithm>
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
constexpr int CONSUMER_REGS = 200, PRODUCER_REGS = 56;
constexpr int DEPTH = 2;                              // super-steps the producers may run ahead of the consumers
constexpr int BAR_FULL = 1, BAR_EMPTY = 5, BAR_CONSUMERS = 9, BAR_GROUP = 10;   // named barriers 1..4 (full), 5..8 (empty)

typedef unsigned long long f2;

// Phase accounting for tools/profile_k2s_phases.cu (never defined in the library build): cycles per warp role and phase.
#ifdef RP_K2S_PROFILE
__device__ unsigned long long g_prof[8][8];
__shared__ unsigned s_prof[8][8];   // per CTA; lane 0 of warp w owns row w; flushed once at the end of the kernel
#define PROF_NOW(x) unsigned x; asm volatile("mov.u32 %0, %%clock;" : "=r"(x) :: "memory")
// ... ordered after the computation of `var` (the compiler is otherwise free to move arithmetic across a clock read)
#define PROF_NOW_AFTER(x, var) unsigned x; asm volatile("mov.u32 %0, %%clock;" : "=r"(x), "+f"(var) :: "memory")
#define PROF_ADD(w, k, t0, t1) do { if ((threadIdx.x & 31) == 0) s_prof[w][k] += (unsigned)max((int)((t1) - (t0)), 0) >> 4; } while (0)
#define PROF_INC(w, k) do { if ((threadIdx.x & 31) == 0) s_prof[w][k] += 1u; } while (0)
#define PROF_FLUSH(w) do { if ((threadIdx.x & 31) == 0) for (int k_ = 0; k_ < 8; k_++) atomicAdd(&g_prof[w][k_], (unsigned long long)s_prof[w][k_]); } while (0)
#else
#define PROF_NOW(x)
#define PROF_NOW_AFTER(x, var)
#define PROF_ADD(w, k, t0, t1)
#define PROF_INC(w, k)
#define PROF_FLUSH(w)
#endif

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

// The same, predicated (warp-uniform predicate; keeps the caller's basic block in one piece).
__device__ __forceinline__ void load_row_if(const float* __restrict__ p, f2 (&ar)[8], bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.u64 {%0, %1}, [%2];\n\t}"
                     : "+l"(ar[2 * q]), "+l"(ar[2 * q + 1])
                     : "r"(sa + 16 * q), "r"((int)pred));
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

// Band mask of a step (built on the host, bits 12..21 of the control word): bit i <-> d = r - c = (2u-1) - c0 + (i - 7),
// valid iff -w+1 <= d <= w. Bits 7-j (row 2u-1) and 8-j (row 2u) are the cells of column c0+j, bits 8 and 9 the two cells
// of column c0-1 (the left neighbour's last).

// One step of a block: template rows 2u-1 and 2u (rp0 = ring slot of row pair u, rp1 = of u+1) against the eight columns.
// full: every cell is inside the band (warp-uniform); otherwise M masks the costs (bits 7-j: row 2u-1, 8-j: row 2u).
__device__ __forceinline__ void block_step(const float* __restrict__ rp0, const float* __restrict__ rp1, bool full, unsigned M, float li1,
                                           float li2p, float li1_prev, const f2 (&bcol)[CB][8], f2 (&ar1)[8], float (&D1)[CB],
                                           float (&D2)[CB], float (&cost2)[CB], float& out1, float& out2, f2 one) {
    f2 ar2[8], acc[CB];
    float cost1[CB];
    // ---- H1: dots of row 2u-1, DP of row 2u-2
    load_row(rp0 + kD, ar2);
    half_step(ar1, bcol, acc, cost2, D1, D2, li2p, li1_prev, one);
    out2 = D2[CB - 1];
#pragma unroll
    for (int j = 0; j < CB; j++) cost1[j] = hsum(acc[j]);
    if (!full) {
#pragma unroll
        for (int j = 0; j < CB; j++)
            if (!((M >> (7 - j)) & 1u)) cost1[j] = INFINITY;
    }
    // ---- H2: dots of row 2u, DP of row 2u-1
    load_row(rp1, ar1);
    half_step(ar2, bcol, acc, cost1, D2, D1, li1, li2p, one);
    out1 = D1[CB - 1];
#pragma unroll
    for (int j = 0; j < CB; j++) cost2[j] = hsum(acc[j]);
    if (!full) {
#pragma unroll
        for (int j = 0; j < CB; j++)
            if (!((M >> (8 - j)) & 1u)) cost2[j] = INFINITY;
    }
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
// Control word of one (warp, step), built on the host (build_stream4_schedule): the consumers' step prologue is one
// uniform load and a few bit tests instead of a serial chain of index arithmetic.
constexpr unsigned CTL_ACTIVE = 1u << 0;        // the warp's block is inside the band at this step
constexpr unsigned CTL_FRESH = 1u << 1;         // first step of a block: no look-ahead row was loaded
constexpr unsigned CTL_FULL = 1u << 2;          // all 16 cells inside the band: no masks
constexpr unsigned CTL_SWITCH = 1u << 3;        // last step of the block (and not the last step of the group)
constexpr unsigned CTL_OK1 = 1u << 4;           // cell (2u-1, c0-1) is inside the band and a left block exists
constexpr unsigned CTL_OK2 = 1u << 5;           // cell (2u,   c0-1) ...
constexpr unsigned CTL_OK2PREV = 1u << 6;       // (fresh steps) cell (2u-2, c0-1) ...
constexpr unsigned CTL_DRAIN_RD = 1u << 7;      // D[2u-2][c0-1] comes from the left warp's drain slot
constexpr unsigned CTL_HAS_LEFT = 1u << 8;      // block > 0
constexpr unsigned CTL_NEXT_EXISTS = 1u << 9;   // (switch steps) block + 4 exists
constexpr unsigned CTL_NEXT_PARITY = 1u << 10;  // (switch steps) its staging slot
constexpr unsigned CTL_NEXT_LAST = 1u << 11;    // (switch steps) it is the pair's last block
constexpr int CTL_MASK_SHIFT = 12;              // 10 bits: band mask
constexpr int CTL_SLOT_SHIFT = 22;              // 4 bits: row pair u & 15

__device__ __forceinline__ void consumer_loop(const DtwPairsArgs& a, int64_t n_groups, const Geometry& g, float* smem, const Stream4Sched& sched) {
    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform, 0..3
    const int m = g.m, n = g.n, n_blocks = g.n_blocks, steps = g.steps;
    const int owner_w = (n_blocks - 1) & (NW - 1);
    const int left_w = (wid + NW - 1) & (NW - 1);
    const f2 one = pk(1.f, 0.f);
    // shared-memory byte offsets, kept opaque so that they stay in registers instead of being recomputed from the thread id
    unsigned ring_o = (unsigned)(lane * RING_PAIR_F * 4);
    unsigned stage_o = (unsigned)((RING_F + lane * STAGE_PAIR_F) * 4);
    unsigned xw_o = (unsigned)((RING_F + STAGE_F + wid * (XS * 2 * 32) + lane) * 4);
    unsigned xr_o = (unsigned)((RING_F + STAGE_F + left_w * (XS * 2 * 32) + lane) * 4);
    unsigned dw_o = (unsigned)((RING_F + STAGE_F + XCH_F + wid * 32 + lane) * 4);
    unsigned dr_o = (unsigned)((RING_F + STAGE_F + XCH_F + left_w * 32 + lane) * 4);
    asm volatile("" : "+r"(ring_o), "+r"(stage_o), "+r"(xw_o), "+r"(xr_o), "+r"(dw_o), "+r"(dr_o));
    char* const sm = reinterpret_cast<char*>(smem);
    const float* const ring_p = reinterpret_cast<const float*>(sm + ring_o);
    const float* const stage_p = reinterpret_cast<const float*>(sm + stage_o);
    float* const xch_w = reinterpret_cast<float*>(sm + xw_o);
    const float* const xch_r = reinterpret_cast<const float*>(sm + xr_o);
    float* const xdrain_w = reinterpret_cast<float*>(sm + dw_o);
    const float* const xdrain_r = reinterpret_cast<const float*>(sm + dr_o);

    // (Rotating the warps' roles from group to group, so that the longest chain of blocks -- warp 0's, 84 of a group's 84
    // steps against 66-72 -- visits every scheduler, was measured: no change, profiles/r02_k2s_phases.txt.)
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        PROF_NOW(pg0);
        const int64_t p = grp * PPG + lane;
        const bool valid = p < a.n_pairs;
        const int64_t pc = valid ? p : a.n_pairs - 1;                       // clamped: every lane computes on real data
        const float* const win = a.win + (a.win_off ? a.win_off[pc] : pc * (int64_t)n * kD);

        // ---- super-step 0: the warp's first block straight from global memory (the producers fill the ring meanwhile)
        bool on_last = wid == n_blocks - 1;
        f2 bcol[CB][8];
        if (wid < n_blocks) {
            load_block_global(win, n, wid, bcol);
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
        if (0 <= g.n_super - 1 - DEPTH) bar_arrive(BAR_EMPTY);   // super-step 0 is over (nothing was read)

        unsigned ctl_next = sched.ctl[wid][1];
        PROF_NOW(pg1);
        PROF_ADD(wid, 0, pg0, pg1);   // prologue
        for (int S = 1; S <= g.n_super; S++) {
            PROF_NOW_AFTER(pw0, out1);
            bar_sync(BAR_FULL + ((S - 1) & 3));   // batch S-1 of the producers, and the neighbours' boundary values
#ifdef RP_K2S_PROFILE
            float probe = *reinterpret_cast<const volatile float*>(xdrain_r);   // BAR.SYNC defers blocking to the next memory access
#endif
            PROF_NOW_AFTER(pw1, probe);
            PROF_ADD(wid, 1, pw0, pw1);   // waiting at the super-step barrier
#pragma unroll 1
            for (int st = 2 * S - 1; st <= min(2 * S, steps); st++) {
                const unsigned ctl = ctl_next;
                ctl_next = sched.ctl[wid][st + 1];   // (the table has one spare entry)
                if (!(ctl & CTL_ACTIVE)) continue;
                const unsigned so = ((ctl >> CTL_SLOT_SHIFT) & 15u) * (SLOT_F * 4);   // byte offset of row pair u in the ring
                const unsigned so1 = (so + SLOT_F * 4) & (SLOTS * SLOT_F * 4 - 1);      // ... of row pair u+1
                const unsigned xo = ((ctl >> CTL_SLOT_SHIFT) & 3u) * 256;               // ... of row pair u in the exchange array
                const float* const rp0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ring_p) + so);
                const float* const rp1 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ring_p) + so1);
                // fresh block: no look-ahead happened, and the step before it was skipped
                load_row_if(rp0, ar1, (ctl & CTL_FRESH) != 0);
                ok2_prev = (ctl & CTL_FRESH) ? (ctl & CTL_OK2PREV) != 0 : ok2_prev;
                // D[2u-2][c0-1]: the left block's H1 of row pair u, or its drain if row pair u-1 was its last (block 0 reads
                // the last warp's values and ignores them)
                const float* x = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xch_r) + xo);
                // (the drain slot is read only in the step that uses it: its owner may be rewriting it in any other step)
                const float xd = (ctl & CTL_DRAIN_RD) ? xdrain_r[0] : 0.f, x0 = x[0], shf1 = x[32];
                const float shf2 = (ctl & CTL_DRAIN_RD) ? xd : x0;
                const float li2p = ok2_prev ? shf2 : dseed;              // left input of row 2u-2 = diagonal input of row 2u-1
                const float li1 = (ctl & CTL_OK1) ? shf1 : INFINITY;     // left input of row 2u-1 = diagonal input of row 2u
                dseed = INFINITY;
                block_step(rp0, rp1, (ctl & CTL_FULL) != 0, (ctl >> CTL_MASK_SHIFT) & 0x3ffu, li1, li2p, li1_prev, bcol, ar1, D1, D2, cost2,
                           out1, out2, one);
                li1_prev = li1;
                ok2_prev = (ctl & CTL_OK2) != 0;
                {
                    float* x = reinterpret_cast<float*>(reinterpret_cast<char*>(xch_w) + xo);
                    x[0] = out2;
                    x[32] = out1;
                }
                if (ctl & CTL_SWITCH) {   // block finished
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
                    if (ctl & CTL_NEXT_EXISTS) load_block_staged(stage_p + ((ctl & CTL_NEXT_PARITY) ? CB * kD : 0), bcol);
                    on_last = (ctl & CTL_NEXT_LAST) != 0;
#pragma unroll
                    for (int j = 0; j < CB; j++) {
                        D1[j] = INFINITY;
                        D2[j] = INFINITY;
                        cost2[j] = INFINITY;
                    }
                    li1_prev = INFINITY;
                    ok2_prev = false;
                }
                PROF_INC(wid, 3);
            }
            PROF_NOW_AFTER(ps1, out1);
            PROF_ADD(wid, 2, pw1, ps1);   // the super-step's (up to) two steps
            // this super-step's ring and stage reads are over (batch S+DEPTH waits for it, if there is one)
            if (S <= g.n_super - 1 - DEPTH) bar_arrive(BAR_EMPTY + (S & 3));
        }

        PROF_NOW(pe0);
        // ---- result: the warp that owns the last block
        // the group's reads of the ring are over: the producers may store the next group's first batches
        if (grp + gridDim.x < n_groups) bar_arrive(BAR_GROUP);
        // the drain of the last row needs the left neighbour's last values: one more consumer-only rendezvous
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_CONSUMERS), "n"(NW * 32) : "memory");
        if (wid == owner_w) {
            const int Bl = n_blocks - 1;
            float res = INFINITY;
            if (on_last) {   // still on the last block: the result cell is inside its band
                if (!(g.last_row & 1)) {
                    // drain: DP of the last step's second row (row 2*half == m-1)
                    float left = dseed;
                    if (ok2_prev) left = sched.res_from_drain ? xdrain_r[0] : xch_r[sched.res_slot * 64];
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
        PROF_NOW(pe1);
        PROF_ADD(wid, 4, pe0, pe1);   // epilogue
        PROF_ADD(wid, 5, pg0, pe1);   // whole group
    }
    PROF_FLUSH(wid);
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
        for (int c = 0; c < g.n_super; c++) {
            // batch c: what the consumers first read in super-step c+1.
            // All loads first (they need no shared memory), then the wait for the ring slots, then scale and store.
            ulonglong2 v[4][2];
            float* dst[4];
            float sign[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned un = sched.unit[c][i];
                v[i][0] = make_ulonglong2(0ull, 0ull);
                v[i][1] = v[i][0];
                dst[i] = nullptr;
                sign[i] = 1.f;
                if (un & 0x8000u) {
                    const int Bq = (un & 0x7fffu) >> 2, j = un & 3u;
                    const int col = min(Bq * CB + 2 * j + my_row, n - 1);
                    ldg32(winf + (size_t)col * kD + (part & 1) * 8, v[i][0], v[i][1]);
                    if (part == 0 && Bq + 1 < n_blocks) prefetch_l2(winf + (size_t)min((Bq + 1) * CB + 2 * j, n - 1) * kD);
                    dst[i] = stage_f + (Bq & 1) * (CB * kD) + 2 * j * kD;
                    sign[i] = -1.f;
                } else if (un) {
                    const int k = (int)un;
                    const float* src = tmplf + (size_t)(2 * k - 2) * kD + part * 8;
                    if (2 * k - 1 + my_row <= m) ldg32(src, v[i][0], v[i][1]);
                    if (part == 0 && 2 * (k + 8) - 1 <= m) prefetch_l2(src + 8 * SLOT_F);
                    dst[i] = ring_f + (k & (SLOTS - 1)) * SLOT_F;
                }
            }
            // the slots this batch overwrites were last read in super-step c-DEPTH at the latest (Stream4Sched invariant),
            // or by the previous group
            PROF_NOW(pp0);
            if (c >= DEPTH) bar_sync(BAR_EMPTY + ((c - DEPTH) & 3));
            else if (c == 0 && grp != (int64_t)blockIdx.x) bar_sync(BAR_GROUP);
            PROF_NOW(pp1);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (sched.unit[c][i]) norm_store(v[i][0], v[i][1], dst[i], sign[i], true);   // uniform; the shuffle inside needs every lane
            __threadfence_block();
            bar_arrive(BAR_FULL + (c & 3));
            PROF_NOW(pp2);
            PROF_ADD(4 + (tid >> 5), 1, pp0, pp1);   // waiting for free slots
            PROF_ADD(4 + (tid >> 5), 2, pp1, pp2);   // waiting for the loads + scale + store
        }
    }
    PROF_FLUSH(4 + (tid >> 5));
}

__global__ void __launch_bounds__(NTHREADS, 2) dtw_pairs_stream4_kernel(DtwPairsArgs a, int64_t n_groups, int window, Stream4Sched sched) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < SMEM_FLOATS; i += NTHREADS) smem[i] = 0.f;
#ifdef RP_K2S_PROFILE
    if (threadIdx.x < 64) (&s_prof[0][0])[threadIdx.x] = 0u;
#endif
    __syncthreads();
    const Geometry g = make_geometry(a, window);
    if (threadIdx.x >= NW * 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        producer_loop(a, n_groups, g, smem, sched);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
        consumer_loop(a, n_groups, g, smem, sched);
    }
}

}  // namespace

// Static schedule of the producer warps (see the kernel header). Batch c (0 .. n_super-1) is stored while the consumers
// run super-step c (steps 2c-1, 2c; c = 0 is their prologue) and is first read in super-step c+1. Invariants, checked
// here and again by tests/test_host_logic.py through rp_debug_stream4_schedule:
//  (1) row pair k is stored before the step that reads it first, look-ahead included: batch <= (st_first(k) - 2) / 2;
//  (2) ring slot k % 16 is free: the last read of row pair k-16 lies DEPTH super-steps before the batch's (the
//      producers run up to DEPTH super-steps ahead of the consumers);
//  (3) the four quarters of window block B >= 4 are stored before the super-step of the step after which its warp
//      switches to it, and DEPTH super-steps after the one in which block B-2 (same staging slot) was read.
// Across groups nothing overlaps: the producers wait for the consumers' end-of-group signal before batch 0.
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
            if (kf > SLOTS && (st_last(kf - SLOTS) + 1) / 2 > c - DEPTH) return false;   // (2)
            s.unit[c][slot++] = (unsigned short)kf++;
        }
        while (qB < n_blocks && slot < 4 && c >= c_due(qB) - 2) {
            if (c > c_due(qB)) return false;                                            // (3) too late
            if (qB - 2 >= NW && (s_sw(qB - 2) + 1) / 2 > c - DEPTH) return false;       // (3) slot still in use
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
    // consumers: one control word per (warp, step)
    auto mask = [&](int u, int c0) {
        const int t = 2 * u - c0 + w - 2;
        const int lo = std::min(std::max(7 - t, 0), 10), hi = std::min(std::max(7 - t + 2 * w, 0), 10);
        return (1u << hi) - (1u << lo);
    };
    for (int wq = 0; wq < NW; wq++) {
        int B = wq;
        for (int st = 1; st <= steps; st++) {
            const int u = st - SIGMA * B;
            const int ufirst = std::max(1, 4 * B + 1 - w / 2), ulast = 4 * B + fin0;
            if (B >= n_blocks || u < ufirst || u > ulast) continue;
            const int c0 = B * CB + 1;
            const unsigned M = mask(u, c0);
            unsigned c = CTL_ACTIVE | (M << CTL_MASK_SHIFT) | ((unsigned)(u & 15) << CTL_SLOT_SHIFT);
            if (u == ufirst) c |= CTL_FRESH;
            if ((M & 0x1ffu) == 0x1ffu) c |= CTL_FULL;
            if (B > 0) {
                c |= CTL_HAS_LEFT;
                if ((M >> 8) & 1u) c |= CTL_OK1;
                if ((M >> 9) & 1u) c |= CTL_OK2;
                if (u == ufirst && ((mask(u - 1, c0) >> 9) & 1u)) c |= CTL_OK2PREV;
                if (u == 4 * (B - 1) + fin0 + 1) c |= CTL_DRAIN_RD;
            }
            if (u == ulast && st < steps) {
                c |= CTL_SWITCH;
                if (B + NW < n_blocks) c |= CTL_NEXT_EXISTS;
                if ((B + NW) & 1) c |= CTL_NEXT_PARITY;
                if (B + NW == n_blocks - 1) c |= CTL_NEXT_LAST;
                s.ctl[wq][st] = c;
                B += NW;
                continue;
            }
            s.ctl[wq][st] = c;
        }
    }
    {
        const int Bl = n_blocks - 1, u = steps + 1 - SIGMA * Bl;
        s.res_from_drain = Bl > 0 && u == 4 * (Bl - 1) + fin0 + 1;
        s.res_slot = u & (XS - 1);
    }
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
