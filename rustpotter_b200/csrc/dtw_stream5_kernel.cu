// K2s v5 — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Contract: reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105 (banded DTW with the
// asymmetric band [r-w, r+w-1], result cell D[m-1][n], cosine distance with similarity 0 for zero
// vectors, cost/(m+n) -> logistic score).
//
// The consumer side is v4's (dtw_stream4_kernel.cu): a lane is one pair of a group of 32, four consumer warps own the
// 8-column window blocks b = warp (mod 4) as a systolic array turned by 90 degrees, a block of negated unit columns lives
// in registers (128 of them), the template streams past it two rows per step (128 FFMA2 with the DP chain of the previous
// row interleaved), warp-uniform control words built on the host drive every step.
//
// What changed is how the data gets there. ncu on v4 (profiles/r02_ncu_k2s_raw.csv): its four PRODUCER warps execute 26 %
// of all instructions (HBM -> registers -> unit length -> shared memory, on the same schedulers and the same FP32 pipe as
// the consumers) and the consumers spend 21 % of their time at the full/empty barrier; the consumers' step body alone,
// fed from shared memory with nothing else running, needs 633 cycles per step at this occupancy against 1383 in the
// kernel (tools/microbench_k2s_ceiling.cu, profiles/r02_k2s_ceiling.txt). So:
//  * Four LOADER warps per CTA replace the producers, one per unit of a batch. Lane = pair: for its unit of the host-built
//    schedule (a template row pair or a quarter of a window block: 128 contiguous bytes of one pair) every lane issues one bulk async copy
//    (cp.async.bulk.shared.global, the TMA path: UBLKCP in SASS) straight into the ring / staging slot; completion is
//    counted in bytes on an mbarrier per batch (SYNCS.ARRIVE.TRANS64). No data passes through registers, ~10 instructions
//    per unit and warp instead of ~160. (Four warps rather than one so that, as in v4, each scheduler's register file hosts
//    exactly one 200-register consumer warp per resident CTA: setmaxnreg can only grow a warp inside its own scheduler.)
//  * The RAW rows are normalised by the consumer that touches them first: the host-built control word says when a row
//    load is the first read of that row by any warp (always block b_min(k) of row pair k); that warp scales the row to
//    unit length in registers (it needs it there anyway) and writes it back, every later reader — at least one
//    super-step barrier away — finds unit rows. 26 instructions per template row per group. Staged window blocks are
//    normalised (and negated) when a warp switches to them, as v4 did for its first block.
//  * The consumers meet at their own 128-thread barrier per super-step (boundary columns), wait for the batch's mbarrier,
//    and release ring slots to the loader through the named "empty" barriers as before.
//
// Shapes: d == 16, uniform m >= 2, n >= 1, no CMN, 3 <= window = max(band, |m-n|) <= 20, at most 238 steps, 16-byte
// aligned pair offsets. Everything else takes the older kernels.
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int CB = 8;                                 // window columns per block
constexpr int NW = 4;                                 // consumer warps per CTA = blocks of one pair in flight
constexpr int NTHREADS = 2 * NW * 32;                 // + four loader warps (one per unit of a batch): warp w sits on scheduler w % 4,
                                                      // so every scheduler hosts one consumer and one loader warp per CTA
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
constexpr int MBAR_F = 8;                             // four 8-byte mbarriers ("batch landed")
constexpr int SMEM_FLOATS = RING_F + STAGE_F + XCH_F + XDRAIN_F + MBAR_F;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;           // ~104 KB: two CTAs per SM
constexpr int MIN_WINDOW = 3, MAX_WINDOW = 20;
constexpr int CONSUMER_REGS = 200, LOADER_REGS = 40;
constexpr int DEPTH = 2;                              // super-steps the producers may run ahead of the consumers
constexpr int BAR_FULL = 1, BAR_EMPTY = 5, BAR_CONSUMERS = 9, BAR_GROUP = 10;   // named barriers 1..4 (full), 5..8 (empty)

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
// ---- mbarrier / bulk-copy (TMA) primitives; shared-memory addresses are 32-bit shared::cta addresses
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {   // one arrival + the bytes the batch will deliver
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk async copy (TMA, UBLKCP): bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes of the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

// The staged copy of a block is raw (bulk-copied): negate and scale to unit length while loading.
__device__ __forceinline__ void load_block_staged(const float* __restrict__ stage, f2 (&bcol)[CB][8]) {
#pragma unroll
    for (int j = 0; j < CB; j++) {
        f2 x[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(stage + j * kD + 4 * q);
            x[2 * q] = v.x;
            x[2 * q + 1] = v.y;
        }
        unit_column(x, bcol[j]);
    }
}

// Reads one 64-byte template row from the ring.
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

// A raw 64-byte template row in registers -> unit length (a zero row stays zero: similarity 0, distance 1) and back to
// the ring, for every later reader. Done by the warp whose load is the first read of that row (CTL_NORM_*).
__device__ __forceinline__ void unit_row_store(f2 (&ar)[8], float* __restrict__ p) {
    f2 n2 = mul2(ar[0], ar[0]);
    f2 n3 = mul2(ar[1], ar[1]);
#pragma unroll
    for (int q = 2; q < 8; q += 2) {
        n2 = fma2(ar[q], ar[q], n2);
        n3 = fma2(ar[q + 1], ar[q + 1], n3);
    }
    const float nn = hsum(n2) + hsum(n3);
    const float sc = nn > 0.f ? rsqrtf(nn) : 0.f;
    const f2 s2 = pk(sc, sc);
#pragma unroll
    for (int q = 0; q < 8; q++) ar[q] = mul2(ar[q], s2);
#pragma unroll
    for (int q = 0; q < 4; q++) *reinterpret_cast<ulonglong2*>(p + 4 * q) = make_ulonglong2(ar[2 * q], ar[2 * q + 1]);
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
__device__ __forceinline__ void block_step(float* __restrict__ rp0, float* __restrict__ rp1, bool norm_a, bool norm_b, bool full, unsigned M, float li1,
                                           float li2p, float li1_prev, const f2 (&bcol)[CB][8], f2 (&ar1)[8], float (&D1)[CB],
                                           float (&D2)[CB], float (&cost2)[CB], float& out1, float& out2, f2 one) {
    f2 ar2[8], acc[CB];
    float cost1[CB];
    // ---- H1: dots of row 2u-1, DP of row 2u-2
    load_row(rp0 + kD, ar2);
    if (norm_a) unit_row_store(ar2, rp0 + kD);   // first read of row 2u by any warp (warp-uniform)
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
    if (norm_b) unit_row_store(ar1, rp1);        // first read of row 2u+1 (look-ahead for the next step)
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
constexpr unsigned CTL_NORM_A = 1u << 26;       // row 2u (loaded in H1) is read here for the first time: scale to unit length, write back
constexpr unsigned CTL_NORM_B = 1u << 27;       // the same for the look-ahead row 2u+1 (loaded in H2)
constexpr unsigned CTL_NORM_F = 1u << 28;       // the same for row 2u-1 loaded by a fresh block
constexpr int CTL_MASK_SHIFT = 12;              // 10 bits: band mask
constexpr int CTL_SLOT_SHIFT = 22;              // 4 bits: row pair u & 15

__device__ __forceinline__ void consumer_loop(const DtwPairsArgs& a, int64_t n_groups, const Geometry& g, float* smem, const Stream4Sched& sched) {
    const unsigned full_bar = (unsigned)__cvta_generic_to_shared(smem + RING_F + STAGE_F + XCH_F + XDRAIN_F);   // four mbarriers
    unsigned batch0 = 0;   // running number of the group's first batch (the loader counts the same way)
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
    float* const ring_p = reinterpret_cast<float*>(sm + ring_o);
    const float* const stage_p = reinterpret_cast<const float*>(sm + stage_o);
    float* const xch_w = reinterpret_cast<float*>(sm + xw_o);
    const float* const xch_r = reinterpret_cast<const float*>(sm + xr_o);
    float* const xdrain_w = reinterpret_cast<float*>(sm + dw_o);
    const float* const xdrain_r = reinterpret_cast<const float*>(sm + dr_o);

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPG + lane;
        const bool valid = p < a.n_pairs;
        const int64_t pc = valid ? p : a.n_pairs - 1;                       // clamped: every lane computes on real data
        const float* const win = a.win + (a.win_off ? a.win_off[pc] : pc * (int64_t)n * kD);

        // ---- super-step 0: the warp's first block straight from global memory (the loader's first batches are in flight)
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
        for (int S = 1; S <= g.n_super; S++) {
            {   // batch S-1 has landed (bulk copies count their bytes on its mbarrier) ...
                const unsigned t = batch0 + (unsigned)(S - 1);
                mbar_wait(full_bar + 8u * (t & 3u), (t >> 2) & 1u);
            }
            // ... and the neighbours' boundary values and first-touch rows of the last super-step are written
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_CONSUMERS), "n"(NW * 32) : "memory");
#pragma unroll 1
            for (int st = 2 * S - 1; st <= min(2 * S, steps); st++) {
                const unsigned ctl = ctl_next;
                ctl_next = sched.ctl[wid][st + 1];   // (the table has one spare entry)
                if (!(ctl & CTL_ACTIVE)) continue;
                const unsigned so = ((ctl >> CTL_SLOT_SHIFT) & 15u) * (SLOT_F * 4);   // byte offset of row pair u in the ring
                const unsigned so1 = (so + SLOT_F * 4) & (SLOTS * SLOT_F * 4 - 1);      // ... of row pair u+1
                const unsigned xo = ((ctl >> CTL_SLOT_SHIFT) & 3u) * 256;               // ... of row pair u in the exchange array
                float* const rp0 = reinterpret_cast<float*>(reinterpret_cast<char*>(ring_p) + so);
                float* const rp1 = reinterpret_cast<float*>(reinterpret_cast<char*>(ring_p) + so1);
                // fresh block: no look-ahead happened, and the step before it was skipped
                load_row_if(rp0, ar1, (ctl & CTL_FRESH) != 0);
                if (ctl & CTL_NORM_F) unit_row_store(ar1, rp0);   // (first read of row 2u-1 by any warp)
                ok2_prev = (ctl & CTL_FRESH) ? (ctl & CTL_OK2PREV) != 0 : ok2_prev;
                // D[2u-2][c0-1]: the left block's H1 of row pair u, or its drain if row pair u-1 was its last (block 0 reads
                // the last warp's values and ignores them)
                const float* x = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xch_r) + xo);
                const float xd = xdrain_r[0], x0 = x[0], shf1 = x[32];
                const float shf2 = (ctl & CTL_DRAIN_RD) ? xd : x0;
                const float li2p = ok2_prev ? shf2 : dseed;              // left input of row 2u-2 = diagonal input of row 2u-1
                const float li1 = (ctl & CTL_OK1) ? shf1 : INFINITY;     // left input of row 2u-1 = diagonal input of row 2u
                dseed = INFINITY;
                block_step(rp0, rp1, (ctl & CTL_NORM_A) != 0, (ctl & CTL_NORM_B) != 0, (ctl & CTL_FULL) != 0, (ctl >> CTL_MASK_SHIFT) & 0x3ffu, li1,
                           li2p, li1_prev, bcol, ar1, D1, D2, cost2, out1, out2, one);
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
                    // the warp's next block from its staging slot (raw; negated and scaled to unit length on the way)
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
            }
            // this super-step's ring and stage accesses are over (batch S+DEPTH waits for it, if there is one); the rows this
            // thread wrote back must be ordered before the bulk copies that will overwrite their slots
            fence_proxy_async();
            if (S <= g.n_super - 1 - DEPTH) bar_arrive(BAR_EMPTY + (S & 3));
        }
        batch0 += (unsigned)g.n_super;

        // ---- result: the warp that owns the last block
        // the group's accesses to the ring are over: the loader may start the next group's first batches
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
    }
}

// ------------------------------------------------------------------------------------------------ loader
// Unit codes of the schedule: 0 nothing; 1 .. kmax: template row pair k; 0x8000 | (B << 2) | j: quarter j of window block B.
// Lane = pair: every unit is 128 contiguous bytes of the lane's pair in global memory and in its ring / staging slot.
__device__ __forceinline__ void loader_loop(const DtwPairsArgs& a, int64_t n_groups, const Geometry& g, float* smem, const Stream4Sched& sched) {
    const int lane = threadIdx.x & 31;
    const int unit = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5) - NW, 0);   // warp-uniform: which of the batch's four units
    const int m = g.m, n = g.n;
    const unsigned smem_a = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned ring_a = smem_a + (unsigned)(lane * RING_PAIR_F * 4);
    const unsigned stage_a = smem_a + (unsigned)((RING_F + lane * STAGE_PAIR_F) * 4);
    const unsigned full_bar = smem_a + (unsigned)((RING_F + STAGE_F + XCH_F + XDRAIN_F) * 4);
    unsigned batch = 0;   // running batch number: mbarrier (batch & 3), phase parity (batch >> 2) & 1

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t pf = min(grp * PPG + lane, a.n_pairs - 1);
        const float* const winf = a.win + pf * (int64_t)n * kD;
        const float* const tmplf = a.tmpl + pf * (int64_t)m * kD;
        for (int c = 0; c < g.n_super; c++, batch++) {
            // the slots this batch overwrites were last touched in super-step c-DEPTH at the latest (Stream4Sched invariant),
            // or by the previous group
            if (c >= DEPTH) bar_sync(BAR_EMPTY + ((c - DEPTH) & 3));
            else if (c == 0 && grp != (int64_t)blockIdx.x) bar_sync(BAR_GROUP);
            const unsigned un = sched.unit[c][unit];
            const float* src = nullptr;
            unsigned dst = 0, bytes = 0;
            if (un & 0x8000u) {   // two window columns (a quarter block); columns >= n are not copied (their cells are never read)
                const int Bq = (un & 0x7fffu) >> 2, j = un & 3u;
                const int c0 = Bq * CB + 2 * j;
                src = winf + (size_t)c0 * kD;
                dst = stage_a + (unsigned)(((Bq & 1) * (CB * kD) + 2 * j * kD) * 4);
                bytes = (unsigned)min(max(n - c0, 0), 2) * (kD * 4);
            } else if (un) {      // template rows 2k-1, 2k (the second one only if it exists)
                const int k = (int)un;
                src = tmplf + (size_t)(2 * k - 2) * kD;
                dst = ring_a + (unsigned)((k & (SLOTS - 1)) * SLOT_F * 4);
                bytes = (unsigned)min(max(m - (2 * k - 2), 0), 2) * (kD * 4);
            }
            const unsigned bar = full_bar + 8u * (batch & 3u);
            if (lane == 0) mbar_expect_tx(bar, bytes * 32u);   // one arrival per loader warp (an empty unit is a plain arrival)
            __syncwarp();
            if (bytes) bulk_g2s(dst, src, bytes, bar);
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) dtw_pairs_stream5_kernel(DtwPairsArgs a, int64_t n_groups, int window, Stream4Sched sched) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < SMEM_FLOATS; i += NTHREADS) smem[i] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned full_bar = (unsigned)__cvta_generic_to_shared(smem + RING_F + STAGE_F + XCH_F + XDRAIN_F);
        for (unsigned i = 0; i < 4; i++) mbar_init(full_bar + 8u * i, (unsigned)NW);   // one arrival per loader warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();   // the zero fill above precedes every bulk copy into the same bytes
    __syncthreads();
    const Geometry g = make_geometry(a, window);
    // 256 threads x 128 registers at launch; the loaders hand most of theirs to the consumers. Warp w lives on scheduler w % 4,
    // so each scheduler's register file holds one consumer and one loader warp of each resident CTA: 2 x (200 + 40) x 32 <= 16 K
    if (threadIdx.x >= NW * 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LOADER_REGS));
        loader_loop(a, n_groups, g, smem, sched);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
        consumer_loop(a, n_groups, g, smem, sched);
    }
}

}  // namespace

bool dtw_pairs_stream5_supported(const DtwPairsArgs& a) {
    // dense layout only: bulk copies need 16-byte aligned sources, which per-pair offsets do not promise
    if (a.d != kD || a.cmn || a.tmpl_len || a.win_len || a.tmpl_off || a.win_off) return false;
    if ((reinterpret_cast<uintptr_t>(a.tmpl) | reinterpret_cast<uintptr_t>(a.win)) & 15u) return false;
    return build_stream4_schedule(a.tmpl_len_max, a.win_len_max, a.band, nullptr, nullptr);
}

cudaError_t launch_dtw_pairs_stream5(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    Stream4Sched sched;
    if (!build_stream4_schedule(m, n, a.band, &sched, nullptr)) return cudaErrorInvalidValue;
    const int64_t n_groups = (a.n_pairs + PPG - 1) / PPG;
    cudaError_t e = cudaFuncSetAttribute(dtw_pairs_stream5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream5_kernel, NTHREADS, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > n_groups) blocks = n_groups;
    dtw_pairs_stream5_kernel<<<(unsigned)blocks, NTHREADS, SMEM_BYTES, stream>>>(a, n_groups, window, sched);
    return cudaGetLastError();
}

}  // namespace rp
