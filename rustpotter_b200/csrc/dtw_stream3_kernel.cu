// K2s v3 — streaming DTW scorer for INDEPENDENT (template, window) pairs, mfcc_size = 16, sm_100a.
//
// Contract: reference src/mfcc/comparator.rs:18-26 over src/mfcc/dtw.rs:56-105 (banded DTW with the
// asymmetric band [r-w, r+w-1], result cell D[m-1][n], cosine distance with similarity 0 for zero
// vectors, cost/(m+n) -> logistic score). Same mapping family as the retired round-1 kernel ("column-block
// systolic array": 5 lanes per pair, 6 pairs per warp, a lane keeps one block of 8 window columns as
// negated unit vectors in registers and walks the template two rows per step; template rows stream
// through a cp.async ring in shared memory), rebuilt around what the round-1 ncu source page showed:
// the dot products ran as dependent FFMA2 chains, the DP chain sat serialised at the end of the step,
// and every step paid for per-cell band masks and divergent bookkeeping.
//
//  * Software pipeline in half-steps. A step handles template rows (2u-1, 2u) of the lane's block.
//    Half-step H1 issues the 64 FFMA2 of row 2u-1 (eight independent accumulator chains, element index
//    outermost) and, interleaved, the eight-cell DP chain of the PREVIOUS step's row 2u-2; H2 issues the
//    FFMA2 of row 2u and the DP chain of row 2u-1. One DP cell (FMNMX3 + FADD) per eight FFMA2, so the
//    chain latency hides under the FMA stream of the same warp.
//  * The band is a pure function of d = r - c (valid iff -w+1 <= d <= w). A step touches ten distinct
//    values of d, so ONE 10-bit mask per step covers both rows' cells and both left-neighbour inputs;
//    it is applied to the cost (cost = +inf outside the band), off the DP chain. A block that is not
//    active simply sees an all-zero mask: there is no per-lane "active" state, and the one fully masked
//    step between two blocks of a lane resets its DP columns to +inf for free.
//  * Lanes are interleaved (lane = 6*l + pair): the eight lanes of a quarter-warp are six pairs at one
//    block position plus two at the next. Ring slots hold row PAIRS at a 160-byte stride and the pair
//    rings sit 2576 bytes apart, which makes the 128-bit template reads 4.6 wavefronts on average
//    (4 is the floor; the round-1 layout took 7.6). The prefetch pointer follows the slowest unfinished
//    block so thirteen slots are live at most.
//  * Template rows are scaled to unit length IN the ring, two row pairs at a time by 24 lanes of the
//    warp (one row each), instead of every lane recomputing the norm of every row it reads: the cost
//    epilogue is one FADD (the 1 of 1 - cos is the accumulator's initial value).
//  * A lane stages its next block in shared memory with cp.async five steps before it switches (two
//    512-byte buffers per pair, block parity selects), so the switch - the only divergent code in the
//    loop - reads shared memory instead of waiting on L2 (round-1 capture: 27 % of all stall samples).
//
// Shapes: d == 16, uniform m >= 2, n >= 1, no CMN, window = max(band, |m-n|) <= 20 (a block is active
// for at most 24 steps of the lane's 25-step period). Everything else takes the older kernels.
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int CB = 8;            // window columns per block
constexpr int L = 5;             // lanes per pair
constexpr int PPW = 6;           // pairs per warp (30 of 32 lanes)
constexpr int SLOTS = 16;        // ring slots (template row pairs) per pair
constexpr int SLOT_F = 40;       // floats per slot: two 64-byte rows + 32 bytes (consecutive slots rotate two bank groups)
constexpr int GROUP_F = 644;     // floats per pair ring (16*40 + 4): consecutive pairs rotate one bank group
constexpr int LAS = 7;           // prefetch distance in steps
constexpr int WAIT_G = LAS - 3;  // cp.async groups that may stay in flight at the top of a step
constexpr int STAGE_F = CB * kD; // floats of one staged window block (512 bytes)
constexpr int STAGE_LEAD = 5;    // a lane stages its next block this many steps before it switches
constexpr int RING_F = PPW * GROUP_F;   // template rings; the staging buffers (two per pair) follow
constexpr int MAX_WINDOW = 20;
constexpr int BIG = 1 << 20;

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
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_if(float* smem_dst, const float* gmem_src, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(s), "l"(gmem_src),
        "r"((int)pred)
        : "memory");
}
__device__ __forceinline__ void st_shared_32B_if(float* p, f2 a, f2 b, f2 c, f2 d, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.shared.v2.b64 [%0], {%1, %2};\n\t@p st.shared.v2.b64 [%0+16], {%3, %4};\n\t}" ::"r"(s),
        "l"(a), "l"(b), "l"(c), "l"(d), "r"((int)pred)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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

// The same block from its shared-memory staging buffer.
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

// Reads one 64-byte (already unit-length) template row from the ring.
__device__ __forceinline__ void load_row(const float* __restrict__ p, f2 (&ar)[8]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + 4 * q);
        ar[2 * q] = v.x;
        ar[2 * q + 1] = v.y;
    }
}

// Scales one 64-byte template row in the ring to unit length (a zero row stays zero: similarity 0, distance 1).
__device__ __forceinline__ void normalize_row(float* __restrict__ p) {
    f2 x[8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + 4 * q);
        x[2 * q] = v.x;
        x[2 * q + 1] = v.y;
    }
    f2 n2 = mul2(x[0], x[0]);
    f2 n3 = mul2(x[1], x[1]);
#pragma unroll
    for (int q = 2; q < 8; q += 2) {
        n2 = fma2(x[q], x[q], n2);
        n3 = fma2(x[q + 1], x[q + 1], n3);
    }
    const float nn = hsum(n2) + hsum(n3);
    const float sc = nn > 0.f ? rsqrtf(nn) : 0.f;
    const f2 s2 = pk(sc, sc);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        ulonglong2 v;
        v.x = mul2(x[2 * q], s2);
        v.y = mul2(x[2 * q + 1], s2);
        *reinterpret_cast<ulonglong2*>(p + 4 * q) = v;
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

__global__ void __launch_bounds__(32) dtw_pairs_stream3_kernel(DtwPairsArgs a, int64_t n_groups, int window) {
    __shared__ __align__(16) float ring_all[RING_F + PPW * 2 * STAGE_F];
    const int lane = threadIdx.x;
    const int l = lane / PPW, g = lane - l * PPW;   // interleaved: lane = 6*l + pair; lanes 30, 31 (l == 5) idle
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int w = window, tw = 2 * window;
    const int n_blocks = (n + CB - 1) / CB;
    const int last_row = m - 1;                 // rows 1 .. m-1 (the result cell is D[m-1][n])
    const int half = (last_row + 1) / 2;        // row pairs that contain a needed row
    const int steps = half + (n_blocks - 1);
    const int fin0 = 4 + (w + 1) / 2;           // block B leaves the band after step 5B + fin0
    const int left_lane = l == 0 ? lane + (L - 1) * PPW : lane - PPW;
    const int owner_lane = ((n_blocks - 1) % L) * PPW + g;        // lane that ends on the last block
    float* ring = ring_all + g * GROUP_F;
    float* stage = ring_all + RING_F + g * 2 * STAGE_F;   // two buffers, block parity selects
    const f2 one = pk(1.f, 0.f);
    // in-ring normalisation of the prologue's first two row pairs: lane i < 24 scales row (i / 6) & 1 of row pair
    // 1 + i / 12 of pair i % 6
    float* norm_base = ring_all + (lane % PPW) * GROUP_F + ((lane / PPW) & 1) * kD;
    const int norm_pair = lane / (2 * PPW);
    // ... and of one row pair per step inside the loop: lane i < 24 scales half (i / 6) & 1 of row (i / 12) of that
    // pair of pair i % 6; its partner lane (the other half of the row) is six lanes away
    const int nt = min(lane / PPW, 3);
    float* half_base = ring_all + (lane % PPW) * GROUP_F + (nt >> 1) * kD + (nt & 1) * 8;
    const int half_partner = (nt & 1) ? lane - PPW : min(lane + PPW, 31);

    for (int i = lane; i < RING_F + PPW * 2 * STAGE_F; i += 32) ring_all[i] = 0.f;
    __syncwarp();

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t p = grp * PPW + g;
        const bool valid = l < L && p < a.n_pairs;
        const float* tmpl = a.tmpl + (a.tmpl_off && valid ? a.tmpl_off[p] : (valid ? p : 0) * (int64_t)m * kD);
        const float* win = a.win + (a.win_off && valid ? a.win_off[p] : (valid ? p : 0) * (int64_t)n * kD);

        // template row pair k (rows 2k-1, 2k; row m is fetched too, rows beyond are never needed)
        auto issue_pair = [&](int k) {
            if (valid && l < 4) {
                float* dst = ring + (k & (SLOTS - 1)) * SLOT_F + 4 * l;
                const float* src = tmpl + (size_t)(2 * k - 2) * kD + 4 * l;
                if (2 * k - 1 <= m) cp_async16(dst, src);
                if (2 * k <= m) cp_async16(dst + kD, src + kD);
            }
        };

        __syncwarp();  // every lane is done with the previous group's ring
#pragma unroll
        for (int k = 1; k <= LAS; k++) {
            issue_pair(k);
            cp_async_commit();
        }
        {   // warm L2 with the heads of the warp's next group (first blocks of the window, first rows of the template)
            const int64_t pn = p + (int64_t)gridDim.x * PPW;
            if (l < L && pn < a.n_pairs && !a.win_off) {
                const float* wn = a.win + pn * (int64_t)n * kD + (size_t)min(l, n_blocks - 1) * CB * kD;
#pragma unroll
                for (int k = 0; k < 4; k++) prefetch_l2(wn + 32 * k);
                if (l < 4) prefetch_l2(a.tmpl + pn * (int64_t)m * kD + 32 * l);
            }
        }
        int bm = 0;            // blocks that have left the band (all lanes of the pair agree)
        int bm_step = fin0 + 1;
        bool adv_prev = true;  // the front st - bm advanced at the end of the previous step

        int B = l;
        bool has_block = valid && B < n_blocks;
        int c0 = has_block ? B * CB + 1 : BIG;          // first column of the block, 1-based
        int u_switch = 4 * B + fin0;                    // last step-local row pair of the block: ceil((c0+7+w)/2)
        // next event of this lane: staging its next block (STAGE_LEAD steps ahead), then the switch itself
        int evt_u = has_block && B + L < n_blocks ? u_switch - STAGE_LEAD : BIG;
        bool evt_stage = true;
        f2 bcol[CB][8];
        if (has_block) {
            load_block_global(win, n, B, bcol);
            if (B + L < n_blocks) {
#pragma unroll
                for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L) * CB * kD + 32 * k);
            }
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
        float dseed = (l == 0) ? 0.f : INFINITY;   // D[0][0], the diagonal input of cell (1,1)
        bool ok2_prev = false;

        cp_async_wait<LAS - 2>();   // pairs 1 and 2 have landed
        __syncwarp();
        if (lane < 4 * PPW) normalize_row(norm_base + ((1 + norm_pair) & (SLOTS - 1)) * SLOT_F);
        __syncwarp();
        f2 ar1[8], ar2[8];
        load_row(ring + ((1 - B) & (SLOTS - 1)) * SLOT_F, ar1);

        for (int st = 1; st <= steps; st++) {
            cp_async_wait<WAIT_G>();    // every pair up to st - bm + 2 has landed for its issuing lane ...
            __syncwarp();               // ... and is visible to the warp
            // row pair st - bm + 2 has landed and no lane reads it in this step: scale it to unit length in the shadow
            // of the FFMA2 stream (24 lanes, half a row each); it is stored at the end of the step
            const bool do_norm = adv_prev;
            float* const hp = half_base + ((st - bm + 2) & (SLOTS - 1)) * SLOT_F;
            f2 hx[4];
            {
                const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(hp);
                const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(hp + 4);
                hx[0] = v0.x, hx[1] = v0.y, hx[2] = v1.x, hx[3] = v1.y;
            }
            const float shf1 = __shfl_sync(0xffffffffu, out1, left_lane);
            const float shf2 = __shfl_sync(0xffffffffu, out2, left_lane);
            const int u = st - B;
            // band mask: bit i <-> d = r - c = (2u-1) - c0 + (i - 7); valid iff -w+1 <= d <= w
            const int t = 2 * u - c0 + w - 2;
            const int lo = min(max(7 - t, 0), 10), hi = min(max(7 - t + tw, 0), 10);
            const unsigned M = u >= 1 ? (1u << hi) - (1u << lo) : 0u;
            const bool ok1 = ((M >> 8) & 1u) && B > 0;
            const bool ok2 = ((M >> 9) & 1u) && B > 0;
            const float li2p = ok2_prev ? shf2 : dseed;   // left input of row 2u-2 = diagonal input of row 2u-1
            const float li1 = ok1 ? shf1 : INFINITY;      // left input of row 2u-1 = diagonal input of row 2u
            dseed = INFINITY;

            // ---- H1: dots of row 2u-1, DP of row 2u-2
            load_row(ring + (u & (SLOTS - 1)) * SLOT_F + kD, ar2);
            f2 acc[CB];
            half_step(ar1, bcol, acc, cost2, D1, D2, li2p, li1_prev, one);
            out2 = D2[CB - 1];
            float cost1[CB];
#pragma unroll
            for (int j = 0; j < CB; j++) cost1[j] = (M >> (7 - j)) & 1u ? hsum(acc[j]) : INFINITY;

            // ---- H2: dots of row 2u, DP of row 2u-1
            load_row(ring + ((u + 1) & (SLOTS - 1)) * SLOT_F, ar1);
            half_step(ar2, bcol, acc, cost1, D2, D1, li1, li2p, one);
            out1 = D1[CB - 1];
#pragma unroll
            for (int j = 0; j < CB; j++) cost2[j] = (M >> (8 - j)) & 1u ? hsum(acc[j]) : INFINITY;
            li1_prev = li1;
            ok2_prev = ok2;
            {
                const f2 hn = fma2(hx[3], hx[3], fma2(hx[2], hx[2], fma2(hx[1], hx[1], mul2(hx[0], hx[0]))));
                const float part = hsum(hn);
                const float nn = part + __shfl_sync(0xffffffffu, part, half_partner);
                const float sc = nn > 0.f ? rsqrtf(nn) : 0.f;
                const f2 s2 = pk(sc, sc);
                st_shared_32B_if(hp, mul2(hx[0], s2), mul2(hx[1], s2), mul2(hx[2], s2), mul2(hx[3], s2), do_norm && lane < 4 * PPW);
            }

            // ---- bookkeeping
            if (u == evt_u) {
                if (evt_stage) {   // stage the next block (its cp.async group lands before the switch)
                    float* dst = stage + ((B + L) & 1) * STAGE_F;
#pragma unroll
                    for (int j = 0; j < CB; j++) {
                        const float* src = win + (size_t)min((B + L) * CB + j, n - 1) * kD;
#pragma unroll
                        for (int q = 0; q < 4; q++) cp_async16(dst + j * kD + 4 * q, src + 4 * q);
                    }
                    evt_u = u_switch;
                    evt_stage = false;
                } else {           // block finished: next block of this lane
                    B += L;
                    c0 = B * CB + 1;
                    u_switch = 4 * B + fin0;
                    load_block_staged(stage + (B & 1) * STAGE_F, bcol);
                    evt_stage = true;
                    evt_u = BIG;
                    if (B + L < n_blocks) {
                        evt_u = u_switch - STAGE_LEAD;
#pragma unroll
                        for (int k = 0; k < 4; k++) prefetch_l2(win + (size_t)(B + L) * CB * kD + 32 * k);
                    }
                }
            }
            {
                // the slowest unfinished block advances every fifth step: the front st - bm then stands still for
                // one step, and neither a new row pair is fetched nor one scaled
                const bool adv = st != bm_step;
                bm += adv ? 0 : 1;
                bm_step += adv ? 0 : L;
                const int k = st - bm + LAS;   // row pair to fetch: rows 2k-1, 2k (row m is fetched too)
                float* dst = ring + (k & (SLOTS - 1)) * SLOT_F + 4 * l;
                const float* src = tmpl + (int64_t)(2 * k - 2) * kD + 4 * l;
                const bool mine = adv && valid && l < 4;
                cp_async16_if(dst, src, mine && 2 * k - 1 <= m);
                cp_async16_if(dst + kD, src + kD, mine && 2 * k <= m);
                adv_prev = adv;
            }
            cp_async_commit();
        }
        cp_async_wait<0>();

        // ---- drain: DP of the last step's second row
        {
            const float shf2 = __shfl_sync(0xffffffffu, out2, left_lane);
            float left = ok2_prev ? shf2 : dseed, diag = li1_prev;
#pragma unroll
            for (int j = 0; j < CB; j++) {
                const float up = D1[j];
                const float v = cost2[j] + min3(up, diag, left);
                diag = up;
                left = v;
                D2[j] = v;
            }
        }
        // result cell D[m-1][n]: column n of the last block, row m-1 = first (odd) or second (even) row of pair `half`
        const int jn = n - ((n_blocks - 1) * CB + 1);
        float res = INFINITY;
#pragma unroll
        for (int j = 0; j < CB; j++)
            if (j == jn) res = (last_row & 1) ? D1[j] : D2[j];
        res = __shfl_sync(0xffffffffu, res, owner_lane);
        if (valid && l == 0) {
            const float normalized = __fdiv_rn(res, (float)(m + n));
            a.out[p] = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
        }
    }
}

}  // namespace

bool dtw_pairs_stream3_supported(const DtwPairsArgs& a) {
    if (a.d != kD || a.cmn || a.tmpl_len || a.win_len) return false;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    if (m < 2 || n < 1) return false;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    return window >= 1 && window <= MAX_WINDOW;
}

cudaError_t launch_dtw_pairs_stream3(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    const int m = a.tmpl_len_max, n = a.win_len_max;
    const int diff = m > n ? m - n : n - m;
    const int window = a.band > diff ? a.band : diff;
    const int64_t n_groups = (a.n_pairs + PPW - 1) / PPW;
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtw_pairs_stream3_kernel, 32, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > n_groups) blocks = n_groups;
    dtw_pairs_stream3_kernel<<<(unsigned)blocks, 32, 0, stream>>>(a, n_groups, window);
    return cudaGetLastError();
}

}  // namespace rp
