// K2p — tuned window scorer for the detector pipeline (mfcc_size <= 16, band_size 1..20), sm_100a.
//
// Same arithmetic contract as the generic kernel in dtw_kernel.cu (reference
// src/wakewords/comp/wakeword_comp.rs:22-37 + src/mfcc/{normalizer,comparator,dtw}.rs), restructured
// around what the detector's workload shares:
//   * every window of a stream is scored against the SAME templates, and
//   * consecutive windows of a stream overlap in all but one frame.
//
// One CTA scores NW = 128 consecutive windows of one stream against one template ("slot"), one
// window per thread (plus one helper warp). In the detector n == m, so window = band_size = W and row
// r touches the 2W columns c in [r-W, r+W-1]. For window j (thread t), column c is frame
// u = t + c - 1 of the CTA's frame tile, CMN'd with that window's own mean mu_j:
//     cost(r, c) = 1 - a^_r . (x_u - mu_j) / |x_u - mu_j|          a^_r = a_r / |a_r| (pre-normalised)
//                = 1 - (G_r[u] - A_r) * inv_c
//   G_r[u] = a^_r . x_u   is independent of the window -> computed ONCE per (row, frame) and shared
//                          through shared memory by the 2W windows that need it (the thread that
//                          has frame u in registers for its own column r+W-1 produces it);
//   A_r    = a^_r . mu_j  one 16-dim dot per row per window;
//   inv_c  = 1/|x_u - mu_j| one per column per window, kept in a rotating register file.
// The DP itself is thread-serial: D[i] (band offset i = c - r + W) is updated in place,
//     D[i] = cost_i + min3(D[i+1] /*(r-1,c)*/, D[i] /*(r-1,c-1)*/, D[i-1] /*(r,c-1), new*/),
// 5 instructions per cell (LDS, FADD, FFMA, FMNMX3, FADD) at full lane utilisation, instead of a
// shuffle wavefront that keeps ~W of 32 lanes busy. 16-dim dots use packed FFMA2/FADD2.
// The returned cell is D[m-1][n] (dtw.rs:101), i.e. band offset W+1 after row m-1; row m is never
// computed. Left-border cells (c < 1) stay +inf by induction once row 1 is special-cased; right-
// border cells (c > m) hold garbage that no valid cell ever reads.
//
// Shapes. mfcc_size < 16 is zero-padded to 16 while the frame tile is staged (zero columns change neither dot
// products nor norms; the unit templates arrive padded). The band is a template parameter W; band_size 5 (the
// reference's default, config.rs:193-208) runs the exact W = 5 instance, any other band_size <= 20 runs the next
// larger W in {5, 8, 12, 20} with the cells outside [r-band, r+band-1] forced to +inf after every row (MASK).
//
// Avg gate (wakeword_comp.rs:85-94: a wakeword whose avg_features score is below avg_threshold is not scored against
// its templates). The engine launches the avg slots first (gate == 1: every CTA ORs "score >= avg_threshold" over its
// 128 windows into tile_pass[stream][window block][wakeword]) and then the template slots (gate == 2: a CTA whose
// tile has no passing window returns at once). A tile with one passing window scores all its windows: a superset of
// what the reference computes, so detections are identical; K3 never reads the scores that were skipped.
#include <cfloat>
#include <cmath>

#include <atomic>
#include <mutex>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kD = 16;
constexpr int kNW = 128;          // windows (DP threads) per CTA
constexpr int kThreads = kNW + 32;  // + one helper warp
constexpr int kXS = 20;           // smem row stride of the frame tile in floats (80 B: LDS.128 conflict-free)

typedef unsigned long long f2;    // packed pair of f32

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
struct Row16 {  // 16 floats as 8 packed pairs
    f2 p[8];
};
__device__ __forceinline__ Row16 lds_row(const float* s) {
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

// Unit-length template rows in constant memory (64 KB = 1024 rows): a template row is the same for every thread of a CTA, so it
// comes through the constant cache into uniform registers instead of as four broadcast LDS.128 per warp and row — a quarter of
// this kernel's shared-memory wavefronts, the pipe that bounds it (ncu r01: LSU data pipe 78 % busy). CT selects that path; the
// launcher keeps the constant copy current and falls back to the shared-memory copy when the templates do not fit.
constexpr int kConstRows = 1024;
__constant__ float4 c_tmpl_unit[kConstRows * 4];

__device__ __forceinline__ Row16 ldc_row(int f4_index) {
    Row16 r;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 v = c_tmpl_unit[f4_index + q];
        r.p[2 * q] = pk(v.x, v.y);
        r.p[2 * q + 1] = pk(v.z, v.w);
    }
    return r;
}

// Shared memory through 32-bit shared-space addresses (the row loop keeps two of them in registers; built from generic
// pointers the compiler re-derived the CTA's shared window base -- S2UR, ULEA, LEA -- at every use: 6 % of the loop).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ Row16 lds_row_sa(unsigned a) {
    Row16 r;
#pragma unroll
    for (int q = 0; q < 4; q++) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.p[2 * q]), "=l"(r.p[2 * q + 1]) : "r"(a + 16 * q));
    return r;
}
// All warps of the CTA, from whichever of the role loops they run (every warp is in one role as a whole). A counted named
// barrier in its non-aligned form: the three loops reach it from different program locations, which neither __syncthreads()
// nor an aligned bar.sync allow (compute-sanitizer synccheck reports the aligned form as divergence; this form is clean).
__device__ __forceinline__ void cta_sync() { asm volatile("barrier.sync 1, %0;" ::"n"(kThreads) : "memory"); }
// 1/sqrt(n2) for the squared norm of a column; 0 for a zero vector (similarity 0, comparator.rs:44-47). MUFU.RSQ without the
// denormal rescue code of rsqrtf (four instructions per row): a squared norm is 0 or far above FLT_MIN.
__device__ __forceinline__ float inv_norm(float n2) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(n2));
    return n2 >= FLT_MIN ? y : 0.f;
}

struct WindowLaunch {              // what one launch covers (device pointers)
    const int32_t* slots;          // [n_slots] slot ids of this launch, or nullptr: slot i = i
    int n_slots;
    const int64_t* unit_off;       // [a.n_slots] offset of the slot's unit template in tmpl_unit (floats, rows of 16)
    int gate;                      // 0: no gate; 1: avg slots, write tile_pass; 2: template slots, read tile_pass
    const int32_t* slot_ww;        // [a.n_slots] wakeword of a slot
    const WakewordMeta* metas;     // [n_wakewords]
    int n_wakewords;
    unsigned char* tile_pass;      // [n_streams][j_blocks][n_wakewords]
    int band;                      // band_size (MASK instances)
    long long const_base;          // CT: float offset in tmpl_unit of the range that c_tmpl_unit currently holds
};

template <int W, bool CT, bool MASK>
// (4 or 6 CTAs per SM for W = 5 -- 91 registers without spills, 64 with 32 bytes of them -- measured within +-1.5 % of 5;
// rotating the helper role over the five warps by CTA measured 8 % slower)
__global__ void __launch_bounds__(kThreads, (W <= 5 ? 5 : W <= 8 ? 3 : 2))
dtw_windows_d16_kernel(DtwWindowsArgs a, WindowLaunch L, const float* __restrict__ tmpl_unit, int x_rows, int j_blocks) {
    constexpr int NB = 2 * W;  // band cells per row
    extern __shared__ __align__(16) float sm[];
    // The arrays the row loop touches sit at compile-time offsets (no address arithmetic on runtime sizes in the loop).
    constexpr int GS = kNW + NB;
    float* Gs = sm;                              // [2][2][kNW + NB] shared dot products: two rows per barrier, double buffered
    float* Os = Gs + 4 * GS;                     // [kThreads/16 + 1][16] segment offsets of the row prefix (16 B aligned)
    float* Xs = Os + (kThreads / 16 + 1) * kD;   // [x_rows][kXS] frame tile (raw frames)
    float* Ts = Xs + (size_t)x_rows * kXS;       // [max_len][16] unit template rows
    const int p_rows = kNW + a.max_len + 1;      // prefix rows 0 .. kNW + m
    const int SEG = (p_rows + kThreads / 16 - 1) / (kThreads / 16);  // rows per prefix segment
    float* Ps = Ts + (size_t)a.max_len * kD;     // [p_rows][16] exclusive row prefix sums within SEG-row segments

    const int tid = threadIdx.x;
    const int64_t cta = blockIdx.x;
    const int si = (int)(cta % L.n_slots);
    const int s = L.slots ? L.slots[si] : si;
    const int64_t rest = cta / L.n_slots;
    const int jb = (int)(rest % j_blocks);
    const int64_t b = rest / j_blocks;
    // gate == 2: no window of this tile passed the wakeword's avg gate (uniform over the CTA)
    if (L.gate == 2 && L.tile_pass[(b * j_blocks + jb) * L.n_wakewords + L.slot_ww[s]] == 0) return;
    const int j0 = a.first_window + jb * kNW;
    const int m = a.slot_len[s];
    const int c_row0 = CT ? (int)((L.unit_off[s] - L.const_base) >> 2) : 0;   // float4 index of the slot's first row in c_tmpl_unit
    // MASK: band offsets i (column c = r - W + i) inside the reference's band [r-band, r+band-1]
    const unsigned long long band_mask = MASK ? (((1ull << (2 * L.band)) - 1ull) << (W - L.band)) : ~0ull;

    // ---- stage the frame tile and the template in shared memory
    {
        const int64_t row0 = (int64_t)a.first_window_row + j0;
        const float* src = a.frames + (b * a.frame_rows + row0) * kD;
        const int64_t avail = a.frame_rows - row0;  // rows of this stream from row0 on
        if (a.d == kD) {
            for (int i = tid; i < x_rows * 4; i += kThreads) {
                const int u = i >> 2, q = i & 3;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < avail) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)u * kD) + q);
                *reinterpret_cast<float4*>(Xs + u * kXS + 4 * q) = v;
            }
        } else {   // mfcc_size < 16: rows of d floats, zero-padded to 16
            const float* srcd = a.frames + (b * a.frame_rows + row0) * a.d;
            for (int i = tid; i < x_rows * kD; i += kThreads) {
                const int u = i >> 4, q = i & 15;
                Xs[u * kXS + q] = (u < avail && q < a.d) ? __ldg(srcd + (size_t)u * a.d + q) : 0.f;
            }
        }
        if (!CT) {
            const float4* ts = reinterpret_cast<const float4*>(tmpl_unit + L.unit_off[s]);
            for (int i = tid; i < m * 4; i += kThreads) reinterpret_cast<float4*>(Ts)[i] = __ldg(ts + i);
        }
    }
    __syncthreads();

    // DP thread (one window) vs helper warp; a DP warp whose 32 windows all lie beyond n_new (tail block) only
    // keeps the barriers company: nothing it would produce is read by a live window (G_r[u] of thread t is
    // consumed by threads t..t+2W-1 only)
    // (the warp index through a lane-0 broadcast: the compiler then knows the role branches are warp-uniform and keeps the
    // template row, the loop counters and the addresses on the uniform datapath inside them)
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool dp = warp_u < kNW / 32 && (j0 + warp_u * 32) < a.n_new;
    const int t = dp ? tid : 0;
    const bool live = dp && (j0 + tid) < a.n_new;

    // ---- per-window mean (normalizer.rs:3-31) from a cooperative prefix sum over the tile rows:
    // Ps[u] = sum of rows < u within the SEG-row segment of u, Os[seg] = sum of all earlier segments;
    // window sum of thread t = P(t+m) - P(t). (Reading the m frames per thread instead costs a third
    // of the kernel's shared-memory bandwidth; the prefix form differs from the reference's
    // sequential sum by O(1e-6) relative, far inside the parity bar.)
    {
        const int seg = tid >> 4, dd = tid & 15;  // 160 threads = 10 segments x 16 coefficients
        float run = 0.f;
        for (int i = 0; i < SEG; i++) {
            const int u = seg * SEG + i;
            if (u < p_rows) {
                Ps[u * kD + dd] = run;
                if (u < x_rows) run += Xs[u * kXS + dd];
            }
        }
        Os[(seg + 1) * kD + dd] = run;
        __syncthreads();
        if (tid < kD) {
            float acc = 0.f;
            Os[tid] = 0.f;
            for (int g = 1; g <= kThreads / 16; g++) {
                acc += Os[g * kD + tid];
                Os[g * kD + tid] = acc;
            }
        }
        __syncthreads();
    }
    f2 nmu[8];  // NEGATED mean, packed
    {
        const int u0 = t, u1 = t + m;
        const Row16 p0 = lds_row(Ps + u0 * kD), p1 = lds_row(Ps + u1 * kD);
        const Row16 o0 = lds_row(Os + (u0 / SEG) * kD), o1 = lds_row(Os + (u1 / SEG) * kD);
        const float fm = (float)m;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float a0, a1, b0, b1, c0, c1, d0, d1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p1.p[q]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(p0.p[q]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(o1.p[q]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(o0.p[q]));
            const float s0 = (a0 - b0) + (c0 - d0), s1 = (a1 - b1) + (c1 - d1);
            nmu[q] = pk(-__fdiv_rn(s0, fm), -__fdiv_rn(s1, fm));
        }
    }

    float D[NB], inv[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        D[i] = INFINITY;
        inv[i] = 0.f;
    }
    D[W] = 0.f;  // D[0][0] seen from row 1 as its (r-1, c-1) neighbour of column 1

    // columns 1 .. W-1 enter the band before row 1 (column r+W-1 enters at row r)
    auto col_inv = [&](int u) -> float {  // 1 / |x_u - mu|, 0 when the vector is 0 (similarity 0)
        const Row16 x = lds_row(Xs + u * kXS);
        f2 nn = 0ull;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const f2 y = add2(x.p[q], nmu[q]);
            nn = fma2(y, y, nn);
        }
        return inv_norm(hsum(nn));
    };
    if (dp) {
#pragma unroll
        for (int c = 1; c < W; c++) inv[c % NB] = col_inv(t + c - 1);
    }

    // ---- the rows. Three loops, one per role of a warp (DP windows / helper / tail warp without a live window), each with one
    // CTA barrier per row: no divergence bookkeeping inside the loop.
    const int last_row = m - 1;  // rows 1 .. m-1 (the result lives in row m-1)
    // Two rows per barrier: [column work of rows r, r+1] barrier [DP of rows r, r+1]; pair p = (r-1)/2 uses G buffers 2(p&1), 2(p&1)+1.
    if (dp) {
        unsigned g_sa = smem_u32(Gs) + 4u * (unsigned)t;                              // &G[0][t]
        unsigned x_sa = smem_u32(Xs) + (unsigned)(kXS * 4) * (unsigned)(t + W - 1);   // frame u = t + r + W - 2 of row r = 1
        asm volatile("" : "+r"(g_sa), "+r"(x_sa));
        // column work of row r (band slot k): own new column c = r+W-1 <-> frame u = t + r + W - 2; returns A = a^_r . mu
        auto column = [&](int r, int k, unsigned g_row) -> float {
            const Row16 ar = CT ? ldc_row(c_row0 + (r - 1) * 4) : lds_row(Ts + (r - 1) * kD);  // the same for the whole CTA
            const Row16 x = lds_row_sa(x_sa);
            x_sa += kXS * 4;
            f2 nn = 0ull, g = 0ull, aa = 0ull;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const f2 y = add2(x.p[q], nmu[q]);
                nn = fma2(y, y, nn);
                g = fma2(ar.p[q], x.p[q], g);
                aa = fma2(ar.p[q], nmu[q], aa);
            }
            inv[(k + W) % NB] = inv_norm(hsum(nn));  // column r+W-1 = (k+1)+W-1 mod NB
            sts32(g_row + (NB - 1) * 4, hsum(g));
            return -hsum(aa);                         // (nmu is negated)
        };
        // inv0: 1/|x - mu| of the row's leftmost band column r-W (its ring slot is the one column r+W takes over)
        auto dp_row = [&](int k, unsigned g_row, float A, bool first, float inv0) {
            if (first) {
                // row 1: columns c < 1 must stay +inf (they would otherwise inherit D[0][0])
#pragma unroll
                for (int i = W; i < NB; i++) {
                    const float sim = (lds32(g_row + i * 4) - A) * inv[(k + i + 1 + NB - W) % NB];
                    const float best = min3(i + 1 < NB ? D[i + 1] : INFINITY, D[i], D[i - 1]);
                    const float v = (1.f - sim) + best;
                    D[i] = (!MASK || ((band_mask >> i) & 1ull)) ? v : INFINITY;
                }
                D[W - 1] = INFINITY;
            } else {
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    // column c = r - W + i  ->  inv slot c % NB = (k + 1 - W + i) mod NB
                    const float sim = (lds32(g_row + i * 4) - A) * (i == 0 ? inv0 : inv[(k + i + 1 + NB - W) % NB]);
                    const float best = min3(i + 1 < NB ? D[i + 1] : INFINITY, D[i], i > 0 ? D[i - 1] : INFINITY);
                    const float v = (1.f - sim) + best;
                    D[i] = (!MASK || ((band_mask >> i) & 1ull)) ? v : INFINITY;
                }
            }
        };
        unsigned g_pair = 0;   // byte offset of the pair's two G buffers: alternates 0, 2 GS floats
        for (int r0 = 0; r0 < last_row; r0 += NB) {
#pragma unroll
            for (int k = 0; k < NB; k += 2) {
                const int r = r0 + k + 1;  // r % NB == (k + 1) % NB because r0 is a multiple of NB
                if (r <= last_row) {       // uniform over the CTA
                    const unsigned g_a = g_sa + g_pair, g_b = g_a + (unsigned)(GS * 4);
                    g_pair ^= (unsigned)(2 * GS * 4);
                    const bool two = r + 1 <= last_row;
                    const float A = column(r, k, g_a);
                    const float inv_left = inv[(k + 1 + W) % NB];   // column r-W, before column r+1+W-1 = r+W takes its slot
                    float A2 = 0.f;
                    if (two) A2 = column(r + 1, k + 1, g_b);
                    cta_sync();
                    dp_row(k, g_a, A, r == 1, inv_left);
                    if (two) dp_row(k + 1, g_b, A2, false, inv[(k + 2 + W) % NB]);
                }
            }
        }
    } else if (warp_u >= kNW / 32) {
        // helper warp: the NB-1 lowest frames of each row's shared range, both rows of a pair
        const int e0 = tid - kNW;
        for (int r = 1; r <= last_row; r += 2) {
            const int rows = r + 1 <= last_row ? 2 : 1;
            float* G = Gs + ((((r - 1) >> 1) & 1) * 2) * GS;
            for (int base = 0; base < rows * (NB - 1); base += 32) {   // uniform trip count; no lane leaves the warp's path
                const int idx = base + e0;
                const bool act = idx < rows * (NB - 1);
                const int rr = (act && idx >= NB - 1) ? 1 : 0, e = act ? idx - rr * (NB - 1) : 0;
                const Row16 ar = CT ? ldc_row(c_row0 + (r + rr - 1) * 4) : lds_row(Ts + (r + rr - 1) * kD);
                const int u = r + rr - W - 1 + e;
                const float g = dot16(ar, lds_row(Xs + max(u, 0) * kXS));
                if (act) G[rr * GS + e] = u >= 0 ? g : 0.f;
            }
            cta_sync();
        }
    } else {
        for (int r = 1; r <= last_row; r += 2) cta_sync();
    }
    bool pass = false;
    if (live) {
        // D[m-1][m] = band offset W+1 of row m-1 (dtw.rs:101); m == 1 has no such cell -> +inf
        const float cost = m >= 2 ? D[W + 1] : INFINITY;
        const float normalized = __fdiv_rn(cost, (float)(2 * m));
        const float score = __fdiv_rn(1.f, 1.f + expf(__fdiv_rn(normalized - a.score_ref, a.score_ref)));
        a.scores[(b * a.n_new + (j0 + tid)) * a.n_slots + s] = score;
        if (L.gate == 1) {   // the comparison K3 makes (score_logic.h judge_window): avg_score < avg_threshold => skip
            const float thr = L.metas[L.slot_ww[s]].avg_threshold;
            pass = thr == 0.f || !(score < thr);
        }
    }
    if (L.gate == 1) {
        const int any = __syncthreads_or(pass ? 1 : 0);
        if (tid == 0) L.tile_pass[(b * j_blocks + jb) * L.n_wakewords + L.slot_ww[s]] = any ? 1 : 0;
    }
}

}  // namespace

// 0 = templates from constant memory when they fit (default), 3 = always from shared memory (A/B measurements, debug only)
std::atomic<int> g_window_kernel{0};
void set_dtw_window_kernel(int v) { g_window_kernel.store(v); }

// Unit-normalised templates are prepared by the engine (tmpl_unit has the layout of a.tmpl; tmpl_floats of them, tmpl_version
// changes whenever their contents do).
namespace {
// Whose templates c_tmpl_unit holds, per device. Handles are independent objects that may be driven from different host
// threads and streams, so ownership changes hands carefully: under a lock, after every kernel on the device has finished, and
// only once the current owner has been silent for kTakeOver launches of the challenger (which scores from shared memory meanwhile).
struct ConstOwner {
    const float* src = nullptr;
    uint64_t version = 0;
    size_t begin = 0, floats = 0;   // the range of src that the constant copy holds (a whole set, or one wakeword's templates)
    const float* challenger = nullptr;
    uint64_t challenger_version = 0;
    int challenger_launches = 0;
};
constexpr int kTakeOver = 64;
ConstOwner g_const_owner[64];
std::mutex g_const_mu;

// Called with g_const_mu held; true: the constant copy is this template set's and stays so until the lock is released.
// [begin, begin + floats) of tmpl_unit is what this launch reads. The owner may move its range from launch to launch (one
// wakeword's templates at a time when the whole set exceeds 64 KB): its launches run on one stream, so the copy is ordered
// between the kernels that read the old and the new range.
bool const_templates_mine(int dev, const float* tmpl_unit, size_t begin, size_t floats, uint64_t version, cudaStream_t stream, cudaError_t* err) {
    ConstOwner& o = g_const_owner[dev];
    if (o.src == tmpl_unit && o.version == version) {
        o.challenger = nullptr;
        if (o.begin != begin || o.floats != floats) {
            *err = cudaMemcpyToSymbolAsync(c_tmpl_unit, tmpl_unit + begin, floats * sizeof(float), 0, cudaMemcpyDeviceToDevice, stream);
            if (*err != cudaSuccess) return false;
            o.begin = begin;
            o.floats = floats;
        }
        return true;
    }
    if (o.src != nullptr) {
        if (o.challenger == tmpl_unit && o.challenger_version == version) {
            o.challenger_launches++;
        } else {
            o.challenger = tmpl_unit;
            o.challenger_version = version;
            o.challenger_launches = 1;
        }
        if (o.challenger_launches < kTakeOver) return false;
        // kernels of the previous owner (any stream) may still read the constant copy
        *err = cudaDeviceSynchronize();
        if (*err != cudaSuccess) return false;
    }
    *err = cudaMemcpyToSymbolAsync(c_tmpl_unit, tmpl_unit + begin, floats * sizeof(float), 0, cudaMemcpyDeviceToDevice, stream);
    if (*err != cudaSuccess) return false;
    o.src = tmpl_unit;
    o.version = version;
    o.begin = begin;
    o.floats = floats;
    o.challenger = nullptr;
    return true;
}

template <int W, bool MASK>
cudaError_t launch_instance(const DtwWindowsArgs& a, WindowLaunch L, const float* tmpl_unit, size_t const_begin, size_t const_floats,
                            uint64_t tmpl_version, int j_blocks, cudaStream_t stream) {
    const int64_t ctas = a.n_streams * (int64_t)j_blocks * L.n_slots;
    if (ctas <= 0) return cudaSuccess;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int x_rows = kNW + a.max_len + W + 1;
    const int p_rows = kNW + a.max_len + 1;
    const size_t bytes = ((size_t)x_rows * kXS + (size_t)a.max_len * kD + 4 * (kNW + 2 * W) + (kThreads / 16 + 1) * kD +
                          (size_t)p_rows * kD) * sizeof(float);
    if (bytes > 200 * 1024) return cudaErrorInvalidValue;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_const_mu);   // (the launch below happens while the ownership is known)
    cudaError_t e0 = cudaSuccess;
    const bool ct = g_window_kernel.load() != 3 && const_floats > 0 && const_floats <= (size_t)kConstRows * kD && dev >= 0 && dev < 64 &&
                    const_templates_mine(dev, tmpl_unit, const_begin, const_floats, tmpl_version, stream, &e0);
    if (e0 != cudaSuccess) return e0;
    L.const_base = (long long)const_begin;
    if (ct) {
        cudaError_t e = cudaFuncSetAttribute(dtw_windows_d16_kernel<W, true, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        dtw_windows_d16_kernel<W, true, MASK><<<(unsigned)ctas, kThreads, bytes, stream>>>(a, L, tmpl_unit, x_rows, j_blocks);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(dtw_windows_d16_kernel<W, false, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    dtw_windows_d16_kernel<W, false, MASK><<<(unsigned)ctas, kThreads, bytes, stream>>>(a, L, tmpl_unit, x_rows, j_blocks);
    return cudaGetLastError();
}
}  // namespace

bool dtw_windows_tuned_supported(int d, int band, int max_slot_len, int window_len) {
    return d >= 1 && d <= kD && band >= 1 && band <= 20 && max_slot_len <= window_len && max_slot_len >= 1;
}

int dtw_windows_tile() { return kNW; }

cudaError_t launch_dtw_windows_d16(const DtwWindowsArgs& a, const WindowGate& g, const float* tmpl_unit, size_t tmpl_floats,
                                   uint64_t tmpl_version, cudaStream_t stream) {
    if (!dtw_windows_tuned_supported(a.d, a.band, a.max_len, a.window_len > 0 ? a.window_len : a.max_len)) return cudaErrorInvalidValue;
    const int j_blocks = (a.n_new - a.first_window + kNW - 1) / kNW;
    if (j_blocks <= 0) return cudaSuccess;
    WindowLaunch L;
    L.slots = g.slots;
    L.n_slots = g.slots ? g.n_slots : a.n_slots;
    L.unit_off = g.unit_off;
    L.gate = g.gate;
    L.slot_ww = g.slot_ww;
    L.metas = g.metas;
    L.n_wakewords = g.n_wakewords;
    L.tile_pass = g.tile_pass;
    L.band = a.band;
    L.const_base = 0;
    // what the constant-memory copy of the templates should hold for this launch: the whole set, one range of it, or nothing
    const size_t cb = g.const_floats > 0 ? (size_t)g.const_begin : 0;
    const size_t cf = g.const_floats > 0 ? (size_t)g.const_floats : (g.const_floats < 0 ? 0 : tmpl_floats);
    if (a.band == 5) return launch_instance<5, false>(a, L, tmpl_unit, cb, cf, tmpl_version, j_blocks, stream);
    if (a.band < 5) return launch_instance<5, true>(a, L, tmpl_unit, cb, cf, tmpl_version, j_blocks, stream);
    if (a.band <= 8) return launch_instance<8, true>(a, L, tmpl_unit, cb, cf, tmpl_version, j_blocks, stream);
    if (a.band <= 12) return launch_instance<12, true>(a, L, tmpl_unit, cb, cf, tmpl_version, j_blocks, stream);
    return launch_instance<20, true>(a, L, tmpl_unit, cb, cf, tmpl_version, j_blocks, stream);
}

}  // namespace rp
