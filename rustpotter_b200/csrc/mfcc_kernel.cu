// K1 — fused per-frame MFCC for sm_100a.
//
// Replaces MfccExtractor::{process_audio_part, pre_emphasis, calculate_magnitude_spectrum,
// calculate_mel_frequency_cepstrum, calculate_mel_frequency_cepstral_coefficients,
// discrete_cosine_transform} (reference src/mfcc/extractor.rs:69-163).
//
// One warp per 30 ms frame (480 samples = 3 hops of 160). The 480-point real DFT is factored
// 480 = 15 x 32:
//   lane p owns samples s = 32*i + n2, i = 0..14, with n2 = bitreverse5(p)
//   (1) per-lane real 15-point DFT over i           (registers, conjugate symmetry)
//   (2) twiddle by W480^(n2*k1)
//   (3) 32-point radix-2 DIT FFT ACROSS lanes for each k1 (input is already in bit-reversed lane
//       order, so the output lands in natural order: lane k2 holds X[k1 + 15*k2])
// Only bins < 240 are needed, i.e. k2 < 16: lanes 0..15 each hold 15 consecutive-stride bins.
// Magnitudes go through shared memory so that filter i (lane i) can accumulate its triangular
// band in ascending bin order — the reference's summation order (extractor.rs:136-144) — then
// ln(x + f32::MIN_POSITIVE) and the un-normalised DCT-II x2 via warp shuffles; c0 is dropped.
//
// Arithmetic notes: pre-emphasis, mel accumulation and the DCT use explicit round-to-nearest
// mul/add (no FMA contraction) to follow the reference's operation order bit for bit; the FFT
// butterflies may contract (the reference's FFT is a third-party crate whose rounding order is
// not part of its contract). No fast-math: denormal mel energies must not flush to zero.
#include <cfloat>
#include <cstdint>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kHop = 160;
constexpr int kBins = 240;
constexpr float kPre = 0.97f;

__device__ __forceinline__ float cos15(int j) {
    constexpr float t[15] = {1.f, 0.913545458f, 0.669130606f, 0.309016994f, -0.104528463f, -0.5f, -0.809016994f,
                             -0.978147601f, -0.978147601f, -0.809016994f, -0.5f, -0.104528463f, 0.309016994f,
                             0.669130606f, 0.913545458f};
    return t[j];
}
__device__ __forceinline__ float sin15(int j) {
    constexpr float t[15] = {0.f, 0.406736643f, 0.743144825f, 0.951056516f, 0.994521895f, 0.866025404f, 0.587785252f,
                             0.207911691f, -0.207911691f, -0.587785252f, -0.866025404f, -0.994521895f, -0.951056516f,
                             -0.743144825f, -0.406736643f};
    return t[j];
}

__device__ __forceinline__ float load_sample(const float* __restrict__ audio, const float* __restrict__ carry, int64_t g) {
    // g: logical sample index relative to the start of this call's audio for the stream
    return g >= 0 ? __ldg(audio + g) : __ldg(carry + (2 * kHop + g));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mfcc_frames_kernel(const float* __restrict__ audio, int64_t audio_stride, const float* __restrict__ carry,
                   int64_t n_frames_total, int frames_per_stream, int sample_offset0, MfccTablesDev t,
                   float* __restrict__ out, int64_t out_stride_frames, int out_row0, float* __restrict__ vad_out) {
    __shared__ float spec[kWarpsPerBlock][kBins];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t fidx = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (fidx >= n_frames_total) return;  // whole warp exits together
    const int64_t b = fidx / frames_per_stream;
    const int j = (int)(fidx - b * frames_per_stream);
    const float* a = audio + b * audio_stride;
    const float* cr = carry ? carry + b * (2 * kHop) : nullptr;
    const int64_t frame_start = (int64_t)kHop * j + sample_offset0;
    const int n2 = (int)(__brev((unsigned)lane) >> 27);
    const int C = t.num_coefficients;

    // ---- load, pre-emphasis (restarts at every hop: extractor.rs:87-97), Hamming (:104-109)
    float v[15];
#pragma unroll
    for (int i = 0; i < 15; i++) {
        const int s = 32 * i + n2;
        const int64_t g = frame_start + s;
        float x = load_sample(a, cr, g);
        float y = x;
        if (s % kHop != 0) y = __fsub_rn(x, __fmul_rn(kPre, load_sample(a, cr, g - 1)));
        v[i] = __fmul_rn(y, __ldg(t.hamming + s));
    }

    // ---- (1) real 15-point DFT over i: Y[k], k = 0..14, Y[15-k] = conj(Y[k])
    float re[15], im[15];
    {
        float sm[8], df[8];
#pragma unroll
        for (int i = 1; i <= 7; i++) {
            sm[i] = v[i] + v[15 - i];
            df[i] = v[i] - v[15 - i];
        }
        float r0 = v[0];
#pragma unroll
        for (int i = 1; i <= 7; i++) r0 += sm[i];
        re[0] = r0;
        im[0] = 0.f;
#pragma unroll
        for (int k = 1; k <= 7; k++) {
            float r = v[0], q = 0.f;
#pragma unroll
            for (int i = 1; i <= 7; i++) {
                r = fmaf(sm[i], cos15((i * k) % 15), r);
                q = fmaf(df[i], sin15((i * k) % 15), q);
            }
            re[k] = r;
            im[k] = -q;
            re[15 - k] = r;
            im[15 - k] = q;
        }
    }

    // ---- (2) twiddle W480^(n2*k1)
#pragma unroll
    for (int k = 1; k < 15; k++) {
        const float2 w = __ldg(t.tw480 + n2 * k);
        const float r = re[k] * w.x - im[k] * w.y;
        const float q = re[k] * w.y + im[k] * w.x;
        re[k] = r;
        im[k] = q;
    }

    // ---- (3) 32-point DIT FFT across lanes, one per k1
    // stage h = 1: W = 1
    {
        const bool bottom = lane & 1;
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const float orr = __shfl_xor_sync(0xffffffffu, re[k], 1);
            const float oi = __shfl_xor_sync(0xffffffffu, im[k], 1);
            re[k] = bottom ? orr - re[k] : re[k] + orr;
            im[k] = bottom ? oi - im[k] : im[k] + oi;
        }
    }
#pragma unroll
    for (int h = 2; h <= 16; h <<= 1) {
        const bool bottom = (lane & h) != 0;
        float2 w = __ldg(t.tw480 + (lane & (h - 1)) * (kBins / h));  // W_{2h}^t = W480^(t*240/h)
        if (bottom) { w.x = -w.x; w.y = -w.y; }
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const float orr = __shfl_xor_sync(0xffffffffu, re[k], h);
            const float oi = __shfl_xor_sync(0xffffffffu, im[k], h);
            const float ar = bottom ? orr : re[k], ai = bottom ? oi : im[k];
            const float br = bottom ? re[k] : orr, bi = bottom ? im[k] : oi;
            re[k] = ar + (br * w.x - bi * w.y);
            im[k] = ai + (br * w.y + bi * w.x);
        }
    }

    // ---- magnitude spectrum (extractor.rs:111-113), bins k1 + 15*lane for lane < 16
    float* sp = spec[warp];
    if (lane < 16) {
#pragma unroll
        for (int k = 0; k < 15; k++)
            sp[k + 15 * lane] = __fsqrt_rn(__fadd_rn(__fmul_rn(re[k], re[k]), __fmul_rn(im[k], im[k])));
    }
    __syncwarp();

    // ---- mel energies (extractor.rs:135-145): filter `lane`, ascending bins; then ln (:125-129)
    float logmel = 0.f;
    if (lane < C) {
        const int lo = __ldg(t.centres + lane);
        int hi = __ldg(t.centres + lane + 2);
        hi = hi < kBins ? hi : kBins;
        const float* fb = t.mel_bank + (size_t)lane * kBins;
        float sum = 0.f;
        for (int k = lo; k < hi; k++) {
            const float ms = sp[k];
            sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(ms, ms), __ldg(fb + k)));
        }
        logmel = logf(__fadd_rn(sum, FLT_MIN));
    }

    // ---- DCT-II x2 (extractor.rs:146-163): c[k] = 2 * sum_n l[n] * cos(pi/C * (n + .5) * k)
    float acc = 0.f;
    const float* drow = t.dct + (size_t)(lane < C ? lane : 0) * C;
    for (int n = 0; n < C; n++) {
        const float ln = __shfl_sync(0xffffffffu, logmel, n);
        acc = __fadd_rn(acc, __fmul_rn(ln, __ldg(drow + n)));
    }
    const float coef = __fmul_rn(2.f, acc);

    // ---- drop c0 (extractor.rs:84) and store
    const int D = C - 1;
    if (lane >= 1 && lane <= D) out[((b * out_stride_frames) + out_row0 + j) * (int64_t)D + (lane - 1)] = coef;

    if (vad_out != nullptr) {  // mean |mfcc| (vad.rs:12), summed in coefficient order
        float s = 0.f;
        for (int k = 1; k <= D; k++) s = __fadd_rn(s, fabsf(__shfl_sync(0xffffffffu, coef, k)));
        if (lane == 0) vad_out[b * frames_per_stream + j] = __fdiv_rn(s, (float)D);
    }
}

// =============================================================================================
// K1 v2 — two frames per warp, TMA-staged samples.
//   * A CTA (8 warps) produces up to 4 tiles of 16 consecutive frames of one stream. A tile's 18 hops of
//     samples are contiguous in HBM and are staged into shared memory by ONE bulk asynchronous copy
//     (cp.async.bulk -> UBLKCP, completion on an mbarrier), double-buffered: tile i+1 is in flight while
//     tile i is transformed; the first tile of a streaming call adds a second bulk copy for the two
//     carried hops.
//   * Warp w packs frame A = 2w (real part) and frame B = 2w+1 (imaginary part) into one complex
//     480-point FFT (same 15 x 32 factorisation as v1) and separates the two spectra afterwards:
//     X_A[k] = (Z[k] + conj Z[480-k]) / 2,  X_B[k] = (Z[k] - conj Z[480-k]) / 2i; Z[480-k] lives in lane
//     31-k2 (index 15-k1), one shuffle pair per bin.
//   * Mel energies from per-segment sums (see mfcc_tables.h): each lane sums one chunk of bins, S = sum p
//     and U = sum p*up_weight; filter i = U_i + (S_{i+1} - U_{i+1}).
//   * DCT: lanes 0..15 produce c1..c16 of frame A, lanes 16..31 those of frame B (reference summation
//     order, un-fused), so the two frames leave as one coalesced 128-byte store.
// Requires mfcc_size <= 16 and 16-byte aligned stream rows; otherwise v1 is used.
constexpr int kFramesPerTile = 16;
constexpr int kTileHops = kFramesPerTile + 2;
constexpr int kTilesPerCta = 4;     // consecutive tiles of one stream per CTA, double-buffered

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
mfcc_frames2_kernel(const float* __restrict__ audio, int64_t audio_stride, const float* __restrict__ carry,
                    int frames_per_stream, int sample_offset0, MfccTablesDev t, float* __restrict__ out,
                    int64_t out_stride_frames, int out_row0, float* __restrict__ vad_out, int ctas_per_stream) {
    __shared__ __align__(16) float xs_all[2][kTileHops * kHop];
    __shared__ float pw[kWarpsPerBlock][2][kBins];
    __shared__ float part_s[kWarpsPerBlock][2][32], part_u[kWarpsPerBlock][2][32];
    __shared__ float lbuf[kWarpsPerBlock][2][32];
    __shared__ __align__(8) unsigned long long bars[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x / ctas_per_stream;
    const int jcta = (int)(blockIdx.x - b * ctas_per_stream) * (kFramesPerTile * kTilesPerCta);
    const int n_tiles = min(kTilesPerCta, (frames_per_stream - jcta + kFramesPerTile - 1) / kFramesPerTile);

    // Stages tile `tile` into buffer `stage`: one bulk async copy (two for the first tile of a streaming
    // call, whose first two hops come from the carry buffer), completion counted in bytes on the mbarrier.
    auto issue_tile = [&](int tile, int stage) {
        const int j0 = jcta + tile * kFramesPerTile;
        const int nf = min(kFramesPerTile, frames_per_stream - j0);
        const int64_t g0 = (int64_t)kHop * j0 + sample_offset0;   // logical sample of xs[0]
        const int n_samples = (nf + 2) * kHop;
        const unsigned bar_a = smem_u32(&bars[stage]);
        float* dst = xs_all[stage];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n_samples * 4) : "memory");
        int from_carry = 0;
        if (g0 < 0) {
            from_carry = (int)(-g0);
            const float* src = carry + b * (2 * kHop) + (2 * kHop - from_carry);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst)), "l"(src), "r"(from_carry * 4), "r"(bar_a) : "memory");
        }
        const float* src = audio + b * audio_stride + (g0 < 0 ? 0 : g0);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + from_carry)), "l"(src), "r"((n_samples - from_carry) * 4), "r"(bar_a) : "memory");
    };

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) issue_tile(0, 0);

    for (int tile = 0; tile < n_tiles; tile++) {
    const int stage = tile & 1;
    // prefetch the next tile into the other buffer (its last readers passed the __syncthreads below)
    if (tid == 0 && tile + 1 < n_tiles) issue_tile(tile + 1, stage ^ 1);
    {   // wait for this tile: the barrier of a stage completes once per use -> parity (tile / 2) & 1
        const unsigned bar_a = smem_u32(&bars[stage]);
        const unsigned parity = (tile >> 1) & 1;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra WAIT_DONE;\n"
            "bra WAIT_LOOP;\n"
            "WAIT_DONE:\n"
            "}\n" ::"r"(bar_a), "r"(parity) : "memory");
    }
    const float* xs = xs_all[stage];
    const int j0 = jcta + tile * kFramesPerTile;
    const int nf = min(kFramesPerTile, frames_per_stream - j0);
    const int fa = 2 * warp;            // frame A of this warp inside the tile (frame B = fa + 1)
    if (fa < nf) {
    const bool has_b = fa + 1 < nf;
    const int n2 = (int)(__brev((unsigned)lane) >> 27);
    const int C = t.num_coefficients, D = C - 1;
    const float* xa = xs + fa * kHop;

    // Frames A and B are packed as real / imaginary part of one complex FFT. Dynamic-range guard: packing
    // lets the weaker frame inherit the rounding noise of the stronger one (relative error ~ eps * amplitude
    // ratio; an exactly silent frame would come out as the other frame's noise floor instead of the zeros
    // the reference computes). If the two frame energies differ by more than 256x, frame A is transformed
    // alone and a second pass (rare, warp-uniform) reloads frame B from shared memory.
    float* pwa = pw[warp][0];
    float* pwb = pw[warp][1];
    bool split = false;
    int pass = 0;
    do {
        // ---- load, pre-emphasis (restarts at every hop), Hamming; the two frames are 160 samples apart
        const float* xp = pass == 0 ? xa : xa + kHop;   // frame that goes to the real slot
        float xr[15], xi[15];
#pragma unroll
        for (int i = 0; i < 15; i++) {
            const int s = 32 * i + n2;
            const float h = __ldg(t.hamming + s);
            const bool second = pass == 0 && has_b;      // imaginary slot <- frame B
            // y = x[s] - 0.97 * x[s-1], restarting at every hop (s % 160 == 0 happens only for n2 == 0): the
            // coefficient is 0 there, which keeps the arithmetic identical (x - 0*prev == x) without a branch
            const float coef = (s % kHop != 0) ? kPre : 0.f;
            const int sp = s > 0 ? s - 1 : 0;
            const float y0 = __fsub_rn(xp[s], __fmul_rn(coef, xp[sp]));
            const float y1 = second ? __fsub_rn(xp[s + kHop], __fmul_rn(coef, xp[sp + kHop])) : 0.f;
            xr[i] = __fmul_rn(y0, h);
            xi[i] = __fmul_rn(y1, h);
        }
        if (pass == 0 && has_b) {
            float ea = 0.f, eb = 0.f;
#pragma unroll
            for (int i = 0; i < 15; i++) {
                ea = fmaf(xr[i], xr[i], ea);
                eb = fmaf(xi[i], xi[i], eb);
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                ea += __shfl_xor_sync(0xffffffffu, ea, o);
                eb += __shfl_xor_sync(0xffffffffu, eb, o);
            }
            split = !(ea <= 256.f * eb && eb <= 256.f * ea);
            if (split) {
#pragma unroll
                for (int i = 0; i < 15; i++) xi[i] = 0.f;
            }
        }
        // ---- per-lane 15-point DFTs of the two real sequences, merged into Z = Y_r + i Y_i
        float re[15], im[15];
        {
            float sa[8], da[8], sb[8], db[8];
#pragma unroll
            for (int i = 1; i <= 7; i++) {
                sa[i] = xr[i] + xr[15 - i];
                da[i] = xr[i] - xr[15 - i];
                sb[i] = xi[i] + xi[15 - i];
                db[i] = xi[i] - xi[15 - i];
            }
            float ra0 = xr[0], rb0 = xi[0];
#pragma unroll
            for (int i = 1; i <= 7; i++) {
                ra0 += sa[i];
                rb0 += sb[i];
            }
            re[0] = ra0;
            im[0] = rb0;
#pragma unroll
            for (int k = 1; k <= 7; k++) {
                float ra = xr[0], qa = 0.f, rb = xi[0], qb = 0.f;
#pragma unroll
                for (int i = 1; i <= 7; i++) {
                    const float c = cos15((i * k) % 15), sn = sin15((i * k) % 15);
                    ra = fmaf(sa[i], c, ra);
                    qa = fmaf(da[i], sn, qa);
                    rb = fmaf(sb[i], c, rb);
                    qb = fmaf(db[i], sn, qb);
                }
                // Y_r[k] = ra - i qa, Y_r[15-k] = ra + i qa (same for the imaginary-slot frame);  Z = Y_r + i Y_i
                re[k] = ra + qb;
                im[k] = rb - qa;
                re[15 - k] = ra - qb;
                im[15 - k] = rb + qa;
            }
        }
        // ---- twiddle W480^(n2*k1)
#pragma unroll
        for (int k = 1; k < 15; k++) {
            const float2 w = __ldg(t.tw480 + n2 * k);
            const float r = re[k] * w.x - im[k] * w.y;
            const float q = re[k] * w.y + im[k] * w.x;
            re[k] = r;
            im[k] = q;
        }
        // ---- 32-point DIT FFT across lanes (lane k2 ends with Z[k1 + 15*k2]).
        // Butterfly X_top = A + wB, X_bot = A - wB without selects: every lane first scales its own value by
        // its stage factor (w on "bottom" lanes, 1 on "top" lanes), the pair swaps, and each lane finishes
        // with one signed FMA: top = own + other, bottom = other - own.
        {
            const float sgn = (lane & 1) ? -1.f : 1.f;
#pragma unroll
            for (int k = 0; k < 15; k++) {
                const float orr = __shfl_xor_sync(0xffffffffu, re[k], 1);
                const float oi = __shfl_xor_sync(0xffffffffu, im[k], 1);
                re[k] = fmaf(sgn, re[k], orr);
                im[k] = fmaf(sgn, im[k], oi);
            }
        }
#pragma unroll
        for (int h = 2; h <= 16; h <<= 1) {
            const bool bottom = (lane & h) != 0;
            float2 w = __ldg(t.tw480 + (lane & (h - 1)) * (kBins / h));
            if (!bottom) w = make_float2(1.f, 0.f);
            const float sgn = bottom ? -1.f : 1.f;
#pragma unroll
            for (int k = 0; k < 15; k++) {
                const float vr = re[k] * w.x - im[k] * w.y;
                const float vi = re[k] * w.y + im[k] * w.x;
                const float orr = __shfl_xor_sync(0xffffffffu, vr, h);
                const float oi = __shfl_xor_sync(0xffffffffu, vi, h);
                re[k] = fmaf(sgn, vr, orr);
                im[k] = fmaf(sgn, vi, oi);
            }
        }
        // ---- separate the two spectra and take |X|^2 for bins < 240 (lanes 0..15)
        float* dst_r = pass == 0 ? pwa : pwb;   // where the real-slot frame's power goes
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const int kk = k == 0 ? 0 : 15 - k;                        // index of Z[480 - bin] in the partner lane
            const int partner = k == 0 ? ((32 - lane) & 31) : 31 - lane;
            const float cr = __shfl_sync(0xffffffffu, re[kk], partner);
            const float ci = __shfl_sync(0xffffffffu, im[kk], partner);
            if (lane < 16) {
                const float ar = re[k] + cr, ai = im[k] - ci;
                dst_r[k + 15 * lane] = 0.25f * (ar * ar + ai * ai);
                if (!split) {
                    const float br = re[k] - cr, bi = im[k] + ci;
                    pwb[k + 15 * lane] = 0.25f * (br * br + bi * bi);
                }
            }
        }
        pass++;
    } while (pass == 1 && split);
    __syncwarp();

    // ---- mel energies: chunk sums -> segment sums -> filters; ln
    const int4 ch = lane < t.n_chunks ? __ldg(t.chunks + lane) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int f = 0; f < 2; f++) {
        const float* p = pw[warp][f];
        float s = 0.f, u = 0.f;
        for (int k = ch.y; k < ch.z; k++) {
            const float e = p[k];
            s += e;
            u = fmaf(e, __ldg(t.up_weight + k), u);
        }
        part_s[warp][f][lane] = s;
        part_u[warp][f][lane] = u;
    }
    __syncwarp();
    const int2 sc = lane <= C ? __ldg(t.seg_chunks + lane) : make_int2(0, 0);
#pragma unroll
    for (int f = 0; f < 2; f++) {
        float s = 0.f, u = 0.f;
        for (int q = 0; q < sc.y; q++) {
            s += part_s[warp][f][sc.x + q];
            u += part_u[warp][f][sc.x + q];
        }
        const float s_next = __shfl_down_sync(0xffffffffu, s, 1);
        const float u_next = __shfl_down_sync(0xffffffffu, u, 1);
        const float mel = u + (s_next - u_next);
        if (lane < C) lbuf[warp][f][lane] = logf(__fadd_rn(mel, FLT_MIN));
    }
    __syncwarp();

    // ---- DCT-II x2, c0 dropped: half-warp per frame, lane (p & 15) -> coefficient k = (p & 15) + 1
    const int f = lane >> 4, kc = (lane & 15) + 1;
    float coef = 0.f;
    if (kc <= D) {
        const float* drow = t.dct + (size_t)kc * C;
        const float* lb = lbuf[warp][f];
        float acc = 0.f;
        for (int n = 0; n < C; n++) acc = __fadd_rn(acc, __fmul_rn(lb[n], __ldg(drow + n)));
        coef = __fmul_rn(2.f, acc);
        if (f == 0 || has_b)
            out[((b * out_stride_frames) + out_row0 + j0 + fa + f) * (int64_t)D + (kc - 1)] = coef;
    }
    if (vad_out != nullptr) {  // mean |mfcc| in coefficient order (vad.rs:12)
        float s0 = 0.f, s1 = 0.f;
        for (int k = 0; k < D; k++) {
            s0 = __fadd_rn(s0, fabsf(__shfl_sync(0xffffffffu, coef, k)));
            s1 = __fadd_rn(s1, fabsf(__shfl_sync(0xffffffffu, coef, 16 + k)));
        }
        if (lane == 0) vad_out[b * frames_per_stream + j0 + fa] = __fdiv_rn(s0, (float)D);
        if (lane == 16 && has_b) vad_out[b * frames_per_stream + j0 + fa + 1] = __fdiv_rn(s1, (float)D);
    }
    }  // fa < nf
    __syncthreads();  // every warp is done with xs_all[stage] before it is refilled
    }  // tiles
}

__global__ void copy_rows_kernel(const float* __restrict__ src, int64_t src_stride, float* __restrict__ dst,
                                 int64_t dst_stride, int64_t n_streams, int64_t row_floats) {
    const int64_t total = n_streams * row_floats;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / row_floats, k = i - b * row_floats;
        dst[b * dst_stride + k] = src[b * src_stride + k];
    }
}

}  // namespace

int g_mfcc_variant = 0;  // 0 automatic, 1 force the one-frame-per-warp kernel
void set_mfcc_variant(int v) { g_mfcc_variant = v; }

cudaError_t launch_mfcc_frames(const float* audio, int64_t audio_stride, const float* carry, int64_t n_streams,
                               int frames_per_stream, int sample_offset0, const MfccTablesDev& t, float* out,
                               int64_t out_stride_frames, int out_row0, float* vad_out, cudaStream_t stream) {
    const int64_t total = n_streams * (int64_t)frames_per_stream;
    if (total <= 0) return cudaSuccess;
    const bool aligned = (audio_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(audio) & 15) == 0) &&
                         (carry == nullptr || (reinterpret_cast<uintptr_t>(carry) & 15) == 0);
    if (g_mfcc_variant != 1 && t.num_coefficients - 1 <= 16 && t.n_chunks > 0 && aligned && sample_offset0 % 4 == 0) {
        const int ctas_per_stream = (frames_per_stream + kFramesPerTile * kTilesPerCta - 1) / (kFramesPerTile * kTilesPerCta);
        const int64_t ctas = n_streams * (int64_t)ctas_per_stream;
        if (ctas <= 0x7fffffffLL) {
            mfcc_frames2_kernel<<<(unsigned)ctas, kWarpsPerBlock * 32, 0, stream>>>(
                audio, audio_stride, carry, frames_per_stream, sample_offset0, t, out, out_stride_frames, out_row0, vad_out,
                ctas_per_stream);
            return cudaGetLastError();
        }
    }
    const int64_t blocks = (total + kWarpsPerBlock - 1) / kWarpsPerBlock;
    mfcc_frames_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, stream>>>(
        audio, audio_stride, carry, total, frames_per_stream, sample_offset0, t, out, out_stride_frames, out_row0, vad_out);
    return cudaGetLastError();
}

cudaError_t launch_copy_rows(const float* src, int64_t src_stride, float* dst, int64_t dst_stride, int64_t n_streams,
                             int64_t row_floats, cudaStream_t stream) {
    const int64_t total = n_streams * row_floats;
    if (total <= 0) return cudaSuccess;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    copy_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, src_stride, dst, dst_stride, n_streams, row_floats);
    return cudaGetLastError();
}

}  // namespace rp
