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

__global__ void copy_rows_kernel(const float* __restrict__ src, int64_t src_stride, float* __restrict__ dst,
                                 int64_t dst_stride, int64_t n_streams, int64_t row_floats) {
    const int64_t total = n_streams * row_floats;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / row_floats, k = i - b * row_floats;
        dst[b * dst_stride + k] = src[b * src_stride + k];
    }
}

}  // namespace

cudaError_t launch_mfcc_frames(const float* audio, int64_t audio_stride, const float* carry, int64_t n_streams,
                               int frames_per_stream, int sample_offset0, const MfccTablesDev& t, float* out,
                               int64_t out_stride_frames, int out_row0, float* vad_out, cudaStream_t stream) {
    const int64_t total = n_streams * (int64_t)frames_per_stream;
    if (total <= 0) return cudaSuccess;
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
