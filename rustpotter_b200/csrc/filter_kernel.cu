// Audio filters of the batched front-end: the optional stage in front of the MFCC kernel.
//
// Replaces GainNormalizerFilter::{get_rms_level, filter} (reference src/audio/gain_normalizer_filter.rs:14-55)
// and BandPassFilter::filter (src/audio/band_pass_filter.rs:19-30) as Rustpotter::process_audio applies them
// to every 30 ms chunk (src/detector.rs:358-371). Both are scalar recurrences over a stream's samples, so the
// parallelism is across streams only: one lane per stream, one warp per 32 streams; a chunk of the 32 streams
// is moved through shared memory with coalesced loads/stores ([32][481] floats, conflict-free per-lane rows).
// Every floating-point operation is evaluated in the reference's order with explicit round-to-nearest
// mul/add (the Rust code never fuses), so the filtered samples are bit-identical to the reference's.
#include <cmath>

#include "kernels.h"

namespace rp {
namespace {

constexpr int kChunk = 480;
constexpr int kPad = kChunk + 1;

__global__ void __launch_bounds__(32) audio_filter_kernel(FilterArgs a) {
    extern __shared__ float tile[];  // [32][481]
    const int lane = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int64_t b = b0 + lane;
    const bool live = b < a.n_streams;
    const int nrows = (int)min((int64_t)32, a.n_streams - b0);

    // per-stream filter state (persists across calls; Rustpotter::reset does not touch the filters)
    float x1 = 0.f, x2 = 0.f, y1 = 0.f, y2 = 0.f;
    int wcount = 0, whead = 0;  // rms window: `wcount` valid entries, oldest at `whead`
    float* win = a.gain_window + (live ? b : 0) * a.window_cap;
    if (live) {
        if (a.band_pass) {
            const float4 st = reinterpret_cast<const float4*>(a.bp_state)[b];
            x1 = st.x; x2 = st.y; y1 = st.z; y2 = st.w;
        }
        if (a.gain) {
            wcount = a.gain_count[2 * b];
            whead = a.gain_count[2 * b + 1];
        }
    }
    const bool gain_active = a.gain && !isnan(a.rms_level_ref);

    for (int c = 0; c < a.n_chunks; c++) {
        // ---- coalesced load of the chunk of 32 streams
        for (int row = 0; row < nrows; row++) {
            const float* src = a.in + (b0 + row) * a.in_stride + (int64_t)c * kChunk;
            for (int i = lane; i < kChunk; i += 32) tile[row * kPad + i] = __ldg(src + i);
        }
        __syncwarp();
        if (live) {
            float* x = tile + lane * kPad;
            // GainNormalizerFilter::get_rms_level (gain_normalizer_filter.rs:49-55) — always computed (detector.rs:358)
            float sum_squared = 0.f;
            for (int i = 0; i < kChunk; i++) sum_squared = __fadd_rn(sum_squared, __fmul_rn(x[i], x[i]));
            const float rms = __fsqrt_rn(__fdiv_rn(sum_squared, (float)kChunk));
            float gain = 1.f;
            if (gain_active && rms != 0.f) {  // gain_normalizer_filter.rs:15-38
                // push, keep at most window_size entries
                if (wcount < a.window_size) {
                    win[(whead + wcount) % a.window_cap] = rms;
                    wcount++;
                } else {
                    win[(whead + wcount) % a.window_cap] = rms;
                    whead = (whead + 1) % a.window_cap;
                }
                float acc = 0.f;
                for (int k = 0; k < wcount; k++) acc = __fadd_rn(acc, win[(whead + k) % a.window_cap]);
                const float frame_rms = __fdiv_rn(acc, (float)wcount);
                gain = __fdiv_rn(a.rms_level_sqrt, __fsqrt_rn(frame_rms));
                gain = __fdiv_rn(roundf(__fmul_rn(gain, 10.f)), 10.f);
                gain = fminf(fmaxf(gain, a.min_gain), a.max_gain);
                if (gain != 1.f)
                    for (int i = 0; i < kChunk; i++) x[i] = fminf(fmaxf(__fmul_rn(x[i], gain), -1.f), 1.f);
            }
            if (a.gains_out) a.gains_out[b * a.n_chunks + c] = gain;
            if (a.band_pass) {  // band_pass_filter.rs:19-30
                for (int i = 0; i < kChunk; i++) {
                    const float xin = x[i];
                    float y = __fadd_rn(__fmul_rn(a.a0, xin), __fmul_rn(a.a1, x1));
                    y = __fadd_rn(y, __fmul_rn(a.a2, x2));
                    y = __fsub_rn(y, __fmul_rn(a.b1, y1));
                    y = __fsub_rn(y, __fmul_rn(a.b2, y2));
                    x2 = x1;
                    x1 = xin;
                    y2 = y1;
                    y1 = y;
                    x[i] = y;
                }
            }
        }
        __syncwarp();
        // ---- coalesced store
        for (int row = 0; row < nrows; row++) {
            float* dst = a.out + (b0 + row) * a.out_stride + (int64_t)c * kChunk;
            for (int i = lane; i < kChunk; i += 32) dst[i] = tile[row * kPad + i];
        }
        __syncwarp();
    }
    if (live) {
        if (a.band_pass) reinterpret_cast<float4*>(a.bp_state)[b] = make_float4(x1, x2, y1, y2);
        if (a.gain) {
            a.gain_count[2 * b] = wcount;
            a.gain_count[2 * b + 1] = whead;
        }
    }
}

}  // namespace

cudaError_t launch_audio_filters(const FilterArgs& a, cudaStream_t stream) {
    if (a.n_streams <= 0 || a.n_chunks <= 0) return cudaSuccess;
    const size_t bytes = 32 * kPad * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(audio_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (a.n_streams + 31) / 32;
    audio_filter_kernel<<<(unsigned)blocks, 32, bytes, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace rp
