// Launchers of the CUDA kernels (defined in mfcc_kernel.cu / dtw_kernel.cu / score_kernel.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "score_logic.h"

namespace rp {

// Device-resident MFCC tables (see mfcc_tables.h)
struct MfccTablesDev {
    const float* hamming = nullptr;   // [480]
    const float2* tw480 = nullptr;    // [480] exp(-2 pi i k / 480)
    const float* mel_bank = nullptr;  // [C][240]
    const int* centres = nullptr;     // [C+2]
    const float* dct = nullptr;       // [C][C]
    int num_coefficients = 0;         // C
    // v2 kernel (segment mel sums)
    const float* up_weight = nullptr; // [240]
    const int4* chunks = nullptr;     // [32]
    const int2* seg_chunks = nullptr; // [C+1]
    int n_chunks = 0;
};

// K1. Frame j of stream b is built from the 480 samples starting at logical sample
// 160*j + sample_offset0 of that stream, where logical sample i is audio[b*audio_stride + i] for
// i >= 0 and carry[b*320 + 320 + i] for -320 <= i < 0 (the two hops kept from the previous call).
// out[(b*out_stride_frames + out_row0 + j) * D + k], k < D = C-1.  vad_out (optional):
// [(b*frames_per_stream + j)] = mean |mfcc| of the frame (src/mfcc/vad.rs:12).
cudaError_t launch_mfcc_frames(const float* audio, int64_t audio_stride, const float* carry, int64_t n_streams,
                               int frames_per_stream, int sample_offset0, const MfccTablesDev& t, float* out,
                               int64_t out_stride_frames, int out_row0, float* vad_out, cudaStream_t stream);

void set_mfcc_variant(int v);  // 0 automatic (v2 where it applies), 1 one-frame-per-warp kernel

// Describes the DTW work of one launch of the generic kernel.
struct DtwPairsArgs {
    const float* tmpl = nullptr;       // template rows
    const int64_t* tmpl_off = nullptr; // per pair offset in floats, or nullptr => p * tmpl_len_max * d
    const int32_t* tmpl_len = nullptr; // per pair rows, or nullptr => tmpl_len_max
    int tmpl_len_max = 0;
    const float* win = nullptr;
    const int64_t* win_off = nullptr;
    const int32_t* win_len = nullptr;
    int win_len_max = 0;
    int64_t n_pairs = 0;
    int d = 0;
    int band = 5;
    float score_ref = 0.22f;
    int cmn = 0;                       // apply MfccNormalizer::normalize to the window
    float* out = nullptr;
};
// K2 generic ("faithful") kernel: one warp per pair, anti-diagonal wavefront, reference operation order.
cudaError_t launch_dtw_pairs_generic(const DtwPairsArgs& a, cudaStream_t stream);
// K2s streaming kernel for independent pairs (dtw_pairs_dispatch.cu): d == 16, uniform lengths, no CMN,
// window = max(band, |m-n|) in 3..20, at most 238 steps.
bool dtw_pairs_stream_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream(const DtwPairsArgs& a, cudaStream_t stream);
// K2s v4 (dtw_stream4_kernel.cu): what its producer warps fetch in which batch; built on the host per (m, n, band).
constexpr int STREAM4_MAX_BATCHES = 120;
constexpr int STREAM4_MAX_STEPS = 2 * STREAM4_MAX_BATCHES;
struct Stream4Sched {
    unsigned short unit[STREAM4_MAX_BATCHES][4];   // producers: what to fetch in batch c
    unsigned ctl[4][STREAM4_MAX_STEPS + 2];        // consumers: control word of (warp, step), see CTL_* in the kernel
    int res_from_drain, res_slot;                  // where the last block finds its left input after the last step
};
bool build_stream4_schedule(int m, int n, int band, Stream4Sched* out, int* n_super_out);

// Pipeline scoring: every new frame j of every stream closes a window; slot s scores the first
// slot_len[s] frames of that window (after CMN) against template s.
struct DtwWindowsArgs {
    const float* frames = nullptr;     // [n_streams][frame_rows][d]
    int64_t frame_rows = 0;            // rows per stream in `frames` (history + new)
    int first_window_row = 0;          // row of the first frame of the window that ends at new frame 0
    int n_new = 0;                     // new frames (= windows) per stream
    int64_t n_streams = 0;
    int d = 0;
    const float* tmpl = nullptr;       // all slots' template rows, packed
    const int64_t* slot_off = nullptr; // [n_slots] offset in floats
    const int32_t* slot_len = nullptr; // [n_slots] rows
    int n_slots = 0;
    int max_len = 0;                   // max slot_len
    int window_len = 0;                // max_mfcc_frames: rows of a stream's window (the longest TEMPLATE; avg_features may be longer)
    int band = 5;
    float score_ref = 0.22f;
    float* scores = nullptr;           // [n_streams][n_new][n_slots]
    int first_window = 0;              // windows j < first_window are not scored (no stream can use them)
};
cudaError_t launch_dtw_windows_generic(const DtwWindowsArgs& a, cudaStream_t stream);
// Tuned variant for d <= 16, band 1..20, slot_len <= window_len (dtw_window_kernel.cu). tmpl_unit: the templates with every
// row scaled to unit length (zero rows stay zero) and zero-padded to 16 floats; g.unit_off[s] = offset of slot s in it.
// Which slots a launch covers and the avg gate (wakeword_comp.rs:85-94) are described by WindowGate (device pointers).
struct WindowGate {
    const int64_t* unit_off = nullptr;   // [n_slots]
    const int32_t* slots = nullptr;      // slot ids of this launch (nullptr: all of a.n_slots)
    int n_slots = 0;
    int gate = 0;                        // 0 none; 1 these are avg slots: write tile_pass; 2 template slots: skip tiles without a pass
    const int32_t* slot_ww = nullptr;    // [n_slots] wakeword of a slot
    const WakewordMeta* metas = nullptr;
    int n_wakewords = 0;
    unsigned char* tile_pass = nullptr;  // [n_streams][ceil((n_new - first_window) / dtw_windows_tile())][n_wakewords]
    // constant-memory copy of the unit templates (64 KB): const_floats == 0: the whole set if it fits; > 0: the range
    // [const_begin, const_begin + const_floats) of tmpl_unit, which must hold every slot of this launch (one wakeword's
    // templates when the set as a whole does not fit); < 0: do not use it for this launch
    int64_t const_begin = 0, const_floats = 0;
};
constexpr int64_t kWindowConstFloats = 1024 * 16;   // capacity of that copy
bool dtw_windows_tuned_supported(int d, int band, int max_slot_len, int window_len);
int dtw_windows_tile();                  // windows per CTA tile (the granularity of the avg gate)
cudaError_t launch_dtw_windows_d16(const DtwWindowsArgs& a, const WindowGate& g, const float* tmpl_unit, size_t tmpl_floats,
                                   uint64_t tmpl_version, cudaStream_t stream);
void set_dtw_window_kernel(int v);  // 0 = templates from constant memory when they fit, 3 = always shared memory (debug)
// Short calls (dtw_cadence_kernel.cu): one warp per (stream, template, three consecutive windows). d <= 16, band <= 5.
int dtw_windows_cadence_max_new();      // the engine takes this kernel when a call brings at most this many windows per stream
bool dtw_windows_cadence_supported(int d, int band, int max_slot_len, int window_len);
cudaError_t launch_dtw_windows_cadence(const DtwWindowsArgs& a, const float* tmpl_unit, const int64_t* unit_off, cudaStream_t stream);

// K3: judge every window, append detections to a compact hit list.
// hit record (floats/ints, stride = 5 + max_templates): [stream, frame, wakeword, avg_score, score, scores...]
struct JudgeArgs {
    const float* scores = nullptr;     // [n_streams][n_new][n_slots]
    int64_t n_streams = 0;
    int n_new = 0;
    int n_slots = 0;
    const WakewordMeta* metas = nullptr;  // device
    int n_wakewords = 0;
    int score_mode = 1;
    int max_templates = 0;
    int* hit_count = nullptr;          // device counter (zeroed by the caller)
    float* hits = nullptr;             // device [capacity][5 + max_templates]
    int64_t capacity = 0;
    int stream_base = 0;               // added to the stream index written into hit records (group offset)
    int first_window = 0;              // windows j < first_window are skipped
};
cudaError_t launch_judge_windows(const JudgeArgs& a, cudaStream_t stream);

// Audio filters in front of the MFCC kernel (filter_kernel.cu); one lane per stream, chunks of 480 samples.
struct FilterArgs {
    const float* in = nullptr;   // [n_streams][in_stride]
    int64_t in_stride = 0;
    float* out = nullptr;        // [n_streams][out_stride] (may alias `in`)
    int64_t out_stride = 0;
    int64_t n_streams = 0;
    int n_chunks = 0;
    // gain normaliser (gain_normalizer_filter.rs)
    int gain = 0;
    float rms_level_ref = 0.f, rms_level_sqrt = 0.f, min_gain = 0.1f, max_gain = 1.f;
    int window_size = 1, window_cap = 1;
    float* gain_window = nullptr;  // [n_streams][window_cap] rms ring
    int* gain_count = nullptr;     // [n_streams][2] entries, head
    float* gains_out = nullptr;    // [n_streams][n_chunks] gain applied to each chunk (nullptr: not recorded)
    // band pass (band_pass_filter.rs)
    int band_pass = 0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
    float* bp_state = nullptr;     // [n_streams][4] x1 x2 y1 y2
};
cudaError_t launch_audio_filters(const FilterArgs& a, cudaStream_t stream);

// Sample decoding in front of the filters / MFCC kernel (audio_types.rs:98-137, encoder.rs:26-50): raw interleaved samples
// of RP_FMT_* (big_endian: byte order of the source) -> channel 0 as f32. in: [n_streams][in_stride_bytes], out: [n_streams][out_stride].
cudaError_t launch_decode_samples(const void* in, int64_t in_stride_bytes, int fmt, int channels, int big_endian, float* out,
                                  int64_t out_stride, int64_t n_streams, int64_t samples_mono, cudaStream_t stream);

// misc small kernels
cudaError_t launch_copy_rows(const float* src, int64_t src_stride, float* dst, int64_t dst_stride, int64_t n_streams,
                             int64_t row_floats, cudaStream_t stream);

}  // namespace rp
