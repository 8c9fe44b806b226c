/*
 * rustpotter_b200.h — C ABI of the B200-native wakeword-scoring path.
 *
 * This is the drop-in boundary for ONE path of GiviMAD/rustpotter v3.0.2: per-frame MFCC extraction
 * (src/mfcc/extractor.rs) and the WakewordRef banded-DTW scorer (src/mfcc/{dtw,comparator,
 * normalizer}.rs, src/wakewords/comp/wakeword_comp.rs) together with the window bookkeeping of
 * src/detector.rs. Every entry point cites the reference interface it replaces. The reference has
 * no FFI today (it is a pure-Rust crate); INTEGRATION.md shows the `extern "C"` block a maintainer
 * would add on the Rust side to bind these symbols.
 *
 * Conventions (mirroring the reference):
 *   - Inputs are borrowed for the duration of the call and never retained.
 *   - Construction / loading returns an error code (<0) and leaves a message retrievable with
 *     rp_last_error() — the analogue of `Result<_, String>`.
 *   - Processing returns 1 = Some(detection), 0 = None, <0 = error; a wrong buffer length or a
 *     detector without wakewords is None (0), never an error (detector.rs:235-237,249-251,348-350).
 *   - Handles are not thread-safe; distinct handles may be used from distinct threads (`Send`).
 *   - There is NO CPU fallback: creating a handle without a usable CUDA device fails with
 *     RP_ERR_CUDA.
 *   - Plain pointers and sizes only; `void* cuda_stream` is a cudaStream_t (NULL = legacy default
 *     stream). Device pointers are marked `_dev`.
 */
#ifndef RUSTPOTTER_B200_H
#define RUSTPOTTER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_OK 0
#define RP_ERR_INVALID (-1)     /* bad argument */
#define RP_ERR_CUDA (-2)        /* no device / CUDA runtime failure */
#define RP_ERR_FORMAT (-3)      /* not a WakewordRef / WakewordV2 CBOR file */
#define RP_ERR_UNSUPPORTED (-4) /* outside this path (WakewordModel files, ...) */
#define RP_ERR_MISMATCH (-5)    /* wakewords with different mfcc_size (detector.rs:308-320) */

#define RP_NAME_MAX 128

/* ---- enums carried as integers -------------------------------------------------------------- */
/* SampleFormat (src/audio/audio_types.rs:3-10) */
enum { RP_FMT_I8 = 0, RP_FMT_I16 = 1, RP_FMT_I32 = 2, RP_FMT_F32 = 3 };
/* Endianness (audio_types.rs:51-57) */
enum { RP_ENDIAN_LITTLE = 0, RP_ENDIAN_BIG = 1, RP_ENDIAN_NATIVE = 2 };
/* ScoreMode (src/config.rs:86-97) */
enum { RP_SCORE_AVERAGE = 0, RP_SCORE_MAX = 1, RP_SCORE_MEDIAN = 2, RP_SCORE_P25 = 3, RP_SCORE_P50 = 4,
       RP_SCORE_P75 = 5, RP_SCORE_P80 = 6, RP_SCORE_P90 = 7, RP_SCORE_P95 = 8 };
/* VADMode (config.rs:134-139); -1 = None */
enum { RP_VAD_NONE = -1, RP_VAD_EASY = 0, RP_VAD_MEDIUM = 1, RP_VAD_HARD = 2 };

/* RustpotterConfig { fmt: AudioFmt, detector: DetectorConfig, filters: FiltersConfig }
 * (src/config.rs:9-29,172-208,31-84,212-219), flattened. rp_config_default() = Default impls. */
typedef struct rp_config {
    uint32_t sample_rate;   /* 16000; other rates go through the FFT resampler (encoder.rs:72-79) on the host first */
    uint32_t sample_format; /* RP_FMT_*; used by rp_process_bytes only */
    uint32_t channels;      /* channel 0 is used (encoder.rs:41-48) */
    uint32_t endianness;    /* RP_ENDIAN_* */
    float avg_threshold;    /* 0.2 */
    float threshold;        /* 0.5 */
    uint64_t min_scores;    /* 5 */
    uint32_t eager;         /* 0 */
    float score_ref;        /* 0.22 */
    uint32_t band_size;     /* 5 */
    uint32_t score_mode;    /* RP_SCORE_MAX */
    int32_t vad_mode;       /* RP_VAD_NONE */
    uint32_t gain_normalizer_enabled; /* 0 */
    uint32_t gain_ref_set;            /* Option<f32> gain_ref: 0 = None */
    float gain_ref;
    float min_gain;                   /* 0.1 */
    float max_gain;                   /* 1.0 */
    uint32_t band_pass_enabled;       /* 0 */
    float low_cutoff;                 /* 80 */
    float high_cutoff;                /* 400 */
} rp_config;

/* RustpotterDetection (src/detector.rs:487-501). `score_names` / `score_values` (the
 * HashMap<String,f32> `scores`) point into storage owned by the handle that produced the detection
 * and stay valid until the next processing call on that handle. */
typedef struct rp_detection {
    char name[RP_NAME_MAX];
    float avg_score;
    float score;
    uint64_t counter;
    float gain;
    uint32_t n_scores;
    const char* const* score_names;
    const float* score_values;
} rp_detection;

const char* rp_version(void);
int rp_device_count(void);               /* CUDA devices visible; 0 when none */
void rp_config_default(rp_config* cfg);  /* RustpotterConfig::default() */
/* Message of the last failed call on this thread (handle creation) or on `handle_or_null`. */
const char* rp_last_error(const void* handle_or_null);

/* =============================================================================================
 * Per-stream drop-in: mirrors `Rustpotter` (src/detector.rs:34-302). One handle = one audio stream.
 * ============================================================================================= */
typedef struct rp_handle rp_handle;

int rp_create(const rp_config* cfg, int device, rp_handle** out);            /* Rustpotter::new            :95  */
void rp_destroy(rp_handle* h);
int rp_add_wakeword_from_buffer(rp_handle* h, const char* key, const uint8_t* buf, size_t len); /*       :152 */
int rp_add_wakeword_from_file(rp_handle* h, const char* key, const char* path);                 /*       :165 */
int rp_remove_wakeword(rp_handle* h, const char* key);                       /* 1 removed / 0 absent      :180 */
int rp_remove_wakewords(rp_handle* h);                                       /*                            :193 */
size_t rp_get_samples_per_frame(const rp_handle* h);                         /*                            :204 */
size_t rp_get_bytes_per_frame(const rp_handle* h);                           /*                            :208 */
int rp_get_partial_detection(const rp_handle* h, rp_detection* out);         /* 1 Some / 0 None           :212 */
float rp_get_rms_level(const rp_handle* h);                                  /*                            :216 */
float rp_get_gain(const rp_handle* h);                                       /*                            :220 */
float rp_get_rms_level_ref(const rp_handle* h);                              /*                            :224 */
int rp_process_bytes(rp_handle* h, const uint8_t* audio_bytes, size_t len, rp_detection* out);  /*       :234 */
/* process_samples<T: Sample> (:245), one symbol per Sample impl (audio_types.rs:98-137) */
int rp_process_samples_i8(rp_handle* h, const int8_t* samples, size_t n, rp_detection* out);
int rp_process_samples_i16(rp_handle* h, const int16_t* samples, size_t n, rp_detection* out);
int rp_process_samples_i32(rp_handle* h, const int32_t* samples, size_t n, rp_detection* out);
int rp_process_samples_f32(rp_handle* h, const float* samples, size_t n, rp_detection* out);
int rp_update_config(rp_handle* h, const rp_config* cfg);                    /* update_config             :257 */
int rp_update_detector_config(rp_handle* h, const rp_config* cfg);           /*                            :265 */
int rp_update_filters_config(rp_handle* h, const rp_config* cfg);            /*                            :283 */
void rp_reset(rp_handle* h);                                                 /*                            :290 */
uint64_t rp_windows_scored(const rp_handle* h); /* calls of run_wakeword_detectors (:433) so far */

/* =============================================================================================
 * Batched front-end: N independent streams ("N Rustpotter structs") scored together on
 * one device. All streams share config and wakewords; stream state is per stream.
 * ============================================================================================= */
typedef struct rp_batch rp_batch;

typedef struct rp_batch_detection {
    int64_t stream;  /* stream index in [0, n_streams) */
    int64_t chunk;   /* 480-sample chunk index within this rp_batch_process call whose
                        process_samples() would have returned the detection */
    rp_detection det;
} rp_batch_detection;

int rp_batch_create(const rp_config* cfg, int64_t n_streams, int device, rp_batch** out);
/* The same over several devices: stream range [i*n/G, (i+1)*n/G) lives on device_ids[i]; templates and
 * tables are replicated, nothing is exchanged between devices (streams are independent `Rustpotter`s), one host thread
 * per device drives its range inside every call. A multi-device batch takes HOST audio. */
int rp_batch_create_multi(const rp_config* cfg, int64_t n_streams, const int* device_ids, int n_devices, rp_batch** out);
int rp_batch_n_devices(const rp_batch* b);
/* Rustpotter::get_samples_per_frame for one stream of the batch (480 * channels at 16 kHz; the resampler's input chunk * channels
 * otherwise, e.g. 1440 at 48 kHz). With sample_rate != 16000 every stream is resampled on the host (one FftFixedInOut each) and
 * the audio must be in host memory. */
size_t rp_batch_samples_per_frame(const rp_batch* b);
void rp_batch_destroy(rp_batch* b);
int rp_batch_add_wakeword_from_buffer(rp_batch* b, const char* key, const uint8_t* buf, size_t len);
int rp_batch_add_wakeword_from_file(rp_batch* b, const char* key, const char* path);
int rp_batch_remove_wakeword(rp_batch* b, const char* key);   /* Rustpotter::remove_wakeword  detector.rs:180 */
int rp_batch_remove_wakewords(rp_batch* b);
/* Use the caller's cudaStream_t for every launch/copy of this batch (default: a private stream). */
int rp_batch_set_cuda_stream(rp_batch* b, void* cuda_stream);
/* Equivalent to calling process_samples(&audio[s][480*c .. 480*(c+1)]) for c = 0..samples_per_stream/480
 * on each stream s. `audio` is [n_streams][samples_per_stream] f32 mono 16 kHz (already encoded:
 * Sample::into_f32), row stride = samples_per_stream; host memory (pinned recommended) when
 * audio_on_device == 0, device memory otherwise. samples_per_stream must be a multiple of 480.
 * *dets / *n_dets: detections in (stream, chunk) order, storage owned by the batch until the next call. */
int rp_batch_process(rp_batch* b, const float* audio, int64_t samples_per_stream, int audio_on_device,
                     const rp_batch_detection** dets, int64_t* n_dets);
/* Rustpotter::process_samples<T> (detector.rs:245-256) for every stream: `audio` is [n_streams][samples_per_stream]
 * samples of `sample_format` (RP_FMT_*: int8_t / int16_t / int32_t / float, native byte order) with the config's
 * `channels` interleaved (channel 0 is used, encoder.rs:41-48); samples_per_stream counts interleaved samples and must
 * be a multiple of 480 * channels. Sample::into_f32 (audio_types.rs:98-137) runs on the device, so an i16 source costs
 * half the host-to-device bytes of f32. Otherwise as rp_batch_process. */
int rp_batch_process_samples(rp_batch* b, const void* audio, int sample_format, int64_t samples_per_stream, int audio_on_device,
                             const rp_batch_detection** dets, int64_t* n_dets);
/* Rustpotter::process_bytes (detector.rs:234-244) for every stream: raw bytes in the config's sample_format /
 * endianness / channels; bytes_per_stream must be a multiple of rp_get_bytes_per_frame(). */
int rp_batch_process_bytes(rp_batch* b, const uint8_t* audio_bytes, int64_t bytes_per_stream, int audio_on_device,
                           const rp_batch_detection** dets, int64_t* n_dets);
int rp_batch_update_config(rp_batch* b, const rp_config* cfg);
void rp_batch_reset(rp_batch* b);
uint64_t rp_batch_windows_scored(const rp_batch* b);      /* sum over streams, cumulative */
int64_t rp_batch_n_streams(const rp_batch* b);
int rp_batch_max_mfcc_frames(const rp_batch* b);          /* longest template over all wakewords */
/* Timing/diagnostic taps for bench.py: milliseconds (CUDA events on the batch's stream) that the
 * last rp_batch_process spent in each stage: [0] H2D, [1] MFCC kernel, [2] DTW/score kernels,
 * [3] D2H of hits, [4] host state machine (wall clock). Returns the number of entries written. */
int rp_batch_last_timings(const rp_batch* b, float* ms, int cap);
/* Number of kernels launched by the last rp_batch_process. */
int rp_batch_last_launches(const rp_batch* b);
/* Avg gate of the last call (wakeword_comp.rs:85-94: templates are scored only when the avg_features score reaches
 * avg_threshold; here per tile of 128 consecutive windows of one stream and wakeword): tiles examined and tiles with a
 * passing window. 0 / 0 when the gate did not run (dense mode, no avg_features, generic kernel). */
int rp_batch_last_gate_stats(const rp_batch* b, int64_t* tiles, int64_t* passed);
/* Parity-test tap: copies the dense per-window scores of the last rp_batch_process to host memory,
 * [n_streams][n_new][n_slots] (n_new = samples_per_stream/160 windows, one per new 10 ms hop; slots per
 * wakeword in insertion order: [avg_features score if present], template 0..T-1). These are the raw
 * MfccComparator::compare outputs BEFORE the avg gate / threshold; windows no stream of the batch could score
 * (the leading hops after a reset of the whole batch) and template scores the avg gate skipped read as NaN, other
 * windows the detector would not score (stream start, after a reset of one stream) are present but meaningless.
 * Returns floats written or <0. */
int64_t rp_batch_copy_last_scores(const rp_batch* b, float* out_host, int64_t cap_floats, int32_t* n_new, int32_t* n_slots);

/* =============================================================================================
 * Wakeword-reference builder: the producer of the templates the hot path scores.
 * ============================================================================================= */
/* WakewordRef::new_from_sample_buffers (rms_median == 0: rms_level = max over samples) /
 * new_from_sample_files (rms_median != 0: rms_level = median over samples) followed by
 * WakewordSave::save_to_buffer — reference src/wakewords/comp/wakeword_ref_build.rs:9-110,
 * src/mfcc/wav_file_extractor.rs:18-91, src/mfcc/averager.rs:5-37, src/wakewords/wakeword_file.rs:10-26.
 * wavs[i] / wav_lens[i]: whole 16 kHz WAV files (PCM int 8/16/32 or float 32, any channel count; other
 * rates need the reference's rubato resampler, which is outside this path -> RP_ERR_UNSUPPORTED).
 * MFCCs are extracted by K1 on CUDA device `device`; averaging and CBOR encoding run on the host.
 * Writes the .rpw bytes to out (out may be NULL to query) and returns their size, or <0. */
int64_t rp_wakeword_build(const char* name, int has_threshold, float threshold, int has_avg_threshold, float avg_threshold,
                          int n_samples, const char* const* sample_names, const uint8_t* const* wavs, const size_t* wav_lens,
                          int mfcc_size, int rms_median, int device, uint8_t* out, size_t out_cap);

/* =============================================================================================
 * Raw kernels (micro-benchmarks, parity tests). All pointers are DEVICE pointers.
 * ============================================================================================= */
/* K1 — MfccExtractor::compute over whole streams (extractor.rs:60-163): a fresh extractor fed
 * samples_per_stream/160 hops emits hops-3 frames. audio_dev [n_streams][samples_per_stream],
 * out_dev [n_streams][samples_per_stream/160 - 3][mfcc_size]. */
int rp_mfcc_frames(const float* audio_dev, int64_t n_streams, int64_t samples_per_stream, int mfcc_size,
                   float* out_dev, void* cuda_stream);
/* K2 — MfccComparator::compare for independent (template, window) pairs (comparator.rs:18-26 +
 * dtw.rs:56-105). Pair p: template rows at tmpl_dev + tmpl_off[p] (tmpl_len[p] rows of d floats),
 * window at win_dev + win_off[p]; offsets in floats. cmn != 0 applies MfccNormalizer::normalize
 * (normalizer.rs:3-31) to the window first, as WakewordComparator::cut_and_normalize_frame does.
 * If tmpl_off/win_off are NULL the layout is dense: pair p at p*m*d (resp. p*n*d) with
 * m = tmpl_len_uniform, n = win_len_uniform. out_dev[p] = score. */
int rp_dtw_scores(const float* tmpl_dev, const int64_t* tmpl_off_dev, const int32_t* tmpl_len_dev, int tmpl_len_uniform,
                  const float* win_dev, const int64_t* win_off_dev, const int32_t* win_len_dev, int win_len_uniform,
                  int64_t n_pairs, int d, int band, float score_ref, int cmn, float* out_dev, void* cuda_stream);
/* Selects the DTW kernel variants (process-wide debug knob for A/B measurements and parity tests): 0 = automatic,
 * 1 = generic reference-order kernels, 2 = tuned kernels, 7 = tuned with the pipeline kernel reading its templates from
 * shared instead of constant memory, 9 = tuned with the pipeline kernel also for short calls (instead of the cadence
 * kernel); 3, 4, 5 and 6 (retired variants) behave like 2. */
int rp_set_dtw_variant(int variant);
/* Avg gate of the batched window scorer: 1 / -1 (default) = score avg_features first and the templates only where the
 * gate can pass, as the reference does; 0 = every template of every window (dense score tensor; parity taps, A/B). The
 * detections are identical in both modes. Process-wide debug knob like rp_set_dtw_variant. */
int rp_set_avg_gate(int mode);
/* Selects the MFCC kernel: 0 = automatic (two-frames-per-warp TMA-staged kernel where it applies), 1 = one frame
 * per warp. For A/B measurements and parity tests. */
int rp_set_mfcc_variant(int variant);
/* The AudioEncoder's resampling stage on its own (encoder.rs:52-60: rubato FftFixedInOut, source rate -> 16 kHz, host code):
 * whole input chunks of `in` (mono f32 at sample_rate_in) through a fresh resampler. Returns the number of output samples
 * written (out may be NULL to query), or <0. *in_chunk (optional) receives the resampler's input chunk length. No GPU needed. */
int64_t rp_resample_to_16k(uint32_t sample_rate_in, const float* in, size_t n_in, float* out, size_t out_cap, size_t* in_chunk);

/* =============================================================================================
 * Host-logic hooks (no GPU needed): used by the CPU test-suite to exercise the wakeword-file
 * reader and the per-stream state machine that consume the kernels' output.
 * ============================================================================================= */
typedef struct rp_wakeword_info {
    char name[RP_NAME_MAX];
    int32_t mfcc_size;
    int32_t n_templates;
    int32_t avg_frames;      /* 0 when avg_features is None */
    int32_t max_frames;      /* longest template */
    int32_t has_threshold, has_avg_threshold;
    float threshold, avg_threshold, rms_level;
    int32_t is_v2;
} rp_wakeword_info;
/* Static producer schedule of the streaming DTW kernel (dtw_stream4_kernel.cu) for uniform template length m,
 * window length n and band: out[batch * 4 + slot] = 0 (nothing), k (template row pair k = rows 2k-1, 2k) or
 * 0x8000 | block << 2 | quarter (a quarter of a window block of 8 columns). Returns the number of batches, or 0
 * when the kernel does not take the shape. Exposed so that the schedule's invariants are tested without a GPU. */
int rp_debug_stream4_schedule(int m, int n, int band, uint16_t* out, size_t out_cap);
/* The same kernel's per-(warp, step) control words: out[warp * (steps + 1) + step], warp 0..3, step 1..steps (bit layout:
 * CTL_* in dtw_stream4_kernel.cu). Returns steps, or 0 when the kernel does not take the shape. */
int rp_debug_stream4_ctl(int m, int n, int band, uint32_t* out, size_t out_cap);
/* Parses a .rpw buffer (WakewordV2 then WakewordRef, detector.rs:152-163). */
int rp_wakeword_inspect(const uint8_t* buf, size_t len, rp_wakeword_info* info);
/* Copies template t (t == -1: avg_features) of a .rpw buffer: name (RP_NAME_MAX bytes) and
 * row-major [frames][mfcc_size] floats; returns frames or <0. out may be NULL to query. */
int rp_wakeword_template(const uint8_t* buf, size_t len, int t, char* name_out, float* out, size_t out_cap_floats);
/* WakewordRef::compute_avg_samples_features + WakewordRef::new + save_to_buffer
 * (wakeword_ref_build.rs:93-110, wakeword_ref.rs:43-66, averager.rs:5-37) from already extracted and
 * normalised template matrices: data[t] is row-major [frames[t]][mfcc_size]. The host half of
 * rp_wakeword_build, exposed so the averager and the CBOR writer can be tested without a GPU.
 * Returns the .rpw size (out may be NULL to query) or <0. */
int64_t rp_wakeword_from_features(const char* name, int has_threshold, float threshold, int has_avg_threshold,
                                  float avg_threshold, int mfcc_size, int n_templates, const char* const* names,
                                  const int32_t* frames, const float* const* data, float rms_level,
                                  uint8_t* out, size_t out_cap);
/* Replays the per-stream state machine of detector.rs:377-454 over a dense score tensor as the
 * kernels produce it. scores: [n_frames][n_slots] where frame i is the i-th frame the extractor
 * emits for a fresh stream (hop i+3) and slots are, per wakeword in insertion order,
 * [avg (only if avg_features is Some)], template 0..T-1 (file order). Windows that the detector
 * would not score are ignored. vad_values: NULL or [n_frames] mean |mfcc| per frame. */
int rp_host_replay(const rp_config* cfg, const uint8_t* const* rpws, const size_t* rpw_lens, int n_rpw,
                   const float* scores, int64_t n_frames, int n_slots, const float* vad_values,
                   rp_batch_detection* out, int64_t out_cap, int64_t* n_out, uint64_t* windows_scored);

#ifdef __cplusplus
}
#endif
#endif /* RUSTPOTTER_B200_H */
