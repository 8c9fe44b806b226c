/*
 * rp_oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a plain C++17 restatement of the reference algorithm
 * (GiviMAD/rustpotter v3.0.2, the MFCC + WakewordRef/DTW scoring path). It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
 * check and time the CUDA path against the reference's arithmetic. Nothing under
 * rustpotter_b200/ may include, link or call anything declared here.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this oracle against every
 * WakewordRef golden of the reference's tests/detector.rs (detector.rs:9-214, the two 48 kHz
 * goldens through the restated rubato resampler included) and against the template matrices stored in the reference's .rpw
 * fixtures (which are the reference's own MFCC+CMN output for the fixture wavs).
 */
#ifndef RP_ORACLE_H
#define RP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field layout as include/rustpotter_b200.h:rp_config so one ctypes Structure serves both. */
typedef struct rpo_config {
    /* AudioFmt (reference src/config.rs:9-29) */
    uint32_t sample_rate;      /* default 16000 */
    uint32_t sample_format;    /* 0=i8 1=i16 2=i32 3=f32 (default 3) */
    uint32_t channels;         /* default 1 */
    uint32_t endianness;       /* 0=little 1=big 2=native (default 0) */
    /* DetectorConfig (src/config.rs:172-208) */
    float avg_threshold;       /* 0.2 */
    float threshold;           /* 0.5 */
    uint64_t min_scores;       /* 5 */
    uint32_t eager;            /* 0 */
    float score_ref;           /* 0.22 */
    uint32_t band_size;        /* 5 */
    uint32_t score_mode;       /* 0=Average 1=Max 2=Median 3=P25 4=P50 5=P75 6=P80 7=P90 8=P95 (default 1) */
    int32_t vad_mode;          /* -1 none, 0 easy, 1 medium, 2 hard */
    /* FiltersConfig (src/config.rs:31-84) */
    uint32_t gain_normalizer_enabled; /* 0 */
    uint32_t gain_ref_set;            /* 0 => use wakeword rms level */
    float gain_ref;
    float min_gain;                   /* 0.1 */
    float max_gain;                   /* 1.0 */
    uint32_t band_pass_enabled;       /* 0 */
    float low_cutoff;                 /* 80 */
    float high_cutoff;                /* 400 */
} rpo_config;

#define RPO_NAME_MAX 128
#define RPO_MAX_SCORES 64

typedef struct rpo_detection {
    char name[RPO_NAME_MAX];
    float avg_score;
    float score;
    uint64_t counter;
    float gain;
    uint32_t n_scores;
    char score_names[RPO_MAX_SCORES][RPO_NAME_MAX];
    float score_values[RPO_MAX_SCORES];
} rpo_detection;

void rpo_config_default(rpo_config* cfg);

/* ---- MFCC (reference src/mfcc/extractor.rs) ---- */
/* Number of frames MfccExtractor emits for a fresh stream of n_samples (multiple of 160 used). */
size_t rpo_mfcc_num_frames(size_t n_samples);
/* Runs a fresh MfccExtractor (set_out_size(mfcc_size)) over the samples in 160-sample hops.
 * out: [frames][mfcc_size]. Returns the number of frames written. */
size_t rpo_mfcc_stream(const float* audio, size_t n_samples, int mfcc_size, float* out);
/* One frame from exactly 480 already pre-emphasised samples (extract_mfccs, extractor.rs:80). */
void rpo_mfcc_frame(const float* samples480, int mfcc_size, float* out);
void rpo_mel_centres(int mfcc_size, int* out /* [mfcc_size+3] */);
void rpo_hamming(float* out480);
void rpo_mel_bank(int mfcc_size, float* out /* [(mfcc_size+1)][240] */);

/* ---- DTW / comparator (src/mfcc/dtw.rs, comparator.rs, normalizer.rs) ---- */
float rpo_dtw_cost(const float* a, int m, const float* b, int n, int d, int band);
float rpo_compare(const float* a, int m, const float* b, int n, int d, int band, float score_ref);
void rpo_normalize(float* frames, int n, int d);
/* Batch of independent compares; pair p uses a + a_off[p] (a_len[p] rows) etc. cmn!=0 applies
 * MfccNormalizer::normalize to the window (b) first, as cut_and_normalize_frame does. */
void rpo_compare_pairs(const float* a, const int64_t* a_off, const int32_t* a_len,
                       const float* b, const int64_t* b_off, const int32_t* b_len,
                       int64_t n_pairs, int d, int band, float score_ref, int cmn,
                       float* out, int n_threads);
/* get_percentile / score aggregation (wakeword_comp.rs:38-49,108-139) */
float rpo_aggregate(const float* scores, int n, int score_mode);

/* ---- Wakeword file (.rpw CBOR; wakeword_ref.rs, wakeword_v2.rs, wakeword_file.rs) ---- */
typedef struct rpo_wakeword rpo_wakeword;
rpo_wakeword* rpo_wakeword_load(const uint8_t* buf, size_t len, char* err, size_t err_len);
void rpo_wakeword_free(rpo_wakeword*);
const char* rpo_wakeword_name(const rpo_wakeword*);
int rpo_wakeword_mfcc_size(const rpo_wakeword*);
int rpo_wakeword_num_templates(const rpo_wakeword*);
const char* rpo_wakeword_template_name(const rpo_wakeword*, int t);
int rpo_wakeword_template_frames(const rpo_wakeword*, int t);        /* t == -1: avg_features (0 if none) */
const float* rpo_wakeword_template_data(const rpo_wakeword*, int t); /* row-major [frames][mfcc_size] copy */
float rpo_wakeword_rms_level(const rpo_wakeword*);
int rpo_wakeword_threshold(const rpo_wakeword*, float* thr);         /* 1 if Some */
int rpo_wakeword_avg_threshold(const rpo_wakeword*, float* thr);
/* Build a synthetic WakewordRef from raw matrices and serialise it as CBOR (used by tests and
 * bench to create D=16 wakewords). Returns bytes written or required size if out==NULL. */
size_t rpo_wakeword_encode(const char* name, int mfcc_size, int n_templates, const char* const* names,
                           const int32_t* frames, const float* const* data,
                           int avg_frames, const float* avg_data, float rms_level,
                           int has_thr, float thr, int has_avg_thr, float avg_thr, int v2,
                           uint8_t* out, size_t out_cap);

/* WakewordRef::new_from_sample_files (from_files != 0: rms_level = median over samples) /
 * new_from_sample_buffers (rms_level = max) — src/wakewords/comp/wakeword_ref_build.rs:9-110, with
 * MfccWavFileExtractor (wav_file_extractor.rs:18-91) and MfccAverager (averager.rs:5-37). `wavs` are whole
 * WAV files (16 kHz). Writes the .rpw CBOR; returns its size (0 + err on failure). */
size_t rpo_wakeword_build(const char* name, int has_thr, float thr, int has_avg_thr, float avg_thr, int n_samples,
                          const char* const* sample_names, const uint8_t* const* wavs, const size_t* wav_lens, int mfcc_size,
                          int from_files, uint8_t* out, size_t out_cap, char* err, size_t err_len);

/* ---- Detector (src/detector.rs) ---- */
typedef struct rpo_detector rpo_detector;
/* rubato FftFixedInOut restated (source rate -> 16 kHz, one channel): whole input chunks through a fresh resampler.
 * Returns output samples written (out may be NULL to query) or -1; *in_chunk = input chunk length. */
int64_t rpo_resample_to_16k(uint32_t fs_in, const float* in, size_t n_in, float* out, size_t out_cap, size_t* in_chunk);
rpo_detector* rpo_detector_new(const rpo_config* cfg, char* err, size_t err_len);
void rpo_detector_free(rpo_detector*);
int rpo_detector_add_wakeword_from_buffer(rpo_detector*, const char* key, const uint8_t* buf, size_t len,
                                          char* err, size_t err_len);
int rpo_detector_remove_wakeword(rpo_detector*, const char* key);
int rpo_detector_remove_wakewords(rpo_detector*);
size_t rpo_detector_samples_per_frame(const rpo_detector*);
size_t rpo_detector_bytes_per_frame(const rpo_detector*);
/* return 1 = detection written to *out, 0 = None */
int rpo_detector_process_bytes(rpo_detector*, const uint8_t* bytes, size_t len, rpo_detection* out);
int rpo_detector_process_f32(rpo_detector*, const float* samples, size_t n, rpo_detection* out);
int rpo_detector_process_i16(rpo_detector*, const int16_t* samples, size_t n, rpo_detection* out);
int rpo_detector_process_i32(rpo_detector*, const int32_t* samples, size_t n, rpo_detection* out);
int rpo_detector_process_i8(rpo_detector*, const int8_t* samples, size_t n, rpo_detection* out);
int rpo_detector_get_partial(const rpo_detector*, rpo_detection* out);
float rpo_detector_rms_level(const rpo_detector*);
float rpo_detector_gain(const rpo_detector*);
float rpo_detector_rms_level_ref(const rpo_detector*);
void rpo_detector_update_config(rpo_detector*, const rpo_config* cfg);
void rpo_detector_reset(rpo_detector*);
uint64_t rpo_detector_windows_scored(const rpo_detector*); /* calls of run_wakeword_detectors */

/* Dense per-window trace for parity tests: for a fresh detector without VAD/filters, feed
 * n_samples (multiple of 480) of f32 mono 16 kHz audio and record, for every scored window in
 * order, [avg_score, score, scores[0..T)] of wakeword `key` computed WITHOUT the avg gate or the
 * threshold (every template is always scored). Returns windows written (<= max_windows). */
size_t rpo_trace_window_scores(const rpo_config* cfg, const uint8_t* rpw, size_t rpw_len,
                               const float* audio, size_t n_samples, float* out, size_t max_windows);

/* ---- CPU baseline: B independent detectors over [B][S] f32 streams (S multiple of 480),
 * n_threads host threads, streams partitioned contiguously. Returns total windows scored;
 * det_counts[b] = detections of stream b; first min(det_counts[b], max_det) detections of
 * stream b are written to dets[b*max_det ...] when dets != NULL. */
uint64_t rpo_run_streams(const rpo_config* cfg, const uint8_t* const* rpws, const size_t* rpw_lens, int n_rpw,
                         const float* audio, int64_t n_streams, int64_t samples_per_stream, int n_threads,
                         int32_t* det_counts, rpo_detection* dets, int max_det);

#ifdef __cplusplus
}
#endif
#endif
