// Device side of one batch of streams: buffers resident in HBM, the per-call launch sequence
// (H2D -> K1 MFCC -> K2 DTW window scores -> K3 judge -> D2H of the compact hit list).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "kernels.h"
#include "mfcc_tables.h"
#include "rp_internal.h"
#include "stream_state.h"

namespace rp {

void cuda_check(cudaError_t e, const char* what);

class DeviceBuffer {
  public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer();
    // grows (never shrinks); contents are NOT preserved on growth
    void reserve(size_t bytes, const char* what);
    void release();
    void swap(DeviceBuffer& o) {
        void* p = p_; p_ = o.p_; o.p_ = p;
        size_t b = bytes_; bytes_ = o.bytes_; o.bytes_ = b;
    }
    template <typename T> T* as() const { return static_cast<T*>(p_); }
    size_t bytes() const { return bytes_; }
  private:
    void* p_ = nullptr;
    size_t bytes_ = 0;
};

// MFCC tables resident on the current device.
struct DeviceMfccTables {
    DeviceBuffer hamming, tw480, mel_bank, centres, dct, up_weight, chunks, seg_chunks;
    MfccTablesDev dev;
    void upload(int mfcc_size, cudaStream_t stream);
};

struct HitRecord {     // one judged detection of K3, host copy
    int32_t stream, frame, wakeword;
    float avg_score, score;
    const float* scores;  // into Engine::hit_host_
};

// One call's audio: [n_streams][samples * channels] interleaved samples of `fmt`, host or device memory.
struct AudioIn {
    const void* data = nullptr;
    int fmt = RP_FMT_F32;      // RP_FMT_*: what `data` holds
    int channels = 1;          // interleaved channels; channel 0 is used (encoder.rs:41-48)
    bool big_endian = false;   // byte order of multi-byte samples in `data`
    bool on_device = false;
    int64_t samples = 0;       // MONO samples per stream (a multiple of 480)
    size_t bytes_per_sample() const { return fmt == RP_FMT_I8 ? 1 : fmt == RP_FMT_I16 ? 2 : 4; }
    size_t bytes_per_stream() const { return (size_t)samples * channels * bytes_per_sample(); }
    bool needs_decode() const { return fmt != RP_FMT_F32 || channels != 1 || big_endian; }
};

class Engine {
  public:
    Engine(int device, int64_t n_streams);
    ~Engine();
    Engine(const Engine&) = delete;

    int device() const { return device_; }
    int64_t n_streams() const { return n_streams_; }
    void set_cuda_stream(cudaStream_t s);  // caller-owned stream (nullptr restores the private one)
    cudaStream_t cuda_stream() const { return stream_; }

    // Uploads templates / tables for the wakeword set; keeps per-stream frame history.
    void configure(const WakewordSet& ws, const rp_config& cfg);
    void set_dtw_variant(int v) { dtw_variant_ = v; }
    // Avg gate of the tuned window kernel (wakeword_comp.rs:85-94): true = avg slots first, template slots only for the
    // window tiles in which some window passes (default); false = every slot of every window (dense score tensor).
    void set_avg_gate(bool on) { avg_gate_ = on; }
    // FiltersConfig (src/config.rs:31-84): fresh filter state for every stream (update_filters_config semantics:
    // the gain reference is NOT re-derived from the wakewords until set_gain_reference is called again).
    void set_filters(const rp_config& cfg);
    // GainNormalizerFilter::set_rms_level_ref (gain_normalizer_filter.rs:42-48), on every wakeword change
    void set_gain_reference(float target_rms_level, int window_size);
    bool gain_filter_enabled() const { return filt_gain_; }
    // gains applied per (stream, chunk) by the last process() when the gain normaliser is enabled
    const std::vector<float>& last_gains() const { return gains_host_; }

    // Scores samples_per_stream/160 new hops per stream. Returns the hits sorted by (stream, frame);
    // `vad` (if want_vad) receives [n_streams][n_hops] mean |mfcc| per new frame.
    // first_window: no stream can use the windows that end before this new hop (fresh / just-reset streams).
    void process(const AudioIn& in, bool want_vad, int first_window, std::vector<HitRecord>& hits, std::vector<float>* vad);
    // Avg-gate statistics of the last process(): (stream, 128-window block, wakeword) tiles and how many of them had a
    // window that passed the avg gate (the template slots of the others were skipped). 0/0 when the gate was off.
    void last_gate_stats(int64_t* tiles, int64_t* passed) const;
    // Host copy of the last call's dense scores; entries of gated-out tiles (never computed) and of the skipped leading
    // windows are NaN.
    void copy_last_scores(float* out_host) const;

    // diagnostics
    float timings_ms[5] = {0, 0, 0, 0, 0};
    int launches = 0;
    // dense scores of the last call [n_streams][n_new][n_slots] (device) — parity tests read it
    const float* last_scores_dev() const { return tscore_.as<float>(); }
    const float* last_frames_dev(int64_t* rows_per_stream, int* first_new_row) const;
    int n_slots() const { return n_slots_; }
    int last_n_new() const { return last_n_new_; }

  private:
    void ensure_frames(int n_new);

    int device_ = 0;
    int64_t n_streams_ = 0;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr, copy_stream_ = nullptr;
    cudaEvent_t ev_[6] = {};
    std::vector<cudaEvent_t> group_ev_;   // [group][4]: copy done, start, after MFCC, after DTW/judge
    int group_streams_ = 256;             // streams per pipeline group (RP_GROUP_STREAMS)
    bool group_streams_fixed_ = false;    // set by RP_GROUP_STREAMS: no automatic enlargement for short calls
    int dtw_variant_ = 0;
    bool avg_gate_ = true;

    // wakeword set on device
    int d_ = 0, max_frames_ = 0, n_slots_ = 0, n_wakewords_ = 0, max_templates_ = 0, max_slot_len_ = 0;
    int band_ = 5, score_mode_ = 1;
    float score_ref_ = 0.22f;
    DeviceBuffer tmpl_, tmpl_unit_, slot_off_, slot_len_, metas_;
    DeviceBuffer unit_off_, avg_slots_, tmpl_slots_, slot_ww_;   // tuned window kernel: padded-template offsets, launch lists
    DeviceBuffer all_slots_;   // identity list 0 .. n_slots-1 (sub-lists = one wakeword's slots)
    int n_avg_slots_ = 0, n_tmpl_slots_ = 0;
    struct WakewordRange {     // where one wakeword sits in the slot lists and in tmpl_unit_
        int slot_begin, n_slots, tmpl_list_begin, n_templates;
        int64_t unit_begin, unit_floats;
    };
    std::vector<WakewordRange> ww_ranges_;
    bool per_wakeword_const_ = false;   // the set exceeds the kernel's constant-memory copy but every wakeword fits: one launch each
    DeviceBuffer tile_pass_;   // [B][j_blocks][n_wakewords] avg-gate verdict per window tile
    int last_j_blocks_ = 0, last_first_window_ = 0;
    bool last_gated_ = false;
    size_t tmpl_unit_floats_ = 0;   // floats in tmpl_unit_
    uint64_t tmpl_version_ = 0;     // changes with every upload of the templates (the window kernel's constant-memory copy follows it)
    // MFCC tables on device
    DeviceMfccTables mfcc_tables_;
    // per-stream state
    DeviceBuffer carry_;       // [B][320] last two hops of audio
    DeviceBuffer frames_[2];   // [B][hist_ + frames_cap_][d]
    int cur_ = 0;
    int hist_ = 0;             // history rows kept in front of the new frames (= max_frames - 1)
    int frames_cap_ = 0;       // new-frame capacity per stream
    // per-call
    DeviceBuffer audio_;       // [B][S] f32 staging (host audio, decoded samples, filter output)
    DeviceBuffer raw_;         // [B][S * channels * bytes] staging for host audio that needs decoding
    DeviceBuffer tscore_;      // [B][n_new][n_slots]
    DeviceBuffer vad_;         // [B][n_new]
    DeviceBuffer hits_;        // [cap][5 + max_templates]
    DeviceBuffer hit_count_;   // int
    int last_n_new_ = 0;
    // audio filters (batched front-end)
    bool filt_gain_ = false, filt_bp_ = false, filt_fixed_ref_ = false;
    float filt_ref_ = 0.f, filt_ref_sqrt_ = 0.f, filt_min_gain_ = 0.1f, filt_max_gain_ = 1.f;
    int filt_window_ = 1;
    float bp_[5] = {0, 0, 0, 0, 0};
    static constexpr int kGainWindowCap = 512;
    DeviceBuffer gain_window_, gain_count_, bp_state_, gains_;
    std::vector<float> gains_host_;
    // pinned host staging for the hit list
    float* hit_host_ = nullptr;
    size_t hit_host_floats_ = 0;
    int* count_host_ = nullptr;
};

}  // namespace rp
