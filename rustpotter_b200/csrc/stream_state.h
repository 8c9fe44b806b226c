// Host side of the path: the wakeword set (what the kernels score against) and the per-stream
// detection state machine that consumes the kernels' window judgements. Restates
// src/detector.rs:290-454 (window bookkeeping, partial detections, countdown, eager/min_scores,
// reset-after-detection) and src/mfcc/vad.rs:11-49 on top of hop-indexed kernel output.
#pragma once

#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "rp_internal.h"
#include "score_logic.h"

namespace rp {

struct SlotRef {
    int wakeword;   // index into WakewordSet::refs
    int tmpl;       // -1 = avg_features
};

// The loaded wakewords, flattened into "slots" (one DTW template each).
struct WakewordSet {
    std::vector<std::string> keys;
    std::vector<WakewordRefData> refs;
    std::vector<WakewordMeta> metas;  // parallel to refs
    std::vector<SlotRef> slots;
    int mfcc_size = 0;
    int max_frames = 0;               // max_mfcc_frames (detector.rs:328-335)
    int max_templates = 0;            // largest n_templates over wakewords
    float target_rms_level = 0.f;     // detector.rs:331-334 (NaN when empty)

    bool empty() const { return refs.empty(); }
    // add_wakeword (detector.rs:304-327): replaces an existing key; rejects a different mfcc_size.
    void add(const std::string& key, WakewordRefData ww);
    bool remove(const std::string& key);
    bool clear();
    // resolve thresholds against the config and rebuild slots/metas
    void rebuild(const rp_config& cfg);
    const FrameMatrix& slot_matrix(int slot) const;
};

// Name of a wakeword and of its templates, snapshotted when the wakeword set changes. A pending partial detection
// keeps its snapshot by value (shared), as the reference keeps `name` and the `scores` keys inside the
// RustpotterDetection (detector.rs:487-501): removing or replacing the wakeword later cannot invalidate it.
struct WakewordNames {
    std::string name;
    std::vector<std::string> templates;
    std::vector<const char*> template_cstrs;   // into `templates`
};

// A window judged as a detection by K3 (or by judge_window on the host).
struct Hit {
    int64_t stream;
    int32_t frame;      // hop index within the current call (0-based)
    int32_t wakeword;
    float avg_score;
    float score;
    const float* scores;  // n_scores values (the wakeword's templates, file order)
    int32_t n_scores;
    std::shared_ptr<const WakewordNames> names;  // of `wakeword` at the time of the hit (may be null in host-only replays)
};

struct PartialDetection {
    int wakeword = -1;                           // index at the time of the hit; only `names` is used afterwards
    std::shared_ptr<const WakewordNames> names;
    float avg_score = 0.f, score = 0.f;
    std::vector<float> scores;
    uint64_t counter = 0;
    float gain = 1.f;
};

// VadDetector (src/mfcc/vad.rs)
struct VadState {
    float mode_value = 2.f;
    int index = 0;
    float window[50];
    int voice_countdown = 0;
    explicit VadState(float mv) : mode_value(mv) { reset(); }
    void reset();
    bool is_voice(float mean_abs_mfcc);
};

struct DetectorParams {
    int max_frames = 0;
    uint64_t min_scores = 5;
    bool eager = false;
    int vad_mode = -1;
};

class StreamState {
  public:
    void configure(const DetectorParams& p);   // (re)creates the VAD, keeps the rest
    void reset();                               // Rustpotter::reset (detector.rs:290-302)
    // Feeds one 10 ms hop. `hit` is the judgement of the window ending at this hop (nullptr: no
    // detection on it). Returns true when a detection is emitted into *out; the detector has then
    // been reset and the caller must drop the remaining hops of the current 30 ms chunk
    // (find_map, detector.rs:372-375).
    bool on_hop(const DetectorParams& p, const Hit* hit, float vad_value, float gain, PartialDetection* out);
    // Equivalent to n consecutive on_hop(p, nullptr, ...) calls when no partial detection exists
    // and VAD is off (nothing can fire); O(1).
    void skip_hops(const DetectorParams& p, int64_t n);
    bool idle() const { return !partial_.has_value(); }
    // Number of coming hops that cannot close a scorable window (extractor warm-up + window fill).
    int64_t hops_until_scorable(const DetectorParams& p) const {
        const int64_t warm = kHopsPerChunk - hops_in_ring_;
        const int64_t fill = p.max_frames - 1 - win_len_;
        return warm + (fill > 0 ? fill : 0);
    }
    const std::optional<PartialDetection>& partial() const { return partial_; }
    uint64_t windows_scored() const { return windows_scored_; }
    void clamp_window(int max_frames);          // after a wakeword change (see DESIGN.md deviations)

  private:
    bool run_detection(const DetectorParams& p, const Hit* hit, float gain, PartialDetection* out);
    int hops_in_ring_ = 0;   // hops held by the extractor ring since reset, saturating at 3
    int64_t win_len_ = 0;    // audio_mfcc_window.len()
    std::optional<PartialDetection> partial_;
    uint64_t countdown_ = 0; // detection_countdown (survives reset, as in the reference)
    std::optional<VadState> vad_;
    uint64_t windows_scored_ = 0;
};

float vad_mode_value(int mode);  // config.rs:141-148

}  // namespace rp
