// Ties the engine (kernels) to the per-stream host state machines; shared by the per-stream handle
// (`Rustpotter`, one stream) and the batched front-end (N streams).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "engine.h"
#include "stream_state.h"

namespace rp {

struct Emitted {
    int64_t stream;
    int64_t chunk;
    PartialDetection det;
};

void validate_detector_config(const rp_config& cfg);  // throws RP_ERR_INVALID
std::shared_ptr<const WakewordNames> make_wakeword_names(const WakewordRefData& r);

class DetectorCore {
  public:
    DetectorCore(const rp_config& cfg, int64_t n_streams, int device);

    // Rustpotter::add_wakeword (detector.rs:304-327)
    void add_wakeword(const std::string& key, const uint8_t* buf, size_t len);
    bool remove_wakeword(const std::string& key);  // :180-189
    bool remove_wakewords();                       // :193-202
    void update_detector_config(const rp_config& cfg);  // :265-282
    void update_filters_config(const rp_config& cfg);   // :283-289 (batched front-end: filters run on the GPU)
    void enable_device_filters(const rp_config& cfg);   // batched front-end only
    void reset();                                       // :290-302

    // in.samples = n_chunks * 480 mono samples per stream; gains: per-chunk gain stamped on detections (nullptr = 1.0)
    // chunk_hops: 10 ms hops that make up one caller-visible chunk (one process_samples call of the reference): 3 without
    // a resampler; with one, whatever FftFixedInOut emits per call (a multiple of 160 samples is required here).
    void process(const AudioIn& in, const float* gains, std::vector<Emitted>& out, int chunk_hops = kHopsPerChunk);
    void process(const float* audio, int64_t samples_per_stream, bool on_device, const float* gains, std::vector<Emitted>& out,
                 int chunk_hops = kHopsPerChunk) {
        AudioIn in;
        in.data = audio;
        in.samples = samples_per_stream;
        in.on_device = on_device;
        process(in, gains, out, chunk_hops);
    }

    // Fills an rp_detection whose pointers refer to the detection's own name snapshot and to `score_store` (caller-owned).
    void fill_detection(const PartialDetection& d, rp_detection* out, std::vector<float>& score_store) const;
    const std::optional<PartialDetection>& partial(int64_t stream) const { return states_[stream].partial(); }
    uint64_t windows_scored() const;

    const rp_config& config() const { return cfg_; }
    const WakewordSet& wakewords() const { return ws_; }
    Engine& engine() { return *engine_; }
    const Engine& engine() const { return *engine_; }
    float host_ms = 0.f;

  private:
    void on_wakeword_change();
    rp_config cfg_;
    WakewordSet ws_;
    DetectorParams params_;
    std::unique_ptr<Engine> engine_;
    std::vector<StreamState> states_;
    std::vector<std::shared_ptr<const WakewordNames>> names_;  // per wakeword; a fresh snapshot after every change
    std::vector<HitRecord> hits_;
    std::vector<float> vad_;
    bool device_filters_ = false;
};

}  // namespace rp
