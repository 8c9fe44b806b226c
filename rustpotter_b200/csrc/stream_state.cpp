#include "stream_state.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace rp {

// ---------------------------------------------------------------- WakewordSet
void WakewordSet::add(const std::string& key, WakewordRefData ww) {
    if (!refs.empty() && refs.front().mfcc_size != ww.mfcc_size)
        throw Error(RP_ERR_MISMATCH, "Usage of wakewords with different mfcc size is not supported, ignoring wakeword");
    for (size_t i = 0; i < keys.size(); i++)
        if (keys[i] == key) {
            refs[i] = std::move(ww);
            return;
        }
    keys.push_back(key);
    refs.push_back(std::move(ww));
}

bool WakewordSet::remove(const std::string& key) {
    for (size_t i = 0; i < keys.size(); i++)
        if (keys[i] == key) {
            keys.erase(keys.begin() + i);
            refs.erase(refs.begin() + i);
            return true;
        }
    return false;
}

bool WakewordSet::clear() {
    bool had = !refs.empty();
    keys.clear();
    refs.clear();
    return had;
}

void WakewordSet::rebuild(const rp_config& cfg) {
    metas.clear();
    slots.clear();
    max_frames = 0;
    max_templates = 0;
    mfcc_size = refs.empty() ? 0 : refs.front().mfcc_size;
    target_rms_level = std::numeric_limits<float>::quiet_NaN();
    for (size_t w = 0; w < refs.size(); w++) {
        const WakewordRefData& r = refs[w];
        WakewordMeta m;
        m.slot_begin = (int)slots.size();
        m.n_templates = (int)r.samples_features.size();
        m.has_avg = r.avg_features.has_value() ? 1 : 0;
        m.threshold = r.threshold.value_or(cfg.threshold);
        m.avg_threshold = r.avg_threshold.value_or(cfg.avg_threshold);
        if (m.has_avg) slots.push_back({(int)w, -1});
        for (int t = 0; t < m.n_templates; t++) slots.push_back({(int)w, t});
        metas.push_back(m);
        max_frames = std::max(max_frames, r.max_frames());
        max_templates = std::max(max_templates, m.n_templates);
        target_rms_level = std::fmax(r.rms_level, target_rms_level);  // f32::max ignores NaN
    }
}

const FrameMatrix& WakewordSet::slot_matrix(int slot) const {
    const SlotRef& s = slots[slot];
    const WakewordRefData& r = refs[s.wakeword];
    return s.tmpl < 0 ? *r.avg_features : r.samples_features[s.tmpl].second;
}

// ---------------------------------------------------------------- VAD (src/mfcc/vad.rs)
float vad_mode_value(int mode) { return mode == 0 ? 2.f : mode == 1 ? 2.5f : 3.f; }

void VadState::reset() {
    for (float& v : window) v = std::numeric_limits<float>::quiet_NaN();
    voice_countdown = 0;
    index = 0;
}

bool VadState::is_voice(float value) {
    window[index] = value;
    index = index >= 49 ? 0 : index + 1;
    float mn = std::numeric_limits<float>::infinity();
    for (float v : window)
        if (!std::isnan(v) && v < mn) mn = v;
    mn = std::fmax(mn, 0.01f);
    float th = mn * mode_value;
    int n_high = 0;
    for (float v : window)
        if (v > th) n_high++;
    if (n_high > 10) voice_countdown = 500;
    if (voice_countdown > 0) {
        voice_countdown--;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------- StreamState
void StreamState::configure(const DetectorParams& p) {
    if (p.vad_mode >= 0) vad_.emplace(vad_mode_value(p.vad_mode));
    else vad_.reset();
}

void StreamState::reset() {
    partial_.reset();
    win_len_ = 0;
    hops_in_ring_ = 0;
    if (vad_) vad_->reset();
}

void StreamState::clamp_window(int max_frames) {
    if (max_frames > 0 && win_len_ > max_frames - 1) win_len_ = max_frames - 1;
}

bool StreamState::on_hop(const DetectorParams& p, const Hit* hit, float vad_value, float gain, PartialDetection* out) {
    // MfccExtractor::process_audio_part (extractor.rs:69-79): a frame is produced only if the ring
    // already held 480 samples before this hop was appended.
    if (hops_in_ring_ < kHopsPerChunk) {
        hops_in_ring_++;
        return false;
    }
    // Rustpotter::process_new_mfccs (detector.rs:377-397)
    bool should_run = partial_.has_value() || !vad_ || vad_->is_voice(vad_value);
    win_len_++;
    if (win_len_ >= p.max_frames) {
        if (should_run && run_detection(p, hit, gain, out)) return true;
    }
    if (win_len_ >= p.max_frames) win_len_--;
    return false;
}

// Rustpotter::run_detection (detector.rs:398-432)
bool StreamState::run_detection(const DetectorParams& p, const Hit* hit, float gain, PartialDetection* out) {
    if (countdown_ != 0) countdown_--;
    if (partial_ && (countdown_ == 0 || (p.eager && partial_->counter >= p.min_scores))) {
        PartialDetection d = std::move(*partial_);
        partial_.reset();
        if (d.counter >= p.min_scores) {
            reset();
            *out = std::move(d);
            return true;
        }
    }
    windows_scored_++;  // run_wakeword_detectors (detector.rs:433)
    if (hit) {
        uint64_t counter = partial_ ? partial_->counter + 1 : 1;
        if (!partial_ || partial_->score < hit->score) {
            PartialDetection d;
            d.wakeword = hit->wakeword;
            d.names = hit->names;
            d.avg_score = hit->avg_score;
            d.score = hit->score;
            d.counter = counter;
            d.gain = gain;
            d.scores.assign(hit->scores, hit->scores + hit->n_scores);
            partial_ = std::move(d);
        } else {
            partial_->counter = counter;
        }
        countdown_ = (uint64_t)(p.max_frames / 2);
    }
    return false;
}

void StreamState::skip_hops(const DetectorParams& p, int64_t n) {
    int64_t warm = std::min<int64_t>(n, kHopsPerChunk - hops_in_ring_);
    hops_in_ring_ += (int)warm;
    int64_t emitted = n - warm;
    if (emitted <= 0) return;
    int64_t fill = std::max<int64_t>(0, (int64_t)p.max_frames - 1 - win_len_);
    int64_t scored = std::max<int64_t>(0, emitted - fill);
    win_len_ = std::min<int64_t>(win_len_ + emitted, std::max<int64_t>(win_len_, (int64_t)p.max_frames - 1));
    countdown_ = countdown_ > (uint64_t)scored ? countdown_ - (uint64_t)scored : 0;
    windows_scored_ += (uint64_t)scored;
}

}  // namespace rp
