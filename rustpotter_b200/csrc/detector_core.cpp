#include "detector_core.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

namespace rp {

DetectorCore::DetectorCore(const rp_config& cfg, int64_t n_streams, int device) : cfg_(cfg) {
    if (cfg.sample_rate < 1000 || cfg.sample_rate > 768000)
        throw Error(RP_ERR_INVALID, "Unsupported sample rate, unable to initialize the resampler");   // encoder.rs:78
    if (cfg.sample_format > RP_FMT_F32 || cfg.channels == 0 || cfg.endianness > RP_ENDIAN_NATIVE)
        throw Error(RP_ERR_INVALID, "invalid configuration value");
    validate_detector_config(cfg);
    engine_ = std::make_unique<Engine>(device, n_streams);
    states_.resize((size_t)n_streams);
    params_.min_scores = cfg.min_scores;
    params_.eager = cfg.eager != 0;
    params_.vad_mode = cfg.vad_mode;
    for (auto& s : states_) s.configure(params_);
}

void DetectorCore::add_wakeword(const std::string& key, const uint8_t* buf, size_t len) {
    WakewordRefData ww = parse_rpw(buf, len);
    const bool first = ws_.empty();
    ws_.add(key, std::move(ww));  // throws RP_ERR_MISMATCH before touching anything
    if (first) reset();           // detector.rs:305-307 (set_out_size happens in Engine::configure)
    on_wakeword_change();
}

bool DetectorCore::remove_wakeword(const std::string& key) {
    if (!ws_.remove(key)) return false;
    on_wakeword_change();
    return true;
}

bool DetectorCore::remove_wakewords() {
    if (!ws_.clear()) return false;
    on_wakeword_change();
    return true;
}

std::shared_ptr<const WakewordNames> make_wakeword_names(const WakewordRefData& r) {
    auto n = std::make_shared<WakewordNames>();
    n->name = r.name;
    for (auto& t : r.samples_features) n->templates.push_back(t.first);
    for (auto& t : n->templates) n->template_cstrs.push_back(t.c_str());
    return n;
}

void DetectorCore::on_wakeword_change() {
    ws_.rebuild(cfg_);
    params_.max_frames = ws_.max_frames;
    engine_->configure(ws_, cfg_);
    names_.clear();
    for (auto& r : ws_.refs) names_.push_back(make_wakeword_names(r));
    // Deviation: the reference would keep scoring a window longer than the new
    // max_mfcc_frames after a removal; here the window is clamped to the new length.
    for (auto& s : states_) s.clamp_window(params_.max_frames);
    if (device_filters_) engine_->set_gain_reference(ws_.target_rms_level, ws_.max_frames / 3);  // detector.rs:336-338
}

// The enum-valued fields of DetectorConfig (config.rs:86-97,134-139): one check for construction and updates.
void validate_detector_config(const rp_config& cfg) {
    if (cfg.score_mode > RP_SCORE_P95) throw Error(RP_ERR_INVALID, "invalid score_mode");
    if (cfg.vad_mode < RP_VAD_NONE || cfg.vad_mode > RP_VAD_HARD) throw Error(RP_ERR_INVALID, "invalid vad_mode");
    if (cfg.band_size > 0xffffu) throw Error(RP_ERR_INVALID, "band_size does not fit the reference's u16");
}

void DetectorCore::update_detector_config(const rp_config& cfg) {
    validate_detector_config(cfg);   // nothing is changed when a value is out of range
    cfg_.avg_threshold = cfg.avg_threshold;
    cfg_.threshold = cfg.threshold;
    cfg_.min_scores = cfg.min_scores;
    cfg_.eager = cfg.eager;
    cfg_.band_size = cfg.band_size;
    cfg_.score_ref = cfg.score_ref;
    cfg_.score_mode = cfg.score_mode;
    cfg_.vad_mode = cfg.vad_mode;
    params_.min_scores = cfg.min_scores;
    params_.eager = cfg.eager != 0;
    params_.vad_mode = cfg.vad_mode;
    for (auto& s : states_) s.configure(params_);
    ws_.rebuild(cfg_);
    engine_->configure(ws_, cfg_);
    reset();
}

void DetectorCore::enable_device_filters(const rp_config& cfg) {
    device_filters_ = true;
    cfg_.gain_normalizer_enabled = cfg.gain_normalizer_enabled;
    cfg_.band_pass_enabled = cfg.band_pass_enabled;
    engine_->set_filters(cfg);
}

void DetectorCore::update_filters_config(const rp_config& cfg) {
    if (device_filters_) enable_device_filters(cfg);  // fresh filters; gain reference waits for the next wakeword change
    reset();
}

void DetectorCore::reset() {
    for (auto& s : states_) s.reset();
}

uint64_t DetectorCore::windows_scored() const {
    uint64_t t = 0;
    for (auto& s : states_) t += s.windows_scored();
    return t;
}

void DetectorCore::process(const AudioIn& in, const float* gains, std::vector<Emitted>& out, int chunk_hops) {
    out.clear();
    if (ws_.empty()) return;  // detector.rs:348-350: audio is dropped, extractor untouched
    const int64_t S = in.samples;
    if (chunk_hops < 1 || S <= 0 || S % ((int64_t)kHopSamples * chunk_hops) != 0)
        throw Error(RP_ERR_INVALID, "samples_per_stream must be a positive multiple of the chunk length (480 mono samples at 16 kHz)");
    if (!in.data || in.channels < 1 || in.fmt < RP_FMT_I8 || in.fmt > RP_FMT_F32) throw Error(RP_ERR_INVALID, "bad audio description");
    const bool vad = params_.vad_mode >= 0;
    // leading hops of this call that no stream can turn into a scored window (e.g. a freshly reset batch)
    int64_t skip = S / kHopSamples;
    for (const StreamState& st : states_) skip = std::min(skip, st.hops_until_scorable(params_));
    engine_->process(in, vad, (int)std::max<int64_t>(skip, 0), hits_, vad ? &vad_ : nullptr);

    const auto t0 = std::chrono::steady_clock::now();
    const int64_t n_hops = S / kHopSamples;
    const int64_t n_chunks = n_hops / chunk_hops;
    const int64_t B = engine_->n_streams();
    const float* dev_gains = (device_filters_ && engine_->gain_filter_enabled()) ? engine_->last_gains().data() : nullptr;
    size_t hi = 0;
    for (int64_t b = 0; b < B; b++) {
        StreamState& st = states_[(size_t)b];
        size_t h0 = hi;
        while (hi < hits_.size() && hits_[hi].stream == b) hi++;
        if (h0 == hi && st.idle() && !vad) {  // nothing can fire on this stream in this call
            st.skip_hops(params_, n_hops);
            continue;
        }
        size_t hp = h0;
        for (int64_t c = 0; c < n_chunks;) {
            if (st.idle() && !vad) {  // jump to the chunk that holds the next judged detection
                while (hp < hi && hits_[hp].frame < c * chunk_hops) hp++;
                const int64_t next = hp < hi ? hits_[hp].frame / chunk_hops : n_chunks;
                if (next > c) {
                    st.skip_hops(params_, (next - c) * chunk_hops);
                    c = next;
                    continue;
                }
            }
            const float gain = gains ? gains[c] : (dev_gains ? dev_gains[(size_t)(b * n_chunks + c)] : 1.f);
            for (int k = 0; k < chunk_hops; k++) {
                const int64_t j = c * chunk_hops + k;
                while (hp < hi && hits_[hp].frame < j) hp++;
                Hit hit;
                const Hit* hptr = nullptr;
                if (hp < hi && hits_[hp].frame == j) {
                    const HitRecord& r = hits_[hp];
                    hit.stream = b;
                    hit.frame = r.frame;
                    hit.wakeword = r.wakeword;
                    hit.avg_score = r.avg_score;
                    hit.score = r.score;
                    hit.scores = r.scores;
                    hit.n_scores = ws_.metas[(size_t)r.wakeword].n_templates;
                    hit.names = names_[(size_t)r.wakeword];
                    hptr = &hit;
                }
                PartialDetection det;
                const float vv = vad ? vad_[(size_t)(b * n_hops + j)] : 0.f;
                if (st.on_hop(params_, hptr, vv, gain, &det)) {
                    out.push_back(Emitted{b, c, std::move(det)});
                    break;  // find_map: the rest of this chunk's frames are dropped (detector.rs:372-375)
                }
            }
            c++;
        }
    }
    host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void DetectorCore::fill_detection(const PartialDetection& d, rp_detection* out, std::vector<float>& score_store) const {
    std::memset(out, 0, sizeof(*out));
    if (!d.names) throw Error(RP_ERR_INVALID, "detection without a wakeword name snapshot");
    std::snprintf(out->name, RP_NAME_MAX, "%s", d.names->name.c_str());
    out->avg_score = d.avg_score;
    out->score = d.score;
    out->counter = d.counter;
    out->gain = d.gain;
    score_store = d.scores;
    // (a replaced wakeword may have fewer templates than the detection has scores: the snapshot is the detection's own)
    out->n_scores = (uint32_t)std::min(score_store.size(), d.names->template_cstrs.size());
    out->score_names = d.names->template_cstrs.data();
    out->score_values = score_store.data();
}

}  // namespace rp
