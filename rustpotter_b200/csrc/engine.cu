#include "engine.h"

#include <atomic>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace rp {

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(RP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

DeviceBuffer::~DeviceBuffer() { release(); }
void DeviceBuffer::release() {
    if (p_) cudaFree(p_);
    p_ = nullptr;
    bytes_ = 0;
}
void DeviceBuffer::reserve(size_t bytes, const char* what) {
    if (bytes <= bytes_) return;
    release();
    cuda_check(cudaMalloc(&p_, bytes), what);
    bytes_ = bytes;
}

template <typename T>
static void upload(DeviceBuffer& buf, const std::vector<T>& v, cudaStream_t s, const char* what) {
    buf.reserve(std::max<size_t>(v.size() * sizeof(T), 16), what);
    if (!v.empty()) cuda_check(cudaMemcpyAsync(buf.as<void>(), v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s), what);
}

void DeviceMfccTables::upload(int mfcc_size, cudaStream_t stream) {
    MfccTables t = build_mfcc_tables(mfcc_size);
    rp::upload(hamming, t.hamming, stream, "hamming");
    rp::upload(tw480, t.tw480, stream, "twiddles");
    rp::upload(mel_bank, t.mel_bank, stream, "mel bank");
    rp::upload(centres, t.centres, stream, "mel centres");
    rp::upload(dct, t.dct, stream, "dct");
    rp::upload(up_weight, t.up_weight, stream, "mel up weights");
    rp::upload(chunks, t.chunks, stream, "mel chunks");
    rp::upload(seg_chunks, t.seg_chunks, stream, "mel segments");
    cuda_check(cudaStreamSynchronize(stream), "mfcc tables");
    dev.hamming = hamming.as<float>();
    dev.tw480 = tw480.as<float2>();
    dev.mel_bank = mel_bank.as<float>();
    dev.centres = centres.as<int>();
    dev.dct = dct.as<float>();
    dev.num_coefficients = t.num_coefficients;
    dev.up_weight = up_weight.as<float>();
    dev.chunks = chunks.as<int4>();
    dev.seg_chunks = seg_chunks.as<int2>();
    dev.n_chunks = t.n_chunks;
}

Engine::Engine(int device, int64_t n_streams) : device_(device), n_streams_(n_streams) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(RP_ERR_CUDA, std::string("no usable CUDA device (this path has no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) throw Error(RP_ERR_INVALID, "device index out of range");
    if (n_streams < 1) throw Error(RP_ERR_INVALID, "n_streams must be >= 1");
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    cuda_check(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    stream_ = own_stream_;
    cuda_check(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    if (const char* g = std::getenv("RP_GROUP_STREAMS")) {
        group_streams_ = std::max(1, std::atoi(g));
        group_streams_fixed_ = true;
    }
    for (auto& ev : ev_) cuda_check(cudaEventCreate(&ev), "cudaEventCreate");
    carry_.reserve((size_t)n_streams_ * 2 * kHopSamples * sizeof(float), "carry");
    cuda_check(cudaMemsetAsync(carry_.as<void>(), 0, carry_.bytes(), stream_), "memset carry");
    hit_count_.reserve(sizeof(int), "hit counter");
    cuda_check(cudaMallocHost(&count_host_, sizeof(int)), "cudaMallocHost");
}

Engine::~Engine() {
    cudaSetDevice(device_);
    if (own_stream_) cudaStreamSynchronize(own_stream_);
    for (auto& ev : ev_)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : group_ev_)
        if (ev) cudaEventDestroy(ev);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (hit_host_) cudaFreeHost(hit_host_);
    if (count_host_) cudaFreeHost(count_host_);
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

void Engine::set_cuda_stream(cudaStream_t s) {
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    stream_ = s ? s : own_stream_;
}

const float* Engine::last_frames_dev(int64_t* rows_per_stream, int* first_new_row) const {
    // process() flips cur_ after moving the history, so the last call wrote frames_[1 - cur_]
    if (rows_per_stream) *rows_per_stream = hist_ + frames_cap_;
    if (first_new_row) *first_new_row = hist_;
    return frames_[1 - cur_].as<float>();
}

void Engine::configure(const WakewordSet& ws, const rp_config& cfg) {
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    band_ = (int)cfg.band_size;
    score_ref_ = cfg.score_ref;
    score_mode_ = (int)cfg.score_mode;
    n_wakewords_ = (int)ws.refs.size();
    n_slots_ = (int)ws.slots.size();
    max_templates_ = ws.max_templates;
    const int old_d = d_, old_hist = hist_;
    d_ = ws.mfcc_size;
    max_frames_ = ws.max_frames;
    if (n_slots_ == 0) return;

    // templates, packed in slot order
    std::vector<float> tmpl;
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    max_slot_len_ = 0;
    for (int s = 0; s < n_slots_; s++) {
        const FrameMatrix& m = ws.slot_matrix(s);
        off.push_back((int64_t)tmpl.size());
        len.push_back(m.rows);
        max_slot_len_ = std::max(max_slot_len_, m.rows);
        tmpl.insert(tmpl.end(), m.v.begin(), m.v.end());
    }
    upload(tmpl_, tmpl, stream_, "templates");
    {   // unit-length rows for the tuned kernels: a / sqrt(|a|^2), zero rows stay zero; rows zero-padded to 16 floats
        // (mfcc_size <= 16) so that one kernel instance serves every width
        const int dp = d_ <= 16 ? 16 : d_;
        std::vector<float> unit((tmpl.size() / (size_t)d_) * (size_t)dp, 0.f);
        for (size_t r = 0, row = 0; r + d_ <= tmpl.size(); r += (size_t)d_, row++) {
            float n2 = 0.f;
            for (int k = 0; k < d_; k++) n2 += tmpl[r + k] * tmpl[r + k];
            const float inv = n2 > 0.f ? 1.f / std::sqrt(n2) : 0.f;
            for (int k = 0; k < d_; k++) unit[row * dp + k] = tmpl[r + k] * inv;
        }
        upload(tmpl_unit_, unit, stream_, "unit templates");
        tmpl_unit_floats_ = unit.size();
        static std::atomic<uint64_t> next_version{1};
        tmpl_version_ = next_version++;
        std::vector<int64_t> uoff;
        for (int64_t o : off) uoff.push_back(o / d_ * dp);
        upload(unit_off_, uoff, stream_, "unit template offsets");
        // launch lists of the avg gate: avg_features slots first, then the template slots
        std::vector<int32_t> avg_slots, tmpl_slots, slot_ww;
        for (int s2 = 0; s2 < n_slots_; s2++) {
            slot_ww.push_back(ws.slots[(size_t)s2].wakeword);
            (ws.slots[(size_t)s2].tmpl < 0 ? avg_slots : tmpl_slots).push_back(s2);
        }
        n_avg_slots_ = (int)avg_slots.size();
        n_tmpl_slots_ = (int)tmpl_slots.size();
        std::vector<int32_t> all_slots((size_t)n_slots_);
        for (int s2 = 0; s2 < n_slots_; s2++) all_slots[(size_t)s2] = s2;
        upload(all_slots_, all_slots, stream_, "slot ids");
        ww_ranges_.clear();
        bool each_fits = true;
        int tl = 0;
        for (size_t w = 0; w < ws.metas.size(); w++) {
            const WakewordMeta& mt = ws.metas[w];
            WakewordRange r;
            r.slot_begin = mt.slot_begin;
            r.n_slots = mt.n_templates + (mt.has_avg ? 1 : 0);
            r.tmpl_list_begin = tl;
            r.n_templates = mt.n_templates;
            tl += mt.n_templates;
            r.unit_begin = uoff[(size_t)r.slot_begin];
            const int last = r.slot_begin + r.n_slots - 1;
            r.unit_floats = uoff[(size_t)last] + (int64_t)len[(size_t)last] * dp - r.unit_begin;
            each_fits = each_fits && r.unit_floats <= kWindowConstFloats;
            ww_ranges_.push_back(r);
        }
        per_wakeword_const_ = (int64_t)unit.size() > kWindowConstFloats && each_fits && ws.metas.size() > 1;
        upload(avg_slots_, avg_slots, stream_, "avg slots");
        upload(tmpl_slots_, tmpl_slots, stream_, "template slots");
        upload(slot_ww_, slot_ww, stream_, "slot wakewords");
    }
    upload(slot_off_, off, stream_, "slot offsets");
    upload(slot_len_, len, stream_, "slot lengths");
    upload(metas_, ws.metas, stream_, "wakeword metas");

    if (d_ != old_d) {  // set_out_size (extractor.rs:47-59): new filter bank, extractor reset
        mfcc_tables_.upload(d_, stream_);
        frames_[0].release();
        frames_[1].release();
        frames_cap_ = 0;
        hist_ = 0;
    }
    // frame history: keep the most recent rows when max_frames changes
    const int new_hist = std::max(max_frames_ - 1, 0);
    if (new_hist != old_hist || frames_cap_ == 0) {
        const int cap = std::max(frames_cap_, kHopsPerChunk);
        DeviceBuffer nb[2];
        const size_t rows = (size_t)new_hist + cap;
        for (auto& b : nb) {
            b.reserve((size_t)n_streams_ * rows * d_ * sizeof(float), "frame buffer");
            cuda_check(cudaMemsetAsync(b.as<void>(), 0, b.bytes(), stream_), "memset frames");
        }
        if (frames_cap_ > 0 && d_ == old_d && old_hist > 0) {
            const int keep = std::min(old_hist, new_hist);
            const size_t old_rows = (size_t)old_hist + frames_cap_;
            if (keep > 0)
                cuda_check(launch_copy_rows(frames_[cur_].as<float>() + (size_t)(old_hist - keep) * d_, (int64_t)old_rows * d_,
                                            nb[0].as<float>() + (size_t)(new_hist - keep) * d_, (int64_t)rows * d_, n_streams_,
                                            (int64_t)keep * d_, stream_), "history move");
        }
        cuda_check(cudaStreamSynchronize(stream_), "sync");
        frames_[0].swap(nb[0]);
        frames_[1].swap(nb[1]);
        cur_ = 0;
        hist_ = new_hist;
        frames_cap_ = cap;
    }
    cuda_check(cudaStreamSynchronize(stream_), "configure");
}

void Engine::set_filters(const rp_config& cfg) {
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    filt_gain_ = cfg.gain_normalizer_enabled != 0;
    filt_bp_ = cfg.band_pass_enabled != 0;
    filt_fixed_ref_ = cfg.gain_ref_set != 0;
    filt_min_gain_ = cfg.min_gain;
    filt_max_gain_ = cfg.max_gain;
    filt_ref_ = filt_fixed_ref_ ? cfg.gain_ref : std::nanf("");
    filt_ref_sqrt_ = filt_fixed_ref_ ? std::sqrt(cfg.gain_ref) : std::nanf("");
    filt_window_ = 1;  // GainNormalizerFilter::new: window_size 1 until set_rms_level_ref
    if (filt_gain_) {
        gain_window_.reserve((size_t)n_streams_ * kGainWindowCap * sizeof(float), "gain window");
        gain_count_.reserve((size_t)n_streams_ * 2 * sizeof(int), "gain window count");
        cuda_check(cudaMemsetAsync(gain_count_.as<void>(), 0, (size_t)n_streams_ * 2 * sizeof(int), stream_), "memset");
    }
    if (filt_bp_) {
        // BandPassFilter::new (band_pass_filter.rs:31-55), f32
        const float pi = 3.14159265358979323846f, sr = (float)kSampleRate;
        const float omega_low = 2.0f * pi * cfg.low_cutoff / sr, omega_high = 2.0f * pi * cfg.high_cutoff / sr;
        const float alpha_low = std::sin(omega_low) / 2.0f, alpha_high = std::sin(omega_high) / 2.0f;
        const float a0 = 1.0f / (1.0f + alpha_high - alpha_low);
        bp_[0] = a0;
        bp_[1] = -2.0f * std::cos(omega_low) * a0;
        bp_[2] = (1.0f - alpha_high - alpha_low) * a0;
        bp_[3] = -2.0f * std::cos(omega_high) * a0;
        bp_[4] = (1.0f - alpha_high + alpha_low) * a0;
        bp_state_.reserve((size_t)n_streams_ * 4 * sizeof(float), "band-pass state");
        cuda_check(cudaMemsetAsync(bp_state_.as<void>(), 0, (size_t)n_streams_ * 4 * sizeof(float), stream_), "memset");
    }
    cuda_check(cudaStreamSynchronize(stream_), "sync");
}

void Engine::set_gain_reference(float target_rms_level, int window_size) {
    if (!filt_fixed_ref_) {
        filt_ref_ = target_rms_level;
        filt_ref_sqrt_ = std::sqrt(target_rms_level);
    }
    filt_window_ = std::min(window_size != 0 ? window_size : 1, kGainWindowCap - 1);
}

void Engine::ensure_frames(int n_new) {
    if (n_new <= frames_cap_) return;
    const size_t old_rows = (size_t)hist_ + frames_cap_, rows = (size_t)hist_ + n_new;
    DeviceBuffer nb[2];
    for (auto& b : nb) b.reserve((size_t)n_streams_ * rows * d_ * sizeof(float), "frame buffer");
    cuda_check(cudaMemsetAsync(nb[0].as<void>(), 0, nb[0].bytes(), stream_), "memset frames");
    if (hist_ > 0)
        cuda_check(launch_copy_rows(frames_[cur_].as<float>(), (int64_t)old_rows * d_, nb[0].as<float>(), (int64_t)rows * d_,
                                    n_streams_, (int64_t)hist_ * d_, stream_), "history move");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    frames_[0].swap(nb[0]);
    frames_[1].swap(nb[1]);
    cur_ = 0;
    frames_cap_ = n_new;
}

void Engine::process(const AudioIn& in, bool want_vad, int first_window, std::vector<HitRecord>& hits, std::vector<float>* vad) {
    hits.clear();
    if (n_slots_ == 0) return;
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    const int64_t S = in.samples;
    const bool on_device = in.on_device;
    const bool decode = in.needs_decode();
    const int n_new = (int)(S / kHopSamples);
    ensure_frames(n_new);
    launches = 0;
    const int64_t rows = (int64_t)hist_ + frames_cap_;
    const int64_t n_windows = n_streams_ * (int64_t)n_new;
    const int stride = 5 + max_templates_;
    tscore_.reserve((size_t)n_windows * n_slots_ * sizeof(float), "window scores");
    hits_.reserve((size_t)n_windows * stride * sizeof(float), "hit list");
    float* vad_dev = nullptr;
    if (want_vad) {
        vad_.reserve((size_t)n_windows * sizeof(float), "vad values");
        vad_dev = vad_.as<float>();
    }
    const bool filters = filt_gain_ || filt_bp_;
    if (!on_device || filters || decode) audio_.reserve((size_t)n_streams_ * S * sizeof(float), "audio staging");
    const size_t raw_stream = in.bytes_per_stream();
    if (!on_device && decode) raw_.reserve((size_t)n_streams_ * raw_stream, "raw audio staging");
    // what K1 (or the filter stage) reads: the caller's f32 device audio as it is, everything else through the f32 staging buffer
    const float* f32_in = (on_device && !decode) ? static_cast<const float*>(in.data) : audio_.as<float>();
    const float* src_all = filters ? audio_.as<float>() : f32_in;
    const uint8_t* raw_all = !decode ? nullptr : (on_device ? static_cast<const uint8_t*>(in.data) : raw_.as<uint8_t>());
    const int n_chunks = (int)(S / kFrameSamples);
    if (filt_gain_) gains_.reserve((size_t)n_streams_ * n_chunks * sizeof(float), "gains");

    // The batch is processed in groups of streams: group g's kernels wait only for group g's H2D copy,
    // so the copy of group g+1 (copy stream) overlaps the kernels of group g, and a group's frames
    // (~33 MB for 512 streams x 10 s) are still in L2 when its DTW kernel reads them.
    // (short calls — the reference's 30 ms cadence — take larger groups: at least ~16 MB of f32 audio per group, so that a
    // chunk-by-chunk caller pays one launch chain per call instead of one per 256 streams)
    int64_t gs = std::min<int64_t>(group_streams_, n_streams_);
    if (!group_streams_fixed_) {
        const int64_t min_streams = ((int64_t)16 << 20) / std::max<int64_t>(1, S * (int64_t)sizeof(float));
        gs = std::min<int64_t>(n_streams_, std::max<int64_t>(gs, min_streams));
    }
    const int n_groups = (int)((n_streams_ + gs - 1) / gs);
    while ((int)group_ev_.size() < 4 * n_groups) {
        cudaEvent_t ev;
        cuda_check(cudaEventCreate(&ev), "cudaEventCreate");
        group_ev_.push_back(ev);
    }
    cuda_check(cudaEventRecord(ev_[0], stream_), "event");
    if (!on_device) {
        cuda_check(cudaStreamWaitEvent(copy_stream_, ev_[0], 0), "wait");
        const uint8_t* host = static_cast<const uint8_t*>(in.data);
        uint8_t* dst = decode ? raw_.as<uint8_t>() : reinterpret_cast<uint8_t*>(audio_.as<float>());
        for (int g = 0; g < n_groups; g++) {
            const int64_t b0 = g * gs, nb = std::min(gs, n_streams_ - b0);
            cuda_check(cudaMemcpyAsync(dst + (size_t)b0 * raw_stream, host + (size_t)b0 * raw_stream, (size_t)nb * raw_stream,
                                       cudaMemcpyHostToDevice, copy_stream_), "H2D audio");
            cuda_check(cudaEventRecord(group_ev_[4 * g], copy_stream_), "event");
        }
    }
    // avg gate: tile verdicts of this call
    const bool tuned = dtw_variant_ != 1 && dtw_windows_tuned_supported(d_, band_, max_slot_len_, max_frames_);
    const int fw = std::min(std::max(first_window, 0), n_new);
    const int tile = dtw_windows_tile();
    const int j_blocks = (n_new - fw + tile - 1) / tile;
    // short calls (the reference's 30 ms cadence: three windows per stream) take the warp-per-window-triple kernel
    const bool cadence = tuned && dtw_variant_ != 9 && n_new - fw > 0 && n_new - fw <= dtw_windows_cadence_max_new() &&
                         dtw_windows_cadence_supported(d_, band_, max_slot_len_, max_frames_);
    const bool gated = tuned && !cadence && avg_gate_ && n_avg_slots_ > 0 && n_tmpl_slots_ > 0 && j_blocks > 0;
    if (gated) tile_pass_.reserve((size_t)n_streams_ * j_blocks * n_wakewords_, "avg-gate tiles");
    last_gated_ = gated;
    last_j_blocks_ = j_blocks;
    last_first_window_ = fw;
    cuda_check(cudaMemsetAsync(hit_count_.as<void>(), 0, sizeof(int), stream_), "memset");
    float* fb_all = frames_[cur_].as<float>();
    for (int g = 0; g < n_groups; g++) {
        const int64_t b0 = g * gs, nb = std::min(gs, n_streams_ - b0);
        if (!on_device) cuda_check(cudaStreamWaitEvent(stream_, group_ev_[4 * g], 0), "wait");
        cuda_check(cudaEventRecord(group_ev_[4 * g + 1], stream_), "event");
        if (decode) {  // Sample::into_f32 + channel 0 (audio_types.rs:98-137, encoder.rs:41-48) -> f32 staging
            cuda_check(launch_decode_samples(raw_all + (size_t)b0 * raw_stream, (int64_t)raw_stream, in.fmt, in.channels, in.big_endian ? 1 : 0,
                                             audio_.as<float>() + b0 * S, S, nb, S, stream_), "decode kernel");
            launches += 1;
        }
        if (filters) {  // gain normaliser / band pass: caller's (or staged) audio -> staging buffer
            FilterArgs fa;
            fa.in = f32_in + b0 * S;
            fa.in_stride = S;
            fa.out = audio_.as<float>() + b0 * S;
            fa.out_stride = S;
            fa.n_streams = nb;
            fa.n_chunks = n_chunks;
            fa.gain = filt_gain_;
            fa.rms_level_ref = filt_ref_;
            fa.rms_level_sqrt = filt_ref_sqrt_;
            fa.min_gain = filt_min_gain_;
            fa.max_gain = filt_max_gain_;
            fa.window_size = filt_window_;
            fa.window_cap = kGainWindowCap;
            fa.gain_window = filt_gain_ ? gain_window_.as<float>() + b0 * kGainWindowCap : nullptr;
            fa.gain_count = filt_gain_ ? gain_count_.as<int>() + b0 * 2 : nullptr;
            fa.gains_out = filt_gain_ ? gains_.as<float>() + b0 * n_chunks : nullptr;
            fa.band_pass = filt_bp_;
            fa.a0 = bp_[0]; fa.a1 = bp_[1]; fa.a2 = bp_[2]; fa.b1 = bp_[3]; fa.b2 = bp_[4];
            fa.bp_state = filt_bp_ ? bp_state_.as<float>() + b0 * 4 : nullptr;
            cuda_check(launch_audio_filters(fa, stream_), "filter kernel");
            launches += 1;
        }
        const float* src = src_all + b0 * S;
        float* fb = fb_all + b0 * rows * d_;
        float* carry = carry_.as<float>() + b0 * 2 * kHopSamples;
        // K1: frame j of this call ends at new hop j and starts two hops earlier (carry or earlier audio)
        cuda_check(launch_mfcc_frames(src, S, carry, nb, n_new, -2 * kHopSamples, mfcc_tables_.dev, fb, rows, hist_,
                                      vad_dev ? vad_dev + b0 * n_new : nullptr, stream_), "mfcc kernel");
        cuda_check(launch_copy_rows(src + (S - 2 * kHopSamples), S, carry, 2 * kHopSamples, nb, 2 * kHopSamples, stream_),
                   "carry update");
        cuda_check(cudaEventRecord(group_ev_[4 * g + 2], stream_), "event");
        // K2: window scores; K3: judgement -> compact hit list
        DtwWindowsArgs wa;
        wa.frames = fb;
        wa.frame_rows = rows;
        wa.first_window_row = hist_ - (max_frames_ - 1);
        wa.n_new = n_new;
        wa.n_streams = nb;
        wa.d = d_;
        wa.tmpl = tmpl_.as<float>();
        wa.slot_off = slot_off_.as<int64_t>();
        wa.slot_len = slot_len_.as<int32_t>();
        wa.n_slots = n_slots_;
        wa.max_len = max_slot_len_;
        wa.window_len = max_frames_;
        wa.band = band_;
        wa.score_ref = score_ref_;
        wa.scores = tscore_.as<float>() + b0 * (int64_t)n_new * n_slots_;
        wa.first_window = fw;
        if (wa.first_window > 0)   // rows no kernel writes read as NaN in rp_batch_copy_last_scores
            cuda_check(cudaMemset2DAsync(wa.scores, (size_t)n_new * n_slots_ * sizeof(float), 0xff,
                                         (size_t)wa.first_window * n_slots_ * sizeof(float), (size_t)nb, stream_), "memset skipped rows");
        // (an avg_features matrix longer than every template scores a window shorter than itself: generic kernel only)
        if (cadence) {
            cuda_check(launch_dtw_windows_cadence(wa, tmpl_unit_.as<float>(), unit_off_.as<int64_t>(), stream_), "dtw cadence kernel");
        } else if (tuned) {
            WindowGate wg;
            wg.unit_off = unit_off_.as<int64_t>();
            wg.slot_ww = slot_ww_.as<int32_t>();
            wg.metas = metas_.as<WakewordMeta>();
            wg.n_wakewords = n_wakewords_;
            auto launch = [&](const char* what) {
                cuda_check(launch_dtw_windows_d16(wa, wg, tmpl_unit_.as<float>(), tmpl_unit_floats_, tmpl_version_, stream_), what);
            };
            if (gated) {
                // wakeword_comp.rs:85-94: avg_features first; templates only where some window of the tile passes the avg gate
                wg.tile_pass = tile_pass_.as<unsigned char>() + (size_t)b0 * j_blocks * n_wakewords_;
                cuda_check(cudaMemsetAsync(wg.tile_pass, 1, (size_t)nb * j_blocks * n_wakewords_, stream_), "memset tiles");
                wg.slots = avg_slots_.as<int32_t>();
                wg.n_slots = n_avg_slots_;
                wg.gate = 1;
                wg.const_floats = per_wakeword_const_ ? -1 : 0;   // (the avg slots of several wakewords: shared-memory templates)
                launch("dtw window kernel (avg)");
                wg.gate = 2;
                if (per_wakeword_const_) {
                    for (const WakewordRange& r : ww_ranges_) {   // one wakeword's templates in constant memory at a time
                        wg.slots = tmpl_slots_.as<int32_t>() + r.tmpl_list_begin;
                        wg.n_slots = r.n_templates;
                        wg.const_begin = r.unit_begin;
                        wg.const_floats = r.unit_floats;
                        launch("dtw window kernel (templates)");
                        launches += 1;
                    }
                } else {
                    wg.slots = tmpl_slots_.as<int32_t>();
                    wg.n_slots = n_tmpl_slots_;
                    launch("dtw window kernel (templates)");
                    launches += 1;
                }
            } else if (per_wakeword_const_) {
                for (const WakewordRange& r : ww_ranges_) {
                    wg.slots = all_slots_.as<int32_t>() + r.slot_begin;
                    wg.n_slots = r.n_slots;
                    wg.const_begin = r.unit_begin;
                    wg.const_floats = r.unit_floats;
                    launch("dtw window kernel");
                    launches += 1;
                }
                launches -= 1;
            } else {
                launch("dtw window kernel");
            }
        } else {
            cuda_check(launch_dtw_windows_generic(wa, stream_), "dtw kernel");
        }
        JudgeArgs ja;
        ja.scores = wa.scores;
        ja.n_streams = nb;
        ja.n_new = n_new;
        ja.n_slots = n_slots_;
        ja.metas = metas_.as<WakewordMeta>();
        ja.n_wakewords = n_wakewords_;
        ja.score_mode = score_mode_;
        ja.max_templates = max_templates_;
        ja.hit_count = hit_count_.as<int>();
        ja.hits = hits_.as<float>();
        ja.capacity = n_windows;
        ja.stream_base = (int)b0;
        ja.first_window = wa.first_window;
        cuda_check(launch_judge_windows(ja, stream_), "judge kernel");
        // next call's history = the last hist_ rows of [history | new frames]
        if (hist_ > 0)
            cuda_check(launch_copy_rows(fb + (size_t)n_new * d_, rows * d_, frames_[1 - cur_].as<float>() + b0 * rows * d_,
                                        rows * d_, nb, (int64_t)hist_ * d_, stream_), "history move");
        cuda_check(cudaEventRecord(group_ev_[4 * g + 3], stream_), "event");
        launches += 5;
    }
    cur_ ^= 1;
    last_n_new_ = n_new;
    cuda_check(cudaEventRecord(ev_[3], stream_), "event");

    // D2H: count, then the records
    cuda_check(cudaMemcpyAsync(count_host_, hit_count_.as<void>(), sizeof(int), cudaMemcpyDeviceToHost, stream_), "D2H count");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    const int n_hits = (int)std::min<int64_t>(*count_host_, n_windows);
    if (n_hits > 0) {
        const size_t need = (size_t)n_hits * stride;
        if (need > hit_host_floats_) {
            if (hit_host_) cudaFreeHost(hit_host_);
            hit_host_ = nullptr;
            hit_host_floats_ = 0;
            cuda_check(cudaMallocHost(&hit_host_, need * 2 * sizeof(float)), "cudaMallocHost hits");
            hit_host_floats_ = need * 2;
        }
        cuda_check(cudaMemcpyAsync(hit_host_, hits_.as<void>(), need * sizeof(float), cudaMemcpyDeviceToHost, stream_), "D2H hits");
    }
    if (filt_gain_) {
        gains_host_.resize((size_t)n_streams_ * n_chunks);
        cuda_check(cudaMemcpyAsync(gains_host_.data(), gains_.as<void>(), gains_host_.size() * sizeof(float), cudaMemcpyDeviceToHost, stream_), "D2H gains");
    }
    if (want_vad && vad) {
        vad->resize((size_t)n_windows);
        cuda_check(cudaMemcpyAsync(vad->data(), vad_dev, (size_t)n_windows * sizeof(float), cudaMemcpyDeviceToHost, stream_), "D2H vad");
    }
    cuda_check(cudaEventRecord(ev_[4], stream_), "event");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    // stage times: [0] H2D (first copy start .. last copy end; overlaps the kernels), [1] MFCC and
    // [2] DTW + judge summed over groups, [3] D2H of the hit list
    timings_ms[0] = timings_ms[1] = timings_ms[2] = 0.f;
    if (!on_device) cudaEventElapsedTime(&timings_ms[0], ev_[0], group_ev_[4 * (n_groups - 1)]);
    for (int g = 0; g < n_groups; g++) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, group_ev_[4 * g + 1], group_ev_[4 * g + 2]);
        cudaEventElapsedTime(&b, group_ev_[4 * g + 2], group_ev_[4 * g + 3]);
        timings_ms[1] += a;
        timings_ms[2] += b;
    }
    cudaEventElapsedTime(&timings_ms[3], ev_[3], ev_[4]);

    hits.resize((size_t)n_hits);
    for (int i = 0; i < n_hits; i++) {
        const float* rec = hit_host_ + (size_t)i * stride;
        HitRecord& h = hits[i];
        std::memcpy(&h.stream, rec + 0, 4);
        std::memcpy(&h.frame, rec + 1, 4);
        std::memcpy(&h.wakeword, rec + 2, 4);
        h.avg_score = rec[3];
        h.score = rec[4];
        h.scores = rec + 5;
    }
    std::sort(hits.begin(), hits.end(), [](const HitRecord& a, const HitRecord& b) {
        return a.stream != b.stream ? a.stream < b.stream : a.frame < b.frame;
    });
}

void Engine::last_gate_stats(int64_t* tiles, int64_t* passed) const {
    int64_t t = 0, p = 0;
    if (last_gated_) {
        cuda_check(cudaSetDevice(device_), "cudaSetDevice");
        std::vector<unsigned char> h((size_t)n_streams_ * last_j_blocks_ * n_wakewords_);
        cuda_check(cudaMemcpy(h.data(), tile_pass_.as<void>(), h.size(), cudaMemcpyDeviceToHost), "D2H avg-gate tiles");
        t = (int64_t)h.size();
        for (unsigned char v : h) p += v ? 1 : 0;
    }
    if (tiles) *tiles = t;
    if (passed) *passed = p;
}

void Engine::copy_last_scores(float* out) const {
    const int64_t n = n_streams_ * (int64_t)last_n_new_ * n_slots_;
    if (n <= 0) return;
    cuda_check(cudaSetDevice(device_), "cudaSetDevice");
    cuda_check(cudaMemcpy(out, tscore_.as<void>(), (size_t)n * sizeof(float), cudaMemcpyDeviceToHost), "D2H scores");
    if (!last_gated_) return;
    // template slots of the tiles the avg gate skipped were never written
    std::vector<unsigned char> h((size_t)n_streams_ * last_j_blocks_ * n_wakewords_);
    cuda_check(cudaMemcpy(h.data(), tile_pass_.as<void>(), h.size(), cudaMemcpyDeviceToHost), "D2H avg-gate tiles");
    std::vector<int32_t> tmpl_slots((size_t)n_tmpl_slots_), slot_ww((size_t)n_slots_);
    cuda_check(cudaMemcpy(tmpl_slots.data(), tmpl_slots_.as<void>(), tmpl_slots.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "D2H");
    cuda_check(cudaMemcpy(slot_ww.data(), slot_ww_.as<void>(), slot_ww.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "D2H");
    const int tile = dtw_windows_tile();
    const float nan = std::nanf("");
    for (int64_t b = 0; b < n_streams_; b++)
        for (int j = last_first_window_; j < last_n_new_; j++) {
            const int jb = (j - last_first_window_) / tile;
            for (int32_t s : tmpl_slots)
                if (!h[(size_t)(b * last_j_blocks_ + jb) * n_wakewords_ + slot_ww[(size_t)s]])
                    out[(b * last_n_new_ + j) * n_slots_ + s] = nan;
        }
}

}  // namespace rp
