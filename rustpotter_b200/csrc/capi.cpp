// extern "C" surface declared in include/rustpotter_b200.h.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <thread>

#include "audio_frontend.h"
#include "detector_core.h"
#include "wakeword_builder.h"
#include "engine.h"
#include "kernels.h"
#include "mfcc_tables.h"
#include "rp_internal.h"
#include "stream_state.h"

using namespace rp;

namespace {
constexpr uint32_t kMagicHandle = 0x52504831;  // "RPH1"
constexpr uint32_t kMagicBatch = 0x52504231;   // "RPB1"
struct HandleBase {
    uint32_t magic;
    std::string error;
};
std::atomic<int> g_dtw_variant{0};   // debug / A-B knob, process-wide (see rp_set_dtw_variant)
}  // namespace

struct rp_handle {
    HandleBase base{kMagicHandle, {}};
    std::unique_ptr<DetectorCore> core;
    AudioIngest ingest;
    std::optional<GainNormalizer> gain_filter;
    std::optional<BandPass> band_pass;
    float rms_level = 0.f, gain = 1.f;
    std::vector<Emitted> emitted;
    std::vector<float> score_store;
    mutable std::vector<float> partial_store;
};

// One shard = the contiguous stream range [begin, begin + count) on one device (streams shard with no collective).
struct BatchShard {
    std::unique_ptr<DetectorCore> core;
    int64_t begin = 0, count = 0;
    int device = 0;
    std::vector<Emitted> emitted;
};
struct rp_batch {
    HandleBase base{kMagicBatch, {}};
    std::vector<BatchShard> shards;
    int64_t n_streams = 0;
    rp_config cfg{};
    std::vector<rp_batch_detection> dets;
    std::vector<std::vector<float>> stores;
    std::vector<float> timings = std::vector<float>(5, 0.f);
    // source rate != 16 kHz (encoder.rs:72-79): one FftFixedInOut per stream, run on the host in front of the device path
    std::vector<std::unique_ptr<FftResampler>> resamplers;
    std::vector<float> resampled;   // [n_streams][calls * output_frames]
    DetectorCore& first() { return *shards.front().core; }
    const DetectorCore& first() const { return *shards.front().core; }
};

namespace {

template <typename H, typename F>
int guarded(H* h, F&& f) {
    try {
        return f();
    } catch (const Error& e) {
        if (h) h->base.error = e.what();
        set_thread_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        if (h) h->base.error = e.what();
        set_thread_error(e.what());
        return RP_ERR_INVALID;
    }
}

void make_filters(rp_handle* h, const rp_config& c) {
    h->gain_filter.reset();
    h->band_pass.reset();
    if (c.gain_normalizer_enabled)
        h->gain_filter.emplace(c.min_gain, c.max_gain, c.gain_ref_set ? std::optional<float>(c.gain_ref) : std::nullopt);
    if (c.band_pass_enabled) h->band_pass.emplace((float)kSampleRate, c.low_cutoff, c.high_cutoff);
}

void after_wakeword_change(rp_handle* h) {  // detector.rs:336-338
    const WakewordSet& ws = h->core->wakewords();
    if (h->gain_filter) h->gain_filter->set_rms_level_ref(ws.target_rms_level, (size_t)(ws.max_frames / 3));
}

// Rustpotter::process_audio (detector.rs:347-376)
int process_audio(rp_handle* h, std::vector<float> audio, rp_detection* out) {
    if (h->core->wakewords().empty()) return 0;
    h->rms_level = GainNormalizer::rms_level(audio);
    if (h->gain_filter) h->gain = h->gain_filter->filter(audio, h->rms_level);
    if (h->band_pass) h->band_pass->filter(audio);
    h->core->process(audio.data(), (int64_t)audio.size(), false, &h->gain, h->emitted, (int)(audio.size() / (size_t)kHopSamples));
    if (h->emitted.empty()) return 0;
    if (out) h->core->fill_detection(h->emitted.front().det, out, h->score_store);
    return 1;
}

template <typename T>
int process_samples(rp_handle* h, const T* samples, size_t n, float max_value, rp_detection* out) {
    if (!h) return RP_ERR_INVALID;
    return guarded(h, [&] {
        if (!samples || n != h->ingest.input_samples_per_frame) return 0;  // detector.rs:249-251
        return process_audio(h, h->ingest.convert(samples, n, max_value), out);
    });
}

// device-resident MFCC tables for the raw kernel entry point, one per (device, mfcc_size)
std::mutex g_table_mutex;
std::map<std::pair<int, int>, std::unique_ptr<DeviceMfccTables>> g_tables;

const MfccTablesDev& tables_for(int device, int mfcc_size, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_table_mutex);
    auto key = std::make_pair(device, mfcc_size);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) return it->second->dev;
    auto c = std::make_unique<DeviceMfccTables>();
    c->upload(mfcc_size, s);
    auto& ref = *c;
    g_tables[key] = std::move(c);
    return ref.dev;
}

}  // namespace

namespace {
std::atomic<int> g_avg_gate{-1};   // -1 default (gate on), 0 dense, 1 gate on: debug / A-B knob, process-wide

// Runs f(shard) on every shard: inline for one shard, one host thread per device otherwise (each thread binds its device).
template <typename F>
void for_each_shard(rp_batch* b, F&& f) {
    if (b->shards.size() == 1) {
        f(b->shards.front());
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> errs(b->shards.size());
    for (size_t i = 0; i < b->shards.size(); i++)
        th.emplace_back([&, i] {
            try {
                f(b->shards[i]);
            } catch (...) {
                errs[i] = std::current_exception();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : errs)
        if (e) std::rethrow_exception(e);
}

int batch_process(rp_batch* b, AudioIn in, const rp_batch_detection** dets, int64_t* n_dets) {
    if (b->shards.size() > 1 && in.on_device)
        throw Error(RP_ERR_INVALID, "a multi-device batch takes host audio (each device copies its own stream range)");
    if (!b->resamplers.empty()) {   // reencode_to_mono_with_sample_rate for every stream, then the 16 kHz path
        if (in.on_device) throw Error(RP_ERR_UNSUPPORTED, "resampling (sample_rate != 16000) takes host audio");
        const size_t nin = b->resamplers.front()->input_frames(), nout = b->resamplers.front()->output_frames();
        if (in.samples <= 0 || (size_t)in.samples % nin != 0)
            throw Error(RP_ERR_INVALID, "samples_per_stream must be a multiple of the resampler's input chunk (rp_batch_samples_per_frame)");
        const size_t calls = (size_t)in.samples / nin;
        b->resampled.resize((size_t)b->n_streams * calls * nout);
        AudioIngest dec;
        dec.fmt = (uint32_t)in.fmt;
        dec.channels = (uint32_t)in.channels;
        dec.endianness = in.big_endian ? RP_ENDIAN_BIG : RP_ENDIAN_LITTLE;
        const uint8_t* src = static_cast<const uint8_t*>(in.data);
        const size_t sb = in.bytes_per_stream(), chunk_bytes = nin * (size_t)in.channels * in.bytes_per_sample();
        const unsigned hw = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        const unsigned nt = (unsigned)std::min<int64_t>(hw, b->n_streams);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                for (int64_t s2 = t; s2 < b->n_streams; s2 += nt)
                    for (size_t c = 0; c < calls; c++) {
                        const std::vector<float> mono = dec.decode_bytes_mono(src + (size_t)s2 * sb + c * chunk_bytes, chunk_bytes);
                        b->resamplers[(size_t)s2]->process(mono.data(), b->resampled.data() + ((size_t)s2 * calls + c) * nout);
                    }
            });
        for (auto& t : th) t.join();
        in = AudioIn();
        in.data = b->resampled.data();
        in.samples = (int64_t)(calls * nout);
    }
    const size_t stream_bytes = in.bytes_per_stream();
    const uint8_t* base = static_cast<const uint8_t*>(in.data);
    const int gate = g_avg_gate.load();
    for_each_shard(b, [&](BatchShard& sh) {
        AudioIn mine = in;
        mine.data = base + (size_t)sh.begin * stream_bytes;
        sh.core->engine().set_dtw_variant(g_dtw_variant.load());
        sh.core->engine().set_avg_gate(gate != 0);
        sh.core->process(mine, nullptr, sh.emitted);
    });
    size_t total = 0;
    for (auto& sh : b->shards) total += sh.emitted.size();
    b->dets.resize(total);
    b->stores.resize(total);
    size_t i = 0;
    for (auto& sh : b->shards)
        for (auto& e : sh.emitted) {   // shards are contiguous and ascending: (stream, chunk) order is kept
            b->dets[i].stream = sh.begin + e.stream;
            b->dets[i].chunk = e.chunk;
            sh.core->fill_detection(e.det, &b->dets[i].det, b->stores[i]);
            i++;
        }
    std::fill(b->timings.begin(), b->timings.end(), 0.f);
    for (auto& sh : b->shards) {   // devices run concurrently: the call's stage time is the slowest shard's
        for (int k = 0; k < 4; k++) b->timings[k] = std::max(b->timings[k], sh.core->engine().timings_ms[k]);
        b->timings[4] = std::max(b->timings[4], sh.core->host_ms);
    }
    if (dets) *dets = b->dets.data();
    if (n_dets) *n_dets = (int64_t)b->dets.size();
    return RP_OK;
}
}  // namespace

extern "C" {

const char* rp_version(void) { return "rustpotter-b200 0.1.0 (path of rustpotter 3.0.2)"; }

int rp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void rp_config_default(rp_config* c) {  // src/config.rs:20-29,43-52,63-71,193-208
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->sample_rate = 16000;
    c->sample_format = RP_FMT_F32;
    c->channels = 1;
    c->endianness = RP_ENDIAN_LITTLE;
    c->avg_threshold = 0.2f;
    c->threshold = 0.5f;
    c->min_scores = 5;
    c->eager = 0;
    c->score_ref = 0.22f;
    c->band_size = 5;
    c->score_mode = RP_SCORE_MAX;
    c->vad_mode = RP_VAD_NONE;
    c->gain_normalizer_enabled = 0;
    c->gain_ref_set = 0;
    c->gain_ref = 0.f;
    c->min_gain = 0.1f;
    c->max_gain = 1.0f;
    c->band_pass_enabled = 0;
    c->low_cutoff = 80.f;
    c->high_cutoff = 400.f;
}

const char* rp_last_error(const void* handle_or_null) {
    if (!handle_or_null) return thread_error();
    const HandleBase* b = static_cast<const HandleBase*>(handle_or_null);
    if (b->magic != kMagicHandle && b->magic != kMagicBatch) return "invalid handle";
    return b->error.c_str();
}

// ------------------------------------------------------------------ per-stream handle
int rp_create(const rp_config* cfg, int device, rp_handle** out) {
    if (!cfg || !out) return RP_ERR_INVALID;
    *out = nullptr;
    return guarded((rp_handle*)nullptr, [&] {
        auto h = std::make_unique<rp_handle>();
        h->core = std::make_unique<DetectorCore>(*cfg, 1, device);
        h->ingest.configure(*cfg);   // AudioEncoder::new: sample decoding and, off 16 kHz, the FFT resampler
        if (h->ingest.output_samples_per_frame() % (size_t)kHopSamples != 0)
            throw Error(RP_ERR_UNSUPPORTED, "this source rate makes the resampler emit chunks that are not whole 10 ms hops");
        make_filters(h.get(), *cfg);
        *out = h.release();
        return RP_OK;
    });
}

void rp_destroy(rp_handle* h) { delete h; }

int rp_add_wakeword_from_buffer(rp_handle* h, const char* key, const uint8_t* buf, size_t len) {
    if (!h || !key || !buf) return RP_ERR_INVALID;
    return guarded(h, [&] {
        h->core->add_wakeword(key, buf, len);
        after_wakeword_change(h);
        return RP_OK;
    });
}

int rp_add_wakeword_from_file(rp_handle* h, const char* key, const char* path) {
    if (!h || !key || !path) return RP_ERR_INVALID;
    return guarded(h, [&] {
        std::vector<uint8_t> b = read_file(path);
        h->core->add_wakeword(key, b.data(), b.size());
        after_wakeword_change(h);
        return RP_OK;
    });
}

int rp_remove_wakeword(rp_handle* h, const char* key) {
    if (!h || !key) return RP_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->core->remove_wakeword(key)) return 0;
        after_wakeword_change(h);
        return 1;
    });
}

int rp_remove_wakewords(rp_handle* h) {
    if (!h) return RP_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->core->remove_wakewords()) return 0;
        after_wakeword_change(h);
        return 1;
    });
}

size_t rp_get_samples_per_frame(const rp_handle* h) { return h ? h->ingest.input_samples_per_frame : 0; }
size_t rp_get_bytes_per_frame(const rp_handle* h) { return h ? h->ingest.input_bytes_per_frame() : 0; }

int rp_get_partial_detection(const rp_handle* h, rp_detection* out) {
    if (!h) return RP_ERR_INVALID;
    const auto& p = h->core->partial(0);
    if (!p) return 0;
    if (out) h->core->fill_detection(*p, out, h->partial_store);
    return 1;
}

float rp_get_rms_level(const rp_handle* h) { return h ? h->rms_level : 0.f; }
float rp_get_gain(const rp_handle* h) { return h ? h->gain : 1.f; }
float rp_get_rms_level_ref(const rp_handle* h) {
    return h && h->gain_filter ? h->gain_filter->rms_level_ref : std::numeric_limits<float>::quiet_NaN();
}

int rp_process_bytes(rp_handle* h, const uint8_t* bytes, size_t len, rp_detection* out) {
    if (!h) return RP_ERR_INVALID;
    return guarded(h, [&] {
        if (!bytes || len != h->ingest.input_bytes_per_frame()) return 0;  // detector.rs:235-237
        return process_audio(h, h->ingest.decode_bytes(bytes, len), out);
    });
}
int rp_process_samples_i8(rp_handle* h, const int8_t* s, size_t n, rp_detection* out) { return process_samples(h, s, n, 127.f, out); }
int rp_process_samples_i16(rp_handle* h, const int16_t* s, size_t n, rp_detection* out) { return process_samples(h, s, n, 32767.f, out); }
int rp_process_samples_i32(rp_handle* h, const int32_t* s, size_t n, rp_detection* out) {
    return process_samples(h, s, n, (float)2147483647, out);
}
int rp_process_samples_f32(rp_handle* h, const float* s, size_t n, rp_detection* out) { return process_samples(h, s, n, 0.f, out); }

int rp_update_detector_config(rp_handle* h, const rp_config* cfg) {
    if (!h || !cfg) return RP_ERR_INVALID;
    return guarded(h, [&] {
        h->core->update_detector_config(*cfg);
        return RP_OK;
    });
}
int rp_update_filters_config(rp_handle* h, const rp_config* cfg) {
    if (!h || !cfg) return RP_ERR_INVALID;
    return guarded(h, [&] {
        // detector.rs:283-289: fresh filters (the gain reference is NOT re-derived from the
        // wakewords until the next wakeword change), then reset()
        make_filters(h, *cfg);
        h->core->reset();
        return RP_OK;
    });
}
int rp_update_config(rp_handle* h, const rp_config* cfg) {
    int r = rp_update_detector_config(h, cfg);
    return r != RP_OK ? r : rp_update_filters_config(h, cfg);
}
void rp_reset(rp_handle* h) {
    if (h) h->core->reset();
}
uint64_t rp_windows_scored(const rp_handle* h) { return h ? h->core->windows_scored() : 0; }

// ------------------------------------------------------------------ batch
int rp_batch_create_multi(const rp_config* cfg, int64_t n_streams, const int* device_ids, int n_devices, rp_batch** out) {
    if (!cfg || !out) return RP_ERR_INVALID;
    *out = nullptr;
    return guarded((rp_batch*)nullptr, [&] {
        if (!device_ids || n_devices < 1) throw Error(RP_ERR_INVALID, "at least one device id is needed");
        if (n_streams < n_devices) throw Error(RP_ERR_INVALID, "fewer streams than devices");
        auto b = std::make_unique<rp_batch>();
        b->n_streams = n_streams;
        b->cfg = *cfg;
        if (cfg->sample_rate != (uint32_t)kSampleRate) {
            for (int64_t i = 0; i < n_streams; i++)
                b->resamplers.push_back(std::make_unique<FftResampler>(cfg->sample_rate, (size_t)kSampleRate, (size_t)kFrameSamples));
            if (b->resamplers.front()->output_frames() != (size_t)kFrameSamples)
                throw Error(RP_ERR_UNSUPPORTED, "the batched front-end needs a source rate whose resampler emits 480-sample chunks "
                                                "(48000, 44100, 32000, 24000, 8000 ...)");
        }
        for (int i = 0; i < n_devices; i++) {   // contiguous ranges [i*B/G, (i+1)*B/G)
            BatchShard sh;
            sh.begin = n_streams * i / n_devices;
            sh.count = n_streams * (i + 1) / n_devices - sh.begin;
            sh.device = device_ids[i];
            sh.core = std::make_unique<DetectorCore>(*cfg, sh.count, sh.device);
            sh.core->enable_device_filters(*cfg);  // gain normaliser / band pass run as a GPU pre-stage
            sh.core->engine().set_dtw_variant(g_dtw_variant.load());
            b->shards.push_back(std::move(sh));
        }
        *out = b.release();
        return RP_OK;
    });
}
int rp_batch_create(const rp_config* cfg, int64_t n_streams, int device, rp_batch** out) {
    return rp_batch_create_multi(cfg, n_streams, &device, 1, out);
}
void rp_batch_destroy(rp_batch* b) { delete b; }
int rp_batch_n_devices(const rp_batch* b) { return b ? (int)b->shards.size() : 0; }
size_t rp_batch_samples_per_frame(const rp_batch* b) {   // get_samples_per_frame (detector.rs:204): one stream's 30 ms chunk
    if (!b) return 0;
    const size_t mono = b->resamplers.empty() ? (size_t)kFrameSamples : b->resamplers.front()->input_frames();
    return mono * b->cfg.channels;
}

int rp_batch_add_wakeword_from_buffer(rp_batch* b, const char* key, const uint8_t* buf, size_t len) {
    if (!b || !key || !buf) return RP_ERR_INVALID;
    return guarded(b, [&] {
        parse_rpw(buf, len);   // a malformed file fails before any shard is touched
        for (auto& sh : b->shards) sh.core->add_wakeword(key, buf, len);
        return RP_OK;
    });
}
int rp_batch_add_wakeword_from_file(rp_batch* b, const char* key, const char* path) {
    if (!b || !key || !path) return RP_ERR_INVALID;
    return guarded(b, [&] {
        std::vector<uint8_t> f = read_file(path);
        parse_rpw(f.data(), f.size());
        for (auto& sh : b->shards) sh.core->add_wakeword(key, f.data(), f.size());
        return RP_OK;
    });
}
int rp_batch_remove_wakeword(rp_batch* b, const char* key) {
    if (!b || !key) return RP_ERR_INVALID;
    return guarded(b, [&] {
        int removed = 0;
        for (auto& sh : b->shards) removed = sh.core->remove_wakeword(key) ? 1 : 0;
        return removed;
    });
}
int rp_batch_remove_wakewords(rp_batch* b) {
    if (!b) return RP_ERR_INVALID;
    return guarded(b, [&] {
        int removed = 0;
        for (auto& sh : b->shards) removed = sh.core->remove_wakewords() ? 1 : 0;
        return removed;
    });
}
int rp_batch_set_cuda_stream(rp_batch* b, void* s) {
    if (!b) return RP_ERR_INVALID;
    return guarded(b, [&] {
        if (b->shards.size() > 1 && s) throw Error(RP_ERR_INVALID, "a multi-device batch runs on its own per-device streams");
        b->first().engine().set_cuda_stream(static_cast<cudaStream_t>(s));
        return RP_OK;
    });
}
int rp_batch_process(rp_batch* b, const float* audio, int64_t S, int on_device, const rp_batch_detection** dets, int64_t* n_dets) {
    if (!b || !audio) return RP_ERR_INVALID;
    return guarded(b, [&] {
        AudioIn in;
        in.data = audio;
        in.samples = S;
        in.on_device = on_device != 0;
        return batch_process(b, in, dets, n_dets);
    });
}
int rp_batch_process_samples(rp_batch* b, const void* audio, int sample_format, int64_t samples_per_stream, int on_device,
                             const rp_batch_detection** dets, int64_t* n_dets) {
    if (!b || !audio) return RP_ERR_INVALID;
    return guarded(b, [&] {
        if (sample_format < RP_FMT_I8 || sample_format > RP_FMT_F32) throw Error(RP_ERR_INVALID, "unknown sample format");
        const int64_t ch = (int64_t)b->cfg.channels;
        const int64_t frame = (int64_t)rp_batch_samples_per_frame(b);
        if (samples_per_stream <= 0 || samples_per_stream % frame != 0)
            throw Error(RP_ERR_INVALID, "samples_per_stream must be a positive multiple of rp_batch_samples_per_frame()");
        AudioIn in;
        in.data = audio;
        in.fmt = sample_format;
        in.channels = (int)ch;
        in.big_endian = false;   // typed samples are in native (little-endian) byte order
        in.samples = samples_per_stream / ch;
        in.on_device = on_device != 0;
        return batch_process(b, in, dets, n_dets);
    });
}
int rp_batch_process_bytes(rp_batch* b, const uint8_t* bytes, int64_t bytes_per_stream, int on_device,
                           const rp_batch_detection** dets, int64_t* n_dets) {
    if (!b || !bytes) return RP_ERR_INVALID;
    return guarded(b, [&] {
        AudioIn in;
        in.data = bytes;
        in.fmt = (int)b->cfg.sample_format;
        in.channels = (int)b->cfg.channels;
        in.big_endian = b->cfg.endianness == RP_ENDIAN_BIG;
        const int64_t frame_bytes = (int64_t)rp_batch_samples_per_frame(b) * (int64_t)in.bytes_per_sample();
        if (bytes_per_stream <= 0 || bytes_per_stream % frame_bytes != 0)
            throw Error(RP_ERR_INVALID, "bytes_per_stream must be a positive multiple of rp_get_bytes_per_frame()");
        in.samples = bytes_per_stream / ((int64_t)in.channels * (int64_t)in.bytes_per_sample());
        in.on_device = on_device != 0;
        return batch_process(b, in, dets, n_dets);
    });
}
int rp_batch_update_config(rp_batch* b, const rp_config* cfg) {
    if (!b || !cfg) return RP_ERR_INVALID;
    return guarded(b, [&] {
        validate_detector_config(*cfg);
        for (auto& sh : b->shards) {
            sh.core->update_detector_config(*cfg);   // update_config = detector config, then filters config (detector.rs:259-262)
            sh.core->update_filters_config(*cfg);
        }
        return RP_OK;
    });
}
void rp_batch_reset(rp_batch* b) {
    if (b)
        for (auto& sh : b->shards) sh.core->reset();
}
uint64_t rp_batch_windows_scored(const rp_batch* b) {
    uint64_t t = 0;
    if (b)
        for (auto& sh : b->shards) t += sh.core->windows_scored();
    return t;
}
int64_t rp_batch_n_streams(const rp_batch* b) { return b ? b->n_streams : 0; }
int rp_batch_max_mfcc_frames(const rp_batch* b) { return b ? b->first().wakewords().max_frames : 0; }
int rp_batch_last_timings(const rp_batch* b, float* ms, int cap) {
    if (!b || !ms) return 0;
    int n = 0;
    for (; n < 5 && n < cap; n++) ms[n] = b->timings[(size_t)n];
    return n;
}
int rp_batch_last_launches(const rp_batch* b) {
    int n = 0;
    if (b)
        for (auto& sh : b->shards) n += sh.core->engine().launches;
    return n;
}
int rp_batch_last_gate_stats(const rp_batch* b, int64_t* tiles, int64_t* passed) {
    if (!b) return RP_ERR_INVALID;
    return guarded(const_cast<rp_batch*>(b), [&] {
        int64_t t = 0, p = 0;
        for (auto& sh : b->shards) {
            int64_t a = 0, c = 0;
            sh.core->engine().last_gate_stats(&a, &c);
            t += a;
            p += c;
        }
        if (tiles) *tiles = t;
        if (passed) *passed = p;
        return RP_OK;
    });
}
int64_t rp_batch_copy_last_scores(const rp_batch* b, float* out, int64_t cap, int32_t* n_new, int32_t* n_slots) {
    if (!b || !out) return RP_ERR_INVALID;
    return guarded(const_cast<rp_batch*>(b), [&]() -> int {
        const Engine& e0 = b->first().engine();
        const int64_t per_stream = (int64_t)e0.last_n_new() * e0.n_slots();
        const int64_t n = b->n_streams * per_stream;
        if (n_new) *n_new = e0.last_n_new();
        if (n_slots) *n_slots = e0.n_slots();
        if (n > cap) throw Error(RP_ERR_INVALID, "output buffer too small");
        for (auto& sh : b->shards) sh.core->engine().copy_last_scores(out + sh.begin * per_stream);
        return (int)std::min<int64_t>(n, 0x7fffffff);
    });
}
int rp_set_avg_gate(int mode) {
    g_avg_gate.store(mode < 0 ? -1 : (mode ? 1 : 0));
    return RP_OK;
}

// ------------------------------------------------------------------ wakeword builder
int64_t rp_wakeword_build(const char* name, int has_threshold, float threshold, int has_avg_threshold, float avg_threshold,
                          int n_samples, const char* const* sample_names, const uint8_t* const* wavs, const size_t* wav_lens,
                          int mfcc_size, int rms_median, int device, uint8_t* out, size_t out_cap) {
    int64_t written = 0;
    const int rc = guarded((rp_handle*)nullptr, [&] {
        if (!name || n_samples < 0 || (n_samples > 0 && (!sample_names || !wavs || !wav_lens)))
            throw Error(RP_ERR_INVALID, "bad argument");
        std::vector<std::pair<std::string, std::pair<const uint8_t*, size_t>>> samples;
        for (int i = 0; i < n_samples; i++) {
            if (!sample_names[i] || !wavs[i]) throw Error(RP_ERR_INVALID, "null sample");
            samples.push_back({sample_names[i], {wavs[i], wav_lens[i]}});
        }
        // MFCC extraction of one sample file = one stream through the K1 kernel
        MfccFn mfcc = [device](const std::vector<float>& mono, int size) {
            FrameMatrix f;
            f.cols = size;
            const int64_t hops = (int64_t)mono.size() / kHopSamples;
            if (hops <= 3) return f;
            cuda_check(cudaSetDevice(device), "cudaSetDevice");
            const MfccTablesDev& t = tables_for(device, size, nullptr);
            DeviceBuffer audio, frames;
            audio.reserve(mono.size() * sizeof(float), "builder audio");
            f.rows = (int)(hops - 3);
            f.v.resize((size_t)f.rows * size);
            frames.reserve(f.v.size() * sizeof(float), "builder frames");
            cuda_check(cudaMemcpy(audio.as<float>(), mono.data(), mono.size() * sizeof(float), cudaMemcpyHostToDevice), "H2D sample");
            cuda_check(launch_mfcc_frames(audio.as<float>(), (int64_t)mono.size(), nullptr, 1, f.rows, kHopSamples, t,
                                          frames.as<float>(), f.rows, 0, nullptr, nullptr), "mfcc kernel");
            cuda_check(cudaMemcpy(f.v.data(), frames.as<float>(), f.v.size() * sizeof(float), cudaMemcpyDeviceToHost), "D2H frames");
            return f;
        };
        const WakewordRefData w = build_wakeword_ref(name, has_threshold ? std::optional<float>(threshold) : std::nullopt,
                                                     has_avg_threshold ? std::optional<float>(avg_threshold) : std::nullopt,
                                                     samples, mfcc_size, rms_median != 0, mfcc);
        const std::vector<uint8_t> bytes = encode_wakeword_ref(w);
        written = (int64_t)bytes.size();
        if (out) {
            if (bytes.size() > out_cap) throw Error(RP_ERR_INVALID, "output buffer too small");
            std::memcpy(out, bytes.data(), bytes.size());
        }
        return RP_OK;
    });
    return rc < 0 ? rc : written;
}

int64_t rp_wakeword_from_features(const char* name, int has_threshold, float threshold, int has_avg_threshold,
                                  float avg_threshold, int mfcc_size, int n_templates, const char* const* names,
                                  const int32_t* frames, const float* const* data, float rms_level,
                                  uint8_t* out, size_t out_cap) {
    int64_t written = 0;
    const int rc = guarded((rp_handle*)nullptr, [&] {
        if (!name || mfcc_size < 1 || n_templates < 0 || (n_templates > 0 && (!names || !frames || !data)))
            throw Error(RP_ERR_INVALID, "bad argument");
        if (n_templates == 0) throw Error(RP_ERR_INVALID, "Can not create an empty wakeword");  // wakeword_ref.rs:52-54
        WakewordRefData w;
        w.name = name;
        if (has_threshold) w.threshold = threshold;
        if (has_avg_threshold) w.avg_threshold = avg_threshold;
        w.rms_level = rms_level;
        w.mfcc_size = mfcc_size;
        for (int t = 0; t < n_templates; t++) {
            if (!names[t] || !data[t] || frames[t] < 1) throw Error(RP_ERR_INVALID, "bad template");
            FrameMatrix f;
            f.rows = frames[t];
            f.cols = mfcc_size;
            f.v.assign(data[t], data[t] + (size_t)f.rows * f.cols);
            w.samples_features.emplace_back(names[t], std::move(f));
        }
        w.avg_features = average_templates(w.samples_features);
        const std::vector<uint8_t> bytes = encode_wakeword_ref(w);
        written = (int64_t)bytes.size();
        if (out) {
            if (bytes.size() > out_cap) throw Error(RP_ERR_INVALID, "output buffer too small");
            std::memcpy(out, bytes.data(), bytes.size());
        }
        return RP_OK;
    });
    return rc < 0 ? rc : written;
}

// ------------------------------------------------------------------ raw kernels
int rp_mfcc_frames(const float* audio_dev, int64_t n_streams, int64_t S, int mfcc_size, float* out_dev, void* cuda_stream) {
    return guarded((rp_handle*)nullptr, [&] {
        if (n_streams < 1 || mfcc_size < 1 || mfcc_size > kMaxMfccSize) throw Error(RP_ERR_INVALID, "bad argument");
        const int64_t hops = S / kHopSamples;
        if (hops <= 3) return RP_OK;  // a fresh extractor emits nothing for the first three hops
        if (!audio_dev || !out_dev) throw Error(RP_ERR_INVALID, "null device pointer");
        int dev = 0;
        cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
        cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
        const MfccTablesDev& t = tables_for(dev, mfcc_size, s);
        // frame i <-> hop i+3, covering samples [160*(i+1), 160*(i+1)+480)  (extractor.rs:69-79)
        cuda_check(launch_mfcc_frames(audio_dev, S, nullptr, n_streams, (int)(hops - 3), kHopSamples, t, out_dev, hops - 3, 0,
                                      nullptr, s), "mfcc kernel");
        return RP_OK;
    });
}

int rp_dtw_scores(const float* tmpl_dev, const int64_t* tmpl_off_dev, const int32_t* tmpl_len_dev, int tmpl_len_uniform,
                  const float* win_dev, const int64_t* win_off_dev, const int32_t* win_len_dev, int win_len_uniform,
                  int64_t n_pairs, int d, int band, float score_ref, int cmn, float* out_dev, void* cuda_stream) {
    return guarded((rp_handle*)nullptr, [&] {
        if (!tmpl_dev || !win_dev || !out_dev || d < 1 || tmpl_len_uniform < 1 || win_len_uniform < 1 || band < 0)
            throw Error(RP_ERR_INVALID, "bad argument");
        DtwPairsArgs a;
        a.tmpl = tmpl_dev;
        a.tmpl_off = tmpl_off_dev;
        a.tmpl_len = tmpl_len_dev;
        a.tmpl_len_max = tmpl_len_uniform;
        a.win = win_dev;
        a.win_off = win_off_dev;
        a.win_len = win_len_dev;
        a.win_len_max = win_len_uniform;
        a.n_pairs = n_pairs;
        a.d = d;
        a.band = band;
        a.score_ref = score_ref;
        a.cmn = cmn;
        a.out = out_dev;
        if (g_dtw_variant.load() != 1 && dtw_pairs_stream_supported(a))
            cuda_check(launch_dtw_pairs_stream(a, static_cast<cudaStream_t>(cuda_stream)), "dtw stream kernel");
        else
            cuda_check(launch_dtw_pairs_generic(a, static_cast<cudaStream_t>(cuda_stream)), "dtw kernel");
        return RP_OK;
    });
}

int rp_set_dtw_variant(int v) {
    // 0 automatic, 1 generic kernels, 2 tuned kernels, 7 tuned with the pipeline kernel reading its templates from shared
    // instead of constant memory; 3..6 (retired variants) behave like 2. Alternatives kept for A/B measurements and parity tests.
    g_dtw_variant.store(v == 9 ? 9 : (v >= 3 ? 2 : v));   // 9: tuned, but the pipeline kernel also for short calls (no cadence kernel)
    set_dtw_window_kernel(v == 7 ? 3 : 0);
    return RP_OK;
}

int rp_set_mfcc_variant(int v) {
    set_mfcc_variant(v);
    return RP_OK;
}

int64_t rp_resample_to_16k(uint32_t sample_rate_in, const float* in, size_t n_in, float* out, size_t out_cap, size_t* in_chunk) {
    int64_t written = 0;
    const int rc = guarded((rp_handle*)nullptr, [&] {
        if (sample_rate_in < 1000 || sample_rate_in > 768000 || (!in && n_in)) throw Error(RP_ERR_INVALID, "bad argument");
        FftResampler r(sample_rate_in, (size_t)kSampleRate, (size_t)kFrameSamples);
        if (in_chunk) *in_chunk = r.input_frames();
        const size_t calls = n_in / r.input_frames();
        written = (int64_t)(calls * r.output_frames());
        if (out) {
            if ((size_t)written > out_cap) throw Error(RP_ERR_INVALID, "output buffer too small");
            for (size_t c = 0; c < calls; c++) r.process(in + c * r.input_frames(), out + c * r.output_frames());
        }
        return RP_OK;
    });
    return rc < 0 ? rc : written;
}

// ------------------------------------------------------------------ host-logic hooks
int rp_debug_stream4_schedule(int m, int n, int band, uint16_t* out, size_t out_cap) {
    Stream4Sched s;
    int n_super = 0;
    if (!build_stream4_schedule(m, n, band, &s, &n_super)) return 0;
    for (int c = 0; c < n_super; c++)
        for (int k = 0; k < 4; k++)
            if (out && (size_t)(c * 4 + k) < out_cap) out[c * 4 + k] = s.unit[c][k];
    return n_super;
}

int rp_debug_stream4_ctl(int m, int n, int band, uint32_t* out, size_t out_cap) {
    Stream4Sched s;
    int n_super = 0;
    if (!build_stream4_schedule(m, n, band, &s, &n_super)) return 0;
    const int n_blocks = (n + 7) / 8, steps = m / 2 + 2 * (n_blocks - 1);
    for (int w = 0; w < 4; w++)
        for (int st = 0; st <= steps; st++)
            if (out && (size_t)(w * (steps + 1) + st) < out_cap) out[w * (steps + 1) + st] = s.ctl[w][st];
    return steps;
}

int rp_wakeword_inspect(const uint8_t* buf, size_t len, rp_wakeword_info* info) {
    if (!buf || !info) return RP_ERR_INVALID;
    return guarded((rp_handle*)nullptr, [&] {
        WakewordRefData w = parse_rpw(buf, len);
        std::memset(info, 0, sizeof(*info));
        std::snprintf(info->name, RP_NAME_MAX, "%s", w.name.c_str());
        info->mfcc_size = w.mfcc_size;
        info->n_templates = (int)w.samples_features.size();
        info->avg_frames = w.avg_features ? w.avg_features->rows : 0;
        info->max_frames = w.max_frames();
        info->has_threshold = w.threshold.has_value();
        info->has_avg_threshold = w.avg_threshold.has_value();
        info->threshold = w.threshold.value_or(0.f);
        info->avg_threshold = w.avg_threshold.value_or(0.f);
        info->rms_level = w.rms_level;
        info->is_v2 = w.is_v2;
        return RP_OK;
    });
}

int rp_wakeword_template(const uint8_t* buf, size_t len, int t, char* name_out, float* out, size_t out_cap) {
    if (!buf) return RP_ERR_INVALID;
    return guarded((rp_handle*)nullptr, [&] {
        WakewordRefData w = parse_rpw(buf, len);
        const FrameMatrix* m = nullptr;
        const char* nm = "";
        if (t < 0) {
            if (!w.avg_features) return 0;
            m = &*w.avg_features;
        } else {
            if (t >= (int)w.samples_features.size()) throw Error(RP_ERR_INVALID, "template index out of range");
            m = &w.samples_features[(size_t)t].second;
            nm = w.samples_features[(size_t)t].first.c_str();
        }
        if (name_out) std::snprintf(name_out, RP_NAME_MAX, "%s", nm);
        if (out) {
            if (out_cap < m->v.size()) throw Error(RP_ERR_INVALID, "output buffer too small");
            std::memcpy(out, m->v.data(), m->v.size() * sizeof(float));
        }
        return m->rows;
    });
}

int rp_host_replay(const rp_config* cfg, const uint8_t* const* rpws, const size_t* rpw_lens, int n_rpw, const float* scores,
                   int64_t n_frames, int n_slots, const float* vad_values, rp_batch_detection* out, int64_t out_cap,
                   int64_t* n_out, uint64_t* windows_scored) {
    if (!cfg || !rpws || !rpw_lens || !scores || !n_out) return RP_ERR_INVALID;
    return guarded((rp_handle*)nullptr, [&] {
        static thread_local WakewordSet ws;                       // keeps names alive for the caller
        static thread_local std::vector<std::vector<float>> stores;
        static thread_local std::vector<std::vector<const char*>> names;
        ws = WakewordSet();
        for (int i = 0; i < n_rpw; i++) ws.add("w" + std::to_string(i), parse_rpw(rpws[i], rpw_lens[i]));
        ws.rebuild(*cfg);
        if ((int)ws.slots.size() != n_slots) throw Error(RP_ERR_INVALID, "n_slots does not match the wakeword files");
        names.clear();
        for (auto& r : ws.refs) {
            std::vector<const char*> n;
            for (auto& t : r.samples_features) n.push_back(t.first.c_str());
            names.push_back(std::move(n));
        }
        DetectorParams p;
        p.max_frames = ws.max_frames;
        p.min_scores = cfg->min_scores;
        p.eager = cfg->eager != 0;
        p.vad_mode = cfg->vad_mode;
        StreamState st;
        st.configure(p);
        stores.clear();
        int64_t n = 0;
        const int64_t total_hops = n_frames + kHopsPerChunk;
        std::vector<float> tmp((size_t)ws.max_templates);
        for (int64_t c = 0; c * kHopsPerChunk < total_hops; c++) {
            for (int k = 0; k < kHopsPerChunk; k++) {
                const int64_t h = c * kHopsPerChunk + k;
                if (h >= total_hops) break;
                Hit hit;
                const Hit* hp = nullptr;
                float vv = 0.f;
                if (h >= kHopsPerChunk) {
                    const float* row = scores + (h - kHopsPerChunk) * (int64_t)n_slots;
                    const Judgement jd = judge_window(row, ws.metas.data(), (int)ws.metas.size(), (int)cfg->score_mode);
                    if (jd.wakeword >= 0) {
                        const WakewordMeta& m = ws.metas[(size_t)jd.wakeword];
                        hit.stream = 0;
                        hit.frame = (int32_t)h;
                        hit.wakeword = jd.wakeword;
                        hit.avg_score = jd.avg_score;
                        hit.score = jd.score;
                        hit.scores = row + m.slot_begin + (m.has_avg ? 1 : 0);
                        hit.n_scores = m.n_templates;
                        hp = &hit;
                    }
                    if (vad_values) vv = vad_values[h - kHopsPerChunk];
                }
                PartialDetection det;
                if (st.on_hop(p, hp, vv, 1.f, &det)) {
                    if (out && n < out_cap) {
                        stores.emplace_back(det.scores);
                        rp_batch_detection& o = out[n];
                        std::memset(&o, 0, sizeof(o));
                        o.stream = 0;
                        o.chunk = c;
                        const WakewordRefData& r = ws.refs[(size_t)det.wakeword];
                        std::snprintf(o.det.name, RP_NAME_MAX, "%s", r.name.c_str());
                        o.det.avg_score = det.avg_score;
                        o.det.score = det.score;
                        o.det.counter = det.counter;
                        o.det.gain = det.gain;
                        o.det.n_scores = (uint32_t)det.scores.size();
                        o.det.score_names = names[(size_t)det.wakeword].data();
                    }
                    n++;
                    break;
                }
            }
        }
        // score_values pointers are taken after the loop: `stores` no longer reallocates
        for (int64_t i = 0; out && i < std::min<int64_t>(n, out_cap); i++) out[i].det.score_values = stores[(size_t)i].data();
        *n_out = n;
        if (windows_scored) *windows_scored = st.windows_scored();
        return RP_OK;
    });
}

}  // extern "C"
