// Wakeword-reference builder: wav samples -> MFCC (K1, on the GPU) -> whole-file CMN ->
// DTW-path averaging -> WakewordRef -> .rpw CBOR.
//
// Replaces WakewordRef::new_from_sample_files / new_from_sample_buffers
// (reference src/wakewords/comp/wakeword_ref_build.rs:9-110), MfccWavFileExtractor::compute_mfccs
// (src/mfcc/wav_file_extractor.rs:18-91), MfccAverager::average (src/mfcc/averager.rs:5-37) with the unbanded
// Dtw::compute_optimal_path + retrieve_optimal_path (src/mfcc/dtw.rs:11-55,106-138) and WakewordSave
// (src/wakewords/wakeword_file.rs:10-26, ciborium). A build-time tool: only the MFCC extraction is heavy and it
// reuses the hot-path kernel; the O(m*n) alignment of a handful of ~100-frame templates runs on the host.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "wakeword_builder.h"

namespace rp {
namespace {

// ---- WAV (what hound + AudioEncoder hand to the extractor): PCM int 8/16/32, IEEE float 32
struct Wav {
    uint32_t rate = 0;
    uint16_t channels = 0, bits = 0;
    bool ieee_float = false;
    const uint8_t* data = nullptr;
    size_t data_len = 0;
};

uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | p[1] << 8); }

Wav read_wav_header(const uint8_t* b, size_t n) {
    if (n < 12 || std::memcmp(b, "RIFF", 4) != 0 || std::memcmp(b + 8, "WAVE", 4) != 0) throw Error(RP_ERR_FORMAT, "not a RIFF/WAVE file");
    Wav w;
    uint16_t tag = 0;
    bool fmt = false;
    for (size_t pos = 12; pos + 8 <= n;) {
        const uint32_t size = le32(b + pos + 4);
        const uint8_t* body = b + pos + 8;
        const size_t room = n - (pos + 8);
        if (std::memcmp(b + pos, "fmt ", 4) == 0 && room >= 16) {
            tag = le16(body);
            w.channels = le16(body + 2);
            w.rate = le32(body + 4);
            w.bits = le16(body + 14);
            if (tag == 0xFFFE && room >= 26) tag = le16(body + 24);
            fmt = true;
        } else if (std::memcmp(b + pos, "data", 4) == 0) {
            if (!fmt) throw Error(RP_ERR_FORMAT, "wav: data chunk before fmt chunk");
            w.ieee_float = tag == 3;
            const bool ok = (tag == 1 && (w.bits == 8 || w.bits == 16 || w.bits == 32)) || (tag == 3 && w.bits == 32);
            if (!ok || w.channels == 0) throw Error(RP_ERR_UNSUPPORTED, "Unsupported wav format");  // wav_file_extractor.rs:93-112
            w.data = body;
            w.data_len = std::min<size_t>(size, room);
            return w;
        }
        pos += 8 + (size_t)size + (size & 1);
    }
    throw Error(RP_ERR_FORMAT, "wav: no data chunk");
}

float sample_to_f32(const Wav& w, size_t i) {  // Sample::into_f32 (audio_types.rs:98-137)
    const uint8_t* p = w.data + i * (w.bits / 8);
    if (w.ieee_float) {
        float f;
        std::memcpy(&f, p, 4);
        return f;
    }
    switch (w.bits) {
        case 8: return (float)(int8_t)(p[0] - 128) / 127.f;  // 8-bit WAV is unsigned; hound yields i8
        case 16: return (float)(int16_t)le16(p) / 32767.f;
        default: return (float)(int32_t)le32(p) / (float)2147483647;
    }
}

float rms_level(const float* x, size_t n) {  // GainNormalizerFilter::get_rms_level
    float sum_squared = 0.f;
    for (size_t i = 0; i < n; i++) sum_squared += x[i] * x[i];
    return std::sqrt(sum_squared / (float)n);
}

bool total_less(float a, float b) {  // f32::total_cmp
    int32_t x, y;
    std::memcpy(&x, &a, 4);
    std::memcpy(&y, &b, 4);
    x ^= (int32_t)((uint32_t)(x >> 31) >> 1);
    y ^= (int32_t)((uint32_t)(y >> 31) >> 1);
    return x < y;
}

float cosine_distance(const float* a, const float* b, int d) {  // comparator.rs:15-17,28-48
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int k = 0; k < d; k++) {
        ab += a[k] * b[k];
        aa += a[k] * a[k];
        bb += b[k] * b[k];
    }
    const float mag = std::sqrt(aa * bb);
    return 1.f - (mag == 0.f ? 0.f : ab / mag);
}

// Dtw::compute_optimal_path + retrieve_optimal_path (unbanded). Returns the path as (row of a, row of b),
// including the reference's quirks: min(m-1, n-1) leading [0,0] entries pre-filled before the walk (they end
// up at the END after the reversal) and the end cell (m-1, n-1) itself never pushed (dtw.rs:106-138).
std::vector<std::pair<int, int>> optimal_path(const FrameMatrix& a, const FrameMatrix& b) {
    const int m = a.rows, n = b.rows, d = a.cols;
    const float inf = std::numeric_limits<float>::infinity();
    std::vector<float> D((size_t)m * n, inf);
    auto at = [&](int r, int c) -> float& { return D[(size_t)r * n + c]; };
    auto dist = [&](int r, int c) { return cosine_distance(a.v.data() + (size_t)r * d, b.v.data() + (size_t)c * d, d); };
    auto min3 = [&](float ins, float del, float mat) { return std::fmin(std::fmin(std::fmin(inf, ins), del), mat); };
    at(0, 0) = dist(0, 0);
    for (int r = 1; r < m; r++) at(r, 0) = dist(r, 0) + at(r - 1, 0);
    for (int c = 1; c < n; c++) at(0, c) = dist(0, c) + at(0, c - 1);
    for (int r = 1; r < m; r++)
        for (int c = 1; c < n; c++) at(r, c) = dist(r, c) + min3(at(r - 1, c), at(r, c - 1), at(r - 1, c - 1));
    int r = m - 1, c = n - 1;
    std::vector<std::pair<int, int>> path((size_t)std::min(r, c), {0, 0});
    while (r > 0 || c > 0) {
        if (r > 0 && c > 0) {
            const float ins = at(r - 1, c), del = at(r, c - 1), mat = at(r - 1, c - 1);
            const float mn = min3(ins, del, mat);
            if (mn == mat) { r--; c--; }
            else if (mn == ins) r--;
            else if (mn == del) c--;
        } else if (r > 0) r--;
        else c--;
        path.push_back({r, c});
    }
    std::reverse(path.begin(), path.end());
    return path;
}

// ---- CBOR writer, ciborium style: definite lengths, floats in the smallest lossless width
struct Cbor {
    std::vector<uint8_t> out;
    void head(int major, uint64_t v) {
        if (v < 24) out.push_back((uint8_t)(major << 5 | v));
        else if (v <= 0xff) { out.push_back((uint8_t)(major << 5 | 24)); out.push_back((uint8_t)v); }
        else if (v <= 0xffff) { out.push_back((uint8_t)(major << 5 | 25)); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); }
        else if (v <= 0xffffffffull) { out.push_back((uint8_t)(major << 5 | 26)); for (int s = 24; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s)); }
        else { out.push_back((uint8_t)(major << 5 | 27)); for (int s = 56; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s)); }
    }
    void text(const std::string& s) {
        head(3, s.size());
        out.insert(out.end(), s.begin(), s.end());
    }
    void null() { out.push_back(0xf6); }
    static bool as_half(float f, uint16_t& h) {  // exact f32 -> f16, if one exists
        uint32_t u;
        std::memcpy(&u, &f, 4);
        const uint32_t sign = (u >> 16) & 0x8000, exp = (u >> 23) & 0xff, man = u & 0x7fffff;
        if (exp == 0xff) {  // inf / nan (ciborium writes NaN as f16 7e00)
            h = (uint16_t)(sign | 0x7c00 | (man ? 0x200 : 0));
            return man == 0 || man == 0x400000;
        }
        if (exp == 0 && man == 0) { h = (uint16_t)sign; return true; }
        const int e = (int)exp - 127;
        if (e > 15) return false;
        if (e >= -14) {  // normal half
            if (man & 0x1fff) return false;
            h = (uint16_t)(sign | (uint32_t)(e + 15) << 10 | man >> 13);
            return true;
        }
        if (e < -24) return false;  // subnormal half: value = m * 2^-24
        const uint32_t full = man | 0x800000;
        const int shift = -e - 14 + 13;
        if (shift > 24 || (full & ((1u << shift) - 1))) return false;
        h = (uint16_t)(sign | full >> shift);
        return true;
    }
    void f32(float f) {
        uint16_t h;
        if (as_half(f, h)) {
            out.push_back(0xf9);
            out.push_back((uint8_t)(h >> 8));
            out.push_back((uint8_t)h);
            return;
        }
        uint32_t u;
        std::memcpy(&u, &f, 4);
        out.push_back(0xfa);
        for (int s = 24; s >= 0; s -= 8) out.push_back((uint8_t)(u >> s));
    }
    void matrix(const FrameMatrix& m) {
        head(4, (uint64_t)m.rows);
        for (int r = 0; r < m.rows; r++) {
            head(4, (uint64_t)m.cols);
            for (int c = 0; c < m.cols; c++) f32(m.v[(size_t)r * m.cols + c]);
        }
    }
};

}  // namespace

// MfccAverager::average after compute_avg_samples_features' ordering (longest first, then name)
std::optional<FrameMatrix> average_templates(const std::vector<std::pair<std::string, FrameMatrix>>& templates) {
    if (templates.size() <= 1) return std::nullopt;
    std::vector<const std::pair<std::string, FrameMatrix>*> order;
    for (auto& t : templates) order.push_back(&t);
    std::stable_sort(order.begin(), order.end(), [](auto* x, auto* y) {
        if (x->second.rows != y->second.rows) return x->second.rows > y->second.rows;
        return x->first < y->first;
    });
    FrameMatrix origin = order[0]->second;
    const int d = origin.cols;
    for (size_t k = 1; k < order.size(); k++) {
        const FrameMatrix& frames = order[k]->second;
        // per (origin row, coefficient): the origin value followed by every aligned value, summed in that order
        std::vector<std::vector<float>> acc((size_t)origin.rows * d);
        for (int x = 0; x < origin.rows; x++)
            for (int j = 0; j < d; j++) acc[(size_t)x * d + j].push_back(origin.v[(size_t)x * d + j]);
        for (auto [x, y] : optimal_path(origin, frames))
            for (int j = 0; j < d; j++) acc[(size_t)x * d + j].push_back(frames.v[(size_t)y * d + j]);
        for (size_t i = 0; i < acc.size(); i++) {
            float sum = 0.f;
            for (float v : acc[i]) sum += v;
            origin.v[i] = sum / (float)acc[i].size();
        }
    }
    return origin;
}

std::vector<uint8_t> encode_wakeword_ref(const WakewordRefData& w) {
    // serde field order of WakewordRef (wakeword_ref.rs:12-20)
    Cbor c;
    c.head(5, 7);
    c.text("name");
    c.text(w.name);
    c.text("avg_features");
    if (w.avg_features) c.matrix(*w.avg_features); else c.null();
    c.text("samples_features");
    c.head(5, w.samples_features.size());
    for (auto& t : w.samples_features) {
        c.text(t.first);
        c.matrix(t.second);
    }
    c.text("threshold");
    if (w.threshold) c.f32(*w.threshold); else c.null();
    c.text("avg_threshold");
    if (w.avg_threshold) c.f32(*w.avg_threshold); else c.null();
    c.text("rms_level");
    c.f32(w.rms_level);
    c.text("mfcc_size");
    c.head(0, (uint64_t)w.mfcc_size);
    return std::move(c.out);
}

WakewordRefData build_wakeword_ref(const std::string& name, std::optional<float> threshold, std::optional<float> avg_threshold,
                                   const std::vector<std::pair<std::string, std::pair<const uint8_t*, size_t>>>& samples,
                                   int mfcc_size, bool rms_median, const MfccFn& mfcc) {
    if (samples.empty()) throw Error(RP_ERR_INVALID, "Can not create an empty wakeword");  // wakeword_ref.rs:52-54
    if (mfcc_size < 1 || mfcc_size > kMaxMfccSize) throw Error(RP_ERR_UNSUPPORTED, "mfcc_size outside 1..31");
    WakewordRefData out;
    out.name = name;
    out.threshold = threshold;
    out.avg_threshold = avg_threshold;
    std::vector<float> sample_rms;
    float rms_max = 0.f;
    for (auto& s : samples) {
        const Wav w = read_wav_header(s.second.first, s.second.second);
        if (w.rate != (uint32_t)kSampleRate)
            throw Error(RP_ERR_UNSUPPORTED, "wav sample rate != 16000 needs the reference's rubato resampler, which is outside this path");
        // 30 ms chunks of interleaved samples -> f32, channel 0 (wav_file_extractor.rs:71-91, encoder.rs:35-50)
        const size_t total = w.data_len / (w.bits / 8);
        const size_t in_frame = (size_t)kFrameSamples * w.channels;
        std::vector<float> mono;
        std::vector<float> rms;
        for (size_t off = 0; off + in_frame <= total; off += in_frame) {
            const size_t base = mono.size();
            for (size_t i = 0; i < in_frame; i += w.channels) mono.push_back(sample_to_f32(w, off + i));
            rms.push_back(rms_level(mono.data() + base, kFrameSamples));
        }
        float level = 0.f;
        if (!rms.empty()) {  // median chunk rms (wav_file_extractor.rs:54-58)
            std::sort(rms.begin(), rms.end(), total_less);
            level = rms[rms.size() / 2];
        }
        FrameMatrix f = mfcc(mono, mfcc_size);
        if (f.rows == 0) throw Error(RP_ERR_INVALID, "sample '" + s.first + "' is too short (needs at least 60 ms)");
        // MfccNormalizer::normalize over the whole file (wav_file_extractor.rs:67; normalizer.rs:3-31)
        for (int j = 0; j < f.cols; j++) {
            float sum = 0.f;
            for (int r = 0; r < f.rows; r++) sum += f.v[(size_t)r * f.cols + j];
            for (int r = 0; r < f.rows; r++) f.v[(size_t)r * f.cols + j] -= sum / (float)f.rows;
        }
        out.samples_features.emplace_back(s.first, std::move(f));
        sample_rms.push_back(level);
        if (level > rms_max) rms_max = level;
    }
    if (rms_median) {  // new_from_sample_files (wakeword_ref_build.rs:78-79)
        std::sort(sample_rms.begin(), sample_rms.end(), total_less);
        out.rms_level = sample_rms[sample_rms.size() / 2];
    } else {           // new_from_sample_buffers (wakeword_ref_build.rs:28-30)
        out.rms_level = rms_max;
    }
    out.avg_features = average_templates(out.samples_features);
    out.mfcc_size = out.samples_features.front().second.cols;  // wakeword_ref.rs:55
    return out;
}

}  // namespace rp
