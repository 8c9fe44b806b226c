// .rpw wakeword files: CBOR (RFC 8949) as written by ciborium 0.2.1 for the reference's serde
// structs WakewordRef (src/wakewords/wakeword_ref.rs:12-20) and WakewordV2 (wakeword_v2.rs:8-16),
// loaded in the order the detector tries them (src/detector.rs:152-163): V2, then Ref; a
// WakewordModel (NN) file is outside this path and reported as RP_ERR_UNSUPPORTED.
//
// Only a reader is needed on the scoring path; it is a small pull parser over a byte cursor.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "rp_internal.h"

namespace rp {
namespace {

enum Major { kUInt = 0, kNInt = 1, kBytes = 2, kText = 3, kArray = 4, kMap = 5, kTag = 6, kSimple = 7 };
constexpr uint64_t kIndefinite = ~0ull;

class Cursor {
  public:
    Cursor(const uint8_t* b, size_t n) : p_(b), end_(b + n) {}

    struct Head {
        int major;
        int info;      // low 5 bits
        uint64_t arg;  // length / value / raw float bits; kIndefinite for info == 31
    };

    Head head() {
        uint8_t b = byte();
        Head h{b >> 5, b & 31, 0};
        if (h.info < 24) h.arg = (uint64_t)h.info;
        else if (h.info == 31) h.arg = kIndefinite;
        else if (h.info <= 27) {
            int n = 1 << (h.info - 24);
            for (int i = 0; i < n; i++) h.arg = (h.arg << 8) | byte();
        } else fail("reserved CBOR additional-information value");
        return h;
    }
    bool at_break() const { return p_ < end_ && *p_ == 0xff; }
    void eat_break() { if (byte() != 0xff) fail("expected CBOR break"); }
    bool at_null() const { return p_ < end_ && (*p_ == 0xf6 || *p_ == 0xf7); }
    void eat() { byte(); }

    std::string text(bool chunk = false) {
        Head h = head();
        if (h.major != kText) fail("expected a text string");
        if (h.arg == kIndefinite) {
            if (chunk) fail("nested indefinite-length string chunk");   // RFC 8949 3.2.3: chunks are definite
            std::string s;
            while (!at_break()) s += text(true);
            eat_break();
            return s;
        }
        need(h.arg);
        std::string s(reinterpret_cast<const char*>(p_), (size_t)h.arg);
        p_ += h.arg;
        return s;
    }

    // f16 / f32 / f64 / integer -> f32
    float number() {
        Head h = head();
        switch (h.major) {
            case kUInt: return (float)h.arg;
            case kNInt: return -1.f - (float)h.arg;
            case kSimple:
                if (h.info == 25) return half((uint16_t)h.arg);
                if (h.info == 26) { uint32_t u = (uint32_t)h.arg; float f; std::memcpy(&f, &u, 4); return f; }
                if (h.info == 27) { double d; std::memcpy(&d, &h.arg, 8); return (float)d; }
                [[fallthrough]];
            default: fail("expected a number");
        }
        return 0.f;
    }

    // iterate a definite or indefinite container: calls fn() once per element (array) / pair (map)
    template <typename F>
    void each(int major, F&& fn) {
        Head h = head();
        if (h.major != major) fail(major == kMap ? "expected a map" : "expected an array");
        if (h.arg == kIndefinite) {
            while (!at_break()) fn();
            eat_break();
        } else {
            for (uint64_t i = 0; i < h.arg; i++) fn();
        }
    }

    // Skips one data item. Nesting is bounded (untrusted input must not overflow the host stack).
    void skip(int depth = 0) {
        if (depth > kMaxDepth) fail("CBOR nesting too deep");
        Head h = head();
        switch (h.major) {
            case kUInt: case kNInt: case kSimple: return;
            case kBytes: case kText:
                if (h.arg == kIndefinite) {
                    while (!at_break()) {   // definite chunks of the same major type only
                        Head ch = head();
                        if (ch.major != h.major || ch.arg == kIndefinite) fail("bad chunk in an indefinite-length string");
                        need(ch.arg);
                        p_ += ch.arg;
                    }
                    eat_break();
                } else { need(h.arg); p_ += h.arg; }
                return;
            case kArray:
                if (h.arg == kIndefinite) { while (!at_break()) skip(depth + 1); eat_break(); }
                else for (uint64_t i = 0; i < h.arg; i++) skip(depth + 1);
                return;
            case kMap:
                if (h.arg == kIndefinite) { while (!at_break()) { skip(depth + 1); skip(depth + 1); } eat_break(); }
                else for (uint64_t i = 0; i < h.arg; i++) { skip(depth + 1); skip(depth + 1); }
                return;
            case kTag: skip(depth + 1); return;
        }
    }
    static constexpr int kMaxDepth = 64;

    [[noreturn]] void fail(const char* what) const { throw Error(RP_ERR_FORMAT, std::string("wakeword file: ") + what); }

  private:
    uint8_t byte() {
        if (p_ >= end_) fail("unexpected end of data");
        return *p_++;
    }
    void need(uint64_t n) const {
        if ((uint64_t)(end_ - p_) < n) fail("string runs past the end of the buffer");
    }
    static float half(uint16_t h) {
        int e = (h >> 10) & 31, m = h & 1023;
        float v = e == 0 ? std::ldexp((float)m, -24)
                  : e == 31 ? (m ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity())
                            : std::ldexp((float)(m | 1024), e - 25);
        return (h & 0x8000) ? -v : v;
    }
    const uint8_t* p_;
    const uint8_t* end_;
};

FrameMatrix read_matrix(Cursor& c) {
    FrameMatrix m;
    c.each(kArray, [&] {
        int cols = 0;
        c.each(kArray, [&] {
            m.v.push_back(c.number());
            cols++;
        });
        if (m.rows == 0) m.cols = cols;
        else if (cols != m.cols) c.fail("ragged feature matrix");
        m.rows++;
    });
    return m;
}

}  // namespace

WakewordRefData parse_rpw(const uint8_t* buf, size_t len) {
    Cursor c(buf, len);
    WakewordRefData w;
    bool saw_name = false, saw_samples = false, saw_rms = false, saw_mfcc = false, saw_enabled = false;
    bool saw_labels = false, saw_weights = false;
    c.each(kMap, [&] {
        std::string key = c.text();
        if (key == "name") { w.name = c.text(); saw_name = true; }
        else if (key == "avg_features") {
            if (c.at_null()) c.eat(); else w.avg_features = read_matrix(c);
        } else if (key == "samples_features") {
            c.each(kMap, [&] {
                std::string tname = c.text();
                w.samples_features.emplace_back(std::move(tname), read_matrix(c));
            });
            saw_samples = true;
        } else if (key == "threshold") {
            if (c.at_null()) c.eat(); else w.threshold = c.number();
        } else if (key == "avg_threshold") {
            if (c.at_null()) c.eat(); else w.avg_threshold = c.number();
        } else if (key == "rms_level") { w.rms_level = c.number(); saw_rms = true; }
        else if (key == "mfcc_size") { w.mfcc_size = (int)c.number(); saw_mfcc = true; }
        else if (key == "enabled") { c.skip(); saw_enabled = true; }
        else {
            if (key == "labels") saw_labels = true;
            if (key == "weights") saw_weights = true;
            c.skip();
        }
    });
    if (saw_labels || saw_weights)
        throw Error(RP_ERR_UNSUPPORTED, "wakeword file is a WakewordModel (neural network): outside the WakewordRef/DTW path");
    if (!saw_name || !saw_samples || !saw_rms) throw Error(RP_ERR_FORMAT, "wakeword file: missing field of WakewordRef");
    if (!saw_mfcc && !saw_enabled) throw Error(RP_ERR_FORMAT, "wakeword file: missing field `mfcc_size`");
    if (w.samples_features.empty()) throw Error(RP_ERR_FORMAT, "wakeword file: no templates");
    int d = w.samples_features.front().second.cols;
    for (auto& t : w.samples_features) {
        if (t.second.rows == 0) throw Error(RP_ERR_FORMAT, "wakeword file: empty template");
        if (t.second.cols != d) throw Error(RP_ERR_FORMAT, "wakeword file: templates with different mfcc size");
    }
    if (w.avg_features && (w.avg_features->rows == 0 || w.avg_features->cols != d))
        throw Error(RP_ERR_FORMAT, "wakeword file: avg_features shape does not match the templates");
    w.is_v2 = !saw_mfcc;
    // WakewordV2 -> WakewordRef takes mfcc_size from the first template (wakeword_v2.rs:22); the
    // comparator itself always uses the template width (wakeword_comp.rs:158-160).
    if (!saw_mfcc) w.mfcc_size = d;
    if (w.mfcc_size != d) throw Error(RP_ERR_FORMAT, "wakeword file: mfcc_size does not match the templates");
    if (d < 1 || d > kMaxMfccSize) throw Error(RP_ERR_UNSUPPORTED, "mfcc_size outside 1..31");
    if ((int)w.samples_features.size() > kMaxTemplates) throw Error(RP_ERR_UNSUPPORTED, "more than 64 templates in one wakeword");
    return w;
}

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error(RP_ERR_INVALID, "Unable to open file " + path);
    std::vector<uint8_t> out;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) out.insert(out.end(), tmp, tmp + n);
    std::fclose(f);
    return out;
}

static thread_local std::string g_thread_error;
void set_thread_error(const std::string& msg) { g_thread_error = msg; }
const char* thread_error() { return g_thread_error.c_str(); }

}  // namespace rp
