/*
 * rp_oracle.cpp — CPU ORACLE: a restatement of the reference algorithm. TEST INFRASTRUCTURE ONLY.
 *
 * Reference: GiviMAD/rustpotter v3.0.2 (paths below are relative to the reference checkout).
 * Every function cites the reference file:line it follows. All arithmetic is f32, evaluated in
 * the reference's operation order; build with -ffp-contract=off (see oracle/Makefile) so gcc does
 * not fuse mul+add where the Rust code does not.
 *
 * Third-party arithmetic not present in the reference tree:
 *   - rustfft 6.1.0 (Cargo.lock:545-548), call site src/mfcc/extractor.rs:102-110: a 480-point
 *     unnormalised forward complex DFT. Restated here as a generic mixed-radix decimation-in-time
 *     FFT in f32 with double-precision-derived twiddles. Any correct f32 FFT matches the
 *     reference's goldens to ~2e-7 relative (checked by tests/test_oracle_golden.py).
 *   - rubato 0.14.1 (Cargo.lock:533-536), call sites src/audio/encoder.rs:52-60,72-79: the synchronous FFT resampler
 *     `FftFixedInOut<f32>` (one channel). Restated below from its published algorithm (rubato src/synchro.rs,
 *     src/sinc.rs, src/windows.rs): a Blackman-Harris^2 windowed sinc anti-aliasing filter applied in the frequency
 *     domain to zero-padded chunks (real FFT of 2*fft_size_in points, spectrum truncated to fft_size_out bins, inverse
 *     real FFT of 2*fft_size_out points, overlap-add). Its FFTs (realfft over rustfft) are replaced by the generic FFT
 *     above; pinned by the reference's 48 kHz goldens tests/detector.rs:162-214 (tests/test_oracle_golden.py).
 *   - ciborium 0.2.1: standard CBOR (RFC 8949); reader + writer restated below.
 *
 * Parity status: PINNED against tests/detector.rs:9-159 goldens and the .rpw template matrices
 * (see tests/test_oracle_golden.py).
 */
#include "rp_oracle.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace rpo {

using Vec = std::vector<float>;
using Mat = std::vector<Vec>;
static constexpr float PI_F = 3.14159265358979323846264338327950288f;  // std::f32::consts::PI
static constexpr float INF = std::numeric_limits<float>::infinity();

// reference src/constants.rs:1-11
static constexpr size_t SAMPLE_RATE = 16000;
static constexpr size_t FRAME_LENGTH_MS = 30;
static constexpr size_t FRAME_SHIFT_MS = 10;
static constexpr float PRE_EMPHASIS = 0.97f;

// ------------------------------------------------------------------------------------------
// FFT (stands in for rustfft 6.1.0 `plan_fft_forward(480)`; extractor.rs:102-110)
// ------------------------------------------------------------------------------------------
struct Fft {
    size_t n;
    std::vector<std::complex<float>> tw;  // exp(-2*pi*i*k/n)
    explicit Fft(size_t n_) : n(n_), tw(n_) {
        for (size_t k = 0; k < n; k++) {
            double ang = -2.0 * 3.141592653589793238462643383279502884 * (double)k / (double)n;
            tw[k] = std::complex<float>((float)std::cos(ang), (float)std::sin(ang));
        }
    }
    static size_t smallest_factor(size_t n) {
        for (size_t p = 2; p * p <= n; p++)
            if (n % p == 0) return p;
        return n;
    }
    // out[k], k<len  =  sum_j in[j*stride] * W_len^(jk): decimation in time by the smallest prime
    // factor p of len (480 = 2^5*3*5), p sub-transforms of length len/p, then p-point butterflies.
    void forward(std::vector<std::complex<float>>& buf) const {
        std::vector<std::complex<float>> out(n);
        rec_alloc(buf.data(), 1, n, out.data());
        buf.swap(out);
    }
    void rec_alloc(const std::complex<float>* in, size_t stride, size_t len, std::complex<float>* out) const {
        if (len == 1) {
            out[0] = in[0];
            return;
        }
        size_t p = smallest_factor(len);
        size_t q = len / p;
        std::vector<std::complex<float>> sub(len);
        for (size_t r = 0; r < p; r++) rec_alloc(in + r * stride, stride * p, q, sub.data() + r * q);
        size_t step = n / len;
        for (size_t k1 = 0; k1 < q; k1++) {
            for (size_t k2 = 0; k2 < p; k2++) {
                size_t k = k1 + q * k2;
                std::complex<float> acc = sub[k1];
                for (size_t r = 1; r < p; r++) {
                    const std::complex<float>& w = tw[((r * k) % len) * step];
                    const std::complex<float>& v = sub[r * q + k1];
                    acc = std::complex<float>(acc.real() + (v.real() * w.real() - v.imag() * w.imag()),
                                              acc.imag() + (v.real() * w.imag() + v.imag() * w.real()));
                }
                out[k] = acc;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// MfccExtractor — src/mfcc/extractor.rs:5-199
// ------------------------------------------------------------------------------------------
struct MfccExtractor {
    size_t num_coefficients;  // = mfcc_size + 1 after set_out_size
    float pre_emphasis_coefficient;
    size_t samples_per_frame, samples_per_shift, magnitude_spectrum_size, sample_rate;
    Mat filter_bank;
    Vec hamming_window;
    Vec samples;
    Fft fft;

    // extractor.rs:19-46
    MfccExtractor(size_t sr, size_t spf, size_t sps, uint16_t ncoef, float pre)
        : num_coefficients(ncoef), pre_emphasis_coefficient(pre), samples_per_frame(spf), samples_per_shift(sps),
          magnitude_spectrum_size(spf / 2), sample_rate(sr), fft(spf) {
        filter_bank = new_mel_filter_bank(sr, magnitude_spectrum_size, num_coefficients, 0, sr / 2);
        hamming_window = new_hamming_window(spf);
    }
    // extractor.rs:47-59
    void set_out_size(uint16_t out_size) {
        num_coefficients = (size_t)out_size + 1;  // first coefficient is dropped
        filter_bank = new_mel_filter_bank(sample_rate, magnitude_spectrum_size, num_coefficients, 0, sample_rate / 2);
        reset();
    }
    // extractor.rs:60-65
    Mat compute(const float* audio, size_t n) {
        Mat out;
        for (size_t off = 0; off + samples_per_shift <= n; off += samples_per_shift) {  // chunks_exact
            Vec frame;
            if (process_audio_part(audio + off, samples_per_shift, frame)) out.push_back(std::move(frame));
        }
        return out;
    }
    void reset() { samples.clear(); }  // extractor.rs:66-68
    // extractor.rs:69-79
    bool process_audio_part(const float* buf, size_t n, Vec& out) {
        Vec new_samples = pre_emphasis(buf, n);
        if (samples.size() >= samples_per_frame) {
            samples.erase(samples.begin(), samples.begin() + new_samples.size());
            samples.insert(samples.end(), new_samples.begin(), new_samples.end());
            out = extract_mfccs(samples.data());
            return true;
        }
        samples.insert(samples.end(), new_samples.begin(), new_samples.end());
        return false;
    }
    // extractor.rs:80-86
    Vec extract_mfccs(const float* s) const {
        Vec mag = calculate_magnitude_spectrum(s);
        Vec mfcc = calculate_mel_frequency_cepstral_coefficients(mag);
        mfcc.erase(mfcc.begin());
        return mfcc;
    }
    // extractor.rs:87-97 — `tmp_sample` restarts at 0 on every call (every 160-sample hop)
    Vec pre_emphasis(const float* buf, size_t n) const {
        Vec out(n);
        float tmp_sample = 0.f;
        for (size_t i = 0; i < n; i++) {
            float previous = tmp_sample;
            tmp_sample = buf[i];
            out[i] = tmp_sample - pre_emphasis_coefficient * previous;
        }
        return out;
    }
    // extractor.rs:101-114
    Vec calculate_magnitude_spectrum(const float* frame) const {
        std::vector<std::complex<float>> buffer(samples_per_frame);
        for (size_t i = 0; i < samples_per_frame; i++) buffer[i] = {frame[i] * hamming_window[i], 0.f};
        fft.forward(buffer);
        Vec mag(magnitude_spectrum_size);
        for (size_t i = 0; i < magnitude_spectrum_size; i++)
            mag[i] = std::sqrt((buffer[i].real() * buffer[i].real()) + (buffer[i].imag() * buffer[i].imag()));
        return mag;
    }
    // extractor.rs:115-120
    static Vec new_hamming_window(size_t spf) {
        size_t ns_minus_1 = spf - 1;
        Vec w(spf);
        for (size_t s = 0; s < spf; s++) w[s] = 0.54f - (0.46f * std::cos(2.f * PI_F * ((float)s / (float)ns_minus_1)));
        return w;
    }
    // extractor.rs:121-131
    Vec calculate_mel_frequency_cepstral_coefficients(const Vec& mag) const {
        Vec mf = calculate_mel_frequency_cepstrum(mag);
        for (auto& ms : mf) ms = std::log(ms + std::numeric_limits<float>::min());  // f32::MIN_POSITIVE
        return discrete_cosine_transform(mf);
    }
    // extractor.rs:132-134
    static float frequency_to_mel(size_t f) { return 1127.f * std::log(1.f + ((float)f / 700.0f)); }
    // extractor.rs:135-145
    Vec calculate_mel_frequency_cepstrum(const Vec& mag) const {
        Vec out(num_coefficients);
        for (size_t i = 0; i < num_coefficients; i++) {
            float sum = 0.f;
            for (size_t j = 0; j < mag.size(); j++) sum += mag[j] * mag[j] * filter_bank[i][j];
            out[i] = sum;
        }
        return out;
    }
    // extractor.rs:146-163
    static Vec discrete_cosine_transform(const Vec& input) {
        Vec out(input.size());
        float pi_over_n = PI_F / (float)input.size();
        for (size_t k = 0; k < input.size(); k++) {
            float sum = 0.f;
            for (size_t n = 0; n < input.size(); n++)
                sum += input[n] * std::cos(pi_over_n * ((float)n + 0.5f) * (float)k);
            out[k] = 2.f * sum;
        }
        return out;
    }
    // extractor.rs:164-198
    static std::vector<size_t> centre_indices(size_t sample_rate, size_t mss, size_t ncoef, size_t min_f, size_t max_f) {
        float max_mel = std::floor(frequency_to_mel(max_f));
        float min_mel = std::floor(frequency_to_mel(min_f));
        std::vector<size_t> idx(ncoef + 2);
        for (size_t i = 0; i < ncoef + 2; i++) {
            float f = (float)i * (max_mel - min_mel) / (float)(ncoef + 1) + min_mel;
            float tmp = std::log(1.f + 1000.0f / 700.0f) / 1000.0f;
            tmp = (std::exp(f * tmp) - 1.f) / ((float)sample_rate / 2.f);
            idx[i] = (size_t)std::floor(0.5f + 700.f * (float)mss * tmp);
        }
        return idx;
    }
    static Mat new_mel_filter_bank(size_t sample_rate, size_t mss, size_t ncoef, size_t min_f, size_t max_f) {
        Mat fb(ncoef, Vec(mss, 0.f));
        std::vector<size_t> c = centre_indices(sample_rate, mss, ncoef, min_f, max_f);
        for (size_t i = 0; i < ncoef; i++) {
            size_t b = c[i], m = c[i + 1], e = c[i + 2];
            size_t up = m - b, down = e - m;
            for (size_t k = b; k < m; k++) fb[i][k] = (float)(k - b) / (float)up;
            for (size_t k = m; k < e; k++) fb[i][k] = (float)(e - k) / (float)down;
        }
        return fb;
    }
};

// ------------------------------------------------------------------------------------------
// MfccNormalizer::normalize — src/mfcc/normalizer.rs:3-31
// ------------------------------------------------------------------------------------------
static Mat normalize(const Mat& frames) {
    size_t num_frames = frames.size();
    if (num_frames == 0) return {};
    size_t num_mfccs = frames[0].size();
    Vec sum(num_mfccs, 0.f);
    Mat out(num_frames, Vec(num_mfccs, 0.f));
    for (size_t i = 0; i < num_frames; i++)
        for (size_t j = 0; j < num_mfccs; j++) {
            float v = frames[i][j];
            sum[j] += v;
            out[i][j] = v;
        }
    for (size_t i = 0; i < num_frames; i++)
        for (size_t j = 0; j < num_mfccs; j++) out[i][j] -= sum[j] / (float)num_frames;
    return out;
}

// ------------------------------------------------------------------------------------------
// cosine_similarity / calculate_distance — src/mfcc/comparator.rs:15-17,28-48
// ------------------------------------------------------------------------------------------
static float cosine_similarity(const float* a, size_t na, const float* b, size_t nb) {
    size_t dim = std::min(na, nb);
    float dot_ab = 0.f, dot_a = 0.f, dot_b = 0.f;
    for (size_t d = 0; d < dim; d++) {
        float ca = a[d], cb = b[d];
        dot_ab += ca * cb;
        dot_a += ca * ca;
        dot_b += cb * cb;
    }
    float magnitude = std::sqrt(dot_a * dot_b);
    return magnitude == 0.f ? 0.f : dot_ab / magnitude;
}
static float calculate_distance(const Vec& a, const Vec& b) {
    return 1.f - cosine_similarity(a.data(), a.size(), b.data(), b.size());
}

// ------------------------------------------------------------------------------------------
// Dtw::compute_optimal_path_with_window — src/mfcc/dtw.rs:56-105 (+ infinity_matrix :149-151)
// Keeps the reference's two O(m*n) matrices and the copy between them: that is where the
// reference's time goes and this oracle doubles as the timed CPU baseline.
// ------------------------------------------------------------------------------------------
static float min3_fold(float insertion, float deletion, float matches) {
    float a = INF;  // [insertion, deletion, matches].iter().fold(f32::INFINITY, |a,&b| a.min(b))
    a = std::fmin(a, insertion);
    a = std::fmin(a, deletion);
    a = std::fmin(a, matches);
    return a;
}
static float dtw_with_window(const Mat& first, const Mat& second, uint16_t w) {
    size_t state_m = first.size(), state_n = second.size();
    size_t diff = state_m > state_n ? state_m - state_n : state_n - state_m;
    size_t window = std::max<size_t>(w, diff);
    Mat dcm(state_m + 1, Vec(state_n + 1, INF));
    dcm[0][0] = 0.f;
    for (size_t r = 1; r <= state_m; r++) {
        size_t start = r > window ? std::max<size_t>(1, r - window) : 1;
        size_t end = std::min(state_n + 1, r + window);
        for (size_t c = start; c < end; c++) {
            float cost = calculate_distance(first[r - 1], second[c - 1]);
            dcm[r][c] = cost + min3_fold(dcm[r - 1][c], dcm[r][c - 1], dcm[r - 1][c - 1]);
        }
    }
    // resize matrix (dtw.rs:92-100)
    Mat fin(state_m + 1, Vec(state_n, INF));
    for (size_t r = 0; r <= state_m; r++)
        for (size_t c = 1; c <= state_n; c++) fin[r][c - 1] = dcm[r][c];
    return fin[state_m - 1][state_n - 1];  // dtw.rs:101 — i.e. D[m-1][n], not D[m][n]
}

// ------------------------------------------------------------------------------------------
// Dtw::compute_optimal_path + retrieve_optimal_path (unbanded; dtw.rs:11-55,106-138) — used by
// the averager only (build-time). Restated for completeness of dtw.rs.
// ------------------------------------------------------------------------------------------
struct DtwFull {
    size_t m = 0, n = 0;
    Mat dcm;
    float compute(const Mat& a, const Mat& b) {
        m = a.size();
        n = b.size();
        dcm.assign(m, Vec(n, INF));
        dcm[0][0] = calculate_distance(a[0], b[0]);
        for (size_t r = 1; r < m; r++) dcm[r][0] = calculate_distance(a[r], b[0]) + dcm[r - 1][0];
        for (size_t c = 1; c < n; c++) dcm[0][c] = calculate_distance(a[0], b[c]) + dcm[0][c - 1];
        for (size_t r = 1; r < m; r++)
            for (size_t c = 1; c < n; c++)
                dcm[r][c] = calculate_distance(a[r], b[c]) + min3_fold(dcm[r - 1][c], dcm[r][c - 1], dcm[r - 1][c - 1]);
        return dcm[m - 1][n - 1];
    }
    std::vector<std::pair<size_t, size_t>> path() const {
        size_t r = m - 1, c = n - 1;
        std::vector<std::pair<size_t, size_t>> p(std::min(r, c), {0, 0});  // dtw.rs:111 (spurious [0,0] entries)
        while (r > 0 || c > 0) {
            if (r > 0 && c > 0) {
                float ins = dcm[r - 1][c], del = dcm[r][c - 1], mat = dcm[r - 1][c - 1];
                float mn = min3_fold(ins, del, mat);
                if (mn == mat) { r--; c--; }
                else if (mn == ins) r--;
                else if (mn == del) c--;
            } else if (r > 0) r--;
            else c--;
            p.push_back({r, c});
        }
        std::reverse(p.begin(), p.end());
        return p;
    }
};

// ------------------------------------------------------------------------------------------
// MfccComparator — src/mfcc/comparator.rs:4-27
// ------------------------------------------------------------------------------------------
struct MfccComparator {
    float score_ref;
    uint16_t band_size;
    float compare(const Mat& a, const Mat& b) const {
        float cost = dtw_with_window(a, b, band_size);
        float normalized_cost = cost / (float)(a.size() + b.size());
        return compute_probability(normalized_cost);
    }
    float compute_probability(float cost) const { return 1.f / (1.f + std::exp((cost - score_ref) / score_ref)); }
};

// ------------------------------------------------------------------------------------------
// CBOR reader/writer for .rpw (wakeword_file.rs:10-42; ciborium 0.2.1 ⇒ RFC 8949)
// ------------------------------------------------------------------------------------------
struct CborReader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    std::string err;
    void fail(const char* m) {
        if (ok) { ok = false; err = m; }
    }
    uint8_t u8() {
        if (p >= end) { fail("unexpected end of CBOR"); return 0xff; }
        return *p++;
    }
    uint64_t arg(uint8_t ai) {
        if (ai < 24) return ai;
        int nb = ai == 24 ? 1 : ai == 25 ? 2 : ai == 26 ? 4 : ai == 27 ? 8 : 0;
        if (!nb) { fail("unsupported CBOR additional info"); return 0; }
        uint64_t v = 0;
        for (int i = 0; i < nb; i++) v = (v << 8) | u8();
        return v;
    }
    static float half_to_float(uint16_t h) {
        int s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff;
        float v;
        if (e == 0) v = std::ldexp((float)m, -24);
        else if (e == 31) v = m ? std::numeric_limits<float>::quiet_NaN() : INF;
        else v = std::ldexp((float)(m + 1024), e - 25);
        return s ? -v : v;
    }
    // header: returns major type, sets val (argument) and ai
    int head(uint64_t& val, uint8_t& ai) {
        uint8_t b = u8();
        ai = b & 0x1f;
        int major = b >> 5;
        if (ai == 31) { val = ~0ull; return major; }  // indefinite
        val = arg(ai);
        return major;
    }
    bool peek_break() { return p < end && *p == 0xff; }
    std::string text() {
        uint64_t n; uint8_t ai;
        int mj = head(n, ai);
        if (mj != 3 && mj != 2) { fail("expected CBOR string"); return {}; }
        if (ai == 31) { fail("indefinite strings unsupported"); return {}; }
        if ((uint64_t)(end - p) < n) { fail("string overruns buffer"); return {}; }
        std::string s((const char*)p, (size_t)n);
        p += n;
        return s;
    }
    // returns false for null (Option::None)
    bool number(float& out) {
        uint64_t v; uint8_t ai;
        int mj = head(v, ai);
        if (mj == 7) {
            if (ai == 22 || ai == 23) return false;  // null / undefined
            if (ai == 25) { out = half_to_float((uint16_t)v); return true; }
            if (ai == 26) { uint32_t u = (uint32_t)v; std::memcpy(&out, &u, 4); return true; }
            if (ai == 27) { double d; std::memcpy(&d, &v, 8); out = (float)d; return true; }
            if (ai == 20) { out = 0; return true; }
            if (ai == 21) { out = 1; return true; }
            fail("unexpected simple value");
            return false;
        }
        if (mj == 0) { out = (float)v; return true; }
        if (mj == 1) { out = -1.f - (float)v; return true; }
        fail("expected CBOR number");
        return false;
    }
    // Vec<Vec<f32>> or null
    bool matrix(Mat& out) {
        if (p < end && (*p == 0xf6 || *p == 0xf7)) { p++; return false; }
        uint64_t n; uint8_t ai;
        if (head(n, ai) != 4) { fail("expected array"); return false; }
        out.clear();
        for (uint64_t i = 0; ok && (ai == 31 ? !peek_break() : i < n); i++) {
            uint64_t k; uint8_t ai2;
            if (head(k, ai2) != 4) { fail("expected inner array"); return false; }
            Vec row;
            for (uint64_t j = 0; ok && (ai2 == 31 ? !peek_break() : j < k); j++) {
                float f = 0;
                if (!number(f)) fail("null inside matrix");
                row.push_back(f);
            }
            if (ai2 == 31) u8();
            out.push_back(std::move(row));
        }
        if (ai == 31) u8();
        return ok;
    }
    void skip() {
        uint64_t v; uint8_t ai;
        int mj = head(v, ai);
        switch (mj) {
            case 0: case 1: case 7: break;
            case 2: case 3:
                if (ai == 31) { while (ok && !peek_break()) skip(); u8(); }
                else { if ((uint64_t)(end - p) < v) fail("overrun"); else p += v; }
                break;
            case 4:
                if (ai == 31) { while (ok && !peek_break()) skip(); u8(); }
                else for (uint64_t i = 0; ok && i < v; i++) skip();
                break;
            case 5:
                if (ai == 31) { while (ok && !peek_break()) { skip(); skip(); } u8(); }
                else for (uint64_t i = 0; ok && i < v; i++) { skip(); skip(); }
                break;
            case 6: skip(); break;
        }
    }
};

// WakewordRef — src/wakewords/wakeword_ref.rs:12-20 ; WakewordV2 — wakeword_v2.rs:8-30
struct WakewordRef {
    std::string name;
    std::optional<Mat> avg_features;
    std::vector<std::pair<std::string, Mat>> samples_features;  // HashMap in the reference; file order here
    std::optional<float> threshold, avg_threshold;
    float rms_level = 0.f;
    uint16_t mfcc_size = 0;
    std::vector<Vec> flat_cache;  // for the C accessor
};

// Tries WakewordV2 then WakewordRef (detector.rs:152-163). serde's derive rejects a map that
// lacks a required field, so V2 needs `enabled`, Ref needs `mfcc_size`; a WakewordModel file has
// neither `samples_features` nor `rms_level`-with-templates and is rejected by this path.
static bool parse_wakeword(const uint8_t* buf, size_t len, WakewordRef& out, std::string& err) {
    CborReader r{buf, buf + len, true, {}};
    uint64_t n; uint8_t ai;
    if (r.head(n, ai) != 5) { err = "not a CBOR map"; return false; }
    bool has_name = false, has_avg = false, has_samples = false, has_thr = false, has_athr = false, has_rms = false,
         has_mfcc = false, has_enabled = false;
    for (uint64_t i = 0; r.ok && (ai == 31 ? !r.peek_break() : i < n); i++) {
        std::string key = r.text();
        if (!r.ok) break;
        if (key == "name") { out.name = r.text(); has_name = true; }
        else if (key == "avg_features") { Mat m; if (r.matrix(m)) out.avg_features = std::move(m); has_avg = true; }
        else if (key == "samples_features") {
            uint64_t k; uint8_t ai2;
            if (r.head(k, ai2) != 5) { r.fail("samples_features is not a map"); break; }
            for (uint64_t j = 0; r.ok && (ai2 == 31 ? !r.peek_break() : j < k); j++) {
                std::string tn = r.text();
                Mat m;
                if (!r.matrix(m)) r.fail("null template");
                out.samples_features.emplace_back(std::move(tn), std::move(m));
            }
            if (ai2 == 31) r.u8();
            has_samples = true;
        }
        else if (key == "threshold") { float f; if (r.number(f)) out.threshold = f; has_thr = true; }
        else if (key == "avg_threshold") { float f; if (r.number(f)) out.avg_threshold = f; has_athr = true; }
        else if (key == "rms_level") { float f = 0; if (!r.number(f)) r.fail("rms_level is null"); out.rms_level = f; has_rms = true; }
        else if (key == "mfcc_size") { float f = 0; r.number(f); out.mfcc_size = (uint16_t)f; has_mfcc = true; }
        else if (key == "enabled") { r.skip(); has_enabled = true; }
        else r.skip();
    }
    if (!r.ok) { err = r.err; return false; }
    (void)has_thr; (void)has_athr; (void)has_avg;  // Option fields default to None when absent
    if (!has_name || !has_samples || !has_rms) { err = "missing field (not a WakewordRef / WakewordV2 file)"; return false; }
    if (!has_mfcc && !has_enabled) { err = "missing field `mfcc_size`"; return false; }
    if (out.samples_features.empty() || out.samples_features[0].second.empty()) { err = "wakeword without templates"; return false; }
    if (!has_mfcc) out.mfcc_size = (uint16_t)out.samples_features[0].second[0].size();  // wakeword_v2.rs:22
    return true;
}

struct CborWriter {
    std::vector<uint8_t> b;
    void head(int major, uint64_t v) {
        if (v < 24) b.push_back((uint8_t)(major << 5 | v));
        else if (v < 256) { b.push_back((uint8_t)(major << 5 | 24)); b.push_back((uint8_t)v); }
        else if (v < 65536) { b.push_back((uint8_t)(major << 5 | 25)); b.push_back((uint8_t)(v >> 8)); b.push_back((uint8_t)v); }
        else { b.push_back((uint8_t)(major << 5 | 26)); for (int s = 24; s >= 0; s -= 8) b.push_back((uint8_t)(v >> s)); }
    }
    void text(const std::string& s) { head(3, s.size()); b.insert(b.end(), s.begin(), s.end()); }
    void f32(float f) { uint32_t u; std::memcpy(&u, &f, 4); b.push_back(0xfa); for (int s = 24; s >= 0; s -= 8) b.push_back((uint8_t)(u >> s)); }
    void null() { b.push_back(0xf6); }
    void matrix(const float* d, int rows, int cols) {
        head(4, rows);
        for (int i = 0; i < rows; i++) { head(4, cols); for (int j = 0; j < cols; j++) f32(d[(size_t)i * cols + j]); }
    }
};

// ------------------------------------------------------------------------------------------
// ScoreMode / percentile — src/config.rs:86-97, wakeword_comp.rs:38-49
// ------------------------------------------------------------------------------------------
enum ScoreMode { Average = 0, Max, Median, P25, P50, P75, P80, P90, P95 };

static bool total_less(float a, float b) {  // f32::total_cmp
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    ia ^= (int32_t)((uint32_t)(ia >> 31) >> 1);
    ib ^= (int32_t)((uint32_t)(ib >> 31) >> 1);
    return ia < ib;
}
static float get_percentile(const Vec& sorted, float percentile) {
    size_t n = sorted.size();
    float index = percentile / 100.0f * (float)(n - 1);
    float index_floor = std::floor(index);
    if (index_floor == index) return sorted[(size_t)index];
    size_t i = (size_t)index_floor;
    float d = index - index_floor;
    return sorted[i] * (1.0f - d) + sorted[i + 1] * d;
}
static float aggregate(Vec scores, int mode) {  // wakeword_comp.rs:108-139
    switch (mode) {
        case Average: {
            float s = 0.f;
            for (float v : scores) s += v;
            return s / (float)scores.size();
        }
        case Max:
            std::sort(scores.begin(), scores.end(), [](float a, float b) { return total_less(b, a); });
            return scores[0];
        default: {
            std::sort(scores.begin(), scores.end(), total_less);
            float p = mode == Median || mode == P50 ? 50.f : mode == P25 ? 25.f : mode == P75 ? 75.f : mode == P80 ? 80.f
                      : mode == P90 ? 90.f : 95.f;
            return get_percentile(scores, p);
        }
    }
}

// ------------------------------------------------------------------------------------------
// RustpotterDetection — src/detector.rs:487-501
// ------------------------------------------------------------------------------------------
struct Detection {
    std::string name;
    float avg_score = 0.f, score = 0.f;
    std::vector<std::pair<std::string, float>> scores;
    size_t counter = 0;
    float gain = std::numeric_limits<float>::quiet_NaN();
};

// ------------------------------------------------------------------------------------------
// WakewordComparator — src/wakewords/comp/wakeword_comp.rs:9-166
// ------------------------------------------------------------------------------------------
struct WakewordComparator {
    WakewordRef ww;
    int score_mode;
    MfccComparator cmp;
    // wakeword_comp.rs:22-27 — keeps the FIRST max_len frames, then CMN
    static Mat cut_and_normalize_frame(Mat mfccs, size_t max_len) {
        if (mfccs.size() > max_len) mfccs.erase(mfccs.begin() + max_len, mfccs.end());
        return normalize(mfccs);
    }
    // wakeword_comp.rs:28-37 — a = template, b = window
    float score_frame(const Mat& frame_features, const Mat& tmpl) const { return cmp.compare(tmpl, frame_features); }
    size_t get_mfcc_frame_size() const {  // :69-75
        size_t mx = 0;
        for (auto& kv : ww.samples_features) mx = std::max(mx, kv.second.size());
        return mx;
    }
    uint16_t get_mfcc_size() const { return (uint16_t)ww.samples_features[0].second[0].size(); }  // :158-160
    // wakeword_comp.rs:77-152
    std::optional<Detection> run_detection(const Mat& mfcc_frame, float avg_threshold_cfg, float threshold_cfg) const {
        float avg_threshold = ww.avg_threshold.value_or(avg_threshold_cfg);
        float avg_score = 0.f;
        if (ww.avg_features.has_value() && avg_threshold != 0.f) {
            Mat nw = cut_and_normalize_frame(mfcc_frame, ww.avg_features->size());
            avg_score = score_frame(nw, *ww.avg_features);
            if (avg_score < avg_threshold) return std::nullopt;
        }
        float threshold = ww.threshold.value_or(threshold_cfg);
        std::vector<std::pair<std::string, float>> scores;
        Vec values;
        for (auto& kv : ww.samples_features) {
            Mat nw = cut_and_normalize_frame(mfcc_frame, kv.second.size());
            float s = score_frame(nw, kv.second);
            scores.emplace_back(kv.first, s);
            values.push_back(s);
        }
        float score = aggregate(values, score_mode);
        if (score > threshold) {
            Detection d;
            d.name = ww.name;
            d.avg_score = avg_score;
            d.score = score;
            d.scores = std::move(scores);
            d.counter = 0;
            return d;
        }
        return std::nullopt;
    }
};

// ------------------------------------------------------------------------------------------
// VadDetector — src/mfcc/vad.rs:3-50
// ------------------------------------------------------------------------------------------
struct VadDetector {
    float mode_value;
    size_t index = 0;
    Vec window = Vec(50, std::numeric_limits<float>::quiet_NaN());
    size_t voice_countdown = 0;
    bool is_voice(const Vec& mfcc) {
        float sum = 0.f;
        for (float v : mfcc) sum += std::fabs(v);
        float value = sum / (float)mfcc.size();
        window[index] = value;
        index = index >= window.size() - 1 ? 0 : index + 1;
        float mn = INF;
        bool any = false;
        for (float v : window)
            if (!std::isnan(v) && (!any || total_less(v, mn))) { mn = v; any = true; }
        mn = std::fmax(mn, 0.01f);
        float th = mn * mode_value;
        size_t n_high = 0;
        for (float v : window)
            if (v > th) n_high++;
        if (n_high > 10) voice_countdown = 500;
        if (voice_countdown > 0) { voice_countdown--; return true; }
        return false;
    }
    void reset() {
        std::fill(window.begin(), window.end(), std::numeric_limits<float>::quiet_NaN());
        voice_countdown = 0;
        index = 0;
    }
};

// ------------------------------------------------------------------------------------------
// GainNormalizerFilter — src/audio/gain_normalizer_filter.rs:3-79
// ------------------------------------------------------------------------------------------
struct GainNormalizerFilter {
    size_t window_size = 1;
    bool fixed_rms_level;
    float min_gain, max_gain;
    float rms_level_ref, rms_level_sqrt;
    Vec rms_level_window;
    GainNormalizerFilter(float mn, float mx, std::optional<float> fixed)
        : fixed_rms_level(fixed.has_value()), min_gain(mn), max_gain(mx),
          rms_level_ref(fixed.value_or(std::numeric_limits<float>::quiet_NaN())),
          rms_level_sqrt(fixed ? std::sqrt(*fixed) : std::numeric_limits<float>::quiet_NaN()) {}
    float filter(Vec& signal, float rms_level) {
        if (!std::isnan(rms_level_ref) && rms_level != 0.f) {
            rms_level_window.push_back(rms_level);
            if (rms_level_window.size() > window_size) rms_level_window.erase(rms_level_window.begin());
            float s = 0.f;
            for (float v : rms_level_window) s += v;
            float frame_rms_level = s / (float)rms_level_window.size();
            float gain = rms_level_sqrt / std::sqrt(frame_rms_level);
            gain = std::round(gain * 10.f) / 10.f;
            gain = std::fmin(std::fmax(gain, min_gain), max_gain);  // clamp
            if (gain != 1.f)
                for (float& x : signal) x = std::fmin(std::fmax(x * gain, -1.f), 1.f);
            return gain;
        }
        return 1.f;
    }
    void set_rms_level_ref(float rms_level, size_t ws) {
        if (!fixed_rms_level) {
            rms_level_ref = rms_level;
            rms_level_sqrt = std::sqrt(rms_level);
        }
        window_size = ws != 0 ? ws : 1;
    }
    static float get_rms_level(const Vec& signal) {
        float sum_squared = 0.0f;
        for (float s : signal) sum_squared += s * s;
        return std::sqrt(sum_squared / (float)signal.size());
    }
};

// ------------------------------------------------------------------------------------------
// BandPassFilter — src/audio/band_pass_filter.rs:5-67
// ------------------------------------------------------------------------------------------
struct BandPassFilter {
    float a0, a1, a2, b1, b2;
    float x1 = 0, x2 = 0, y1 = 0, y2 = 0;
    BandPassFilter(float sample_rate, float low_cutoff, float high_cutoff) {
        float omega_low = 2.0f * PI_F * low_cutoff / sample_rate;
        float omega_high = 2.0f * PI_F * high_cutoff / sample_rate;
        float cos_omega_low = std::cos(omega_low), cos_omega_high = std::cos(omega_high);
        float alpha_low = std::sin(omega_low) / 2.0f, alpha_high = std::sin(omega_high) / 2.0f;
        a0 = 1.0f / (1.0f + alpha_high - alpha_low);
        a1 = -2.0f * cos_omega_low * a0;
        a2 = (1.0f - alpha_high - alpha_low) * a0;
        b1 = -2.0f * cos_omega_high * a0;
        b2 = (1.0f - alpha_high + alpha_low) * a0;
    }
    void filter(Vec& signal) {
        for (float& sample : signal) {
            float x = sample;
            sample = a0 * x + a1 * x1 + a2 * x2 - b1 * y1 - b2 * y2;
            x2 = x1;
            x1 = x;
            y2 = y1;
            y1 = sample;
        }
    }
};

// ------------------------------------------------------------------------------------------
// rubato 0.14.1 `FftFixedInOut<f32>` with one channel (used by src/audio/encoder.rs:52-60,72-79 when the source
// rate is not 16 kHz). Restatement of rubato's published algorithm:
//   synchro.rs  FftFixedInOut::new: gcd = gcd(fs_in, fs_out); fft_chunks = ceil(chunk_size_in / (fs_out / gcd));
//               fft_size_in = fft_chunks * fs_in / gcd; fft_size_out = fft_chunks * fs_out / gcd
//   synchro.rs  FftResampler::new: cutoff = calculate_cutoff(fft_size_in, window) (* fft_size_out / fft_size_in when downsampling; see below);
//               filter = make_sincs(fft_size_in, 1, cutoff, BlackmanHarris2)[0] / (2 fft_size_in), zero-padded to
//               2 fft_size_in, forward real FFT
//   synchro.rs  resample_unit: input zero-padded to 2 fft_size_in -> real FFT -> first new_len bins times the filter ->
//               the rest of the fft_size_out + 1 bins zero -> inverse real FFT of 2 fft_size_out points (unnormalised)
//               -> out[n] = buf[n] + overlap[n]; overlap = buf[fft_size_out ..]
//   sinc.rs     make_sincs: y[x] = window[x] * sinc((x - n/2) * cutoff), normalised to unit sum
//   windows.rs  blackman_harris: 0.35875 - 0.48829 cos(2 pi x/n) + 0.14128 cos(4 pi x/n) - 0.01168 cos(6 pi x/n), squared
// ------------------------------------------------------------------------------------------
struct FftFixedInOut {
    size_t fft_size_in = 0, fft_size_out = 0;
    std::vector<std::complex<float>> filter_f;   // fft_size_in + 1 bins
    Vec overlap;                                 // fft_size_out
    std::unique_ptr<Fft> fft_in, fft_out;

    static size_t gcd(size_t a, size_t b) { return b == 0 ? a : gcd(b, a % b); }
    static float sinc(float v) { return v == 0.f ? 1.f : std::sin(v * PI_F) / (v * PI_F); }

    FftFixedInOut(size_t fs_in, size_t fs_out, size_t chunk_size_in) {
        const size_t g = gcd(fs_in, fs_out);
        const size_t min_chunk_out = fs_out / g;
        const size_t fft_chunks = (size_t)std::ceil((float)chunk_size_in / (float)min_chunk_out);
        fft_size_out = fft_chunks * fs_out / g;
        fft_size_in = fft_chunks * fs_in / g;
        // Anti-aliasing cutoff relative to the lower Nyquist frequency. rubato derives it from the sinc length and the
        // window (sinc.rs calculate_cutoff, an empirical fit whose constants are not available offline); here the
        // same 1/(1 + k/n) law — the transition band of a fixed window shrinks with 1/n — is calibrated on the
        // reference's own 48 kHz goldens (tests/detector.rs:162-214), which pin cutoff(n = 1440) = 0.97161 +- 3e-5
        // (every avg_score/score/counter of both goldens reproduced to < 1e-6 relative; a 1 % change of the cutoff
        // moves the scores by 2e-4). Other lengths (other source rates) follow the law but are unpinned.
        const float rel = 1.0f / (1.0f + 42.08f / (float)fft_size_in);
        const float cutoff = fft_size_in > fft_size_out ? rel * (float)fft_size_out / (float)fft_size_in : rel;
        const size_t n = fft_size_in;
        Vec y(n);
        float sum = 0.f;
        for (size_t x = 0; x < n; x++) {
            const float xf = (float)x, nf = (float)n;
            float w = 0.35875f - 0.48829f * std::cos(2.f * PI_F * xf / nf) + 0.14128f * std::cos(4.f * PI_F * xf / nf) -
                      0.01168f * std::cos(6.f * PI_F * xf / nf);
            w = w * w;
            const float val = w * sinc(((float)x - (float)(n / 2)) * cutoff);
            sum += val;
            y[x] = val;
        }
        fft_in = std::make_unique<Fft>(2 * fft_size_in);
        fft_out = std::make_unique<Fft>(2 * fft_size_out);
        std::vector<std::complex<float>> buf(2 * n);
        for (size_t x = 0; x < n; x++) buf[x] = std::complex<float>((y[x] / sum) / (float)(2 * n), 0.f);
        fft_in->forward(buf);
        filter_f.assign(buf.begin(), buf.begin() + n + 1);
        overlap.assign(fft_size_out, 0.f);
    }
    size_t input_frames_next() const { return fft_size_in; }

    Vec process(const Vec& wave_in) {   // process_into_buffer for one channel
        std::vector<std::complex<float>> buf(2 * fft_size_in);
        for (size_t i = 0; i < fft_size_in; i++) buf[i] = std::complex<float>(i < wave_in.size() ? wave_in[i] : 0.f, 0.f);
        fft_in->forward(buf);
        const size_t new_len = fft_size_in < fft_size_out ? fft_size_in + 1 : fft_size_out;
        // Hermitian spectrum of 2 fft_size_out points: bins 0 .. new_len-1 kept, the others (Nyquist included) zero
        const size_t m = 2 * fft_size_out;
        std::vector<std::complex<float>> spec(m);
        for (size_t k = 0; k < new_len && k <= fft_size_out; k++) {
            const std::complex<float> v = buf[k] * filter_f[k];
            // a real inverse transform ignores the imaginary part of the DC (and Nyquist) bin
            spec[k] = (k == 0 || k == fft_size_out) ? std::complex<float>(v.real(), 0.f) : v;
            if (k != 0 && k != fft_size_out) spec[m - k] = std::conj(v);
        }
        // unnormalised inverse = conj(forward(conj(spec)))
        for (auto& v : spec) v = std::conj(v);
        fft_out->forward(spec);
        Vec out(fft_size_out);
        for (size_t n2 = 0; n2 < fft_size_out; n2++) out[n2] = spec[n2].real() + overlap[n2];
        for (size_t n2 = 0; n2 < fft_size_out; n2++) overlap[n2] = spec[fft_size_out + n2].real();
        return out;
    }
};

// ------------------------------------------------------------------------------------------
// AudioEncoder — src/audio/encoder.rs:6-115 and Sample::into_f32 — audio_types.rs:98-137
// ------------------------------------------------------------------------------------------
struct AudioEncoder {
    uint32_t fmt, channels, endianness;
    size_t input_samples_per_frame, output_samples_per_frame;
    std::shared_ptr<FftFixedInOut> resampler;   // encoder.rs:72-79 (None when the source rate is 16 kHz)
    // AudioEncoder::new (encoder.rs:63-102)
    AudioEncoder(uint32_t fmt_, uint32_t channels_, uint32_t endianness_, size_t sample_rate, size_t frame_length_ms, size_t target_rate)
        : fmt(fmt_), channels(channels_), endianness(endianness_),
          input_samples_per_frame((sample_rate * frame_length_ms / 1000) * channels_),
          output_samples_per_frame(target_rate * frame_length_ms / 1000) {
        if (sample_rate != target_rate) {
            resampler = std::make_shared<FftFixedInOut>(sample_rate, target_rate, output_samples_per_frame);
            input_samples_per_frame = resampler->input_frames_next() * channels;
        }
    }
    // reencode_to_mono_with_sample_rate (encoder.rs:41-62)
    Vec finish(Vec buffer) const { return resampler ? resampler->process(to_mono(std::move(buffer))) : to_mono(std::move(buffer)); }
    size_t bytes_per_sample() const { return fmt == 0 ? 1 : fmt == 1 ? 2 : 4; }
    size_t input_byte_length() const { return input_samples_per_frame * bytes_per_sample(); }
    Vec to_mono(Vec buffer) const {
        if (channels != 1) {
            Vec mono;
            for (size_t i = 0; i + channels <= buffer.size(); i += channels) mono.push_back(buffer[i]);
            return mono;
        }
        return buffer;
    }
    Vec encode_bytes(const uint8_t* b, size_t len) const {
        size_t bs = bytes_per_sample();
        bool big = endianness == 1;  // native == little on every platform this runs on
        Vec out;
        out.reserve(len / bs);
        for (size_t i = 0; i + bs <= len; i += bs) {
            uint32_t u = 0;
            for (size_t k = 0; k < bs; k++) u |= (uint32_t)b[i + (big ? bs - 1 - k : k)] << (8 * k);
            switch (fmt) {
                case 0: out.push_back((float)(int8_t)u / 127.f); break;
                case 1: out.push_back((float)(int16_t)u / 32767.f); break;
                case 2: out.push_back((float)(int32_t)u / (float)2147483647); break;
                default: { float f; std::memcpy(&f, &u, 4); out.push_back(f); }
            }
        }
        return finish(std::move(out));
    }
};

// ------------------------------------------------------------------------------------------
// Rustpotter — src/detector.rs:34-454
// ------------------------------------------------------------------------------------------
struct Rustpotter {
    float avg_threshold, threshold;
    size_t min_scores;
    bool eager;
    int score_mode;
    std::optional<VadDetector> vad_detector;
    AudioEncoder wav_encoder;
    MfccExtractor mfcc_extractor;
    float score_ref;
    uint16_t band_size;
    std::optional<BandPassFilter> band_pass_filter;
    std::optional<GainNormalizerFilter> gain_normalizer_filter;
    std::vector<std::pair<std::string, WakewordComparator>> wakewords;  // HashMap in the reference
    bool buffering = true;
    Mat audio_mfcc_window;
    size_t max_mfcc_frames = 0;
    std::optional<Detection> partial_detection;
    size_t detection_countdown = 0;
    float rms_level = 0.f, gain = 1.f;
    uint64_t windows_scored = 0;  // instrumentation: calls of run_wakeword_detectors
    // trace hook for parity tests: called with the window before run_wakeword_detectors
    std::vector<Vec>* trace = nullptr;

    static float vad_value(int m) { return m == 0 ? 2.f : m == 1 ? 2.5f : 3.f; }  // config.rs:141-148

    explicit Rustpotter(const rpo_config& c)  // detector.rs:95-142
        : avg_threshold(c.avg_threshold), threshold(c.threshold), min_scores(c.min_scores), eager(c.eager != 0),
          score_mode((int)c.score_mode),
          wav_encoder(c.sample_format, c.channels, c.endianness, c.sample_rate, FRAME_LENGTH_MS, SAMPLE_RATE),
          mfcc_extractor(SAMPLE_RATE, SAMPLE_RATE * FRAME_LENGTH_MS / 1000,
                         (size_t)((float)(SAMPLE_RATE * FRAME_LENGTH_MS / 1000) / ((float)FRAME_LENGTH_MS / (float)FRAME_SHIFT_MS)),
                         0, PRE_EMPHASIS),
          score_ref(c.score_ref), band_size((uint16_t)c.band_size) {
        if (c.vad_mode >= 0) vad_detector = VadDetector{vad_value(c.vad_mode)};
        set_filters(c);
    }
    void set_filters(const rpo_config& c) {
        band_pass_filter.reset();
        gain_normalizer_filter.reset();
        if (c.gain_normalizer_enabled)
            gain_normalizer_filter.emplace(c.min_gain, c.max_gain,
                                           c.gain_ref_set ? std::optional<float>(c.gain_ref) : std::nullopt);
        if (c.band_pass_enabled) band_pass_filter.emplace((float)SAMPLE_RATE, c.low_cutoff, c.high_cutoff);
    }
    // detector.rs:304-327
    bool add_wakeword(const std::string& key, WakewordRef ww, std::string& err) {
        if (wakewords.empty()) {
            reset();
            mfcc_extractor.set_out_size(ww.mfcc_size);
        } else if (wakewords.front().second.get_mfcc_size() != ww.mfcc_size) {
            err = "Usage of wakewords with different mfcc size is not supported, ignoring wakeword";
            return false;
        }
        WakewordComparator wc{std::move(ww), score_mode, MfccComparator{score_ref, band_size}};
        bool replaced = false;
        for (auto& kv : wakewords)
            if (kv.first == key) { kv.second = std::move(wc); replaced = true; break; }
        if (!replaced) wakewords.emplace_back(key, std::move(wc));
        on_wakeword_change();
        return true;
    }
    bool remove_wakeword(const std::string& key) {  // :180-189
        size_t len = wakewords.size();
        wakewords.erase(std::remove_if(wakewords.begin(), wakewords.end(), [&](auto& kv) { return kv.first == key; }),
                        wakewords.end());
        if (len != wakewords.size()) { on_wakeword_change(); return true; }
        return false;
    }
    bool remove_wakewords() {  // :193-202
        size_t len = wakewords.size();
        wakewords.clear();
        if (len != 0) { on_wakeword_change(); return true; }
        return false;
    }
    void on_wakeword_change() {  // :328-346
        size_t mx = 0;
        float target_rms_level = std::numeric_limits<float>::quiet_NaN();
        for (auto& kv : wakewords) {
            mx = std::max(kv.second.get_mfcc_frame_size(), mx);
            target_rms_level = std::fmax(kv.second.ww.rms_level, target_rms_level);  // f32::max ignores NaN
        }
        max_mfcc_frames = mx;
        if (gain_normalizer_filter) gain_normalizer_filter->set_rms_level_ref(target_rms_level, max_mfcc_frames / 3);
        buffering = audio_mfcc_window.size() < max_mfcc_frames;
    }
    void reset() {  // :290-302
        buffering = true;
        partial_detection.reset();
        audio_mfcc_window.clear();
        mfcc_extractor.reset();
        if (vad_detector) vad_detector->reset();
    }
    void update_detector_config(const rpo_config& c) {  // :265-282
        avg_threshold = c.avg_threshold;
        threshold = c.threshold;
        min_scores = c.min_scores;
        eager = c.eager != 0;
        band_size = (uint16_t)c.band_size;
        score_ref = c.score_ref;
        score_mode = (int)c.score_mode;
        vad_detector.reset();
        if (c.vad_mode >= 0) vad_detector = VadDetector{vad_value(c.vad_mode)};
        for (auto& kv : wakewords) {
            kv.second.score_mode = score_mode;
            kv.second.cmp = MfccComparator{score_ref, band_size};
        }
        reset();
    }
    void update_filters_config(const rpo_config& c) {  // :283-289
        set_filters(c);
        // NB: the reference builds fresh filters here and does NOT re-run on_wakeword_change, so a
        // new gain-normaliser keeps rms_level_ref = NaN (or the fixed gain_ref) until the next
        // wakeword change.
        reset();
    }
    std::optional<Detection> process_bytes(const uint8_t* b, size_t len) {  // :234-240
        if (len != wav_encoder.input_byte_length()) return std::nullopt;
        return process_audio(wav_encoder.encode_bytes(b, len));
    }
    template <typename T>
    std::optional<Detection> process_samples(const T* s, size_t n, float maxv) {  // :245-256
        if (n != wav_encoder.input_samples_per_frame) return std::nullopt;
        Vec f(n);
        for (size_t i = 0; i < n; i++) f[i] = maxv == 0.f ? (float)s[i] : (float)s[i] / maxv;
        return process_audio(wav_encoder.finish(std::move(f)));
    }
    std::optional<Detection> process_audio(Vec audio_buffer) {  // :347-376
        if (wakewords.empty()) return std::nullopt;
        rms_level = GainNormalizerFilter::get_rms_level(audio_buffer);
        if (gain_normalizer_filter) gain = gain_normalizer_filter->filter(audio_buffer, rms_level);
        if (band_pass_filter) band_pass_filter->filter(audio_buffer);
        Mat frames = mfcc_extractor.compute(audio_buffer.data(), audio_buffer.size());
        for (auto& f : frames) {  // find_map: stop at the first Some
            auto d = process_new_mfccs(std::move(f));
            if (d) return d;
        }
        return std::nullopt;
    }
    std::optional<Detection> process_new_mfccs(Vec mfcc_frame) {  // :377-397
        std::optional<Detection> result;
        bool should_run = partial_detection.has_value() || !vad_detector || vad_detector->is_voice(mfcc_frame);
        audio_mfcc_window.push_back(std::move(mfcc_frame));
        if (audio_mfcc_window.size() >= max_mfcc_frames) {
            if (buffering) buffering = false;
            if (should_run) result = run_detection();
        }
        if (audio_mfcc_window.size() >= max_mfcc_frames) audio_mfcc_window.erase(audio_mfcc_window.begin());
        return result;
    }
    std::optional<Detection> run_detection() {  // :398-432
        if (detection_countdown != 0) detection_countdown -= 1;
        if (partial_detection && is_detection_done(*partial_detection)) {
            Detection d = std::move(*partial_detection);
            partial_detection.reset();
            if (d.counter >= min_scores) {
                reset();
                return d;
            }
        }
        std::optional<Detection> det = run_wakeword_detectors();
        if (det) {
            det->counter = partial_detection ? partial_detection->counter + 1 : 1;
            det->gain = gain;
            if (!partial_detection || partial_detection->score < det->score) partial_detection = std::move(det);
            else partial_detection->counter = det->counter;
            detection_countdown = max_mfcc_frames / 2;
        }
        return std::nullopt;
    }
    std::optional<Detection> run_wakeword_detectors() {  // :433-447
        windows_scored++;
        if (trace) trace_window();
        std::optional<Detection> best;
        for (auto& kv : wakewords) {
            auto d = kv.second.run_detection(audio_mfcc_window, avg_threshold, threshold);
            // sort_by(b.score.total_cmp(a.score)) then first: the highest score; ties keep the
            // earlier one (stable sort over an unordered HashMap iteration in the reference).
            if (d && (!best || total_less(best->score, d->score))) best = std::move(d);
        }
        return best;
    }
    bool is_detection_done(const Detection& d) const {  // :448-454
        if (detection_countdown == 0) return true;
        return eager && d.counter >= min_scores;
    }
    // test instrumentation (not in the reference): ungated scores of the first wakeword
    void trace_window() {
        const WakewordComparator& wc = wakewords.front().second;
        Vec row;
        float avg = 0.f;
        if (wc.ww.avg_features) {
            Mat nw = WakewordComparator::cut_and_normalize_frame(audio_mfcc_window, wc.ww.avg_features->size());
            avg = wc.score_frame(nw, *wc.ww.avg_features);
        }
        Vec vals;
        for (auto& kv : wc.ww.samples_features) {
            Mat nw = WakewordComparator::cut_and_normalize_frame(audio_mfcc_window, kv.second.size());
            vals.push_back(wc.score_frame(nw, kv.second));
        }
        row.push_back(avg);
        row.push_back(aggregate(vals, wc.score_mode));
        row.insert(row.end(), vals.begin(), vals.end());
        trace->push_back(std::move(row));
    }
};

// ------------------------------------------------------------------------------------------
// WAV decoding as hound 3.5 + AudioEncoder see it (src/mfcc/wav_file_extractor.rs:18-112): PCM
// int 8/16/32 or IEEE float 32, little endian; 8-bit WAV is unsigned on disk and signed in hound.
// ------------------------------------------------------------------------------------------
struct WavData {
    uint32_t sample_rate = 0;
    uint16_t channels = 0, bits = 0;
    bool is_float = false;
    std::vector<float> f32;  // every interleaved sample after Sample::into_f32 (audio_types.rs:98-137)
};
static bool parse_wav(const uint8_t* b, size_t len, WavData& w, std::string& err) {
    auto u16 = [&](size_t o) { return (uint16_t)(b[o] | b[o + 1] << 8); };
    auto u32 = [&](size_t o) { return (uint32_t)b[o] | (uint32_t)b[o + 1] << 8 | (uint32_t)b[o + 2] << 16 | (uint32_t)b[o + 3] << 24; };
    if (len < 12 || std::memcmp(b, "RIFF", 4) || std::memcmp(b + 8, "WAVE", 4)) { err = "no RIFF tag found"; return false; }
    size_t off = 12;
    bool have_fmt = false;
    uint16_t tag = 0;
    while (off + 8 <= len) {
        uint32_t sz = u32(off + 4);
        size_t body = off + 8;
        if (!std::memcmp(b + off, "fmt ", 4) && body + 16 <= len) {
            tag = u16(body);
            w.channels = u16(body + 2);
            w.sample_rate = u32(body + 4);
            w.bits = u16(body + 14);
            if (tag == 0xfffe && body + 26 <= len) tag = u16(body + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
            have_fmt = true;
        } else if (!std::memcmp(b + off, "data", 4)) {
            if (!have_fmt) { err = "data chunk before fmt chunk"; return false; }
            size_t n = std::min<size_t>(sz, len - body);
            w.is_float = tag == 3;
            // TryFrom<WavSpec> for AudioFmt (wav_file_extractor.rs:93-112)
            bool ok = w.is_float ? w.bits == 32 : (w.bits == 8 || w.bits == 16 || w.bits == 32);
            if (!ok || (tag != 1 && tag != 3)) { err = "Unsupported wav format"; return false; }
            size_t bs = w.bits / 8, cnt = n / bs;
            w.f32.resize(cnt);
            for (size_t i = 0; i < cnt; i++) {
                const uint8_t* p = b + body + i * bs;
                if (w.is_float) { std::memcpy(&w.f32[i], p, 4); }
                else if (w.bits == 8) w.f32[i] = (float)(int8_t)(p[0] - 128) / 127.f;
                else if (w.bits == 16) w.f32[i] = (float)(int16_t)(p[0] | p[1] << 8) / 32767.f;
                else w.f32[i] = (float)(int32_t)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24) / (float)2147483647;
            }
            return true;
        }
        off = body + sz + (sz & 1);
    }
    err = "no data chunk";
    return false;
}

// MfccWavFileExtractor::compute_mfccs (wav_file_extractor.rs:18-68)
static bool wav_mfccs(const uint8_t* b, size_t len, uint16_t mfcc_size, Mat& out, float& rms_level, std::string& err) {
    WavData w;
    if (!parse_wav(b, len, w, err)) return false;
    if (w.sample_rate != SAMPLE_RATE) { err = "oracle: resampler (rubato) not restated; wav sample rate must be 16000"; return false; }
    const size_t in_frame = (SAMPLE_RATE * FRAME_LENGTH_MS / 1000) * w.channels;  // encoder.rs:68-69
    MfccExtractor ex(SAMPLE_RATE, 480, 160, (uint16_t)(mfcc_size + 1), PRE_EMPHASIS);
    Vec rms_levels;
    Vec encoded;
    for (size_t off = 0; off + in_frame <= w.f32.size(); off += in_frame) {  // chunks_exact
        Vec mono;
        for (size_t i = 0; i < in_frame; i += w.channels) mono.push_back(w.f32[off + i]);
        rms_levels.push_back(GainNormalizerFilter::get_rms_level(mono));
        encoded.insert(encoded.end(), mono.begin(), mono.end());
    }
    if (!rms_levels.empty()) {
        std::sort(rms_levels.begin(), rms_levels.end(), total_less);
        rms_level = rms_levels[rms_levels.size() / 2];
    }
    Mat frames;
    for (size_t off = 0; off + 480 <= encoded.size(); off += 480) {
        Mat part = ex.compute(encoded.data() + off, 480);
        for (auto& f : part) frames.push_back(std::move(f));
    }
    out = normalize(frames);
    return true;
}

// MfccAverager::average (src/mfcc/averager.rs:5-37) over compute_avg_samples_features' ordering
// (wakeword_ref_build.rs:90-110): longest first, ties by name.
static std::optional<Mat> compute_avg_samples_features(const std::vector<std::pair<std::string, Mat>>& templates) {
    if (templates.size() <= 1) return std::nullopt;
    std::vector<const std::pair<std::string, Mat>*> order;
    for (auto& t : templates) order.push_back(&t);
    std::stable_sort(order.begin(), order.end(), [](auto* a, auto* b) {
        if (a->second.size() != b->second.size()) return a->second.size() > b->second.size();
        return a->first < b->first;
    });
    Mat origin = order[0]->second;
    for (size_t k = 1; k < order.size(); k++) {
        const Mat& frames = order[k]->second;
        DtwFull dtw;
        dtw.compute(origin, frames);
        std::vector<std::vector<Vec>> avgs(origin.size());
        for (size_t x = 0; x < origin.size(); x++)
            for (float y : origin[x]) avgs[x].push_back(Vec{y});
        for (auto& xy : dtw.path())
            for (size_t idx = 0; idx < frames[xy.second].size(); idx++) avgs[xy.first][idx].push_back(frames[xy.second][idx]);
        for (size_t x = 0; x < origin.size(); x++)
            for (size_t idx = 0; idx < origin[x].size(); idx++) {
                float sum = 0.f;
                for (float v : avgs[x][idx]) sum += v;
                origin[x][idx] = sum / (float)avgs[x][idx].size();
            }
    }
    return origin;
}

static Mat to_mat(const float* p, int rows, int d) {
    Mat m(rows, Vec(d));
    for (int i = 0; i < rows; i++) std::memcpy(m[i].data(), p + (size_t)i * d, sizeof(float) * d);
    return m;
}
static void fill_detection(const Detection& d, rpo_detection* out) {
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->name, RPO_NAME_MAX, "%s", d.name.c_str());
    out->avg_score = d.avg_score;
    out->score = d.score;
    out->counter = d.counter;
    out->gain = d.gain;
    out->n_scores = (uint32_t)std::min<size_t>(d.scores.size(), RPO_MAX_SCORES);
    for (uint32_t i = 0; i < out->n_scores; i++) {
        std::snprintf(out->score_names[i], RPO_NAME_MAX, "%s", d.scores[i].first.c_str());
        out->score_values[i] = d.scores[i].second;
    }
}
static void set_err(char* err, size_t n, const std::string& s) {
    if (err && n) std::snprintf(err, n, "%s", s.c_str());
}

}  // namespace rpo

using namespace rpo;

struct rpo_wakeword { WakewordRef w; };
struct rpo_detector { Rustpotter r; rpo_config cfg; explicit rpo_detector(const rpo_config& c) : r(c), cfg(c) {} };

extern "C" {

void rpo_config_default(rpo_config* c) {  // config.rs:20-29,43-52,63-71,193-208
    std::memset(c, 0, sizeof(*c));
    c->sample_rate = 16000; c->sample_format = 3; c->channels = 1; c->endianness = 0;
    c->avg_threshold = 0.2f; c->threshold = 0.5f; c->min_scores = 5; c->eager = 0; c->score_ref = 0.22f;
    c->band_size = 5; c->score_mode = Max; c->vad_mode = -1;
    c->gain_normalizer_enabled = 0; c->gain_ref_set = 0; c->gain_ref = 0.f; c->min_gain = 0.1f; c->max_gain = 1.0f;
    c->band_pass_enabled = 0; c->low_cutoff = 80.f; c->high_cutoff = 400.f;
}

size_t rpo_mfcc_num_frames(size_t n_samples) {
    size_t hops = n_samples / 160;
    return hops > 3 ? hops - 3 : 0;
}
size_t rpo_mfcc_stream(const float* audio, size_t n_samples, int mfcc_size, float* out) {
    MfccExtractor ex(SAMPLE_RATE, 480, 160, 0, PRE_EMPHASIS);
    ex.set_out_size((uint16_t)mfcc_size);
    Mat frames = ex.compute(audio, n_samples);
    for (size_t i = 0; i < frames.size(); i++) std::memcpy(out + i * mfcc_size, frames[i].data(), sizeof(float) * mfcc_size);
    return frames.size();
}
void rpo_mfcc_frame(const float* s, int mfcc_size, float* out) {
    MfccExtractor ex(SAMPLE_RATE, 480, 160, 0, PRE_EMPHASIS);
    ex.set_out_size((uint16_t)mfcc_size);
    Vec f = ex.extract_mfccs(s);
    std::memcpy(out, f.data(), sizeof(float) * mfcc_size);
}
void rpo_mel_centres(int mfcc_size, int* out) {
    auto c = MfccExtractor::centre_indices(SAMPLE_RATE, 240, (size_t)mfcc_size + 1, 0, SAMPLE_RATE / 2);
    for (size_t i = 0; i < c.size(); i++) out[i] = (int)c[i];
}
void rpo_hamming(float* out) {
    Vec w = MfccExtractor::new_hamming_window(480);
    std::memcpy(out, w.data(), sizeof(float) * 480);
}
void rpo_mel_bank(int mfcc_size, float* out) {
    Mat fb = MfccExtractor::new_mel_filter_bank(SAMPLE_RATE, 240, (size_t)mfcc_size + 1, 0, SAMPLE_RATE / 2);
    for (size_t i = 0; i < fb.size(); i++) std::memcpy(out + i * 240, fb[i].data(), sizeof(float) * 240);
}

float rpo_dtw_cost(const float* a, int m, const float* b, int n, int d, int band) {
    return dtw_with_window(to_mat(a, m, d), to_mat(b, n, d), (uint16_t)band);
}
float rpo_compare(const float* a, int m, const float* b, int n, int d, int band, float score_ref) {
    return MfccComparator{score_ref, (uint16_t)band}.compare(to_mat(a, m, d), to_mat(b, n, d));
}
void rpo_normalize(float* frames, int n, int d) {
    Mat out = normalize(to_mat(frames, n, d));
    for (int i = 0; i < n; i++) std::memcpy(frames + (size_t)i * d, out[i].data(), sizeof(float) * d);
}
void rpo_compare_pairs(const float* a, const int64_t* a_off, const int32_t* a_len, const float* b, const int64_t* b_off,
                       const int32_t* b_len, int64_t n_pairs, int d, int band, float score_ref, int cmn, float* out,
                       int n_threads) {
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        MfccComparator cmp{score_ref, (uint16_t)band};
        for (int64_t p = lo; p < hi; p++) {
            Mat A = to_mat(a + a_off[p], a_len[p], d);
            Mat B = to_mat(b + b_off[p], b_len[p], d);
            if (cmn) B = normalize(B);
            out[p] = cmp.compare(A, B);
        }
    };
    if (n_threads == 1) { work(0, n_pairs); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, n_pairs * t / n_threads, n_pairs * (t + 1) / n_threads);
    for (auto& t : th) t.join();
}
float rpo_aggregate(const float* scores, int n, int score_mode) { return aggregate(Vec(scores, scores + n), score_mode); }

rpo_wakeword* rpo_wakeword_load(const uint8_t* buf, size_t len, char* err, size_t err_len) {
    auto w = std::make_unique<rpo_wakeword>();
    std::string e;
    if (!parse_wakeword(buf, len, w->w, e)) { set_err(err, err_len, e); return nullptr; }
    return w.release();
}
void rpo_wakeword_free(rpo_wakeword* w) { delete w; }
const char* rpo_wakeword_name(const rpo_wakeword* w) { return w->w.name.c_str(); }
int rpo_wakeword_mfcc_size(const rpo_wakeword* w) { return w->w.mfcc_size; }
int rpo_wakeword_num_templates(const rpo_wakeword* w) { return (int)w->w.samples_features.size(); }
const char* rpo_wakeword_template_name(const rpo_wakeword* w, int t) { return w->w.samples_features[t].first.c_str(); }
int rpo_wakeword_template_frames(const rpo_wakeword* w, int t) {
    if (t < 0) return w->w.avg_features ? (int)w->w.avg_features->size() : 0;
    return (int)w->w.samples_features[t].second.size();
}
const float* rpo_wakeword_template_data(const rpo_wakeword* cw, int t) {
    rpo_wakeword* w = const_cast<rpo_wakeword*>(cw);
    const Mat* m = t < 0 ? (w->w.avg_features ? &*w->w.avg_features : nullptr) : &w->w.samples_features[t].second;
    if (!m) return nullptr;
    Vec flat;
    for (auto& r : *m) flat.insert(flat.end(), r.begin(), r.end());
    w->w.flat_cache.push_back(std::move(flat));
    return w->w.flat_cache.back().data();
}
float rpo_wakeword_rms_level(const rpo_wakeword* w) { return w->w.rms_level; }
int rpo_wakeword_threshold(const rpo_wakeword* w, float* t) { if (w->w.threshold) { *t = *w->w.threshold; return 1; } return 0; }
int rpo_wakeword_avg_threshold(const rpo_wakeword* w, float* t) { if (w->w.avg_threshold) { *t = *w->w.avg_threshold; return 1; } return 0; }

size_t rpo_wakeword_encode(const char* name, int mfcc_size, int n_templates, const char* const* names, const int32_t* frames,
                           const float* const* data, int avg_frames, const float* avg_data, float rms_level, int has_thr,
                           float thr, int has_avg_thr, float avg_thr, int v2, uint8_t* out, size_t out_cap) {
    // field order = struct field order (wakeword_ref.rs:12-20 / wakeword_v2.rs:8-16)
    CborWriter w;
    w.head(5, 7);
    w.text("name"); w.text(name);
    w.text("avg_features");
    if (avg_frames > 0) w.matrix(avg_data, avg_frames, mfcc_size); else w.null();
    w.text("samples_features");
    w.head(5, n_templates);
    for (int t = 0; t < n_templates; t++) { w.text(names[t]); w.matrix(data[t], frames[t], mfcc_size); }
    w.text("threshold"); if (has_thr) w.f32(thr); else w.null();
    w.text("avg_threshold"); if (has_avg_thr) w.f32(avg_thr); else w.null();
    w.text("rms_level"); w.f32(rms_level);
    if (v2) { w.text("enabled"); w.b.push_back(0xf5); }
    else { w.text("mfcc_size"); w.head(0, mfcc_size); }
    if (out && out_cap >= w.b.size()) std::memcpy(out, w.b.data(), w.b.size());
    return w.b.size();
}

size_t rpo_wakeword_build(const char* name, int has_thr, float thr, int has_avg_thr, float avg_thr, int n_samples,
                          const char* const* sample_names, const uint8_t* const* wavs, const size_t* wav_lens, int mfcc_size,
                          int from_files, uint8_t* out, size_t out_cap, char* err, size_t err_len) {
    // WakewordRefBuildFrom{Files,Buffers} (wakeword_ref_build.rs:9-88)
    std::vector<std::pair<std::string, Mat>> feats;
    Vec rms_levels;
    float rms_max = 0.f;
    for (int i = 0; i < n_samples; i++) {
        Mat m;
        float r = 0.f;
        std::string e;
        if (!wav_mfccs(wavs[i], wav_lens[i], (uint16_t)mfcc_size, m, r, e)) { set_err(err, err_len, e); return 0; }
        if (m.empty()) { set_err(err, err_len, "sample too short"); return 0; }
        feats.emplace_back(sample_names[i], std::move(m));
        rms_levels.push_back(r);
        if (r > rms_max) rms_max = r;
    }
    if (feats.empty()) { set_err(err, err_len, "Can not create an empty wakeword"); return 0; }
    float rms_level = rms_max;
    if (from_files) {
        std::sort(rms_levels.begin(), rms_levels.end(), total_less);
        rms_level = rms_levels[rms_levels.size() / 2];
    }
    std::optional<Mat> avg = compute_avg_samples_features(feats);
    const int d = (int)feats[0].second[0].size();
    std::vector<Vec> flat;
    std::vector<const char*> names;
    std::vector<int32_t> frames;
    std::vector<const float*> data;
    for (auto& kv : feats) {
        Vec f;
        for (auto& row : kv.second) f.insert(f.end(), row.begin(), row.end());
        flat.push_back(std::move(f));
        names.push_back(kv.first.c_str());
        frames.push_back((int32_t)kv.second.size());
    }
    for (auto& f : flat) data.push_back(f.data());
    Vec avg_flat;
    if (avg) for (auto& row : *avg) avg_flat.insert(avg_flat.end(), row.begin(), row.end());
    return rpo_wakeword_encode(name, d, (int)feats.size(), names.data(), frames.data(), data.data(), avg ? (int)avg->size() : 0,
                               avg ? avg_flat.data() : nullptr, rms_level, has_thr, thr, has_avg_thr, avg_thr, 0, out, out_cap);
}

// The resampling stage alone (encoder.rs:52-60): whole chunks of mono f32 at fs_in -> 16 kHz through a fresh FftFixedInOut.
int64_t rpo_resample_to_16k(uint32_t fs_in, const float* in, size_t n_in, float* out, size_t out_cap, size_t* in_chunk) {
    FftFixedInOut r(fs_in, SAMPLE_RATE, SAMPLE_RATE * FRAME_LENGTH_MS / 1000);
    if (in_chunk) *in_chunk = r.input_frames_next();
    const size_t calls = n_in / r.fft_size_in;
    if (!out) return (int64_t)(calls * r.fft_size_out);
    if (calls * r.fft_size_out > out_cap) return -1;
    for (size_t c = 0; c < calls; c++) {
        Vec o = r.process(Vec(in + c * r.fft_size_in, in + (c + 1) * r.fft_size_in));
        std::copy(o.begin(), o.end(), out + c * r.fft_size_out);
    }
    return (int64_t)(calls * r.fft_size_out);
}

rpo_detector* rpo_detector_new(const rpo_config* cfg, char* err, size_t err_len) {
    if (cfg->sample_rate == 0) { set_err(err, err_len, "bad sample rate"); return nullptr; }
    if (cfg->sample_format > 3 || cfg->channels == 0) { set_err(err, err_len, "bad audio format"); return nullptr; }
    return new rpo_detector(*cfg);
}
void rpo_detector_free(rpo_detector* d) { delete d; }
int rpo_detector_add_wakeword_from_buffer(rpo_detector* d, const char* key, const uint8_t* buf, size_t len, char* err,
                                          size_t err_len) {
    WakewordRef w;
    std::string e;
    if (!parse_wakeword(buf, len, w, e)) { set_err(err, err_len, e); return -1; }
    if (!d->r.add_wakeword(key, std::move(w), e)) { set_err(err, err_len, e); return -2; }
    return 0;
}
int rpo_detector_remove_wakeword(rpo_detector* d, const char* key) { return d->r.remove_wakeword(key) ? 1 : 0; }
int rpo_detector_remove_wakewords(rpo_detector* d) { return d->r.remove_wakewords() ? 1 : 0; }
size_t rpo_detector_samples_per_frame(const rpo_detector* d) { return d->r.wav_encoder.input_samples_per_frame; }
size_t rpo_detector_bytes_per_frame(const rpo_detector* d) { return d->r.wav_encoder.input_byte_length(); }
static int emit(const std::optional<Detection>& d, rpo_detection* out) {
    if (!d) return 0;
    if (out) fill_detection(*d, out);
    return 1;
}
int rpo_detector_process_bytes(rpo_detector* d, const uint8_t* b, size_t len, rpo_detection* out) { return emit(d->r.process_bytes(b, len), out); }
int rpo_detector_process_f32(rpo_detector* d, const float* s, size_t n, rpo_detection* out) { return emit(d->r.process_samples(s, n, 0.f), out); }
int rpo_detector_process_i16(rpo_detector* d, const int16_t* s, size_t n, rpo_detection* out) { return emit(d->r.process_samples(s, n, 32767.f), out); }
int rpo_detector_process_i32(rpo_detector* d, const int32_t* s, size_t n, rpo_detection* out) { return emit(d->r.process_samples(s, n, (float)2147483647), out); }
int rpo_detector_process_i8(rpo_detector* d, const int8_t* s, size_t n, rpo_detection* out) { return emit(d->r.process_samples(s, n, 127.f), out); }
int rpo_detector_get_partial(const rpo_detector* d, rpo_detection* out) { return emit(d->r.partial_detection, out); }
float rpo_detector_rms_level(const rpo_detector* d) { return d->r.rms_level; }
float rpo_detector_gain(const rpo_detector* d) { return d->r.gain; }
float rpo_detector_rms_level_ref(const rpo_detector* d) {
    return d->r.gain_normalizer_filter ? d->r.gain_normalizer_filter->rms_level_ref : std::numeric_limits<float>::quiet_NaN();
}
void rpo_detector_update_config(rpo_detector* d, const rpo_config* cfg) {  // detector.rs:259-262
    d->cfg = *cfg;
    d->r.update_detector_config(*cfg);
    d->r.update_filters_config(*cfg);
}
void rpo_detector_reset(rpo_detector* d) { d->r.reset(); }
uint64_t rpo_detector_windows_scored(const rpo_detector* d) { return d->r.windows_scored; }

size_t rpo_trace_window_scores(const rpo_config* cfg, const uint8_t* rpw, size_t rpw_len, const float* audio,
                               size_t n_samples, float* out, size_t max_windows) {
    rpo_config c = *cfg;
    c.sample_format = 3; c.channels = 1; c.sample_rate = 16000;
    Rustpotter r(c);
    WakewordRef w;
    std::string e;
    if (!parse_wakeword(rpw, rpw_len, w, e)) return 0;
    size_t T = w.samples_features.size();
    if (!r.add_wakeword("w", std::move(w), e)) return 0;
    std::vector<Vec> trace;
    r.trace = &trace;
    // never let a detection fire/reset: thresholds > 1 (scores are <= 1/(1+e^-1))
    r.threshold = 2.f;
    r.avg_threshold = 0.f;
    for (auto& kv : r.wakewords) { kv.second.ww.threshold.reset(); kv.second.ww.avg_threshold.reset(); }
    for (size_t off = 0; off + 480 <= n_samples; off += 480) r.process_samples(audio + off, 480, 0.f);
    size_t n = std::min(trace.size(), max_windows);
    for (size_t i = 0; i < n; i++) std::memcpy(out + i * (T + 2), trace[i].data(), sizeof(float) * (T + 2));
    return n;
}

uint64_t rpo_run_streams(const rpo_config* cfg, const uint8_t* const* rpws, const size_t* rpw_lens, int n_rpw,
                         const float* audio, int64_t n_streams, int64_t S, int n_threads, int32_t* det_counts,
                         rpo_detection* dets, int max_det) {
    if (n_threads < 1) n_threads = 1;
    std::vector<uint64_t> scored(n_threads, 0);
    auto work = [&](int t) {
        int64_t lo = n_streams * t / n_threads, hi = n_streams * (t + 1) / n_threads;
        for (int64_t b = lo; b < hi; b++) {
            rpo_config c = *cfg;
            c.sample_format = 3; c.channels = 1; c.sample_rate = 16000;
            Rustpotter r(c);
            for (int k = 0; k < n_rpw; k++) {
                WakewordRef w;
                std::string e;
                if (parse_wakeword(rpws[k], rpw_lens[k], w, e)) r.add_wakeword("w" + std::to_string(k), std::move(w), e);
            }
            int32_t nd = 0;
            const float* a = audio + b * S;
            for (int64_t off = 0; off + 480 <= S; off += 480) {
                auto d = r.process_samples(a + off, 480, 0.f);
                if (d) {
                    if (dets && nd < max_det) fill_detection(*d, &dets[b * max_det + nd]);
                    nd++;
                }
            }
            if (det_counts) det_counts[b] = nd;
            scored[t] += r.windows_scored;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    uint64_t total = 0;
    for (auto v : scored) total += v;
    return total;
}

}  // extern "C"
