// Host-side audio ingest in front of the MFCC kernel for the per-stream handle: sample decoding
// (reference src/audio/encoder.rs:26-50,105-115 and Sample::into_f32, audio_types.rs:98-137) and
// the two optional, sequential-per-stream filters (src/audio/gain_normalizer_filter.rs:14-55,
// src/audio/band_pass_filter.rs:19-55). These are SURVEY §8(f) "next" rows 1-2: scalar per-stream
// recurrences that run before the hot path; they stay on the host here.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <vector>

#include "resampler.h"
#include "rp_internal.h"

namespace rp {

struct AudioIngest {
    uint32_t fmt = RP_FMT_F32, channels = 1, endianness = RP_ENDIAN_LITTLE;
    size_t input_samples_per_frame = kFrameSamples;  // (sample_rate*30/1000) * channels (encoder.rs:68-69)
    std::shared_ptr<FftResampler> resampler;         // encoder.rs:72-79: None when the source rate is 16 kHz
    // AudioEncoder::new (encoder.rs:63-102)
    void configure(const rp_config& c) {
        fmt = c.sample_format;
        channels = c.channels;
        endianness = c.endianness;
        input_samples_per_frame = (size_t)(c.sample_rate * 30 / 1000) * c.channels;
        resampler.reset();
        if (c.sample_rate != (uint32_t)kSampleRate) {
            resampler = std::make_shared<FftResampler>(c.sample_rate, (size_t)kSampleRate, (size_t)kFrameSamples);
            input_samples_per_frame = resampler->input_frames() * c.channels;
        }
    }
    // samples per call that reach the MFCC extractor (480 without a resampler)
    size_t output_samples_per_frame() const { return resampler ? resampler->output_frames() : (size_t)kFrameSamples; }
    // reencode_to_mono_with_sample_rate (encoder.rs:41-62) after the channel pick
    std::vector<float> finish(std::vector<float> mono) const {
        if (!resampler) return mono;
        std::vector<float> out(resampler->output_frames());
        resampler->process(mono.data(), out.data());
        return out;
    }
    size_t bytes_per_sample() const { return fmt == RP_FMT_I8 ? 1 : fmt == RP_FMT_I16 ? 2 : 4; }
    size_t input_bytes_per_frame() const { return input_samples_per_frame * bytes_per_sample(); }

    // keep channel 0 of every interleaved group (encoder.rs:41-48)
    void to_mono(std::vector<float>& v) const {
        if (channels == 1) return;
        size_t n = v.size() / channels;
        for (size_t i = 0; i < n; i++) v[i] = v[i * channels];
        v.resize(n);
    }
    std::vector<float> decode_bytes(const uint8_t* b, size_t len) const { return finish(decode_bytes_mono(b, len)); }
    // encode_audio_bytes + channel pick, before any resampling
    std::vector<float> decode_bytes_mono(const uint8_t* b, size_t len) const {
        const size_t bs = bytes_per_sample();
        const bool big = endianness == RP_ENDIAN_BIG;  // native == little on the platforms this targets
        std::vector<float> out(len / bs);
        for (size_t i = 0; i < out.size(); i++) {
            uint32_t u = 0;
            for (size_t k = 0; k < bs; k++) u |= (uint32_t)b[i * bs + (big ? bs - 1 - k : k)] << (8 * k);
            switch (fmt) {
                case RP_FMT_I8: out[i] = (float)(int8_t)u / 127.f; break;
                case RP_FMT_I16: out[i] = (float)(int16_t)u / 32767.f; break;
                case RP_FMT_I32: out[i] = (float)(int32_t)u / (float)2147483647; break;
                default: std::memcpy(&out[i], &u, 4);
            }
        }
        to_mono(out);
        return out;
    }
    template <typename T>
    std::vector<float> convert(const T* s, size_t n, float max_value) const {
        std::vector<float> out(n);
        for (size_t i = 0; i < n; i++) out[i] = max_value == 0.f ? (float)s[i] : (float)s[i] / max_value;
        to_mono(out);
        return finish(std::move(out));
    }
};

// GainNormalizerFilter (gain_normalizer_filter.rs)
struct GainNormalizer {
    size_t window_size = 1;
    bool fixed_rms_level = false;
    float min_gain = 0.1f, max_gain = 1.f;
    float rms_level_ref = std::numeric_limits<float>::quiet_NaN();
    float rms_level_sqrt = std::numeric_limits<float>::quiet_NaN();
    std::vector<float> rms_level_window;

    GainNormalizer(float mn, float mx, std::optional<float> fixed) : fixed_rms_level(fixed.has_value()), min_gain(mn), max_gain(mx) {
        if (fixed) {
            rms_level_ref = *fixed;
            rms_level_sqrt = std::sqrt(*fixed);
        }
    }
    static float rms_level(const std::vector<float>& signal) {
        float sum_squared = 0.f;
        for (float s : signal) sum_squared += s * s;
        return std::sqrt(sum_squared / (float)signal.size());
    }
    void set_rms_level_ref(float rms, size_t ws) {
        if (!fixed_rms_level) {
            rms_level_ref = rms;
            rms_level_sqrt = std::sqrt(rms);
        }
        window_size = ws != 0 ? ws : 1;
    }
    float filter(std::vector<float>& signal, float rms) {
        if (std::isnan(rms_level_ref) || rms == 0.f) return 1.f;
        rms_level_window.push_back(rms);
        if (rms_level_window.size() > window_size) rms_level_window.erase(rms_level_window.begin());
        float acc = 0.f;
        for (float v : rms_level_window) acc += v;
        const float frame_rms = acc / (float)rms_level_window.size();
        float gain = rms_level_sqrt / std::sqrt(frame_rms);
        gain = std::round(gain * 10.f) / 10.f;
        gain = std::fmin(std::fmax(gain, min_gain), max_gain);
        if (gain != 1.f)
            for (float& x : signal) x = std::fmin(std::fmax(x * gain, -1.f), 1.f);
        return gain;
    }
};

// BandPassFilter (band_pass_filter.rs)
struct BandPass {
    float a0, a1, a2, b1, b2;
    float x1 = 0.f, x2 = 0.f, y1 = 0.f, y2 = 0.f;
    BandPass(float sample_rate, float low_cutoff, float high_cutoff) {
        const float pi = 3.14159265358979323846f;
        const float omega_low = 2.0f * pi * low_cutoff / sample_rate;
        const float omega_high = 2.0f * pi * high_cutoff / sample_rate;
        const float alpha_low = std::sin(omega_low) / 2.0f, alpha_high = std::sin(omega_high) / 2.0f;
        a0 = 1.0f / (1.0f + alpha_high - alpha_low);
        a1 = -2.0f * std::cos(omega_low) * a0;
        a2 = (1.0f - alpha_high - alpha_low) * a0;
        b1 = -2.0f * std::cos(omega_high) * a0;
        b2 = (1.0f - alpha_high + alpha_low) * a0;
    }
    void filter(std::vector<float>& signal) {
        for (float& sample : signal) {
            const float x = sample;
            sample = a0 * x + a1 * x1 + a2 * x2 - b1 * y1 - b2 * y2;
            x2 = x1;
            x1 = x;
            y2 = y1;
            y1 = sample;
        }
    }
};

}  // namespace rp
