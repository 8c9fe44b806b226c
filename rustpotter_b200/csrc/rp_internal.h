// Internal declarations shared by the host-side sources of librustpotter_b200.so.
#pragma once

#include <cstddef>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rustpotter_b200.h"

namespace rp {

// reference src/constants.rs:1-11
constexpr int kSampleRate = 16000;
constexpr int kFrameSamples = 480;  // 30 ms
constexpr int kHopSamples = 160;    // 10 ms
constexpr int kHopsPerChunk = 3;
constexpr int kSpectrumBins = 240;
constexpr float kPreEmphasis = 0.97f;
constexpr int kMaxTemplates = 64;   // per wakeword (the reference recommends 3..8)
constexpr int kMaxMfccSize = 31;    // num_coefficients = mfcc_size + 1 <= 32 lanes

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// A row-major frames x mfcc_size matrix (Vec<Vec<f32>> in the reference).
struct FrameMatrix {
    int rows = 0, cols = 0;
    std::vector<float> v;
};

// WakewordRef (src/wakewords/wakeword_ref.rs:12-20); also what WakewordV2 converts into
// (wakeword_v2.rs:18-30). Templates keep file order (the reference uses a HashMap).
struct WakewordRefData {
    std::string name;
    std::optional<FrameMatrix> avg_features;
    std::vector<std::pair<std::string, FrameMatrix>> samples_features;
    std::optional<float> threshold, avg_threshold;
    float rms_level = 0.f;
    int mfcc_size = 0;
    bool is_v2 = false;
    int max_frames() const {
        int mx = 0;
        for (auto& t : samples_features) mx = t.second.rows > mx ? t.second.rows : mx;
        return mx;
    }
};

// rpw_reader.cpp — throws rp::Error(RP_ERR_FORMAT / RP_ERR_UNSUPPORTED)
WakewordRefData parse_rpw(const uint8_t* buf, size_t len);
std::vector<uint8_t> read_file(const std::string& path);

void set_thread_error(const std::string& msg);
const char* thread_error();

}  // namespace rp
