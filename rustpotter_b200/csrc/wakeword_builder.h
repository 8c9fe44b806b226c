// Wakeword-reference builder — see wakeword_builder.cpp.
#pragma once

#include <functional>

#include "rp_internal.h"

namespace rp {

// MFCC frames of a fresh extractor run over mono 16 kHz samples (a multiple of 480): [hops - 3][mfcc_size].
// capi.cpp binds this to the K1 kernel; the builder has no CPU path of its own.
using MfccFn = std::function<FrameMatrix(const std::vector<float>& mono, int mfcc_size)>;

// samples: (name, (wav bytes, length)). rms_median: new_from_sample_files semantics, else new_from_sample_buffers.
WakewordRefData build_wakeword_ref(const std::string& name, std::optional<float> threshold, std::optional<float> avg_threshold,
                                   const std::vector<std::pair<std::string, std::pair<const uint8_t*, size_t>>>& samples,
                                   int mfcc_size, bool rms_median, const MfccFn& mfcc);

// WakewordRef::compute_avg_samples_features (wakeword_ref_build.rs:93-110): None for a single template
std::optional<FrameMatrix> average_templates(const std::vector<std::pair<std::string, FrameMatrix>>& templates);

// WakewordSave::save_to_buffer for WakewordRef (ciborium layout)
std::vector<uint8_t> encode_wakeword_ref(const WakewordRefData& w);

}  // namespace rp
