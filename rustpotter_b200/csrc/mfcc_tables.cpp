#include "mfcc_tables.h"

#include <algorithm>
#include <cmath>

#include "rp_internal.h"

namespace rp {

static constexpr float kPiF = 3.14159265358979323846f;  // std::f32::consts::PI

MfccTables build_mfcc_tables(int mfcc_size) {
    MfccTables t;
    const int C = mfcc_size + 1;  // set_out_size: first coefficient is dropped (extractor.rs:47-48)
    t.num_coefficients = C;

    // new_hamming_window (extractor.rs:115-120): 0.54 - 0.46*cos(2*pi*(s/479)), all f32
    t.hamming.resize(kFrameSamples);
    for (int s = 0; s < kFrameSamples; s++)
        t.hamming[s] = 0.54f - (0.46f * std::cos(2.f * kPiF * ((float)s / (float)(kFrameSamples - 1))));

    // forward DFT twiddles (the reference delegates to rustfft, which derives them in f64)
    t.tw480.resize(2 * kFrameSamples);
    for (int k = 0; k < kFrameSamples; k++) {
        double a = 2.0 * 3.14159265358979323846264338327950288 * (double)k / (double)kFrameSamples;
        t.tw480[2 * k] = (float)std::cos(a);
        t.tw480[2 * k + 1] = (float)-std::sin(a);
    }

    // new_mel_filter_bank (extractor.rs:164-198)
    const float nyquist = (float)kSampleRate / 2.f;
    const float max_mel = std::floor(1127.f * std::log(1.f + ((float)(kSampleRate / 2) / 700.0f)));
    const float min_mel = std::floor(1127.f * std::log(1.f + (0.f / 700.0f)));
    t.centres.resize(C + 2);
    for (int i = 0; i < C + 2; i++) {
        float f = (float)i * (max_mel - min_mel) / (float)(C + 1) + min_mel;
        float tmp = std::log(1.f + 1000.0f / 700.0f) / 1000.0f;
        tmp = (std::exp(f * tmp) - 1.f) / nyquist;
        t.centres[i] = (int)std::floor(0.5f + 700.f * (float)kSpectrumBins * tmp);
    }
    t.mel_bank.assign((size_t)C * kSpectrumBins, 0.f);
    for (int i = 0; i < C; i++) {
        int b = t.centres[i], c = t.centres[i + 1], e = t.centres[i + 2];
        for (int k = b; k < c && k < kSpectrumBins; k++) t.mel_bank[(size_t)i * kSpectrumBins + k] = (float)(k - b) / (float)(c - b);
        for (int k = c; k < e && k < kSpectrumBins; k++) t.mel_bank[(size_t)i * kSpectrumBins + k] = (float)(e - k) / (float)(e - c);
    }

    // discrete_cosine_transform (extractor.rs:146-163): cos(pi_over_n * (n + 0.5) * k)
    t.dct.resize((size_t)C * C);
    const float pi_over_n = kPiF / (float)C;
    for (int k = 0; k < C; k++)
        for (int n = 0; n < C; n++) t.dct[(size_t)k * C + n] = std::cos(pi_over_n * ((float)n + 0.5f) * (float)k);
    // ---- segment form of the mel bank for the v2 kernel
    t.up_weight.assign(kSpectrumBins, 0.f);
    const int n_seg = C + 1;
    std::vector<int> seg_lo(n_seg), seg_hi(n_seg);
    for (int j = 0; j < n_seg; j++) {
        seg_lo[j] = std::min(t.centres[j], kSpectrumBins);
        seg_hi[j] = std::min(t.centres[j + 1], kSpectrumBins);
        for (int k = seg_lo[j]; k < seg_hi[j]; k++) t.up_weight[k] = (float)(k - t.centres[j]) / (float)(t.centres[j + 1] - t.centres[j]);
    }
    t.chunks.assign(32 * 4, 0);
    t.seg_chunks.assign((size_t)n_seg * 2, 0);
    t.n_chunks = 0;
    if (n_seg > 32) return t;  // one chunk per lane at least: the two-frames-per-warp kernel needs mfcc_size <= 16 anyway
    int max_len = 1;
    for (;; max_len++) {  // smallest chunk length that fits all segments into 32 chunks
        int total = 0;
        for (int j = 0; j < n_seg; j++) total += std::max(1, (seg_hi[j] - seg_lo[j] + max_len - 1) / max_len);
        if (total <= 32) break;
    }
    int c = 0;
    for (int j = 0; j < n_seg; j++) {
        const int len = seg_hi[j] - seg_lo[j];
        const int parts = std::max(1, (len + max_len - 1) / max_len);
        t.seg_chunks[2 * j] = c;
        t.seg_chunks[2 * j + 1] = parts;
        for (int q = 0; q < parts; q++, c++) {
            t.chunks[4 * c] = j;
            t.chunks[4 * c + 1] = seg_lo[j] + (int)((long long)len * q / parts);
            t.chunks[4 * c + 2] = seg_lo[j] + (int)((long long)len * (q + 1) / parts);
        }
    }
    t.n_chunks = c;
    return t;
}

}  // namespace rp
