// Host-built constant tables of the MFCC front-end, following the reference formulas
// (src/mfcc/extractor.rs:115-120 Hamming, :164-198 mel bank, :146-163 DCT) in f32.
#pragma once
#include <vector>

namespace rp {

struct MfccTables {
    int num_coefficients = 0;          // C = mfcc_size + 1 filters / DCT points
    std::vector<float> hamming;        // [480]
    std::vector<float> tw480;          // [480][2]: exp(-2*pi*i*k/480) = (cos, -sin), from double
    std::vector<int> centres;          // [C + 2] filter centre bins
    std::vector<float> mel_bank;       // [C][240] dense triangular weights
    std::vector<float> dct;            // [C][C]: cos((pi/C) * (n + 0.5) * k), row k
    // v2 kernel: mel energies from per-segment sums. Segment j = bins [centres[j], centres[j+1]);
    // up_weight[k] = filter j's rising slope at bin k of segment j; the falling slope of filter j-1 on the same
    // segment is 1 - up_weight[k]. Segments are cut into <= 32 chunks of near-equal length (one per lane).
    std::vector<float> up_weight;      // [240]
    std::vector<int> chunks;           // [32][4]: segment, first bin, end bin, unused
    std::vector<int> seg_chunks;       // [C+1][2]: first chunk, chunk count
    int n_chunks = 0;
};

MfccTables build_mfcc_tables(int mfcc_size);

}  // namespace rp
