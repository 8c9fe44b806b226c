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
};

MfccTables build_mfcc_tables(int mfcc_size);

}  // namespace rp
