// Which kernel scores a launch of independent (template, window) pairs (rp_dtw_scores).
//
//   v4 (dtw_stream4_kernel.cu)  d == 16, uniform lengths, no CMN, window = max(band, |m-n|) in 3..20, <= 238 steps:
//                               warps of a CTA as the systolic array, producer/consumer warpgroups
//   generic (dtw_kernel.cu)     everything else (other widths, CMN, ragged pairs, windows below 3 or above 20, templates of
//                               more than ~470 rows), reference operation order
//
// The earlier streaming kernels (round 1's row-per-step kernels, and v3 with the lanes of a pair as the systolic array) are
// retired: one tuned kernel for the shapes the detector's templates have, one reference-order kernel for everything else.
#include "kernels.h"

namespace rp {

bool dtw_pairs_stream4_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream4(const DtwPairsArgs& a, cudaStream_t stream);

bool dtw_pairs_stream_supported(const DtwPairsArgs& a) { return dtw_pairs_stream4_supported(a); }

cudaError_t launch_dtw_pairs_stream(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    if (dtw_pairs_stream4_supported(a)) return launch_dtw_pairs_stream4(a, stream);
    return cudaErrorInvalidValue;
}

}  // namespace rp
