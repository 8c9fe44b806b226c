// Which kernel scores a launch of independent (template, window) pairs (rp_dtw_scores).
//
//   v4 (dtw_stream4_kernel.cu)  d == 16, uniform lengths, no CMN, window = max(band, |m-n|) in 3..20, <= 238 steps:
//                               warps of a CTA as the systolic array, producer/consumer warpgroups
//   v3 (dtw_stream3_kernel.cu)  the same arithmetic with the lanes of a pair as the systolic array: windows 1..20, any length
//   generic (dtw_kernel.cu)     everything else (other widths, CMN, ragged pairs, wider windows), reference operation order
//
// The round-1 row-per-step streaming kernels (windows up to 23) are retired: v3/v4 cover their shapes up to window 20,
// windows 21..23 take the generic kernel.
#include "kernels.h"

namespace rp {

bool dtw_pairs_stream3_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream3(const DtwPairsArgs& a, cudaStream_t stream);
bool dtw_pairs_stream4_supported(const DtwPairsArgs& a);
cudaError_t launch_dtw_pairs_stream4(const DtwPairsArgs& a, cudaStream_t stream);

namespace {
int g_stream_kernel = 0;   // 0 = newest kernel that takes the shape, 3 = v3 even where v4 applies (A/B, parity tests; debug only)
}
void set_dtw_stream_rows(int v) { g_stream_kernel = v; }

bool dtw_pairs_stream_supported(const DtwPairsArgs& a) {
    return (g_stream_kernel == 0 && dtw_pairs_stream4_supported(a)) || dtw_pairs_stream3_supported(a);
}

cudaError_t launch_dtw_pairs_stream(const DtwPairsArgs& a, cudaStream_t stream) {
    if (a.n_pairs <= 0) return cudaSuccess;
    if (g_stream_kernel == 0 && dtw_pairs_stream4_supported(a)) return launch_dtw_pairs_stream4(a, stream);
    if (dtw_pairs_stream3_supported(a)) return launch_dtw_pairs_stream3(a, stream);
    return cudaErrorInvalidValue;
}

}  // namespace rp
