// Sample decoding on the device: the batched front-end's `process_bytes` / `process_samples<T>` ingest.
//
// Replaces, for a whole batch of streams, Sample::into_f32 (reference src/audio/audio_types.rs:98-137:
// i8 -> v/127, i16 -> v/32767, i32 -> v/2147483647 (as f32), f32 -> v), the byte decoding of
// encode_audio_bytes (src/audio/encoder.rs:105-115, little/big/native endian) and the channel-0 pick of
// reencode_to_mono_with_sample_rate (encoder.rs:41-48). Every operation is exact in f32 (one int -> float
// conversion, one IEEE division), so the result equals the host conversion bit for bit.
//
// HBM-bound byte work: each thread handles four consecutive mono samples of one stream (one 8-byte load for
// mono i16, one 16-byte store), streams along blockIdx.y, grid sized from the sample count.
#include <algorithm>
#include <cstdint>

#include "kernels.h"
#include "rp_internal.h"

namespace rp {
namespace {

template <int BYTES>
__device__ __forceinline__ uint32_t load_sample(const uint8_t* p, bool big) {
    uint32_t u = 0;
#pragma unroll
    for (int k = 0; k < BYTES; k++) u |= (uint32_t)p[big ? BYTES - 1 - k : k] << (8 * k);
    return u;
}

__device__ __forceinline__ float to_f32(uint32_t u, int fmt) {
    switch (fmt) {
        case RP_FMT_I8: return __fdiv_rn((float)(int8_t)u, 127.f);
        case RP_FMT_I16: return __fdiv_rn((float)(int16_t)u, 32767.f);
        case RP_FMT_I32: return __fdiv_rn((float)(int32_t)u, 2147483648.f);   // i32::MAX as f32 rounds to 2^31
        default: return __uint_as_float(u);
    }
}

template <int BYTES>
__global__ void __launch_bounds__(256) decode_samples_kernel(const uint8_t* __restrict__ in, int64_t in_stride, int fmt, int channels,
                                                             int big, float* __restrict__ out, int64_t out_stride, int64_t samples) {
    const int64_t b = blockIdx.y;
    const uint8_t* src = in + b * in_stride;
    float* dst = out + b * out_stride;
    const int64_t quads = (samples + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = q * 4;
        float v[4];
        if (BYTES == 2 && channels == 1 && !big && i0 + 4 <= samples && ((reinterpret_cast<uintptr_t>(src) & 7) == 0)) {
            const uint2 w = __ldg(reinterpret_cast<const uint2*>(src) + q);   // four little-endian i16
            v[0] = to_f32(w.x & 0xffffu, fmt);
            v[1] = to_f32(w.x >> 16, fmt);
            v[2] = to_f32(w.y & 0xffffu, fmt);
            v[3] = to_f32(w.y >> 16, fmt);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t i = i0 + k;
                v[k] = i < samples ? to_f32(load_sample<BYTES>(src + (size_t)i * channels * BYTES, big != 0), fmt) : 0.f;
            }
        }
        if (i0 + 4 <= samples && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            reinterpret_cast<float4*>(dst)[q] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int k = 0; k < 4 && i0 + k < samples; k++) dst[i0 + k] = v[k];
        }
    }
}

}  // namespace

cudaError_t launch_decode_samples(const void* in, int64_t in_stride_bytes, int fmt, int channels, int big_endian, float* out,
                                  int64_t out_stride, int64_t n_streams, int64_t samples_mono, cudaStream_t stream) {
    if (n_streams <= 0 || samples_mono <= 0) return cudaSuccess;
    if (channels < 1 || fmt < RP_FMT_I8 || fmt > RP_FMT_F32) return cudaErrorInvalidValue;
    const int64_t quads = (samples_mono + 3) / 4;
    const unsigned bx = (unsigned)std::min<int64_t>((quads + 255) / 256, 1024);
    const uint8_t* p = static_cast<const uint8_t*>(in);
    for (int64_t b0 = 0; b0 < n_streams; b0 += 65535) {   // gridDim.y limit
        const dim3 grid(bx, (unsigned)std::min<int64_t>(65535, n_streams - b0));
        const uint8_t* pin = p + b0 * in_stride_bytes;
        float* pout = out + b0 * out_stride;
        if (fmt == RP_FMT_I8) decode_samples_kernel<1><<<grid, 256, 0, stream>>>(pin, in_stride_bytes, fmt, channels, big_endian, pout, out_stride, samples_mono);
        else if (fmt == RP_FMT_I16) decode_samples_kernel<2><<<grid, 256, 0, stream>>>(pin, in_stride_bytes, fmt, channels, big_endian, pout, out_stride, samples_mono);
        else decode_samples_kernel<4><<<grid, 256, 0, stream>>>(pin, in_stride_bytes, fmt, channels, big_endian, pout, out_stride, samples_mono);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace rp
