// Source-rate conversion in front of the hot path: what the reference does with rubato 0.14.1's
// `FftFixedInOut<f32>` when AudioFmt::sample_rate is not 16 kHz (src/audio/encoder.rs:52-60,63-102).
//
// Algorithm (rubato synchro.rs / sinc.rs / windows.rs, one channel): chunks of fft_size_in samples are zero-padded to
// 2*fft_size_in, transformed, multiplied with the spectrum of a Blackman-Harris^2 windowed sinc of fft_size_in taps,
// cut (or extended) to the fft_size_out + 1 bins of a 2*fft_size_out-point spectrum, transformed back and overlap-added;
// fft_size_in/out = ceil(chunk / (fs_out/g)) * fs_in/g resp. fs_out/g with g = gcd(fs_in, fs_out).
// This is host code: a sequential per-stream filter in front of K1, like the gain normaliser of the per-stream handle.
// The transforms are an iterative mixed-radix Stockham FFT (complex, f32 data, twiddles rounded from f64).
//
// Parity: the reference's 48 kHz goldens (tests/detector.rs:162-214) are reproduced through rp_process_samples_f32 to
// < 1e-5 relative (tests/test_gpu_parity.py). The anti-aliasing cutoff follows the 1/(1 + k/n) law calibrated on those
// goldens (see oracle/rp_oracle.cpp FftFixedInOut); source rates other than 48 kHz use the same law and are unpinned.
#pragma once

#include <cmath>
#include <complex>
#include <cstddef>
#include <memory>
#include <vector>

namespace rp {

class StockhamFft {   // unnormalised forward DFT of n complex points, n = product of small primes
  public:
    explicit StockhamFft(size_t n) : n_(n), a_(n), b_(n) {
        size_t m = n;
        for (size_t p : {4u, 2u, 3u, 5u}) while (m % p == 0) { radix_.push_back(p); m /= p; }
        for (size_t p = 7; m > 1; p += 2) while (m % p == 0) { radix_.push_back(p); m /= p; }
        tw_.resize(n);
        const double pi = 3.141592653589793238462643383279502884;
        for (size_t k = 0; k < n; k++) {
            const double ang = -2.0 * pi * (double)k / (double)n;
            tw_[k] = std::complex<float>((float)std::cos(ang), (float)std::sin(ang));
        }
    }
    size_t size() const { return n_; }
    // in/out: n complex values; forward transform (exp(-2 pi i jk/n)), out may alias in
    void forward(const std::complex<float>* in, std::complex<float>* out) {
        std::complex<float>* x = a_.data();
        std::complex<float>* y = b_.data();
        for (size_t i = 0; i < n_; i++) x[i] = in[i];
        size_t l = 1;          // product of the radices already applied (stride of the outputs)
        size_t m = n_;         // length of the remaining sub-transforms
        for (size_t p : radix_) {
            m /= p;
            // Stockham autosort step (decimation in frequency): the m*p-point transforms become p transforms of m points;
            // y[q + l*(p*k + r)] = W_{mp}^{k r} * sum_j W_p^{j r} x[q + l*(k + m*j)]
            for (size_t k = 0; k < m; k++) {
                for (size_t q = 0; q < l; q++) {
                    // inputs x[q + l*(k + m*j)], j < p
                    for (size_t r = 0; r < p; r++) {
                        std::complex<float> acc(0.f, 0.f);
                        for (size_t j = 0; j < p; j++) {
                            const std::complex<float> v = x[q + l * (k + m * j)];
                            const std::complex<float> w = tw_[((j * r) % p) * (n_ / p)];
                            acc += v * w;
                        }
                        // twiddle between stages: W_{m*p}^{k r}
                        const std::complex<float> t = tw_[((k * r) % (m * p)) * (n_ / (m * p))];
                        y[q + l * (r + p * k)] = acc * t;
                    }
                }
            }
            std::swap(x, y);
            l *= p;
        }
        for (size_t i = 0; i < n_; i++) out[i] = x[i];   // (the autosort indexing leaves the result in natural order)
    }

  private:
    size_t n_;
    std::vector<size_t> radix_;
    std::vector<std::complex<float>> tw_, a_, b_;
};

class FftResampler {
  public:
    FftResampler(size_t fs_in, size_t fs_out, size_t chunk_size_in) {
        const size_t g = gcd(fs_in, fs_out);
        const size_t min_chunk_out = fs_out / g;
        const size_t fft_chunks = (chunk_size_in + min_chunk_out - 1) / min_chunk_out;
        n_out_ = fft_chunks * fs_out / g;
        n_in_ = fft_chunks * fs_in / g;
        fwd_ = std::make_unique<StockhamFft>(2 * n_in_);
        inv_ = std::make_unique<StockhamFft>(2 * n_out_);
        // anti-aliasing low-pass: Blackman-Harris^2 windowed sinc, unit DC gain, scaled by 1 / (2 n_in)
        const float rel = 1.0f / (1.0f + 42.08f / (float)n_in_);
        const float cutoff = n_in_ > n_out_ ? rel * (float)n_out_ / (float)n_in_ : rel;
        const float pi = 3.14159265358979323846f;
        std::vector<float> y(n_in_);
        float sum = 0.f;
        for (size_t x = 0; x < n_in_; x++) {
            const float xf = (float)x, nf = (float)n_in_;
            float w = 0.35875f - 0.48829f * std::cos(2.f * pi * xf / nf) + 0.14128f * std::cos(4.f * pi * xf / nf) -
                      0.01168f * std::cos(6.f * pi * xf / nf);
            w *= w;
            const float v = ((float)x - (float)(n_in_ / 2)) * cutoff;
            const float val = w * (v == 0.f ? 1.f : std::sin(v * pi) / (v * pi));
            sum += val;
            y[x] = val;
        }
        buf_.assign(2 * n_in_, std::complex<float>(0.f, 0.f));
        for (size_t x = 0; x < n_in_; x++) buf_[x] = std::complex<float>((y[x] / sum) / (float)(2 * n_in_), 0.f);
        spec_.resize(2 * n_in_);
        fwd_->forward(buf_.data(), spec_.data());
        filter_.assign(spec_.begin(), spec_.begin() + (long)n_in_ + 1);
        overlap_.assign(n_out_, 0.f);
        out_spec_.resize(2 * n_out_);
        out_time_.resize(2 * n_out_);
    }
    size_t input_frames() const { return n_in_; }    // FftFixedInOut::input_frames_next
    size_t output_frames() const { return n_out_; }
    void reset() { overlap_.assign(n_out_, 0.f); }

    // in: input_frames() mono samples -> out: output_frames() samples
    void process(const float* in, float* out) {
        for (size_t i = 0; i < n_in_; i++) buf_[i] = std::complex<float>(in[i], 0.f);
        for (size_t i = n_in_; i < 2 * n_in_; i++) buf_[i] = std::complex<float>(0.f, 0.f);
        fwd_->forward(buf_.data(), spec_.data());
        const size_t new_len = n_in_ < n_out_ ? n_in_ + 1 : n_out_;
        const size_t m = 2 * n_out_;
        for (auto& v : out_spec_) v = std::complex<float>(0.f, 0.f);
        for (size_t k = 0; k < new_len && k <= n_out_; k++) {
            const std::complex<float> v = spec_[k] * filter_[k];
            if (k == 0 || k == n_out_) {
                out_spec_[k] = std::complex<float>(v.real(), 0.f);   // a real inverse transform ignores these imaginary parts
            } else {   // Hermitian half, stored conjugated: inverse(X) = conj(forward(conj(X)))
                out_spec_[k] = std::conj(v);
                out_spec_[m - k] = v;
            }
        }
        inv_->forward(out_spec_.data(), out_time_.data());
        for (size_t n = 0; n < n_out_; n++) out[n] = out_time_[n].real() + overlap_[n];
        for (size_t n = 0; n < n_out_; n++) overlap_[n] = out_time_[n_out_ + n].real();
    }

  private:
    static size_t gcd(size_t a, size_t b) { return b == 0 ? a : gcd(b, a % b); }
    size_t n_in_ = 0, n_out_ = 0;
    std::unique_ptr<StockhamFft> fwd_, inv_;
    std::vector<std::complex<float>> filter_, buf_, spec_, out_spec_, out_time_;
    std::vector<float> overlap_;
};

}  // namespace rp
