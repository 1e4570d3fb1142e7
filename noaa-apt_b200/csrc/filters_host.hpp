// Host-side filter design and unit arithmetic of the decode path.
//
// Product code (not the oracle): this is what designs the taps the CUDA kernels
// run with.  It restates, in f32 and in the reference's expression order,
//   frequency.rs:58-117  (Freq / Rate)
//   misc.rs:20-57        (bessel_i0)
//   filters.rs:48-196    (NoFilter, Lowpass, LowpassDcRemoval, kaiser, product)
// Compiled with -ffp-contract=off so that every f32 operation rounds on its own,
// as Rust's do; sinf/cosf/powf come from the platform libm like Rust's std.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "aptb200.h"

namespace aptb200 {

// frequency.rs:30-87.  A discrete-time frequency held as a fraction of pi rad/sample.
struct Freq {
    float pi_rad_;
    static Freq pi_rad(float f) { return Freq{f}; }
    static Freq hz(float f, uint32_t rate);      // 2*f / rate
    float get_pi_rad() const { return pi_rad_; }
    float get_rad() const;                       // pi_rad * PI
    Freq operator/(float d) const { return Freq{pi_rad_ / d}; }
    Freq operator-(Freq o) const { return Freq{pi_rad_ - o.pi_rad_}; }
};

float bessel_i0(float x);

// filters.rs:144-183
std::vector<float> kaiser(float atten, Freq delta_w);

// Filter::design for the three filters of filters.rs.  Returns APT_ERR_BAD_ARG where the
// reference panics ("Kaiser window length should be odd" cannot trigger: kaiser() forces odd;
// a non-positive length -- atten <= 8 or negative delta_w -- has no defined result).
int design(const apt_filter &f, std::vector<float> &taps);

// Filter::resample, filters.rs:90-94 / 134-138.
void resample_filter(apt_filter &f, uint32_t input_rate, uint32_t output_rate);

// L, M of dsp::resample_with_filter (dsp.rs:73-75) and the RateOverflow check (dsp.rs:82-91).
struct Ratio {
    uint32_t l, m;
};
int resample_ratio(uint32_t input_rate, uint32_t output_rate, Ratio &r);

// Number of outputs of fast_resampling (dsp.rs:230-279): ceil((len*L - off) / M), off = (N-1)/2.
uint64_t polyphase_len(uint64_t len, uint32_t l, uint32_t m, size_t ntaps);

// decode::generate_sync_frame, decode.rs:171-199.
int sync_frame(uint32_t work_rate, std::vector<int8_t> &frame);

}  // namespace aptb200
