// Host-side filter design (see filters_host.hpp).  Build with -ffp-contract=off.
#include "filters_host.hpp"

#include <cmath>
#include <numeric>

namespace aptb200 {

namespace {
// std::f32::consts::PI
constexpr float kPi = 3.14159265358979323846f;
// decode.rs:14
constexpr uint32_t kFinalRate = 4160;

// misc.rs:20-41 -- 1 / (k! * 2^k)^2; bessel_i0 only ever reads k = 1..8 (misc.rs:49).
constexpr float kBesselCoef[9] = {
    1.0f,
    0.25f,
    0.015625f,
    0.00043402777777777775f,
    6.781684027777777e-06f,
    6.781684027777778e-08f,
    4.709502797067901e-10f,
    2.4028075495244395e-12f,
    9.385966990329842e-15f,
};
}  // namespace

Freq Freq::hz(float f, uint32_t rate) {
    // frequency.rs:68-72: `2. * f / rate.get_hz() as f32`
    return Freq{2.f * f / static_cast<float>(rate)};
}

float Freq::get_rad() const {
    // frequency.rs:75-77
    return pi_rad_ * kPi;
}

float bessel_i0(float x) {
    // misc.rs:47-57: Horner in x^2 from k = 8 down to 1, then + 1
    const float x2 = x * x;
    float acc = 0.f;
    for (int k = 8; k > 0; --k) {
        acc += kBesselCoef[k];
        acc *= x2;
    }
    return acc + 1.f;
}

std::vector<float> kaiser(float atten, Freq delta_w) {
    // filters.rs:154-161
    float beta;
    if (atten > 50.f) {
        beta = 0.1102f * (atten - 8.7f);
    } else if (atten < 21.f) {
        beta = 0.f;
    } else {
        beta = 0.5842f * std::pow(atten - 21.f, 0.4f) + 0.07886f * (atten - 21.f);
    }

    // filters.rs:164-167
    const float flen = std::ceil((atten - 8.f) / (2.285f * delta_w.get_rad()));
    std::vector<float> window;
    if (!(flen >= 0.f) || flen > 1.0e8f) return window;  // reference: nonsense / OOM
    int32_t length = static_cast<int32_t>(flen) + 1;
    if (length % 2 == 0) length += 1;

    window.reserve(static_cast<size_t>(length));
    const float denom = bessel_i0(beta);
    const float half_len = static_cast<float>(length) / 2.f;   // m / 2.
    const int32_t half = (length - 1) / 2;
    for (int32_t n = -half; n <= half; ++n) {
        // filters.rs:171-175
        const float q = static_cast<float>(n) / half_len;
        window.push_back(bessel_i0(beta * std::sqrt(1.f - q * q)) / denom);
    }
    return window;
}

int design(const apt_filter &f, std::vector<float> &taps) {
    taps.clear();
    switch (f.kind) {
    case APT_FILTER_NONE:
        taps.push_back(1.f);  // filters.rs:49-51
        return APT_OK;
    case APT_FILTER_LOWPASS:
    case APT_FILTER_LOWPASS_DC:
        break;
    default:
        return APT_ERR_BAD_ARG;
    }

    const Freq cutout = Freq::pi_rad(f.cutout_pi);
    const Freq delta_w = Freq::pi_rad(f.delta_w_pi);
    std::vector<float> window = kaiser(f.atten, delta_w);
    if (window.empty() || window.size() % 2 == 0) return APT_ERR_BAD_ARG;  // filters.rs:68-70

    const int32_t m = static_cast<int32_t>(window.size());
    const int32_t half = (m - 1) / 2;
    taps.resize(window.size());
    if (f.kind == APT_FILTER_LOWPASS) {
        // filters.rs:76-83
        for (int32_t n = -half; n <= half; ++n) {
            float v;
            if (n == 0) {
                v = cutout.get_pi_rad();
            } else {
                const float nf = static_cast<float>(n);
                v = std::sin(nf * kPi * cutout.get_pi_rad()) / (nf * kPi);
            }
            taps[static_cast<size_t>(n + half)] = v;
        }
    } else {
        // filters.rs:117-127: sinc(cutout) - sinc(delta_w / 2)
        const float lo = (delta_w / 2.f).get_pi_rad();
        for (int32_t n = -half; n <= half; ++n) {
            float v;
            if (n == 0) {
                v = cutout.get_pi_rad() - lo;
            } else {
                const float nf = static_cast<float>(n);
                v = std::sin(nf * kPi * cutout.get_pi_rad()) / (nf * kPi)
                  - std::sin(nf * kPi * lo) / (nf * kPi);
            }
            taps[static_cast<size_t>(n + half)] = v;
        }
    }
    // product(), filters.rs:186-196
    for (size_t i = 0; i < taps.size(); ++i) taps[i] *= window[i];
    return APT_OK;
}

void resample_filter(apt_filter &f, uint32_t input_rate, uint32_t output_rate) {
    if (f.kind == APT_FILTER_NONE) return;  // filters.rs:53
    const float ratio = static_cast<float>(output_rate) / static_cast<float>(input_rate);
    f.cutout_pi /= ratio;
    f.delta_w_pi /= ratio;
}

int resample_ratio(uint32_t input_rate, uint32_t output_rate, Ratio &r) {
    if (output_rate == 0) return APT_ERR_RESAMPLE_TO_ZERO;  // dsp.rs:69-71
    if (input_rate == 0) return APT_ERR_BAD_ARG;            // m == 0: the reference divides by zero
    const uint32_t g = std::gcd(input_rate, output_rate);   // dsp.rs:73
    r.l = output_rate / g;
    r.m = input_rate / g;
    if (r.l > 1) {
        // input_rate.checked_mul(l), dsp.rs:82-91
        const uint64_t wide = static_cast<uint64_t>(input_rate) * r.l;
        if (wide > UINT32_MAX) return APT_ERR_RATE_OVERFLOW;
    }
    return APT_OK;
}

uint64_t polyphase_len(uint64_t len, uint32_t l, uint32_t m, size_t ntaps) {
    const uint64_t il = len * static_cast<uint64_t>(l);
    const uint64_t off = (static_cast<uint64_t>(ntaps) - 1) / 2;
    return il > off ? (il - off + m - 1) / m : 0;
}

int sync_frame(uint32_t work_rate, std::vector<int8_t> &frame) {
    frame.clear();
    if (work_rate == 0 || work_rate % kFinalRate != 0) return APT_ERR_WORK_RATE;  // decode.rs:172-176
    const size_t pw = work_rate / kFinalRate;
    const size_t pulse = 2 * pw;
    // 1 low pulse, 7 x (low pulse, high pulse), then 8 px low  (decode.rs:188-198)
    frame.insert(frame.end(), pulse, int8_t(-1));
    for (int c = 0; c < 7; ++c) {
        frame.insert(frame.end(), pulse, int8_t(-1));
        frame.insert(frame.end(), pulse, int8_t(1));
    }
    frame.insert(frame.end(), 8 * pw, int8_t(-1));
    return APT_OK;
}

}  // namespace aptb200
