// noaa_apt.hpp -- C++ host-side mirror of the reference's module surface for the decode path, layered on
// the C ABI (aptb200.h).  Same names, argument meaning and error behaviour as the Rust crate:
//
//   noaa_apt::decode(ctx, settings, signal, input_rate, sync)        decode.rs:43-49  (re-exported noaa_apt.rs:5)
//   noaa_apt::dsp::{resample_with_filter, resample, demodulate, filter}   dsp.rs:62,132,350,386
//   noaa_apt::filters::{Filter, NoFilter, Lowpass, LowpassDcRemoval}      filters.rs:10-46
//   noaa_apt::{Freq, Rate}                                                frequency.rs:30-117
//   noaa_apt::Context (status callback)                                   context.rs:100-129
//   noaa_apt::config::Settings (DSP fields)                               config.rs:85-98
//   noaa_apt::err::Error                                                  err.rs:9-44  (err::Result<T> -> exceptions)
//
// Header only; link with libaptb200.so.  The Rust toolchain is not available in the build image, so this is
// the compiled host side above the ABI; rust/ holds the equivalent Rust shim as source.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "aptb200.h"

namespace noaa_apt {

using Signal = std::vector<float>;   // dsp.rs:16

namespace err {
enum class Kind { Internal, RateOverflow, InvalidInput, Cuda };
struct Error : std::runtime_error {
    Kind kind;
    int status;
    Error(Kind k, int st, const std::string &msg) : std::runtime_error(msg), kind(k), status(st) {}
};
inline void check(int st) {
    if (st == APT_OK) return;
    std::string msg = apt_last_error();
    if (msg.empty()) msg = apt_strerror(st);
    switch (st) {
    case APT_ERR_RESAMPLE_TO_ZERO:
    case APT_ERR_TOO_SHORT:
    case APT_ERR_FEW_SYNC_FRAMES:
    case APT_ERR_WORK_RATE:
    case APT_ERR_EMPTY_RESULT: throw Error(Kind::Internal, st, msg);
    case APT_ERR_RATE_OVERFLOW: throw Error(Kind::RateOverflow, st, msg);
    case APT_ERR_CUDA:
    case APT_ERR_NOMEM: throw Error(Kind::Cuda, st, msg);
    default: throw Error(Kind::InvalidInput, st, msg);
    }
}
}  // namespace err

struct Rate {   // frequency.rs:98-117
    uint32_t value;
    static Rate hz(uint32_t r) { return Rate{r}; }
    uint32_t get_hz() const { return value; }
    bool operator==(Rate o) const { return value == o.value; }
};

struct Freq {   // frequency.rs:30-87 (held as a fraction of pi rad/sample)
    float value;
    static Freq pi_rad(float f) { return Freq{f}; }
    static Freq hz(float f, Rate rate) { return Freq{apt_freq_hz(f, rate.get_hz())}; }
    float get_pi_rad() const { return value; }
    Freq operator/(float d) const { return Freq{value / d}; }
    bool operator==(Freq o) const { return value == o.value; }
};

class Context {   // context.rs:100-129; the per-step WAV export is not part of the fast path
  public:
    using Callback = std::function<void(float, const std::string &)>;
    explicit Context(Callback cb = nullptr) : cb_(std::move(cb)) {}
    static Context decode(Callback cb = nullptr) { return Context(std::move(cb)); }
    static Context resample(Callback cb = nullptr) { return Context(std::move(cb)); }
    void status(float progress, const std::string &description) {
        if (cb_) cb_(progress, description);
    }
    static void trampoline(float progress, const char *description, void *user) {
        static_cast<Context *>(user)->status(progress, description ? description : "");
    }
    static constexpr bool export_steps = false, export_resample_filtered = false;

  private:
    Callback cb_;
};

namespace config {
struct Settings {   // config.rs:85-98, defaults = "standard" profile (default_settings.toml:108-116)
    uint32_t work_rate = 12480;
    float resample_atten = 30.f, resample_delta_freq = 1000.f, resample_cutout = 4800.f, demodulation_atten = 25.f;
    float wav_resample_atten = 40.f, wav_resample_delta_freq = 0.1f;
    apt_settings to_c() const {
        return apt_settings{work_rate, resample_atten, resample_delta_freq, resample_cutout, demodulation_atten};
    }
    static Settings profile(const std::string &name) {
        apt_settings c;
        err::check(apt_profile_settings(name.c_str(), &c));
        Settings s;
        s.work_rate = c.work_rate;
        s.resample_atten = c.resample_atten;
        s.resample_delta_freq = c.resample_delta_freq;
        s.resample_cutout = c.resample_cutout;
        s.demodulation_atten = c.demodulation_atten;
        return s;
    }
};
}  // namespace config

namespace filters {
struct Filter {   // filters.rs:10-16
    virtual ~Filter() = default;
    virtual apt_filter to_c() const = 0;
    virtual void resample(Rate /*input_rate*/, Rate /*output_rate*/) {}
    Signal design() const {
        const apt_filter f = to_c();
        size_t n = 0;
        err::check(apt_filter_design(&f, nullptr, 0, &n));
        Signal taps(n);
        err::check(apt_filter_design(&f, taps.data(), taps.size(), &n));
        return taps;
    }
};
struct NoFilter : Filter {   // filters.rs:48-54
    apt_filter to_c() const override { return apt_filter{APT_FILTER_NONE, 0.f, 0.f, 0.f}; }
};
struct Lowpass : Filter {   // filters.rs:56-95
    Freq cutout;
    float atten;
    Freq delta_w;
    Lowpass(Freq c, float a, Freq d) : cutout(c), atten(a), delta_w(d) {}
    apt_filter to_c() const override { return apt_filter{APT_FILTER_LOWPASS, cutout.value, atten, delta_w.value}; }
    void resample(Rate in, Rate out) override {
        apt_filter f = to_c();
        apt_filter_resample(&f, in.get_hz(), out.get_hz());
        cutout.value = f.cutout_pi;
        delta_w.value = f.delta_w_pi;
    }
};
struct LowpassDcRemoval : Lowpass {   // filters.rs:97-139
    using Lowpass::Lowpass;
    apt_filter to_c() const override { return apt_filter{APT_FILTER_LOWPASS_DC, cutout.value, atten, delta_w.value}; }
};
}  // namespace filters

namespace dsp {
using noaa_apt::Freq;
using noaa_apt::Rate;
using noaa_apt::Signal;

inline Signal resample_with_filter(Context &, const Signal &signal, Rate input_rate, Rate output_rate,
                                   const filters::Filter &filt) {   // dsp.rs:62-126
    const apt_filter f = filt.to_c();
    uint64_t n = 0;
    err::check(apt_resample_len(signal.size(), input_rate.get_hz(), output_rate.get_hz(), &f, &n));
    Signal out(n);
    err::check(apt_resample_with_filter(signal.data(), signal.size(), input_rate.get_hz(), output_rate.get_hz(), &f,
                                        out.data(), out.size(), &n));
    out.resize(n);
    return out;
}
inline Signal resample(Context &ctx, const Signal &signal, Rate input_rate, Rate output_rate, float atten,
                       Freq delta_w) {   // dsp.rs:132-162
    const float cut_hz = output_rate.get_hz() > input_rate.get_hz() ? static_cast<float>(input_rate.get_hz()) / 2.f
                                                                    : static_cast<float>(output_rate.get_hz()) / 2.f;
    return resample_with_filter(ctx, signal, input_rate, output_rate,
                                filters::Lowpass(Freq::hz(cut_hz, input_rate), atten, delta_w));
}
inline Signal demodulate(Context &, const Signal &signal, Freq carrier_freq) {   // dsp.rs:350-383
    Signal out(signal.size());
    err::check(apt_demodulate(signal.data(), signal.size(), carrier_freq.get_pi_rad(), out.data()));
    return out;
}
inline Signal filter(Context &, const Signal &signal, const filters::Filter &filt) {   // dsp.rs:386-410
    const apt_filter f = filt.to_c();
    Signal out(signal.size());
    err::check(apt_filter_signal(signal.data(), signal.size(), &f, out.data()));
    return out;
}
}  // namespace dsp

// noaa_apt::decode == decode::decode, decode.rs:43-162
inline Signal decode(Context &ctx, const config::Settings &settings, const Signal &signal, Rate input_rate, bool sync) {
    const apt_settings s = settings.to_c();
    uint64_t bound = 0, n = 0;
    err::check(apt_decode_len_bound(signal.size(), input_rate.get_hz(), &s, &bound));
    Signal out(bound ? bound : 1);
    err::check(apt_decode(signal.data(), signal.size(), input_rate.get_hz(), &s, sync ? 1 : 0, out.data(), out.size(),
                          &n, &Context::trampoline, &ctx));
    out.resize(n);
    return out;
}

// decode::find_sync, decode.rs:204-263 (private in the crate; exposed for tests)
inline std::vector<uint64_t> find_sync(Context &, const Signal &signal, Rate work_rate) {
    std::vector<uint64_t> pos(signal.size() / 64 + 8);
    size_t n = 0;
    err::check(apt_find_sync(signal.data(), signal.size(), work_rate.get_hz(), pos.data(), pos.size(), &n, nullptr));
    pos.resize(n);
    return pos;
}

}  // namespace noaa_apt
