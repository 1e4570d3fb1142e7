// Compiled and run by tests/test_cpp_host.py: the C++ host mirror (include/noaa_apt.hpp) above the C ABI.
// Without a GPU it checks the host-side behaviour (filter design, error mapping); with one it also decodes.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "noaa_apt.hpp"

using namespace noaa_apt;

#define EXPECT(cond)                                                            \
    do {                                                                        \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    Context ctx;
    // filters.rs:368-372, :377-413
    EXPECT(filters::NoFilter().design() == Signal{1.f});
    filters::Lowpass lp(Freq::hz(123.f, Rate::hz(1000)), 40.f, Freq::hz(12.f, Rate::hz(1000)));
    lp.resample(Rate::hz(1000), Rate::hz(3000));
    EXPECT(lp.cutout == Freq::hz(123.f, Rate::hz(3000)) && lp.delta_w == Freq::hz(12.f, Rate::hz(3000)));
    const float cut = 4160.f / 12480.f;
    EXPECT(filters::Lowpass(Freq::pi_rad(cut), 25.f, Freq::pi_rad(cut) / 5.f).design().size() == 37);
    // dsp.rs:420-434 -> err::Error::RateOverflow
    try {
        dsp::resample_with_filter(ctx, Signal(1000, 0.f), Rate::hz(99371), Rate::hz(93911), filters::NoFilter());
        EXPECT(false);
    } catch (const err::Error &e) { EXPECT(e.kind == err::Kind::RateOverflow); }
    // decode.rs:79-83 -> Internal("Got less than 10 rows ...")
    try {
        decode(ctx, config::Settings(), Signal(20000, 0.f), Rate::hz(11025), true);
        EXPECT(false);
    } catch (const err::Error &e) {
        EXPECT(e.kind == err::Kind::Internal && e.status == APT_ERR_TOO_SHORT);
        EXPECT(std::string(e.what()).find("less than 10 rows") != std::string::npos);
    }
    if (!gpu) {
        // no device: compute entry points fail loudly
        if (apt_device_count() == 0) {
            try {
                dsp::demodulate(ctx, Signal(100, 1.f), Freq::hz(2400.f, Rate::hz(12480)));
                EXPECT(false);
            } catch (const err::Error &e) { EXPECT(e.kind == err::Kind::Cuda); }
        }
        std::printf("OK host\n");
        return 0;
    }
    // a synthetic AM signal: sync-like square modulation every half second
    const uint32_t rate = 48000;
    Signal x(rate * 12);
    for (size_t i = 0; i < x.size(); ++i) {
        const double t = double(i) / rate;
        const double line = std::fmod(t, 0.5);
        const double px = line < 0.0094 ? (std::fmod(line * 4160.0, 4.0) < 2.0 ? 0.1 : 1.0) : 0.5 + 0.3 * std::sin(40.0 * line);
        x[i] = float(std::lrint(20000.0 * px * std::sin(2.0 * M_PI * 2400.0 * t) + 150.0 * std::sin(12345.678 * i)));
    }
    int calls = 0;
    Context c2([&](float, const std::string &) { ++calls; });
    const Signal rows = decode(c2, config::Settings(), x, Rate::hz(rate), true);
    EXPECT(rows.size() % 2080 == 0 && rows.size() >= 15 * 2080);
    EXPECT(rows[0] == 0.f && calls == 5);
    const Signal e = dsp::demodulate(ctx, Signal(1000, 3.f), Freq::hz(2400.f, Rate::hz(12480)));
    EXPECT(e[0] == 0.f && e[1] > 0.f);
    std::printf("OK gpu %zu rows\n", rows.size() / 2080);
    return 0;
}
