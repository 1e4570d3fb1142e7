// C ABI (include/aptb200.h) over the decoder object and the stage kernels.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "aptb200.h"
#include "common.hpp"
#include "decoder.hpp"
#include "filters_host.hpp"
#include "launch.hpp"

using namespace aptb200;

namespace aptb200 {
int decoder_enqueue(apt_decoder *d, const void *in, int format, uint64_t n, int sync, float *rows_out,
                    const void *host_chunked);
int run_find_sync(apt_decoder *d, uint64_t nwork);
int ensure_legacy_sync(apt_decoder *d);
int redo_sync_legacy(apt_decoder *d);
int materialise_stages(apt_decoder *d);
extern std::atomic<int> g_jobs_in_flight[64];
}  // namespace aptb200

namespace {

inline size_t sample_bytes(int format) { return format == APT_PCM16 ? 2 : 4; }

// RAII device buffer for the one-shot stage entry points.
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        APT_CUDA(cudaMalloc(&p, bytes ? bytes : 4));
        return APT_OK;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};

int require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(APT_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return APT_OK;
}

int free_decoder_buffers(apt_decoder *d) {
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (void *p : {(void *)d->d_h, (void *)d->d_lp, (void *)d->d_one, (void *)d->d_guard, d->d_in, (void *)d->d_r,
                    (void *)d->d_e, (void *)d->d_f, (void *)d->d_corr, (void *)d->d_aligned, (void *)d->d_root_list,
                    (void *)d->d_root_count, (void *)d->d_pos, (void *)d->d_res, (void *)d->d_out, d->d_pick, (void *)d->d_tile_taps, (void *)d->d_tile_xs, (void *)d->d_conv,
                    (void *)d->d_ph_table, (void *)d->d_ph_xs, (void *)d->d_ctl, (void *)d->d_desc, (void *)d->d_pool, (void *)d->d_roots2, (void *)d->d_tile_base, (void *)d->d_by_id})
        if (p) cudaFree(p);
    if (d->h_res) cudaFreeHost(d->h_res);
    if (d->h_out) cudaFreeHost(d->h_out);
    if (d->h_post) cudaFreeHost(d->h_post);
    for (void *p : {(void *)d->d_post, (void *)d->d_tel, (void *)d->d_out8})
        if (p) cudaFree(p);
    d->own_stager.reset();
    for (auto e : d->ev_begin) cudaEventDestroy(e);
    for (auto e : d->ev_end) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (d->ev_copied[i]) cudaEventDestroy(d->ev_copied[i]);
        if (d->ev_free[i]) cudaEventDestroy(d->ev_free[i]);
    }
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->stream) cudaStreamDestroy(d->stream);
    return APT_OK;
}

// Allocates what find_sync needs for up to max_work work-rate samples: the picker's buffers always, the fused stage's
// record pool when the plan has one (the legacy f / corr / per-block root lists come on first use: ensure_legacy_sync).
int alloc_sync_buffers(apt_decoder *d) {
    const Plan &p = d->plan;
    if (getenv("APTB200_SEQUENTIAL_PICK")) d->use_parallel_pick = false;
    if (getenv("APTB200_GENERIC_LOWPASS")) d->use_fused_lowpass = false;
    if (!p.work_multiple || d->max_work <= p.guard.size()) return APT_OK;
    if (p.dist > 16384) return APT_OK;   // k_roots holds two blocks of min_distance floats in shared memory (work_rate <=
                                         // 9 * 4160): decoding with sync is refused at submit, --no-sync still works
    d->max_corr = d->max_work - p.guard.size();
    d->max_blocks = static_cast<uint32_t>((d->max_corr + p.dist - 1) / p.dist);
    d->max_positions = static_cast<uint32_t>(d->max_work / p.row + 4);
    d->use_records = d->use_fused_lowpass && !getenv("APTB200_LEGACY_SYNC") &&
                     lowpass_corr_supported(static_cast<u32>(p.lp.size()), p.dec);
    if (d->use_records) {
        d->tile_w = records_tile(p.dec);
        d->max_tiles = static_cast<uint32_t>((d->max_corr + d->tile_w - 1) / d->tile_w);
        // Record pool: every tile owns a region of kRegion records (a noisy recording has ~200 per tile of 1920 positions: a
        // third), tiles with more take theirs from a shared overflow area behind the regions (one atomic, rare).  A
        // recording that does not fit even that (silence, ramps: every position a record) is decoded by the exact-order legacy
        // kernels (kSyncRedo).  APTB200_RECORD_POOL=<records> shrinks the pool (tests of that path): no regions then.
        constexpr uint64_t kRegion = 640;
        uint64_t region = kRegion;
        if (const char *e = getenv("APTB200_RECORD_REGION")) region = strtoull(e, nullptr, 10);
        uint64_t cap = static_cast<uint64_t>(d->max_tiles) * region + std::max<uint64_t>(d->max_corr / 8, 1u << 16);
        if (const char *e = getenv("APTB200_RECORD_POOL")) {
            cap = std::max<uint64_t>(strtoull(e, nullptr, 10), 64);
            if (static_cast<uint64_t>(d->max_tiles) * region > cap / 2) region = 0;
        }
        if (cap > (1u << 30)) {                  // hours at once: cap the pool, give the regions what is left of it
            cap = 1u << 30;
            region = std::min<uint64_t>(region, cap / 2 / std::max<uint32_t>(d->max_tiles, 1));
        }
        d->pool_cap = static_cast<uint32_t>(cap);
        d->pool_region = static_cast<uint32_t>(region);
        APT_CUDA(cudaMalloc(&d->d_ctl, sizeof(SyncCtl)));
        APT_CUDA(cudaMemset(d->d_ctl, 0, sizeof(SyncCtl)));
        APT_CUDA(cudaMalloc(&d->d_desc, static_cast<size_t>(d->max_tiles) * sizeof(TileDesc)));
        APT_CUDA(cudaMalloc(&d->d_pool, static_cast<size_t>(d->pool_cap) * sizeof(Rec)));
        APT_CUDA(cudaMalloc(&d->d_roots2, static_cast<size_t>(d->pool_cap) * sizeof(u32)));
        APT_CUDA(cudaMalloc(&d->d_by_id, static_cast<size_t>(d->pool_cap) * sizeof(u32)));
        APT_CUDA(cudaMalloc(&d->d_tile_base, static_cast<size_t>(d->max_tiles) * sizeof(u32)));
    }
    const uint32_t count_blocks = std::max(d->max_blocks, d->max_tiles);
    APT_CUDA(cudaMalloc(&d->d_guard, p.guard.size()));
    APT_CUDA(cudaMemcpy(d->d_guard, p.guard.data(), p.guard.size(), cudaMemcpyHostToDevice));
    APT_CUDA(cudaMalloc(&d->d_root_count, static_cast<size_t>(count_blocks) * sizeof(u32)));
    APT_CUDA(cudaMalloc(&d->d_pos, static_cast<size_t>(d->max_positions) * sizeof(u32)));
    // parallel picker: room for every row-aligned start plus ~63 roots per row (noisy recordings have ~35);
    // beyond that the kernel falls back to the sequential walk by itself.
    const u32 cap = static_cast<u32>(std::min<uint64_t>(static_cast<uint64_t>(d->max_positions) * 64 + 65536, 1u << 28));
    APT_CUDA(cudaMalloc(&d->d_pick, pick_scratch_bytes(count_blocks, d->max_positions, cap)));
    d->pick = pick_scratch_carve(d->d_pick, count_blocks, d->max_positions, cap);
    APT_CUDA(cudaMemset(d->pick.ticket, 0, 8));
    if (d->d_ctl) d->pick.ticket = &d->d_ctl->pad[0];   // zeroed with the control block at the start of every job
    return APT_OK;
}

}  // namespace

// ============================================================================ misc / status

extern "C" const char *apt_strerror(int status) {
    switch (status) {
    case APT_OK: return "ok";
    case APT_ERR_RESAMPLE_TO_ZERO: return "Can't resample to 0Hz";
    case APT_ERR_TOO_SHORT: return "Got less than 10 rows of samples, audio file is too short";
    case APT_ERR_FEW_SYNC_FRAMES: return "Found less than 5 sync frames, audio file is too short or too noisy";
    case APT_ERR_WORK_RATE: return "work_rate is not multiple of FINAL_RATE";
    case APT_ERR_RATE_OVERFLOW: return "Can't resample, looks like the sample rates do not have a big divisor in common";
    case APT_ERR_CUDA: return "CUDA error or no CUDA device (no CPU fallback)";
    case APT_ERR_BAD_ARG: return "invalid argument";
    case APT_ERR_NOMEM: return "out of memory";
    case APT_ERR_CAPACITY: return "output buffer too small";
    case APT_ERR_EMPTY_RESULT: return "Got zero samples after resampling, audio file too short or output sampling frequency too low";
    case APT_ERR_IO: return "WAV file cannot be opened, parsed or written";
    default: return "unknown status";
    }
}

extern "C" const char *apt_last_error(void) { return last_error_slot().c_str(); }
extern "C" int apt_abi_version(void) { return APTB200_ABI_VERSION; }

extern "C" int apt_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

extern "C" void apt_default_settings(apt_settings *s) {
    if (!s) return;
    // default_settings.toml:108-116
    s->work_rate = 12480;
    s->resample_atten = 30.f;
    s->resample_delta_freq = 1000.f;
    s->resample_cutout = 4800.f;
    s->demodulation_atten = 25.f;
}

extern "C" int apt_profile_settings(const char *profile, apt_settings *s) {
    if (!profile || !s) return fail(APT_ERR_BAD_ARG, "null argument");
    if (!strcmp(profile, "standard")) {
        apt_default_settings(s);
    } else if (!strcmp(profile, "fast")) {      // default_settings.toml:120-128
        *s = apt_settings{16640, 30.f, 3000.f, 4800.f, 23.f};
    } else if (!strcmp(profile, "slow")) {      // default_settings.toml:132-140
        *s = apt_settings{20800, 40.f, 500.f, 4800.f, 25.f};
    } else {
        return fail(APT_ERR_BAD_ARG, "unknown profile '%s'", profile);
    }
    return APT_OK;
}

// ================================================================================ filters

extern "C" float apt_freq_hz(float f_hz, uint32_t rate_hz) { return Freq::hz(f_hz, rate_hz).get_pi_rad(); }
extern "C" float apt_bessel_i0(float x) { return bessel_i0(x); }

extern "C" void apt_filter_resample(apt_filter *f, uint32_t input_rate, uint32_t output_rate) {
    if (f) resample_filter(*f, input_rate, output_rate);
}

extern "C" int apt_filter_design(const apt_filter *f, float *out, size_t cap, size_t *n) {
    if (!f || !n) return fail(APT_ERR_BAD_ARG, "null argument");
    std::vector<float> taps;
    int st = design(*f, taps);
    if (st != APT_OK) return fail(st, "filter cannot be designed (kind %d, atten %g, delta_w %g)", f->kind,
                                  (double)f->atten, (double)f->delta_w_pi);
    *n = taps.size();
    if (out && cap) memcpy(out, taps.data(), std::min(cap, taps.size()) * sizeof(float));
    return APT_OK;
}

extern "C" int apt_generate_sync_frame(uint32_t work_rate, int8_t *out, size_t cap, size_t *n) {
    if (!n) return fail(APT_ERR_BAD_ARG, "null argument");
    std::vector<int8_t> g;
    int st = sync_frame(work_rate, g);
    if (st != APT_OK) return fail(st, "work_rate is not multiple of FINAL_RATE");
    *n = g.size();
    if (out && cap) memcpy(out, g.data(), std::min(cap, g.size()));
    return APT_OK;
}

// ============================================================================ stage: resample

namespace {

struct ResamplePlan {
    Ratio r{};
    bool polyphase = false;
    std::vector<float> taps;
    uint64_t nout = 0;
};

int plan_resample(uint64_t n, uint32_t in_rate, uint32_t out_rate, const apt_filter *f, ResamplePlan &rp) {
    if (!f) return fail(APT_ERR_BAD_ARG, "null filter");
    int st = resample_ratio(in_rate, out_rate, rp.r);
    if (st == APT_ERR_RESAMPLE_TO_ZERO) return fail(st, "Can't resample to 0Hz");
    if (st == APT_ERR_RATE_OVERFLOW)
        return fail(st, "Can't resample, looks like the sample rates do not have a big divisor in common. "
                        "input_rate: %u, output_rate: %u, l: %u, m: %u", in_rate, out_rate, rp.r.l, rp.r.m);
    if (st != APT_OK) return fail(st, "invalid input rate");
    rp.polyphase = rp.r.l > 1;
    apt_filter ff = *f;
    if (rp.polyphase) resample_filter(ff, in_rate, in_rate * rp.r.l);
    st = design(ff, rp.taps);
    if (st != APT_OK) return fail(st, "filter cannot be designed");
    rp.nout = rp.polyphase ? polyphase_len(n, rp.r.l, rp.r.m, rp.taps.size()) : n / rp.r.m;
    return APT_OK;
}

}  // namespace

extern "C" int apt_resample_len(uint64_t n, uint32_t input_rate, uint32_t output_rate, const apt_filter *f,
                                uint64_t *nout) {
    if (!nout) return fail(APT_ERR_BAD_ARG, "null argument");
    ResamplePlan rp;
    APT_TRY(plan_resample(n, input_rate, output_rate, f, rp));
    *nout = rp.nout;
    return APT_OK;
}

extern "C" int apt_resample_with_filter(const float *signal, uint64_t n, uint32_t input_rate, uint32_t output_rate,
                                        const apt_filter *f, float *out, uint64_t cap, uint64_t *nout) {
    if (!nout || (!signal && n)) return fail(APT_ERR_BAD_ARG, "null argument");
    ResamplePlan rp;
    APT_TRY(plan_resample(n, input_rate, output_rate, f, rp));
    *nout = rp.nout;
    if (rp.nout > cap || (!out && rp.nout)) return fail(APT_ERR_CAPACITY, "output needs %llu floats", (unsigned long long)rp.nout);
    if (rp.nout == 0) return APT_OK;
    APT_TRY(require_device());
    DevBuf dx, dh, dy;
    APT_TRY(dx.alloc(n * sizeof(float)));
    APT_TRY(dh.alloc(rp.taps.size() * sizeof(float)));
    APT_TRY(dy.alloc(rp.nout * sizeof(float)));
    APT_CUDA(cudaMemcpy(dx.p, signal, n * sizeof(float), cudaMemcpyHostToDevice));
    APT_CUDA(cudaMemcpy(dh.p, rp.taps.data(), rp.taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    const LaunchCtx c{nullptr, 148};
    if (rp.polyphase) {
        const uint64_t off2 = 2 * ((static_cast<uint64_t>(rp.taps.size()) - 1) / 2);
        TilePlan tp{};
        std::vector<float> tt;
        std::vector<u32> xs;
        UtPlan up{};
        std::vector<float> us;
        if (!getenv("APTB200_GENERIC_RESAMPLER") && make_ut_plan(rp.r.l, rp.r.m, rp.taps, up, us)) {
            APT_TRY(launch_polyphase_ut(c, dx.as<float>(), n, dh.as<float>(), up, us, rp.nout, 0, 0, false, 0.f, 1.f, dy.as<float>()));
            APT_CUDA(cudaDeviceSynchronize());
        } else if (!getenv("APTB200_GENERIC_RESAMPLER") && make_tile_plan(rp.r.l, rp.r.m, rp.taps, tp, tt, xs)) {
            DevBuf dt, dg;
            APT_TRY(dt.alloc(tt.size() * sizeof(float)));
            APT_TRY(dg.alloc(xs.size() * sizeof(u32)));
            APT_CUDA(cudaMemcpy(dt.p, tt.data(), tt.size() * sizeof(float), cudaMemcpyHostToDevice));
            APT_CUDA(cudaMemcpy(dg.p, xs.data(), xs.size() * sizeof(u32), cudaMemcpyHostToDevice));
            APT_TRY(launch_polyphase_tiled(c, dx.as<float>(), n, dt.as<float>(), dg.as<u32>(), tp, rp.nout, 0, 0, false, 0.f,
                                           1.f, dy.as<float>()));
            APT_CUDA(cudaDeviceSynchronize());
        } else {
            PhPlan pp{};
            std::vector<float> ptab;
            std::vector<unsigned short> pxs;
            if (!getenv("APTB200_GENERIC_RESAMPLER") && make_ph_plan(rp.r.l, rp.r.m, rp.taps, pp, ptab, pxs)) {
                DevBuf dt, dg;
                APT_TRY(dt.alloc(ptab.size() * sizeof(float)));
                APT_TRY(dg.alloc(pxs.size() * sizeof(unsigned short)));
                APT_CUDA(cudaMemcpy(dt.p, ptab.data(), ptab.size() * sizeof(float), cudaMemcpyHostToDevice));
                APT_CUDA(cudaMemcpy(dg.p, pxs.data(), pxs.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
                APT_TRY(launch_polyphase_ph(c, dx.p, APT_F32, n, dt.as<float>(), dg.as<unsigned short>(), pp, rp.nout, 0, 0, false, 0.f,
                                            1.f, dy.as<float>()));
                APT_CUDA(cudaDeviceSynchronize());
            } else {
                APT_TRY(launch_polyphase(c, dx.p, APT_F32, n, dh.as<float>(), rp.r.l, rp.r.m, off2, 0, rp.nout, false, 0.f, 1.f,
                                         dy.as<float>()));
            }
        }
    } else {
        APT_TRY(launch_fir_decimate(c, dx.p, APT_F32, dh.as<float>(), static_cast<u32>(rp.taps.size()), rp.r.m,
                                    rp.nout, dy.as<float>()));
    }
    APT_CUDA(cudaMemcpy(out, dy.p, rp.nout * sizeof(float), cudaMemcpyDeviceToHost));
    return APT_OK;
}

extern "C" int apt_resample(const float *signal, uint64_t n, uint32_t input_rate, uint32_t output_rate, float atten,
                            float delta_w_pi, float *out, uint64_t cap, uint64_t *nout) {
    if (input_rate == 0) return fail(APT_ERR_BAD_ARG, "invalid input rate");
    // dsp.rs:140-149: keep what fits below the lower of the two Nyquist frequencies
    const float cut_hz = output_rate > input_rate ? static_cast<float>(input_rate) / 2.f
                                                  : static_cast<float>(output_rate) / 2.f;
    apt_filter f{APT_FILTER_LOWPASS, Freq::hz(cut_hz, input_rate).get_pi_rad(), atten, delta_w_pi};
    return apt_resample_with_filter(signal, n, input_rate, output_rate, &f, out, cap, nout);
}

// ===================================================================== stage: demod / filter

extern "C" int apt_demodulate(const float *signal, uint64_t n, float carrier_pi, float *out) {
    if (n == 0) return fail(APT_ERR_BAD_ARG, "empty signal");   // signal[0] panics, dsp.rs:367
    if (!signal || !out) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_TRY(require_device());
    const float phi = 2.f * Freq::pi_rad(carrier_pi).get_rad();
    const float cosphi2 = std::cos(phi) * 2.f;
    const float sinphi = std::sin(phi);
    DevBuf dx, dy;
    APT_TRY(dx.alloc(n * sizeof(float)));
    APT_TRY(dy.alloc(n * sizeof(float)));
    APT_CUDA(cudaMemcpy(dx.p, signal, n * sizeof(float), cudaMemcpyHostToDevice));
    APT_TRY(launch_envelope(LaunchCtx{nullptr, 148}, dx.as<float>(), n, cosphi2, sinphi, dy.as<float>()));
    APT_CUDA(cudaMemcpy(out, dy.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return APT_OK;
}

extern "C" int apt_filter_taps(const float *signal, uint64_t n, const float *coeff, size_t ncoeff, float *out) {
    if ((!signal || !out) && n) return fail(APT_ERR_BAD_ARG, "null argument");
    if (!coeff && ncoeff) return fail(APT_ERR_BAD_ARG, "null coefficients");
    if (n == 0) return APT_OK;
    APT_TRY(require_device());
    DevBuf dx, dc, dy;
    APT_TRY(dx.alloc(n * sizeof(float)));
    APT_TRY(dc.alloc(ncoeff * sizeof(float)));
    APT_TRY(dy.alloc(n * sizeof(float)));
    APT_CUDA(cudaMemcpy(dx.p, signal, n * sizeof(float), cudaMemcpyHostToDevice));
    if (ncoeff) APT_CUDA(cudaMemcpy(dc.p, coeff, ncoeff * sizeof(float), cudaMemcpyHostToDevice));
    APT_TRY(launch_fir_decimate(LaunchCtx{nullptr, 148}, dx.p, APT_F32, dc.as<float>(), static_cast<u32>(ncoeff), 1, n,
                                dy.as<float>()));
    APT_CUDA(cudaMemcpy(out, dy.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return APT_OK;
}

extern "C" int apt_filter_signal(const float *signal, uint64_t n, const apt_filter *f, float *out) {
    if (!f) return fail(APT_ERR_BAD_ARG, "null filter");
    std::vector<float> taps;
    int st = design(*f, taps);
    if (st != APT_OK) return fail(st, "filter cannot be designed");
    return apt_filter_taps(signal, n, taps.data(), taps.size(), out);
}

// ============================================================================ stage: find_sync

extern "C" int apt_find_sync(const float *signal, uint64_t n, uint32_t work_rate, uint64_t *positions, size_t cap,
                             size_t *npositions, float *corr) {
    if (!signal || !npositions) return fail(APT_ERR_BAD_ARG, "null argument");
    apt_decoder d;   // a decoder shell that only owns the sync workspaces
    d.plan.st.work_rate = work_rate;
    int st = sync_frame(work_rate, d.plan.guard);
    if (st != APT_OK) return fail(st, "work_rate is not multiple of FINAL_RATE");
    if (work_rate > UINT32_MAX / kPxPerRow) return fail(APT_ERR_BAD_ARG, "work_rate too large");
    d.plan.work_multiple = true;
    d.plan.row = kPxPerRow * work_rate / kFinalRate;
    d.plan.dist = static_cast<uint32_t>(static_cast<uint64_t>(d.plan.row) * 8 / 10);
    if (n < d.plan.guard.size()) return fail(APT_ERR_BAD_ARG, "signal shorter than the sync frame");   // decode.rs:225 underflow
    if (n >= (1ull << 32)) return fail(APT_ERR_BAD_ARG, "signal too long");
    APT_TRY(require_device());
    struct Guard {
        apt_decoder *d;
        ~Guard() { free_decoder_buffers(d); }
    } guard{&d};
    APT_CUDA(cudaGetDevice(&d.device));
    APT_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    d.max_work = n;
    size_t got = 0;
    if (n == d.plan.guard.size()) {
        // empty correlation: the peak list is just the seed (0, 0.0) (decode.rs:208-209)
        if (positions && cap > 0) positions[0] = 0;
        *npositions = 1;
        return APT_OK;
    }
    d.use_fused_lowpass = false;            // the stage entry point takes an already filtered signal: exact-order kernels
    APT_TRY(alloc_sync_buffers(&d));
    APT_TRY(ensure_legacy_sync(&d));
    APT_CUDA(cudaMalloc(&d.d_res, sizeof(SyncResult)));
    APT_CUDA(cudaMemsetAsync(d.d_res, 0, sizeof(SyncResult), d.stream));
    APT_CUDA(cudaMemcpyAsync(d.d_f, signal, n * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    APT_TRY(run_find_sync(&d, n));
    SyncResult res;
    APT_CUDA(cudaMemcpyAsync(&res, d.d_res, sizeof(res), cudaMemcpyDeviceToHost, d.stream));
    APT_CUDA(cudaStreamSynchronize(d.stream));
    got = res.n_peaks;
    *npositions = got;
    if (positions) {
        std::vector<u32> tmp(got);
        APT_CUDA(cudaMemcpy(tmp.data(), d.d_pos, got * sizeof(u32), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < std::min(got, cap); ++i) positions[i] = tmp[i];
    }
    if (corr) APT_CUDA(cudaMemcpy(corr, d.d_corr, (n - d.plan.guard.size()) * sizeof(float), cudaMemcpyDeviceToHost));
    if (positions && got > cap) return fail(APT_ERR_CAPACITY, "positions needs %zu entries", got);
    return APT_OK;
}

// ================================================================================= decoder

extern "C" int apt_decode_len_bound(uint64_t n, uint32_t input_rate, const apt_settings *s, uint64_t *bound) {
    if (!s || !bound) return fail(APT_ERR_BAD_ARG, "null argument");
    Plan p;
    APT_TRY(make_plan(input_rate, *s, p));
    *bound = plan_out_bound(p, n);
    return APT_OK;
}

extern "C" int apt_decoder_create(int device, uint32_t input_rate, const apt_settings *s, uint64_t max_samples,
                                  apt_decoder **dec) {
    if (!s || !dec) return fail(APT_ERR_BAD_ARG, "null argument");
    *dec = nullptr;
    std::unique_ptr<apt_decoder> d(new (std::nothrow) apt_decoder);
    if (!d) return fail(APT_ERR_NOMEM, "out of host memory");
    APT_TRY(make_plan(input_rate, *s, d->plan));
    APT_TRY(require_device());
    const Plan &p = d->plan;
    d->device = device;
    APT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    APT_CUDA(cudaGetDeviceProperties(&prop, device));
    d->sm_count = prop.multiProcessorCount;

    struct Rollback {
        apt_decoder *d;
        bool armed = true;
        ~Rollback() {
            if (armed) free_decoder_buffers(d);
        }
    } rollback{d.get()};

    APT_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    d->max_samples = max_samples;
    {
        // staging for host submits: whole recording up to 64 Mi samples, else two chunks of 32 Mi (APTB200_CHUNK_SAMPLES overrides)
        uint64_t chunk = 32ull << 20;
        if (const char *e = getenv("APTB200_CHUNK_SAMPLES")) chunk = std::max<uint64_t>(strtoull(e, nullptr, 10), 1u << 16);
        chunk &= ~7ull;   // both staging halves stay 16-byte aligned for f32 and PCM16 samples
        d->chunk_samples = (max_samples > 2 * chunk && p.first_polyphase) ? chunk : 0;
        if (getenv("APTB200_CHUNK_SAMPLES") && max_samples > chunk && p.first_polyphase) d->chunk_samples = chunk;
    }
    d->max_work = plan_work_len(p, max_samples);
    d->max_out = plan_out_bound(p, max_samples);
    if (d->max_work >= (1ull << 32))
        return fail(APT_ERR_BAD_ARG, "recording too long: %llu work-rate samples (limit 2^32-1)",
                    (unsigned long long)d->max_work);

    const float one = 1.f;
    APT_CUDA(cudaMalloc(&d->d_h, p.h.size() * sizeof(float)));
    APT_CUDA(cudaMemcpy(d->d_h, p.h.data(), p.h.size() * sizeof(float), cudaMemcpyHostToDevice));
    APT_CUDA(cudaMalloc(&d->d_lp, p.lp.size() * sizeof(float)));
    APT_CUDA(cudaMemcpy(d->d_lp, p.lp.data(), p.lp.size() * sizeof(float), cudaMemcpyHostToDevice));
    APT_CUDA(cudaMalloc(&d->d_one, sizeof(float)));
    APT_CUDA(cudaMemcpy(d->d_one, &one, sizeof(float), cudaMemcpyHostToDevice));
    if (p.ph) {
        APT_CUDA(cudaMalloc(&d->d_ph_table, p.ph_table.size() * sizeof(float)));
        APT_CUDA(cudaMemcpy(d->d_ph_table, p.ph_table.data(), p.ph_table.size() * sizeof(float), cudaMemcpyHostToDevice));
        APT_CUDA(cudaMalloc(&d->d_ph_xs, p.ph_xs.size() * sizeof(unsigned short)));
        APT_CUDA(cudaMemcpy(d->d_ph_xs, p.ph_xs.data(), p.ph_xs.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    }
    if (p.tiled) {
        APT_CUDA(cudaMalloc(&d->d_tile_taps, p.tile_taps.size() * sizeof(float)));
        APT_CUDA(cudaMemcpy(d->d_tile_taps, p.tile_taps.data(), p.tile_taps.size() * sizeof(float), cudaMemcpyHostToDevice));
        APT_CUDA(cudaMalloc(&d->d_tile_xs, p.tile_xs.size() * sizeof(u32)));
        APT_CUDA(cudaMemcpy(d->d_tile_xs, p.tile_xs.data(), p.tile_xs.size() * sizeof(u32), cudaMemcpyHostToDevice));
    }
    const size_t work_bytes = std::max<uint64_t>(d->max_work, 1) * sizeof(float);
    if (!p.first_polyphase) APT_CUDA(cudaMalloc(&d->d_r, work_bytes));
    APT_CUDA(cudaMalloc(&d->d_e, work_bytes));
    APT_TRY(alloc_sync_buffers(d.get()));
    if (!d->use_records) APT_TRY(ensure_legacy_sync(d.get()));
    APT_CUDA(cudaMalloc(&d->d_res, sizeof(SyncResult)));
    APT_CUDA(cudaMemset(d->d_res, 0, sizeof(SyncResult)));
    APT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&d->h_res), sizeof(SyncResult), cudaHostAllocDefault));
    memset(d->h_res, 0, sizeof(SyncResult));
    rollback.armed = false;
    *dec = d.release();
    return APT_OK;
}

extern "C" void apt_decoder_destroy(apt_decoder *dec) {
    if (!dec) return;
    free_decoder_buffers(dec);
    delete dec;
}

namespace {

// contrast bounds + u8 map of the rows the job leaves in d_out; the head of the control block follows the job to the host
int enqueue_image_stage(apt_decoder *d, unsigned char *out8) {
    const Plan &p = d->plan;
    const LaunchCtx c{d->stream, d->sm_count};
    const u32 max_rows = static_cast<u32>(std::max<uint64_t>(d->max_out / kPxPerRow, 1));
    const u32 fixed = d->job_sync ? 0u : static_cast<u32>(d->job_fixed_out / kPxPerRow);
    (void)p;
    APT_TRY(launch_image_stage(c, d->d_out, d->job_sync ? d->d_res : nullptr, fixed, max_rows, kPxPerRow, d->image_contrast,
                               d->image_percent, d->d_post, d->d_tel, d->d_tel + max_rows, d->d_tel + 2 * static_cast<size_t>(max_rows),
                               nullptr, out8, false));
    d->launches += d->image_contrast == APT_CONTRAST_PERCENT ? 4 : 3;
    APT_CUDA(cudaMemcpyAsync(d->h_post, d->d_post, offsetof(PostCtl, buckets), cudaMemcpyDeviceToHost, d->stream));
    return APT_OK;
}

int submit_common(apt_decoder *d, const void *signal, int format, uint64_t n, int sync, float *out, uint64_t cap,
                  bool host) {
    if (!d) return fail(APT_ERR_BAD_ARG, "null decoder");
    if (d->in_flight) return fail(APT_ERR_BAD_ARG, "decoder already has a job in flight; call apt_decoder_wait first");
    if (format != APT_F32 && format != APT_PCM16) return fail(APT_ERR_BAD_ARG, "unknown sample format %d", format);
    if (n == 0 || !signal) return fail(APT_ERR_BAD_ARG, "empty signal");   // signal[0] panics in demodulate, dsp.rs:367
    if (n > d->max_samples)
        return fail(APT_ERR_BAD_ARG, "recording of %llu samples exceeds the decoder's max_samples %llu",
                    (unsigned long long)n, (unsigned long long)d->max_samples);
    const Plan &p = d->plan;
    APT_CUDA(cudaSetDevice(d->device));

    d->job_status = APT_OK;
    d->job_host = host;
    d->job_sync = sync != 0;
    d->job_n = n;
    d->job_out = out;
    d->job_cap = cap;
    d->job_fixed_out = 0;

    const uint64_t nwork = plan_work_len(p, n);
    d->job_work = nwork;
    if (nwork < 10ull * p.row) {   // decode.rs:79-83
        d->job_status = fail(APT_ERR_TOO_SHORT, "Got less than 10 rows of samples, audio file is too short");
        return d->job_status;
    }
    if (d->job_sync && p.work_multiple && !d->d_pos)
        return fail(APT_ERR_BAD_ARG, "work_rate %u is too high for the sync picker (min_distance %u > 16384); decode without "
                                     "sync or use a work_rate <= 37440", p.st.work_rate, p.dist);
    const uint64_t need = d->job_sync ? (nwork / p.row) * kPxPerRow : plan_out_bound(p, n);
    if (!out || cap < need) return fail(APT_ERR_CAPACITY, "output needs room for %llu %s", (unsigned long long)need,
                                        d->image_contrast >= 0 ? "bytes" : "floats");
    d->job_image = d->image_contrast >= 0;
    if (d->job_image) {
        if (!p.work_multiple) return fail(APT_ERR_BAD_ARG, "the image stage needs rows of 2080 pixels (work_rate multiple of 4160)");
        const uint64_t max_rows = std::max<uint64_t>(d->max_out / kPxPerRow, 1);
        if (!d->d_post) {
            APT_CUDA(cudaMalloc(&d->d_post, sizeof(PostCtl)));
            APT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&d->h_post), sizeof(PostCtl), cudaHostAllocDefault));
            APT_CUDA(cudaMalloc(&d->d_tel, 3 * max_rows * sizeof(float)));
        }
        if (!d->d_out) APT_CUDA(cudaMalloc(&d->d_out, std::max<uint64_t>(d->max_out, 1) * sizeof(float)));
        if (host && !d->d_out8) APT_CUDA(cudaMalloc(&d->d_out8, std::max<uint64_t>(d->max_out, 1)));
    }
    const size_t elem = d->job_image ? 1 : sizeof(float);

    const void *dev_in = signal;
    float *rows_dst = out;
    const void *host_chunked = nullptr;
    d->job_in_pageable = d->job_out_pageable = false;
    if (host) {
        d->job_in_pageable = is_pageable(signal);
        d->job_out_pageable = is_pageable(out);
        if ((d->job_in_pageable || d->job_out_pageable) && !d->stager && !getenv("APTB200_NO_STAGER")) {
            int threads = 8;
            if (const char *e = getenv("APTB200_COPY_THREADS")) threads = std::max(1, atoi(e));
            d->own_stager.reset(new (std::nothrow) HostStager(d->device, 16u << 20, 3, threads));
            if (d->own_stager && !d->own_stager->ok()) d->own_stager.reset();
            d->stager = d->own_stager.get();
        }
        if (d->job_out_pageable && !d->h_out)
            APT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&d->h_out), std::max<uint64_t>(d->max_out, 1) * sizeof(float),
                                   cudaHostAllocDefault));   // sized for f32 rows: also holds the u8 image
        // Recordings longer than the chunk size are uploaded in chunks that overlap the resampling; shorter ones
        // (and the L == 1 first stage) are staged whole.
        const bool chunked = d->chunk_samples != 0 && n > d->chunk_samples && p.first_polyphase;
        if (d->chunk_samples != 0 && n > d->chunk_samples && !chunked)
            return fail(APT_ERR_BAD_ARG, "recording longer than the staging buffer and the first stage is not polyphase");
        const uint64_t stage_samples = d->chunk_samples ? 2 * d->chunk_samples : d->max_samples;
        if (!d->d_in) APT_CUDA(cudaMalloc(&d->d_in, std::max<uint64_t>(stage_samples, 1) * sizeof(float)));
        if (!d->d_out) APT_CUDA(cudaMalloc(&d->d_out, std::max<uint64_t>(d->max_out, 1) * sizeof(float)));
        if (chunked) {
            if (!d->copy_stream) {
                APT_CUDA(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
                for (int i = 0; i < 2; ++i) {
                    APT_CUDA(cudaEventCreateWithFlags(&d->ev_copied[i], cudaEventDisableTiming));
                    APT_CUDA(cudaEventCreateWithFlags(&d->ev_free[i], cudaEventDisableTiming));
                }
            }
            if (format == APT_PCM16 && d->conv_cap < d->chunk_samples) {
                if (d->d_conv) APT_CUDA(cudaFree(d->d_conv));
                d->d_conv = nullptr;
                d->conv_cap = 0;
                APT_CUDA(cudaMalloc(&d->d_conv, d->chunk_samples * sizeof(float)));
                d->conv_cap = d->chunk_samples;
            }
            host_chunked = signal;
            dev_in = nullptr;
        } else if (d->job_in_pageable && d->stager) {
            APT_CUDA(d->stager->upload(d->d_in, signal, n * sample_bytes(format), d->stream));
            dev_in = d->d_in;
        } else {
            APT_CUDA(cudaMemcpyAsync(d->d_in, signal, n * sample_bytes(format), cudaMemcpyHostToDevice, d->stream));
            dev_in = d->d_in;
        }
        rows_dst = d->d_out;
    }
    if (d->job_image) rows_dst = d->d_out;      // the f32 rows are an intermediate of the image job
    d->job_rows_src = rows_dst;
    int st = decoder_enqueue(d, dev_in, format, n, sync, rows_dst, host_chunked);
    if (st == APT_OK && d->job_image) st = enqueue_image_stage(d, host ? d->d_out8 : reinterpret_cast<unsigned char *>(out));
    if (st != APT_OK) {
        cudaStreamSynchronize(d->stream);
        d->job_status = st;
        return st;
    }
    if (d->job_sync) APT_CUDA(cudaMemcpyAsync(d->h_res, d->d_res, sizeof(SyncResult), cudaMemcpyDeviceToHost, d->stream));
    d->job_d2h_floats = 0;
    if (host) {
        // the rows come back with the job: when syncing n_rows is only known on the device, so the copy covers the bound
        // (at most one row more than is produced)
        d->job_d2h_floats = d->job_sync ? need : d->job_fixed_out;
        if (d->job_d2h_floats)
            APT_CUDA(cudaMemcpyAsync(d->job_out_pageable ? static_cast<void *>(d->h_out) : static_cast<void *>(out),
                                     d->job_image ? static_cast<const void *>(d->d_out8) : static_cast<const void *>(d->d_out),
                                     d->job_d2h_floats * elem, cudaMemcpyDeviceToHost, d->stream));
    }
    d->in_flight = true;
    if (d->device >= 0 && d->device < 64) g_jobs_in_flight[d->device].fetch_add(1, std::memory_order_relaxed);
    return APT_OK;
}

}  // namespace

extern "C" int apt_decoder_submit_device(apt_decoder *dec, const void *signal, int format, uint64_t n, int sync,
                                         float *out, uint64_t cap) {
    return submit_common(dec, signal, format, n, sync, out, cap, false);
}

extern "C" int apt_decoder_submit_host(apt_decoder *dec, const void *signal, int format, uint64_t n, int sync,
                                       float *out, uint64_t cap) {
    return submit_common(dec, signal, format, n, sync, out, cap, true);
}

extern "C" int apt_decoder_wait(apt_decoder *d, uint64_t *nout) {
    if (!d) return fail(APT_ERR_BAD_ARG, "null decoder");
    if (nout) *nout = 0;
    if (!d->in_flight) return d->job_status;
    APT_CUDA(cudaSetDevice(d->device));
    d->in_flight = false;
    if (d->device >= 0 && d->device < 64) g_jobs_in_flight[d->device].fetch_sub(1, std::memory_order_relaxed);
    static const bool trace = getenv("APTB200_TRACE_HOST") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    APT_CUDA(cudaStreamSynchronize(d->stream));
    const auto t_synced = std::chrono::steady_clock::now();
    uint64_t produced = d->job_fixed_out;
    d->last_work = d->job_work;
    d->last_peaks = 0;
    d->last_fused = d->job_fused;
    if (d->job_sync) {
        if (d->h_res->status == kSyncRedo) {
            // the record pool of the fused stage overflowed (silence, ramps: every index a record): the sync stage runs
            // again with the legacy kernels on the envelope that is still in d_e
            APT_TRY(redo_sync_legacy(d));
            if (d->job_image)
                APT_TRY(enqueue_image_stage(d, d->job_host ? d->d_out8 : reinterpret_cast<unsigned char *>(d->job_out)));
            APT_CUDA(cudaMemcpyAsync(d->h_res, d->d_res, sizeof(SyncResult), cudaMemcpyDeviceToHost, d->stream));
            if (d->job_host && d->job_d2h_floats)
                APT_CUDA(cudaMemcpyAsync(d->job_out_pageable ? static_cast<void *>(d->h_out) : static_cast<void *>(d->job_out),
                                         d->job_image ? static_cast<const void *>(d->d_out8) : static_cast<const void *>(d->d_out),
                                         d->job_d2h_floats * (d->job_image ? 1 : sizeof(float)), cudaMemcpyDeviceToHost, d->stream));
            APT_CUDA(cudaStreamSynchronize(d->stream));
            d->last_fused = false;
        }
        const SyncResult res = *d->h_res;
        d->last_peaks = res.n_peaks;
        if (res.status != APT_OK) {
            d->job_status = fail(static_cast<int>(res.status),
                                 "Found less than 5 sync frames, audio file is too short or too noisy");
            return d->job_status;
        }
        produced = static_cast<uint64_t>(res.n_rows) * kPxPerRow;
    }
    const size_t elem = d->job_image ? 1 : sizeof(float);
    if (d->job_host && d->job_out_pageable && produced) {
        if (d->stager) d->stager->scatter(d->job_out, d->h_out, produced * elem);
        else memcpy(d->job_out, d->h_out, produced * elem);
    }
    if (trace)
        fprintf(stderr, "[aptb200 host] wait: stream sync %.2f ms, copy-out %.2f ms (%s)\n",
                std::chrono::duration<double>(t_synced - t_begin).count() * 1e3,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_synced).count() * 1e3,
                d->job_out_pageable ? "pageable" : "direct");
    if (d->job_image) {
        const PostCtl &pc = *d->h_post;
        apt_image_info &ii = d->last_image;
        ii.low = pc.low;
        ii.high = pc.high;
        ii.rows = pc.rows;
        ii.telemetry_row = pc.telemetry_row;
        memcpy(ii.wedges_a, pc.wedges_a, sizeof(ii.wedges_a));
        memcpy(ii.wedges_b, pc.wedges_b, sizeof(ii.wedges_b));
        if (pc.status == 3) d->job_status = fail(APT_ERR_TOO_SHORT, "Recording too short for telemetry decoding");
        else if (pc.status != 0) d->job_status = fail(APT_ERR_BAD_ARG, "image stage: %s", pc.status == 1 ? "empty image" :
                                                      pc.status == 2 ? "no contrast bucket found" : "telemetry frame runs off the image");
        if (d->job_status != APT_OK) return d->job_status;
    }
    d->last_rows = d->plan.work_multiple ? produced / kPxPerRow : 0;
    d->last_out = produced;
    if (d->profiling) {
        d->kernel_ms.assign(d->ev_used, 0.f);
        for (int i = 0; i < d->ev_used; ++i) cudaEventElapsedTime(&d->kernel_ms[i], d->ev_begin[i], d->ev_end[i]);
    }
    if (nout) *nout = produced;
    return APT_OK;
}

extern "C" int apt_decoder_set_image_mode(apt_decoder *d, int contrast, float percent) {
    if (!d) return fail(APT_ERR_BAD_ARG, "null decoder");
    if (d->in_flight) return fail(APT_ERR_BAD_ARG, "decoder has a job in flight");
    if (contrast > APT_CONTRAST_TELEMETRY) return fail(APT_ERR_BAD_ARG, "unknown contrast mode %d", contrast);
    if (contrast == APT_CONTRAST_PERCENT && !(percent >= 0.f && percent <= 1.f))
        return fail(APT_ERR_BAD_ARG, "Percent given should be between 0 and 1");      // misc.rs:120-124
    d->image_contrast = contrast;
    d->image_percent = percent;
    return APT_OK;
}

extern "C" int apt_decoder_image_info(apt_decoder *d, apt_image_info *info) {
    if (!d || !info) return fail(APT_ERR_BAD_ARG, "null argument");
    *info = d->last_image;
    return APT_OK;
}

extern "C" int apt_decoder_last_sync(apt_decoder *d, uint64_t *positions, size_t cap, size_t *npositions) {
    if (!d || !npositions) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_CUDA(cudaSetDevice(d->device));
    const size_t got = d->last_peaks;
    *npositions = got;
    if (positions && got) {
        std::vector<u32> tmp(got);
        APT_CUDA(cudaMemcpy(tmp.data(), d->d_pos, got * sizeof(u32), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < std::min(got, cap); ++i) positions[i] = tmp[i];
    }
    return APT_OK;
}

extern "C" int apt_decoder_last_counts(apt_decoder *d, uint64_t *n_work, uint64_t *n_rows, uint64_t *n_peaks) {
    if (!d) return fail(APT_ERR_BAD_ARG, "null decoder");
    if (n_work) *n_work = d->last_work;
    if (n_rows) *n_rows = d->last_rows;
    if (n_peaks) *n_peaks = d->last_peaks;
    return APT_OK;
}

extern "C" int apt_decoder_last_root_count(apt_decoder *d, uint64_t *n_roots) {
    if (!d || !n_roots) return fail(APT_ERR_BAD_ARG, "null argument");
    *n_roots = d->h_res ? d->h_res->n_roots : 0;
    return APT_OK;
}

extern "C" int apt_decoder_last_roots(apt_decoder *d, uint64_t *roots, size_t cap, size_t *nroots) {
    if (!d || !nroots) return fail(APT_ERR_BAD_ARG, "null argument");
    *nroots = 0;
    if (!d->d_root_count || d->last_work <= d->plan.guard.size()) return APT_OK;
    APT_CUDA(cudaSetDevice(d->device));
    const uint64_t ncorr = d->last_work - d->plan.guard.size();
    const bool fused = d->last_fused && d->d_desc;
    const uint32_t block = fused ? d->tile_w : d->plan.dist;
    const uint32_t nb = static_cast<uint32_t>((ncorr + block - 1) / block);
    std::vector<u32> counts(nb);
    APT_CUDA(cudaMemcpy(counts.data(), d->d_root_count, nb * sizeof(u32), cudaMemcpyDeviceToHost));
    std::vector<TileDesc> desc;
    if (fused) {
        desc.resize(nb);
        APT_CUDA(cudaMemcpy(desc.data(), d->d_desc, nb * sizeof(TileDesc), cudaMemcpyDeviceToHost));
    }
    size_t total = 0;
    std::vector<u32> tmp;
    for (uint32_t b = 0; b < nb; ++b) {
        if (roots && counts[b]) {
            tmp.resize(counts[b]);
            const u32 *src = fused ? d->d_roots2 + desc[b].off : d->d_root_list + static_cast<size_t>(b) * block;
            APT_CUDA(cudaMemcpy(tmp.data(), src, counts[b] * sizeof(u32), cudaMemcpyDeviceToHost));
            for (u32 i = 0; i < counts[b]; ++i)
                if (total + i < cap) roots[total + i] = tmp[i];
        }
        total += counts[b];
    }
    *nroots = total;
    return APT_OK;
}

extern "C" int apt_decoder_read_stage(apt_decoder *d, int which, float *out, uint64_t cap, uint64_t *n) {
    if (!d || !n) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_CUDA(cudaSetDevice(d->device));
    const float *src = nullptr;
    uint64_t len = d->last_work;
    if (which != 0 && d->last_fused && len) APT_TRY(materialise_stages(d));   // f / corr were never written: compute them now
    switch (which) {
    case 0: src = d->d_e; break;
    case 1: src = d->d_f; break;
    case 2:
        src = d->d_corr;
        len = d->last_work > d->plan.guard.size() ? d->last_work - d->plan.guard.size() : 0;
        break;
    default: return fail(APT_ERR_BAD_ARG, "unknown stage %d", which);
    }
    if (!src) len = 0;
    *n = len;
    if (out && len) {
        if (cap < len) return fail(APT_ERR_CAPACITY, "stage needs %llu floats", (unsigned long long)len);
        APT_CUDA(cudaMemcpy(out, src, len * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return APT_OK;
}

extern "C" int apt_decoder_set_profiling(apt_decoder *d, int enabled) {
    if (!d) return fail(APT_ERR_BAD_ARG, "null decoder");
    d->profiling = enabled != 0;
    return APT_OK;
}

extern "C" int apt_decoder_kernel_count(apt_decoder *d) { return d ? static_cast<int>(d->kernel_ms.size()) : 0; }

extern "C" const char *apt_decoder_kernel_name(apt_decoder *d, int i) {
    if (!d || i < 0 || i >= static_cast<int>(d->kernel_names.size())) return "";
    return d->kernel_names[i].c_str();
}

extern "C" int apt_decoder_kernel_ms(apt_decoder *d, float *ms, int cap, int *count) {
    if (!d || !count) return fail(APT_ERR_BAD_ARG, "null argument");
    *count = static_cast<int>(d->kernel_ms.size());
    for (int i = 0; i < std::min(cap, *count); ++i) ms[i] = d->kernel_ms[i];
    return APT_OK;
}

extern "C" void *apt_decoder_stream(apt_decoder *d) { return d ? static_cast<void *>(d->stream) : nullptr; }
extern "C" uint64_t apt_decoder_launch_count(apt_decoder *d) { return d ? d->launches : 0; }

// ============================================================================ memory helpers

extern "C" int apt_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_TRY(require_device());
    APT_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 4, cudaHostAllocPortable));
    return APT_OK;
}

extern "C" void apt_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

extern "C" int apt_device_alloc(int device, void **ptr, size_t bytes) {
    if (!ptr) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_TRY(require_device());
    APT_CUDA(cudaSetDevice(device));
    APT_CUDA(cudaMalloc(ptr, bytes ? bytes : 4));
    return APT_OK;
}

extern "C" void apt_device_free(int device, void *ptr) {
    if (!ptr) return;
    cudaSetDevice(device);
    cudaFree(ptr);
}

extern "C" int apt_memcpy_h2d(int device, void *dst, const void *src, size_t bytes) {
    APT_CUDA(cudaSetDevice(device));
    APT_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return APT_OK;
}

extern "C" int apt_memcpy_d2h(int device, void *dst, const void *src, size_t bytes) {
    APT_CUDA(cudaSetDevice(device));
    APT_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return APT_OK;
}

// ============================================================================ one-shot decode

namespace {

// apt_decode() is what the Rust shim binds (rust/decode.rs): one call per recording, ordinary pageable buffers.  Creating a
// decoder costs ~14 cudaMalloc + a stream + pinned staging, so finished decoders are parked here, keyed by what their plan
// depends on, and the next call with the same (device, rate, settings) takes one over.  apt_cache_clear() frees them.
struct CacheEntry {
    int device;
    uint32_t rate;
    apt_settings st;
    apt_decoder *dec;
    uint64_t stamp;
};
std::mutex g_cache_mutex;
std::vector<CacheEntry> g_cache;
uint64_t g_cache_stamp = 0;
constexpr size_t kCacheMaxIdle = 8;

bool same_settings(const apt_settings &a, const apt_settings &b) { return memcmp(&a, &b, sizeof(a)) == 0; }

apt_decoder *cache_take(int device, uint32_t rate, const apt_settings &st, uint64_t n) {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    for (size_t i = 0; i < g_cache.size(); ++i) {
        const CacheEntry &e = g_cache[i];
        if (e.device == device && e.rate == rate && same_settings(e.st, st) && e.dec->max_samples >= n) {
            apt_decoder *d = e.dec;
            g_cache.erase(g_cache.begin() + static_cast<long>(i));
            return d;
        }
    }
    return nullptr;
}

void cache_put(int device, uint32_t rate, const apt_settings &st, apt_decoder *d) {
    apt_decoder *victim = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mutex);
        // one idle decoder per key is enough for a sequential caller: a smaller one of the same key is replaced
        for (size_t i = 0; i < g_cache.size(); ++i) {
            CacheEntry &e = g_cache[i];
            if (e.device == device && e.rate == rate && same_settings(e.st, st) && e.dec->max_samples <= d->max_samples) {
                victim = e.dec;
                g_cache.erase(g_cache.begin() + static_cast<long>(i));
                break;
            }
        }
        if (!victim && g_cache.size() >= kCacheMaxIdle) {
            size_t lru = 0;
            for (size_t i = 1; i < g_cache.size(); ++i)
                if (g_cache[i].stamp < g_cache[lru].stamp) lru = i;
            victim = g_cache[lru].dec;
            g_cache.erase(g_cache.begin() + static_cast<long>(lru));
        }
        g_cache.push_back(CacheEntry{device, rate, st, d, ++g_cache_stamp});
    }
    if (victim) apt_decoder_destroy(victim);
}

// batch feeders park all their decoders (several per key)
void cache_put_multi(int device, uint32_t rate, const apt_settings &st, apt_decoder *d) {
    apt_decoder *victim = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mutex);
        if (g_cache.size() >= kCacheMaxIdle) {
            size_t lru = 0;
            for (size_t i = 1; i < g_cache.size(); ++i)
                if (g_cache[i].stamp < g_cache[lru].stamp) lru = i;
            victim = g_cache[lru].dec;
            g_cache.erase(g_cache.begin() + static_cast<long>(lru));
        }
        g_cache.push_back(CacheEntry{device, rate, st, d, ++g_cache_stamp});
    }
    if (victim) apt_decoder_destroy(victim);
}

int decode_oneshot(const void *signal, int format, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
                   float *out, uint64_t cap, uint64_t *nout, apt_status_cb cb, void *user, int contrast = -1,
                   float percent = 0.f, apt_image_info *info = nullptr) {
    if (!s || !nout) return fail(APT_ERR_BAD_ARG, "null argument");
    *nout = 0;
    if (n == 0 || !signal) return fail(APT_ERR_BAD_ARG, "empty signal");
    int device = 0;
    {
        // errors that do not need a device come first, in the reference's order
        Plan p;
        APT_TRY(make_plan(input_rate, *s, p));
        if (plan_work_len(p, n) < 10ull * p.row)
            return fail(APT_ERR_TOO_SHORT, "Got less than 10 rows of samples, audio file is too short");
        if (sync && !p.work_multiple) return fail(APT_ERR_WORK_RATE, "work_rate is not multiple of FINAL_RATE");
    }
    APT_TRY(require_device());
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    static const bool use_cache = getenv("APTB200_NO_CACHE") == nullptr;
    apt_decoder *d = use_cache ? cache_take(device, input_rate, *s, n) : nullptr;
    if (!d) {
        // head-room so that the next, slightly longer recording of a series reuses the decoder
        const uint64_t room = use_cache ? ((n + n / 8 + (1u << 20)) & ~((1ull << 20) - 1)) : n;
        APT_TRY(apt_decoder_create(device, input_rate, s, room, &d));
    }
    d->cb = cb;
    d->cb_user = user;
    d->image_contrast = contrast;
    d->image_percent = percent;
    int st = apt_decoder_submit_host(d, signal, format, n, sync, out, cap);
    if (st == APT_OK) st = apt_decoder_wait(d, nout);
    if (st == APT_OK && info) *info = d->last_image;
    d->image_contrast = -1;
    d->cb = nullptr;
    d->cb_user = nullptr;
    if (use_cache) cache_put(device, input_rate, *s, d);
    else apt_decoder_destroy(d);
    return st;
}

}  // namespace

extern "C" int apt_bind_thread_to_device(int device) {
    if (require_device() != APT_OK) return 0;
    return bind_thread_to_device(device) ? 1 : 0;
}

extern "C" void apt_cache_clear(void) {
    std::vector<CacheEntry> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache_mutex);
        drop.swap(g_cache);
    }
    for (auto &e : drop) apt_decoder_destroy(e.dec);
}

extern "C" int apt_decode(const float *signal, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
                          float *out, uint64_t cap, uint64_t *nout, apt_status_cb cb, void *user) {
    return decode_oneshot(signal, APT_F32, n, input_rate, s, sync, out, cap, nout, cb, user);
}

extern "C" int apt_decode_pcm16(const int16_t *pcm, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
                                float *out, uint64_t cap, uint64_t *nout, apt_status_cb cb, void *user) {
    return decode_oneshot(pcm, APT_PCM16, n, input_rate, s, sync, out, cap, nout, cb, user);
}

// ============================================================================== image stage

extern "C" int apt_decode_image_u8(const void *signal, int format, uint64_t n, uint32_t input_rate, const apt_settings *s,
                                   int sync, int contrast, float percent, uint8_t *out, uint64_t cap, uint64_t *nout,
                                   apt_image_info *info, apt_status_cb cb, void *user) {
    if (contrast < 0 || contrast > APT_CONTRAST_TELEMETRY) return fail(APT_ERR_BAD_ARG, "unknown contrast mode %d", contrast);
    if (contrast == APT_CONTRAST_PERCENT && !(percent >= 0.f && percent <= 1.f))
        return fail(APT_ERR_BAD_ARG, "Percent given should be between 0 and 1");
    return decode_oneshot(signal, format, n, input_rate, s, sync, reinterpret_cast<float *>(out), cap, nout, cb, user, contrast,
                          percent, info);
}

namespace {

// rows on the host -> device, image-stage kernels, results back (stage entry points; one call, no retained state)
int image_stage_host(const float *signal, uint64_t n, int contrast, float percent, const float *bounds, apt_image_info *info,
                     uint8_t *out, float *mean_a, float *mean_b, float *variance) {
    if (!signal || n == 0) return fail(APT_ERR_BAD_ARG, "empty signal");   // get_min of an empty vector is an error, dsp.rs:21-25
    APT_TRY(require_device());
    const uint64_t rows = n / kPxPerRow;
    const bool whole = rows * kPxPerRow == n && rows > 0;
    if (!whole && (contrast == APT_CONTRAST_TELEMETRY || mean_a)) return fail(APT_ERR_BAD_ARG, "not a whole number of 2080-px rows");
    DevBuf dx, dctl, dtel, dout, dbounds;
    APT_TRY(dx.alloc(n * sizeof(float)));
    APT_TRY(dctl.alloc(sizeof(PostCtl)));
    APT_TRY(dtel.alloc(3 * std::max<uint64_t>(rows, 1) * sizeof(float)));
    if (out) APT_TRY(dout.alloc(n));
    APT_CUDA(cudaMemcpy(dx.p, signal, n * sizeof(float), cudaMemcpyHostToDevice));
    if (bounds) {
        APT_TRY(dbounds.alloc(2 * sizeof(float)));
        APT_CUDA(cudaMemcpy(dbounds.p, bounds, 2 * sizeof(float), cudaMemcpyHostToDevice));
    }
    const LaunchCtx c{nullptr, 148};
    // a partial last row cannot occur in an image; min/max, percent and the map treat the signal as one row of n pixels then
    const u32 nrows = whole ? static_cast<u32>(rows) : 1u;
    const u32 px = whole ? kPxPerRow : static_cast<u32>(n);
    if (!whole && n >= (1ull << 32)) return fail(APT_ERR_BAD_ARG, "signal too long");
    float *tel = dtel.as<float>();
    APT_TRY(launch_image_stage(c, dx.as<float>(), nullptr, nrows, nrows, px, contrast, percent, dctl.as<PostCtl>(), tel,
                               tel + std::max<uint64_t>(rows, 1), tel + 2 * std::max<uint64_t>(rows, 1), bounds ? dbounds.as<float>() : nullptr,
                               out ? dout.as<unsigned char>() : nullptr, out == nullptr));
    APT_CUDA(cudaDeviceSynchronize());
    if (!bounds) {
        PostCtl pc;
        APT_CUDA(cudaMemcpy(&pc, dctl.p, offsetof(PostCtl, buckets), cudaMemcpyDeviceToHost));
        if (pc.status == 3) return fail(APT_ERR_TOO_SHORT, "Recording too short for telemetry decoding");
        if (pc.status != 0) return fail(APT_ERR_BAD_ARG, "image stage failed (status %u)", pc.status);
        if (info) {
            info->low = pc.low;
            info->high = pc.high;
            info->rows = whole ? rows : 0;
            info->telemetry_row = pc.telemetry_row;
            memcpy(info->wedges_a, pc.wedges_a, sizeof(info->wedges_a));
            memcpy(info->wedges_b, pc.wedges_b, sizeof(info->wedges_b));
        }
    }
    if (out) APT_CUDA(cudaMemcpy(out, dout.p, n, cudaMemcpyDeviceToHost));
    if (mean_a) APT_CUDA(cudaMemcpy(mean_a, tel, rows * sizeof(float), cudaMemcpyDeviceToHost));
    if (mean_b) APT_CUDA(cudaMemcpy(mean_b, tel + rows, rows * sizeof(float), cudaMemcpyDeviceToHost));
    if (variance) APT_CUDA(cudaMemcpy(variance, tel + 2 * rows, rows * sizeof(float), cudaMemcpyDeviceToHost));
    return APT_OK;
}

}  // namespace

extern "C" int apt_map_signal_u8(const float *signal, uint64_t n, float low, float high, uint8_t *out) {
    if (n == 0) return APT_OK;
    if (!out) return fail(APT_ERR_BAD_ARG, "null argument");
    const float b[2] = {low, high};
    return image_stage_host(signal, n, APT_CONTRAST_MINMAX, 0.f, b, nullptr, out, nullptr, nullptr, nullptr);
}

extern "C" int apt_contrast_bounds(const float *signal, uint64_t n, int contrast, float percent, apt_image_info *info) {
    if (!info) return fail(APT_ERR_BAD_ARG, "null argument");
    if (contrast < 0 || contrast > APT_CONTRAST_TELEMETRY) return fail(APT_ERR_BAD_ARG, "unknown contrast mode %d", contrast);
    if (contrast == APT_CONTRAST_PERCENT && !(percent >= 0.f && percent <= 1.f))
        return fail(APT_ERR_BAD_ARG, "Percent given should be between 0 and 1");
    memset(info, 0, sizeof(*info));
    return image_stage_host(signal, n, contrast, percent, nullptr, info, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int apt_telemetry_rows(const float *signal, uint64_t n, float *mean_a, float *mean_b, float *variance) {
    if (!mean_a || !mean_b || !variance) return fail(APT_ERR_BAD_ARG, "null argument");
    return image_stage_host(signal, n, APT_CONTRAST_MINMAX, 0.f, nullptr, nullptr, nullptr, mean_a, mean_b, variance);
}

// ======================================================================= resample tool (WAV -> WAV)

extern "C" int apt_quantize_i16(const float *signal, uint64_t n, int16_t *out) {
    if (n == 0 || !signal) return fail(APT_ERR_BAD_ARG, "Can't get maximum of a zero length vector");   // dsp.rs:21-25
    if (!out) return fail(APT_ERR_BAD_ARG, "null argument");
    APT_TRY(require_device());
    DevBuf dx, dctl, dout;
    APT_TRY(dx.alloc(n * sizeof(float)));
    APT_TRY(dctl.alloc(sizeof(PostCtl)));
    APT_TRY(dout.alloc(n * sizeof(int16_t)));
    APT_CUDA(cudaMemcpy(dx.p, signal, n * sizeof(float), cudaMemcpyHostToDevice));
    APT_TRY(launch_quantize_i16(LaunchCtx{nullptr, 148}, dx.as<float>(), n, dctl.as<PostCtl>(), dout.as<short>()));
    APT_CUDA(cudaMemcpy(out, dout.p, n * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return APT_OK;
}

extern "C" int apt_resample_wav(const char *input_path, const char *output_path, uint32_t output_rate, float atten,
                                float delta_w_pi, uint64_t *nout) {
    if (nout) *nout = 0;
    apt_wav_info wi{};
    APT_TRY(apt_wav_info_read(input_path, &wi));
    std::vector<float> x(wi.frames);
    uint64_t n = 0;
    uint32_t rate = 0;
    APT_TRY(apt_wav_load(input_path, x.data(), x.size(), &n, &rate));
    if (rate == 0) return fail(APT_ERR_BAD_ARG, "invalid input rate");
    const float cut_hz = output_rate > rate ? static_cast<float>(rate) / 2.f : static_cast<float>(output_rate) / 2.f;   // dsp.rs:140-149
    apt_filter f{APT_FILTER_LOWPASS, Freq::hz(cut_hz, rate).get_pi_rad(), atten, delta_w_pi};
    uint64_t ny = 0;
    APT_TRY(apt_resample_len(n, rate, output_rate, &f, &ny));
    if (ny == 0)   // resample.rs:46-52
        return fail(APT_ERR_EMPTY_RESULT, "Got zero samples after resampling, audio file too short or output sampling frequency too low");
    std::vector<float> y(ny);
    APT_TRY(apt_resample_with_filter(x.data(), n, rate, output_rate, &f, y.data(), y.size(), &ny));
    std::vector<int16_t> q(ny);
    APT_TRY(apt_quantize_i16(y.data(), ny, q.data()));
    APT_TRY(apt_wav_write_i16(output_path, q.data(), ny, output_rate));
    if (nout) *nout = ny;
    return APT_OK;
}

// ============================================================================= introspection

extern "C" int apt_tile_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_tile_info *info,
                             float *tile_taps, size_t cap_taps, uint32_t *group_xs, size_t cap_groups) {
    if (!taps || !info) return fail(APT_ERR_BAD_ARG, "null argument");
    std::vector<float> h(taps, taps + ntaps), tt;
    std::vector<u32> xs;
    TilePlan tp{};
    memset(info, 0, sizeof(*info));
    if (!make_tile_plan(l, m, h, tp, tt, xs)) return APT_OK;
    *info = apt_tile_info{1, tp.groups, tp.p_out, tp.p_in, tp.usteps, tp.row_len, tp.qt, tp.smem_bytes, 4,
                          tp.slice_stride, tp.half_taps, tp.shift, tp.iters, tp.group_stride, tp.ctas_per_sm,
                          tp.pair_pitch, tp.halves, tp.rows_per_copy};
    if (tile_taps) memcpy(tile_taps, tt.data(), std::min(cap_taps, tt.size()) * sizeof(float));
    if (group_xs) memcpy(group_xs, xs.data(), std::min(cap_groups, xs.size()) * sizeof(u32));
    return APT_OK;
}

extern "C" int apt_ut_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_ut_info *info, float *stream,
                           size_t cap_stream) {
    if (!taps || !info) return fail(APT_ERR_BAD_ARG, "null argument");
    std::vector<float> h(taps, taps + ntaps), st;
    UtPlan up{};
    memset(info, 0, sizeof(*info));
    if (!make_ut_plan(l, m, h, up, st)) return APT_OK;
    info->usable = 1;
    info->l = up.l; info->m = up.m; info->np = up.np; info->q = up.q; info->rows_per_block = up.rb; info->vec = up.vec;
    info->back = up.back; info->chunks = up.chunks; info->slot_floats = up.slot_floats; info->slot_stride = up.slot_stride;
    info->nslot = up.nslot; info->warps = up.warps; info->smem_bytes = up.smem_bytes; info->nvec = up.nvec;
    info->stream_b = up.stream_b; info->halo_u0 = up.halo_u0; info->halo_n = up.halo_n; info->chunk_len = kUtChunk;
    for (int p = 0; p < 8; ++p) {
        info->cs[p] = up.cs[p];
        info->ce[p] = up.ce[p];
    }
    if (stream) memcpy(stream, st.data(), std::min(cap_stream, st.size()) * sizeof(float));
    return APT_OK;
}

extern "C" int apt_ph_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_ph_info *info, float *table,
                           size_t cap_table, uint16_t *xs, size_t cap_xs) {
    if (!taps || !info) return fail(APT_ERR_BAD_ARG, "null argument");
    std::vector<float> h(taps, taps + ntaps), tb;
    std::vector<unsigned short> x;
    PhPlan pp{};
    memset(info, 0, sizeof(*info));
    if (!make_ph_plan(l, m, h, pp, tb, x)) return APT_OK;
    *info = apt_ph_info{1, pp.l, pp.m, pp.j, pp.jpad, pp.pitch, pp.row_len, pp.smem_bytes};
    if (table) memcpy(table, tb.data(), std::min(cap_table, tb.size()) * sizeof(float));
    if (xs) memcpy(xs, x.data(), std::min(cap_xs, x.size()) * sizeof(uint16_t));
    return APT_OK;
}

// ===================================================================================== batch

extern "C" int apt_decoder_poll(apt_decoder *d) {
    if (!d || !d->in_flight) return 1;
    cudaSetDevice(d->device);
    const cudaError_t e = cudaStreamQuery(d->stream);
    if (e == cudaErrorNotReady) return 0;
    return 1;   // finished (or failed: wait() reports it)
}

// One feeder thread per device: binds itself to the device's NUMA node, owns `streams_per_device` decoders and one copy
// stager, keeps every decoder busy and reaps the jobs in the order they complete.  Nothing is shared between devices.
extern "C" int apt_decode_batch(const void *const *signals, int format, const uint64_t *lens, int count,
                                uint32_t input_rate, const apt_settings *s, int sync, float *const *outs,
                                const uint64_t *caps, uint64_t *nouts, int *statuses, const int *devices, int ndevices,
                                int streams_per_device) {
    if (!signals || !lens || !s || !outs || !caps || !nouts || count < 0)
        return fail(APT_ERR_BAD_ARG, "null argument");
    if (count == 0) return APT_OK;
    APT_TRY(require_device());
    int default_device = 0;
    if (!devices || ndevices <= 0) {
        devices = &default_device;
        ndevices = 1;
    }
    if (streams_per_device <= 0) streams_per_device = 4;
    uint64_t max_len = 0;
    for (int i = 0; i < count; ++i) {
        max_len = std::max(max_len, lens[i]);
        nouts[i] = 0;
        if (statuses) statuses[i] = APT_ERR_CUDA;      // overwritten by the job's own status; never left "ok" by default
    }
    static const bool use_cache = getenv("APTB200_NO_CACHE") == nullptr;
    std::vector<int> local_status(count, APT_ERR_CUDA);
    std::vector<std::string> messages(ndevices);
    std::vector<int> fatal(ndevices, APT_OK);

    const int streams_wanted = streams_per_device;
    auto feeder = [&](int g) {
        const int device = devices[g];
        int streams_per_device = streams_wanted;       // per feeder: one device running out of memory does not limit the others
        bind_thread_to_device(device);
        std::vector<apt_decoder *> decs;
        std::vector<int> job_of;
        int next = g;                                  // recording i -> device i % G   
        int in_flight = 0;
        auto reap = [&](size_t k) {
            uint64_t got = 0;
            const int st = apt_decoder_wait(decs[k], &got);
            const int job = job_of[k];
            nouts[job] = got;
            local_status[job] = st;
            job_of[k] = -1;
            --in_flight;
        };
        auto fail_rest = [&](int st) {
            fatal[g] = st;
            messages[g] = apt_last_error();
            for (; next < count; next += ndevices) local_status[next] = st;
        };
        while (next < count || in_flight > 0) {
            // a free decoder takes the next recording of this device
            size_t free_k = decs.size();
            for (size_t k = 0; k < decs.size(); ++k)
                if (job_of[k] < 0) { free_k = k; break; }
            if (next < count && (free_k < decs.size() || static_cast<int>(decs.size()) < streams_per_device)) {
                if (free_k == decs.size()) {
                    apt_decoder *d = cache_take(device, input_rate, *s, max_len);     // parked by an earlier call
                    const int st = d ? APT_OK : apt_decoder_create(device, input_rate, s, max_len, &d);
                    if (st != APT_OK) {
                        if (decs.empty()) { fail_rest(st); continue; }
                        streams_per_device = static_cast<int>(decs.size());   // out of memory: make do with what exists
                        continue;
                    }
                    // one stager (ring + copy threads) serves all decoders of this feeder: the first decoder's own
                    if (decs.empty()) {
                        if (!d->own_stager) {
                            int threads = 8;
                            if (const char *e = getenv("APTB200_COPY_THREADS")) threads = std::max(1, atoi(e));
                            d->own_stager.reset(new (std::nothrow) HostStager(device, 16u << 20, 3, threads));
                            if (d->own_stager && !d->own_stager->ok()) d->own_stager.reset();
                        }
                        d->stager = d->own_stager.get();
                    } else {
                        d->stager = decs[0]->stager;
                    }
                    d->image_contrast = -1;
                    decs.push_back(d);
                    job_of.push_back(-1);
                }
                const int job = next;
                next += ndevices;
                const int st = apt_decoder_submit_host(decs[free_k], signals[job], format, lens[job], sync, outs[job], caps[job]);
                if (st == APT_OK) {
                    job_of[free_k] = job;
                    ++in_flight;
                } else {
                    local_status[job] = st;
                }
                continue;
            }
            // every decoder is busy (or nothing is left to submit): take whichever job has finished, else the oldest
            bool reaped = false;
            for (size_t k = 0; k < decs.size() && !reaped; ++k)
                if (job_of[k] >= 0 && apt_decoder_poll(decs[k])) { reap(k); reaped = true; }
            if (!reaped) {
                size_t oldest = decs.size();
                for (size_t k = 0; k < decs.size(); ++k)
                    if (job_of[k] >= 0 && (oldest == decs.size() || job_of[k] < job_of[oldest])) oldest = k;
                if (oldest < decs.size()) reap(oldest);
            }
        }
        for (auto *d : decs) {
            d->stager = d->own_stager.get();           // nullptr unless it owns one: never keep a pointer into another decoder
            if (use_cache) cache_put_multi(device, input_rate, *s, d);
            else apt_decoder_destroy(d);
        }
    };

    if (ndevices == 1) {
        feeder(0);
    } else {
        std::vector<std::thread> threads;
        for (int g = 0; g < ndevices; ++g) threads.emplace_back(feeder, g);
        for (auto &t : threads) t.join();
    }
    int first_error = APT_OK;
    for (int i = 0; i < count; ++i) {
        if (statuses) statuses[i] = local_status[i];
        if (local_status[i] != APT_OK && first_error == APT_OK) first_error = local_status[i];
    }
    for (int g = 0; g < ndevices; ++g)
        if (fatal[g] != APT_OK) return fail(fatal[g], "device %d: %s", devices[g], messages[g].c_str());
    return first_error;
}
