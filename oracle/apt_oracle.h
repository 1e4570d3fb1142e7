/*
 * apt_oracle.h -- CPU ORACLE for the noaa-apt decode hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C, scalar-f32, single-threaded restatement of the reference's
 * (martinber/noaa-apt v1.4.1) signal-to-image path.  It exists so the CUDA path
 * can be checked against the reference's arithmetic; it is NOT part of the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under noaa-apt_b200/ links or
 * calls it.
 *
 * PARITY STATUS: the reference is 100 % Rust and no Rust toolchain exists in
 * the build image or on the GPU boxes, so the reference itself cannot be run.
 * The oracle is pinned against every vector the reference's own tests hold for
 * this path (sync-frame goldens decode.rs:270-319, Bessel KATs misc.rs:494-513,
 * Freq conversions frequency.rs:325-416, FIR ripple properties
 * filters.rs:243-366, RateOverflow dsp.rs:420-434, zero-input smoke
 * dsp.rs:440-468).  The reference has NO test that pins the output of
 * fast_resampling / demodulate / filter / find_sync / decode, so for those:
 * "parity unpinned" -- correctness rests on this file following the Rust
 * source line by line (every function cites the lines it restates).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c
 * into an FMA, and every f32 op rounds individually).
 */
#ifndef APT_ORACLE_H
#define APT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* err.rs:9-44 -- only the variants reachable on this path, plus BAD_ARG for
 * the inputs on which the reference panics (empty signal dsp.rs:367, rate 0). */
enum {
    ORACLE_OK = 0,
    ORACLE_ERR_RESAMPLE_TO_ZERO = 1,   /* Internal("Can't resample to 0Hz")            dsp.rs:69-71   */
    ORACLE_ERR_TOO_SHORT = 2,          /* Internal("Got less than 10 rows ...")         decode.rs:79-83 */
    ORACLE_ERR_FEW_SYNC_FRAMES = 3,    /* Internal("Found less than 5 sync frames ...") decode.rs:112-118 */
    ORACLE_ERR_WORK_RATE = 4,          /* Internal("work_rate is not multiple ...")     decode.rs:172-176 */
    ORACLE_ERR_RATE_OVERFLOW = 5,      /* RateOverflow                                  dsp.rs:82-91   */
    ORACLE_ERR_BAD_ARG = 7,            /* reference would panic                                         */
    ORACLE_ERR_NOMEM = 8
};

/* filters.rs:18-46 */
enum { ORACLE_FILTER_NONE = 0, ORACLE_FILTER_LOWPASS = 1, ORACLE_FILTER_LOWPASS_DC = 2 };

/* The DSP fields of config::Settings (config.rs:85-98). */
typedef struct {
    uint32_t work_rate;
    float resample_atten;
    float resample_delta_freq;
    float resample_cutout;
    float demodulation_atten;
} oracle_settings;

/* default_settings.toml:108-116 ("standard" profile). */
void oracle_default_settings(oracle_settings *s);

void oracle_free(void *p);

/* frequency.rs:58-87 */
float oracle_freq_hz(float f, uint32_t rate);        /* -> pi_rad */
float oracle_freq_rad(float f);                      /* -> pi_rad */
float oracle_freq_get_rad(float pi_rad);
float oracle_freq_get_hz(float pi_rad, uint32_t rate);

/* misc.rs:47-57 */
float oracle_bessel_i0(float x);

/* filters.rs:144-183; returns a malloc'd window, length in *n. */
float *oracle_kaiser(float atten, float delta_w_pi, size_t *n);

/* filters.rs:48-139; kind selects NoFilter / Lowpass / LowpassDcRemoval. */
float *oracle_design(int kind, float cutout_pi, float atten, float delta_w_pi, size_t *n);

/* dsp.rs:186-289 (export_resample_filtered = false branch). */
float *oracle_fast_resampling(const float *x, uint64_t len, uint32_t l, uint32_t m,
                              const float *coeff, size_t ncoeff, uint64_t *nout);

/* dsp.rs:294-307 */
float *oracle_decimate(const float *x, uint64_t len, uint32_t m, uint64_t *nout);

/* dsp.rs:350-383; carrier given as Freq.pi_rad. */
int oracle_demodulate(const float *x, uint64_t len, float carrier_pi_rad, float *out);

/* dsp.rs:386-410 */
int oracle_filter(const float *x, uint64_t len, const float *coeff, size_t ncoeff, float *out);

/* dsp.rs:62-126 */
int oracle_resample_with_filter(const float *x, uint64_t len, uint32_t in_rate, uint32_t out_rate,
                                int kind, float cutout_pi, float atten, float delta_w_pi,
                                float **out, uint64_t *nout);

/* dsp.rs:132-162 */
int oracle_resample(const float *x, uint64_t len, uint32_t in_rate, uint32_t out_rate,
                    float atten, float delta_w_pi, float **out, uint64_t *nout);

/* decode.rs:171-199; returns malloc'd +-1 template. */
int oracle_generate_sync_frame(uint32_t work_rate, int8_t **out, size_t *n);

/* decode.rs:204-263; positions malloc'd.  If corr_out != NULL it receives the
 * len - guard_len correlation values (the export_steps branch, decode.rs:235). */
int oracle_find_sync(const float *x, uint64_t len, uint32_t work_rate,
                     uint64_t **pos, size_t *npos, float *corr_out);

/* decode.rs:43-162 */
int oracle_decode(const float *x, uint64_t len, uint32_t in_rate, const oracle_settings *s,
                  int sync, float **out, uint64_t *nout);

/* Same as oracle_decode but also hands back the intermediate signals
 * (what Context::step would dump): any pointer may be NULL. */
typedef struct {
    float *resampled;   uint64_t n_resampled;   /* "resample_decimated" */
    float *demodulated; uint64_t n_demodulated; /* "demodulation_result" */
    float *filtered;    uint64_t n_filtered;    /* "filter_result" */
    uint64_t *sync_pos; size_t n_sync_pos;      /* find_sync positions */
    float *aligned;     uint64_t n_aligned;     /* "sync_result" */
} oracle_steps;
int oracle_decode_steps(const float *x, uint64_t len, uint32_t in_rate, const oracle_settings *s,
                        int sync, float **out, uint64_t *nout, oracle_steps *steps);
void oracle_steps_free(oracle_steps *steps);

/* wav.rs:31-40: PCM16 sample -> f32 is a plain `as f32` cast. */
void oracle_pcm16_to_f32(const int16_t *in, uint64_t n, float *out);

/* noaa_apt.rs:249-259 map_signal_u8 (next-row (f)3, used by tests of the u8 path). */
void oracle_map_signal_u8(const float *x, uint64_t n, float low, float high, uint8_t *out);

/* wav.rs:71-85: normalise by max and quantise to i16 (resample tool path). */
int oracle_quantize_i16(const float *x, uint64_t n, int16_t *out);

/* Next-row (f)3 -- contrast bounds and telemetry statistics of the decoded image:
 * dsp.rs:20-54 get_min/get_max; misc.rs:119-175 percent() (and its 1000 bucket counts);
 * telemetry.rs:147-170 per-row band means + variance; telemetry.rs:125-243 + :30-66 frame search and wedge values. */
int oracle_minmax(const float *x, uint64_t n, float *low, float *high);
int oracle_percent_buckets(const float *x, uint64_t n, uint32_t *buckets, float *mn, float *mx);
int oracle_percent(const float *x, uint64_t n, float percent, float *low, float *high);
int oracle_telemetry_rows(const float *x, uint64_t n, float *mean_a, float *mean_b, float *variance);
int oracle_read_telemetry(const float *x, uint64_t n, float *wedges_a, float *wedges_b, uint64_t *best_row);

#ifdef __cplusplus
}
#endif
#endif
