/*
 * aptb200.h -- C ABI of the B200-native APT decode path.
 *
 * Drop-in boundary for the signal-to-image hot path of martinber/noaa-apt
 * v1.4.1 (src/{dsp,filters,decode,resample}.rs).  The reference has no FFI
 * boundary today (it is a single Rust binary crate), so every entry point
 * below names the Rust function whose body it replaces; the Rust-side shim
 * that binds them is in rust/ and INTEGRATION.md.  All signatures are plain
 * C: pointers, sizes, PODs -- no CUDA, torch or C++ types.
 *
 * Conventions
 *   - "host" entry points take ordinary host pointers, are synchronous and
 *     re-entrant (no hidden global state; each call picks its CUDA device).
 *   - `apt_decoder` is the explicit-state variant: it owns a device, a stream,
 *     device workspaces and pinned staging, and offers submit/wait so that a
 *     batch of independent recordings can be kept in flight one per stream.
 *   - Every function returns an `apt_status`.  A short human-readable message
 *     for the last failure on the calling thread is in apt_last_error().
 *   - There is NO CPU fallback: without a usable CUDA device the compute entry
 *     points fail with APT_ERR_CUDA.
 *   - Arithmetic is f32 like the reference; results match the reference's
 *     scalar CPU path to <= 1e-5 of the stage's max |value| (FMA contraction
 *     is the only difference), and sync positions match exactly.
 */
#ifndef APTB200_H
#define APTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APTB200_ABI_VERSION 1

/* ---------------------------------------------------------------- status */

/* Maps onto err::Error (err.rs:9-44).  1-4 are the four Error::Internal
 * conditions reachable on this path, 5 is Error::RateOverflow. */
typedef enum apt_status {
    APT_OK = 0,
    APT_ERR_RESAMPLE_TO_ZERO = 1, /* Internal("Can't resample to 0Hz")              dsp.rs:69-71     */
    APT_ERR_TOO_SHORT = 2,        /* Internal("Got less than 10 rows of samples..") decode.rs:79-83  */
    APT_ERR_FEW_SYNC_FRAMES = 3,  /* Internal("Found less than 5 sync frames...")   decode.rs:112-118*/
    APT_ERR_WORK_RATE = 4,        /* Internal("work_rate is not multiple of ...")   decode.rs:172-176*/
    APT_ERR_RATE_OVERFLOW = 5,    /* RateOverflow(...)                               dsp.rs:82-91     */
    APT_ERR_CUDA = 6,             /* CUDA runtime/driver failure, or no device (no CPU fallback)      */
    APT_ERR_BAD_ARG = 7,          /* inputs on which the reference panics (empty signal dsp.rs:367,
                                     even Kaiser window filters.rs:68-70, rate 0) or NULL pointers    */
    APT_ERR_NOMEM = 8,            /* host or device allocation failed                                 */
    APT_ERR_CAPACITY = 9,         /* caller's output buffer is too small (required size returned)     */
    APT_ERR_EMPTY_RESULT = 10,    /* Internal("Got zero samples after resampling...") resample.rs:46-52 */
    APT_ERR_IO = 11               /* Io / WavOpen (err.rs:11-17): a WAV file cannot be opened, parsed or written       */
} apt_status;

const char *apt_strerror(int status);
/* Message of the last failure on this thread ("" if none); valid until the next call on this thread. */
const char *apt_last_error(void);
int apt_abi_version(void);
/* Number of usable CUDA devices (0 if none / no driver).  Never fails. */
int apt_device_count(void);

/* -------------------------------------------------------------- settings */

/* The DSP fields of config::Settings (config.rs:85-98) -- the only part of the
 * settings that crosses the boundary (read at decode.rs:55-75,98). */
typedef struct apt_settings {
    uint32_t work_rate;           /* Hz */
    float resample_atten;         /* dB, positive */
    float resample_delta_freq;    /* Hz */
    float resample_cutout;        /* Hz */
    float demodulation_atten;     /* dB, positive */
} apt_settings;

/* "standard" profile, default_settings.toml:108-116. */
void apt_default_settings(apt_settings *s);
/* profile: "standard" | "fast" | "slow" (default_settings.toml:108-140). */
int apt_profile_settings(const char *profile, apt_settings *s);

/* --------------------------------------------------------------- filters */

/* filters.rs:18-46.  Frequencies are Freq values in fractions of pi rad/sample
 * (Freq::get_pi_rad, frequency.rs:80-82). */
typedef enum apt_filter_kind {
    APT_FILTER_NONE = 0,          /* filters::NoFilter          */
    APT_FILTER_LOWPASS = 1,       /* filters::Lowpass           */
    APT_FILTER_LOWPASS_DC = 2     /* filters::LowpassDcRemoval  */
} apt_filter_kind;

typedef struct apt_filter {
    int kind;                     /* apt_filter_kind */
    float cutout_pi;              /* cutout.get_pi_rad()  */
    float atten;                  /* dB, positive         */
    float delta_w_pi;             /* delta_w.get_pi_rad() */
} apt_filter;

/* Freq::hz(f, rate).get_pi_rad() -- frequency.rs:68-72. */
float apt_freq_hz(float f_hz, uint32_t rate_hz);
/* misc::bessel_i0 -- misc.rs:47-57. */
float apt_bessel_i0(float x);
/* Filter::resample(input_rate, output_rate) -- filters.rs:90-94,134-138 (no-op for NoFilter). */
void apt_filter_resample(apt_filter *f, uint32_t input_rate, uint32_t output_rate);
/* Filter::design() -- filters.rs:49-51,57-88,98-132 (+ kaiser :144-183).  Host code, no GPU.
 * Writes min(*n, cap) taps to out (out may be NULL when cap == 0) and the tap count to *n. */
int apt_filter_design(const apt_filter *f, float *out, size_t cap, size_t *n);

/* -------------------------------------------- stage entry points (host) */

/* Output length of dsp::resample_with_filter for `n` input samples (exact). */
int apt_resample_len(uint64_t n, uint32_t input_rate, uint32_t output_rate, const apt_filter *f,
                     uint64_t *nout);

/* dsp::resample_with_filter -- dsp.rs:62-126 (polyphase fast_resampling :186-289 when L > 1,
 * filter + decimate :105-123 when L == 1). */
int apt_resample_with_filter(const float *signal, uint64_t n, uint32_t input_rate, uint32_t output_rate,
                             const apt_filter *f, float *out, uint64_t cap, uint64_t *nout);

/* dsp::resample -- dsp.rs:132-162 (Lowpass with cutout = min(in,out)/2); the WAV->WAV tool path. */
int apt_resample(const float *signal, uint64_t n, uint32_t input_rate, uint32_t output_rate,
                 float atten, float delta_w_pi, float *out, uint64_t cap, uint64_t *nout);

/* dsp::demodulate -- dsp.rs:350-383.  carrier_pi = carrier_freq.get_pi_rad(). */
int apt_demodulate(const float *signal, uint64_t n, float carrier_pi, float *out);

/* dsp::filter -- dsp.rs:386-410 (causal, strict i > j). */
int apt_filter_signal(const float *signal, uint64_t n, const apt_filter *f, float *out);
/* Same with explicit coefficients. */
int apt_filter_taps(const float *signal, uint64_t n, const float *coeff, size_t ncoeff, float *out);

/* decode::generate_sync_frame -- decode.rs:171-199. */
int apt_generate_sync_frame(uint32_t work_rate, int8_t *out, size_t cap, size_t *n);

/* decode::find_sync -- decode.rs:204-263.  If corr != NULL it receives the n - guard_len
 * cross-correlation values (the Context::export_steps branch, decode.rs:235-237). */
int apt_find_sync(const float *signal, uint64_t n, uint32_t work_rate,
                  uint64_t *positions, size_t cap, size_t *npositions, float *corr);

/* ------------------------------------------------------- decode (host) */

/* Context::status(progress, description) -- context.rs:127-129.  Fired on the calling thread
 * at the reference's five points (decode.rs:63,87,93,107/136,154). */
typedef void (*apt_status_cb)(float progress, const char *description, void *user);

/* Upper bound, in floats, of decode()'s output for n input samples (rows * 2080). */
int apt_decode_len_bound(uint64_t n, uint32_t input_rate, const apt_settings *s, uint64_t *bound);

/* noaa_apt::decode == decode::decode -- decode.rs:43-162.
 * out receives rows*2080 f32 pixels; *nout the count.  cap in floats. */
int apt_decode(const float *signal, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
               float *out, uint64_t cap, uint64_t *nout, apt_status_cb cb, void *user);

/* apt_decode / apt_decode_pcm16 keep the decoder they used (plan, device workspaces, pinned staging) parked per
 * (device, rate, settings) for the next call; this frees the parked ones.  Thread-safe. */
void apt_cache_clear(void);

/* Same, taking the PCM16 samples of the WAV directly: half the host->device bytes; the `as f32` cast of
 * wav::load_wav (wav.rs:31-40) happens on the device -- inside the resampler's load for the phase-major and generic
 * kernels (11025, 22050, 44100 Hz ...), as one conversion kernel in front of the uniform-tap / tiled kernels (48000,
 * 96000 Hz: +43 MB of HBM writes and reads per 15 minutes, which the PCIe saving outweighs 20 times). */
int apt_decode_pcm16(const int16_t *pcm, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
                     float *out, uint64_t cap, uint64_t *nout, apt_status_cb cb, void *user);

/* ------------------------------------------------------ WAV files and the resample tool */

/* wav::load_wav -- wav.rs:11-56 (the `hound` reader restated for what this path needs): integer PCM of 8/16/24/32 bits or
 * 32-bit float, any number of channels; channel 0 is kept and integer samples are cast with `as f32` (raw values). */
typedef struct apt_wav_info {
    uint32_t sample_rate, channels, bits_per_sample, is_float;
    uint64_t frames;
} apt_wav_info;
int apt_wav_info_read(const char *path, apt_wav_info *info);
int apt_wav_load(const char *path, float *out, uint64_t cap, uint64_t *n, uint32_t *sample_rate);
/* The 16-bit samples as they are (for apt_decode_pcm16 / APT_PCM16: half the PCIe bytes, cast on the device). */
int apt_wav_load_pcm16(const char *path, int16_t *out, uint64_t cap, uint64_t *n, uint32_t *sample_rate);
/* 16-bit mono PCM file, what resample.rs:53-66 writes. */
int apt_wav_write_i16(const char *path, const int16_t *samples, uint64_t n, uint32_t sample_rate);
/* wav::write_wav's conversion for 16-bit files -- wav.rs:71-85: (sample / max * 32767.0) as i16 with max = dsp::get_max,
 * on the device; bit-identical to the reference given the same signal. */
int apt_quantize_i16(const float *signal, uint64_t n, int16_t *out);
/* resample::resample -- resample.rs:17-71: load WAV, dsp::resample (Lowpass, atten dB, delta_w in pi rad/sample) on the
 * GPU, normalise + quantise on the GPU, write a 16-bit WAV.  (The modification-time copy of resample.rs:30,68 is file
 * plumbing and stays with the caller.)  *nout receives the number of samples written. */
int apt_resample_wav(const char *input_path, const char *output_path, uint32_t output_rate, float atten, float delta_w_pi,
                     uint64_t *nout);

/* ------------------------------------------ image stage (after decode, on the device) */

/* The front of noaa_apt::process (noaa_apt.rs:132-190): contrast bounds, then map_signal_u8 (noaa_apt.rs:249-259).  The
 * rows stay on the device until they are u8 (4x fewer bytes leave the GPU).  Given the same f32 rows the results are
 * bit-identical to the reference's (every f32 operation rounded on its own, the reference's summation order). */
typedef enum apt_contrast {
    APT_CONTRAST_MINMAX = 0,      /* Contrast::MinMax: dsp::get_min / get_max             noaa_apt.rs:157-163, dsp.rs:20-54  */
    APT_CONTRAST_PERCENT = 1,     /* Contrast::Percent(p): misc::percent, 1000 buckets    misc.rs:119-175                    */
    APT_CONTRAST_TELEMETRY = 2    /* Contrast::Telemetry: wedges 9 / 8 of the best frame  noaa_apt.rs:141-149, telemetry.rs  */
} apt_contrast;

typedef struct apt_image_info {
    float low, high;              /* contrast bounds: `low` maps to 0, `high` to 255 */
    uint64_t rows;                /* image height (width is 2080) */
    uint64_t telemetry_row;       /* telemetry: row where the best frame starts (telemetry.rs:192-221) */
    float wedges_a[16], wedges_b[16];   /* telemetry: Telemetry::values_a / values_b (telemetry.rs:30-66) */
} apt_image_info;

/* decode() + contrast + map_signal_u8 in one call: out receives rows*2080 u8 pixels (cap in bytes >= apt_decode_len_bound).
 * format: apt_sample_format of `signal` (f32 Signal or the WAV's PCM16).  Errors of the image stage: the reference's
 * Internal("Recording too short for telemetry decoding") and its panics (empty image, frame running off the image) come
 * back as APT_ERR_TOO_SHORT / APT_ERR_BAD_ARG. */
int apt_decode_image_u8(const void *signal, int format, uint64_t n, uint32_t input_rate, const apt_settings *s, int sync,
                        int contrast, float percent, uint8_t *out, uint64_t cap, uint64_t *nout, apt_image_info *info,
                        apt_status_cb cb, void *user);

/* Stage entry points on host buffers (rows = decode()'s output, n = rows*2080 values). */
/* noaa_apt::map_signal_u8 -- noaa_apt.rs:249-259. */
int apt_map_signal_u8(const float *signal, uint64_t n, float low, float high, uint8_t *out);
/* Contrast bounds of an image: MinMax / misc::percent / telemetry (info receives bounds, wedges, frame row). */
int apt_contrast_bounds(const float *signal, uint64_t n, int contrast, float percent, apt_image_info *info);
/* telemetry.rs:147-170: per-row means of the two telemetry bands and their pooled variance (n / 2080 values each). */
int apt_telemetry_rows(const float *signal, uint64_t n, float *mean_a, float *mean_b, float *variance);

/* ----------------------------------------------- decoder object (streams) */

typedef struct apt_decoder apt_decoder;

typedef enum apt_sample_format { APT_F32 = 0, APT_PCM16 = 1 } apt_sample_format;

/* Creates a decoder bound to CUDA device `device` for recordings of `input_rate` Hz of up to
 * `max_samples` samples.  Designs both filters, uploads the taps and allocates every workspace
 * (nothing is allocated afterwards). */
int apt_decoder_create(int device, uint32_t input_rate, const apt_settings *s, uint64_t max_samples,
                       apt_decoder **dec);
void apt_decoder_destroy(apt_decoder *dec);

/* Enqueue one decode on the decoder's stream and return immediately.
 *   *_device: `signal` and `out` are device pointers on the decoder's device (inputs resident in HBM).
 *   *_host:   host pointers; the H2D of the samples and the D2H of the rows are part of the job
 *             (asynchronous when the buffers are pinned, e.g. from apt_host_alloc).
 * `cap` is the capacity of `out` in floats (>= apt_decode_len_bound).  One job in flight per decoder. */
int apt_decoder_submit_device(apt_decoder *dec, const void *signal, int format, uint64_t n, int sync,
                              float *out, uint64_t cap);
int apt_decoder_submit_host(apt_decoder *dec, const void *signal, int format, uint64_t n, int sync,
                            float *out, uint64_t cap);
/* Image mode of a decoder: contrast >= 0 (apt_contrast) makes every following submit produce the u8 image instead of the
 * f32 rows -- `out` is then a uint8_t buffer and `cap` / *nout count bytes; contrast < 0 switches back to f32 rows. */
int apt_decoder_set_image_mode(apt_decoder *dec, int contrast, float percent);
/* Contrast bounds / telemetry of the last image job (after apt_decoder_wait). */
int apt_decoder_image_info(apt_decoder *dec, apt_image_info *info);
/* 1 if the decoder's job has finished (apt_decoder_wait will not block) or none is in flight, else 0. */
int apt_decoder_poll(apt_decoder *dec);
/* Block until the job is finished; returns its status (APT_ERR_FEW_SYNC_FRAMES etc.) and *nout. */
int apt_decoder_wait(apt_decoder *dec, uint64_t *nout);

/* Introspection after a wait(): sync positions found (decode.rs:110), N_w, rows. */
int apt_decoder_last_sync(apt_decoder *dec, uint64_t *positions, size_t cap, size_t *npositions);
int apt_decoder_last_counts(apt_decoder *dec, uint64_t *n_work, uint64_t *n_rows, uint64_t *n_peaks);
/* Number of "roots" of the sync correlation the peak picker worked on in the last job (diagnostic). */
int apt_decoder_last_root_count(apt_decoder *dec, uint64_t *n_roots);
/* The roots themselves, ascending (a root is a correlation index p with no larger value in (p, p + min_distance]: the
 * only places a sync position can land -- DESIGN.md "Peak picker").  Diagnostic; valid after a wait() of a sync job. */
int apt_decoder_last_roots(apt_decoder *dec, uint64_t *roots, size_t cap, size_t *nroots);
/* Copy an intermediate signal of the last job back to the host (what Context::step would dump):
 * which = 0 "demodulation_result" input i.e. resample+envelope output, 1 "filter_result",
 * 2 "sync_correlation". */
int apt_decoder_read_stage(apt_decoder *dec, int which, float *out, uint64_t cap, uint64_t *n);

/* Per-kernel device time of the LAST job in milliseconds (CUDA events on the decoder's stream);
 * enable before submitting.  Names via apt_decoder_kernel_name(i).  Profiling adds event records
 * to the stream, so use it for roofline accounting, not inside a throughput measurement. */
int apt_decoder_set_profiling(apt_decoder *dec, int enabled);
int apt_decoder_kernel_count(apt_decoder *dec);
const char *apt_decoder_kernel_name(apt_decoder *dec, int i);
int apt_decoder_kernel_ms(apt_decoder *dec, float *ms, int cap, int *count);
/* The decoder's cudaStream_t, as an opaque pointer (for event timing by the caller). */
void *apt_decoder_stream(apt_decoder *dec);
/* Kernel launches issued by this decoder since creation. */
uint64_t apt_decoder_launch_count(apt_decoder *dec);

/* Binds the calling thread to the CPUs of the NUMA node CUDA device `device` hangs off (sysfs local_cpulist), so that
 * pinned buffers it allocates afterwards and its copy loops are local to the GPU's PCIe root.  1 if bound, 0 if the
 * topology is unknown / APTB200_NO_AFFINITY is set.  The library binds its own feeder and copy threads this way. */
int apt_bind_thread_to_device(int device);

/* Pinned host memory for asynchronous submit_host (cudaHostAlloc / cudaFreeHost). */
int apt_host_alloc(void **ptr, size_t bytes);
void apt_host_free(void *ptr);
/* Device memory on `device` (cudaMalloc / cudaFree) for submit_device users without a CUDA binding. */
int apt_device_alloc(int device, void **ptr, size_t bytes);
void apt_device_free(int device, void *ptr);
int apt_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int apt_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);

/* ----------------------------------------------------------- introspection */

/* Geometry of the tiled sm_100a resampler for a ratio L/M and a tap set (DESIGN.md "Tiled polyphase
 * kernel"); host code, no GPU.  usable == 0: the shape falls back to the generic kernel.  Optional
 * outputs: tile_taps (groups*group_stride floats; per group one sub-table per slice lane, slice_stride floats
 * apart, holding per loop iteration a 32-float record: taps of half A [4 samples][4 outputs], then half B) and
 * group_xs (groups entries: first input sample of each group relative to its row). */
typedef struct apt_tile_info {
    uint32_t usable, groups, p_out, p_in, usteps, row_len, rows_per_tile, smem_bytes, slices, slice_stride,
        half_taps, shift, iters, group_stride, ctas_per_sm, pair_pitch, halves, rows_per_copy;
} apt_tile_info;
int apt_tile_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_tile_info *info,
                  float *tile_taps, size_t cap_taps, uint32_t *group_xs, size_t cap_groups);

/* Geometry and tap stream of the uniform-tap resampler (noaa-apt_b200/csrc/kernels_ut.cuh), the kernel that serves
 * fast_resampling + demodulate when L == 13 (48/96/192 kHz -> 12 480 Hz): the taps travel as a kernel parameter.
 * Host logic only.  usable == 0: (l, m, taps) does not fit and the tiled / generic kernels are used.
 * A row q is the L outputs L*q .. L*q+L-1; pair p = outputs 2p, 2p+1; a chunk is chunk_len (8) consecutive samples of the row's
 * window (sample u of row q is signal[q*M + u]); pair p is active in chunks [cs[p], ce[p]).  Two warps share a block
 * of rows_per_block rows: role 0 owns pairs [0, ceil(np/2)), role 1 the rest.  `stream` (nvec float4 = 4*nvec floats;
 * role 1 starts at float4 index stream_b) holds, in the order the kernel's loop consumes them, per (chunk, active
 * pair) the 2*chunk_len taps {T[C*c][2p], T[C*c][2p+1], T[C*c+1][2p], ... T[C*c+C-1][2p+1]}, C = chunk_len, T[u][r] = h[u*L - r*M]; the order is
 * ramp-up (pairs pb..pb+a-1 over chunks [cs[pb+a-1], cs[pb+a])), steady (all pairs over [cs[last], ce[pb])), ramp-down
 * (pairs pb+a.. over [ce[pb+a-1], ce[pb+a])). */
typedef struct apt_ut_info {
    uint32_t usable, l, m, np, q, rows_per_block, vec, back, chunks, slot_floats, slot_stride, nslot, warps, smem_bytes,
        nvec, stream_b, halo_u0, halo_n, chunk_len;
    uint32_t cs[8], ce[8];
} apt_ut_info;
int apt_ut_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_ut_info *info, float *stream,
                size_t cap_stream);

/* Geometry and phase table of the phase-major resampler (noaa-apt_b200/csrc/kernels_ph.cuh), the kernel that serves
 * fast_resampling + demodulate for large interpolation factors (11025 / 22050 / 44100 Hz -> 12 480 Hz: L = 832 / 416 / 208).
 * Host logic only.  usable == 0: the shape does not fit.  Four consecutive phases (group g = r / 4) share a window of
 * jpad samples that starts at signal[m*q + xs[g] - 4]; output k = l*q + r is
 *     sum_{i < jpad} table[(g*jpad + i)*4 + r%4] * signal[m*q + xs[g] - 4 + i]   (samples outside the signal count as zero). */
typedef struct apt_ph_info {
    uint32_t usable, l, m, j, jpad, pitch, row_len, smem_bytes;
} apt_ph_info;
int apt_ph_plan(uint32_t l, uint32_t m, const float *taps, size_t ntaps, apt_ph_info *info, float *table, size_t cap_table,
                uint16_t *xs, size_t cap_xs);

/* ------------------------------------------------------------------ batch */

/* Decode `count` independent recordings (all at `input_rate`, same settings), sharded
 * recording i -> device devices[i % ndevices]; one feeder thread per device (bound to the device's NUMA node)
 * keeps streams_per_device jobs in flight and takes them back in completion order;
 * no inter-device communication.  signals[i]/lens[i] are host buffers; outs[i] (capacity caps[i]
 * floats) receive the rows, nouts[i] the counts and statuses[i] the per-recording status.
 * Returns APT_OK if every recording decoded, else the first failing status. */
int apt_decode_batch(const void *const *signals, int format, const uint64_t *lens, int count,
                     uint32_t input_rate, const apt_settings *s, int sync,
                     float *const *outs, const uint64_t *caps, uint64_t *nouts, int *statuses,
                     const int *devices, int ndevices, int streams_per_device);

#ifdef __cplusplus
}
#endif
#endif /* APTB200_H */
