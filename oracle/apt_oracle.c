/*
 * apt_oracle.c -- CPU ORACLE (test infrastructure, not product; see apt_oracle.h).
 *
 * Scalar f32 restatement of martinber/noaa-apt v1.4.1, src/{frequency,misc,
 * filters,dsp,decode}.rs.  All citations are file:line into the reference.
 * "parity unpinned" for fast_resampling/demodulate/filter/find_sync/decode:
 * the reference holds no golden output for them and cannot be run here.
 *
 * Rules followed:
 *   - every f32 operation is rounded on its own (compile with -ffp-contract=off);
 *   - x.powi(2) == x*x;  `n as f32` is round-to-nearest;
 *   - sin/cos/powf are the platform libm's (Rust std forwards to them on Linux);
 *   - indices in fast_resampling are u64 (dsp.rs:194-206).
 */
#include "apt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::f32::consts::PI */
static const float PI_F = 3.14159265358979323846f;

/* decode.rs:14-38 */
#define FINAL_RATE 4160u
#define PX_PER_ROW 2080u
#define CARRIER_FREQ 2400u

void oracle_free(void *p) { free(p); }

void oracle_default_settings(oracle_settings *s) {
    /* default_settings.toml:108-116 */
    s->work_rate = 12480;
    s->resample_atten = 30.f;
    s->resample_delta_freq = 1000.f;
    s->resample_cutout = 4800.f;
    s->demodulation_atten = 25.f;
}

/* ------------------------------------------------------------------ Freq */

/* frequency.rs:68-72: pi_rad = 2. * f / rate as f32 */
float oracle_freq_hz(float f, uint32_t rate) { return 2.f * f / (float)rate; }
/* frequency.rs:58-60 */
float oracle_freq_rad(float f) { return f / PI_F; }
/* frequency.rs:75-77 */
float oracle_freq_get_rad(float pi_rad) { return pi_rad * PI_F; }
/* frequency.rs:85-87: pi_rad * rate as f32 / 2. */
float oracle_freq_get_hz(float pi_rad, uint32_t rate) { return pi_rad * (float)rate / 2.f; }

/* ---------------------------------------------------------------- Bessel */

/* misc.rs:20-41, 1 / (n! * 2^n)^2 */
static const float BESSEL_TABLE[9] = {
    /* the reference table has 20 entries; bessel_i0 reads only [1..=8] (limit = 8, misc.rs:49) */
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

/* misc.rs:47-57 */
float oracle_bessel_i0(float x) {
    float result = 0.f;
    for (int k = 8; k >= 1; --k) {
        result += BESSEL_TABLE[k];
        result *= x * x;
    }
    return result + 1.f;
}

/* --------------------------------------------------------------- filters */

/* filters.rs:144-183 */
float *oracle_kaiser(float atten, float delta_w_pi, size_t *n) {
    float beta;
    if (atten > 50.f) {
        beta = 0.1102f * (atten - 8.7f);
    } else if (atten < 21.f) {
        beta = 0.f;
    } else {
        beta = 0.5842f * powf(atten - 21.f, 0.4f) + 0.07886f * (atten - 21.f);
    }

    /* filters.rs:164-167 */
    int32_t length = (int32_t)ceilf((atten - 8.f) / (2.285f * oracle_freq_get_rad(delta_w_pi))) + 1;
    if (length % 2 == 0) length += 1;
    if (length <= 0) { *n = 0; return NULL; }

    float *window = (float *)malloc((size_t)length * sizeof(float));
    if (!window) { *n = 0; return NULL; }

    size_t i = 0;
    for (int32_t k = -(length - 1) / 2; k <= (length - 1) / 2; ++k) {
        float nf = (float)k;
        float m = (float)length;
        float q = nf / (m / 2.f);
        window[i++] = oracle_bessel_i0(beta * sqrtf(1.f - q * q)) / oracle_bessel_i0(beta);
    }
    *n = (size_t)length;
    return window;
}

/* filters.rs:48-54 (NoFilter), 56-88 (Lowpass), 97-132 (LowpassDcRemoval), 186-196 (product) */
float *oracle_design(int kind, float cutout_pi, float atten, float delta_w_pi, size_t *n) {
    if (kind == ORACLE_FILTER_NONE) {
        float *f = (float *)malloc(sizeof(float));
        if (!f) { *n = 0; return NULL; }
        f[0] = 1.f;
        *n = 1;
        return f;
    }

    size_t wn = 0;
    float *window = oracle_kaiser(atten, delta_w_pi, &wn);
    if (!window) { *n = 0; return NULL; }

    float *filter = (float *)malloc(wn * sizeof(float));
    if (!filter) { free(window); *n = 0; return NULL; }

    int32_t m = (int32_t)wn;
    size_t i = 0;
    if (kind == ORACLE_FILTER_LOWPASS) {
        for (int32_t k = -(m - 1) / 2; k <= (m - 1) / 2; ++k) {
            if (k == 0) {
                filter[i++] = cutout_pi;
            } else {
                float nf = (float)k;
                filter[i++] = sinf(nf * PI_F * cutout_pi) / (nf * PI_F);
            }
        }
    } else {
        float half = delta_w_pi / 2.f; /* (self.delta_w / 2.).get_pi_rad() */
        for (int32_t k = -(m - 1) / 2; k <= (m - 1) / 2; ++k) {
            if (k == 0) {
                filter[i++] = cutout_pi - half;
            } else {
                float nf = (float)k;
                filter[i++] = sinf(nf * PI_F * cutout_pi) / (nf * PI_F)
                            - sinf(nf * PI_F * half) / (nf * PI_F);
            }
        }
    }
    for (size_t j = 0; j < wn; ++j) filter[j] *= window[j];
    free(window);
    *n = wn;
    return filter;
}

/* ------------------------------------------------------------------- dsp */

/* dsp.rs:186-289 with context.export_resample_filtered == false */
float *oracle_fast_resampling(const float *signal, uint64_t len, uint32_t l32, uint32_t m32,
                              const float *coeff, size_t ncoeff, uint64_t *nout) {
    uint64_t l = l32, m = m32;
    uint64_t interpolated_len = len * l;                  /* dsp.rs:203 */
    uint64_t offset = ((uint64_t)ncoeff - 1) / 2;         /* dsp.rs:226 */

    /* Vec::with_capacity(interpolated_len / m) is only a capacity (dsp.rs:206-208);
     * the exact count is ceil((interpolated_len - offset) / m). */
    uint64_t count = interpolated_len > offset ? (interpolated_len - offset + m - 1) / m : 0;
    float *output = (float *)malloc((size_t)(count ? count : 1) * sizeof(float));
    if (!output) { *nout = 0; return NULL; }

    uint64_t k = 0;
    uint64_t t = offset;                                  /* dsp.rs:230 */
    while (t < interpolated_len) {                        /* dsp.rs:234 */
        uint64_t n;
        if (t > offset) {                                 /* dsp.rs:237-248 */
            n = t - offset;
            uint64_t rem = n % l;
            if (rem != 0) n += l - rem;
        } else {
            n = 0;
        }
        float sum = 0.f;                                  /* dsp.rs:252 */
        uint64_t x = n / l;
        while (n <= t + offset) {                         /* dsp.rs:254-263 */
            if (x < len) sum += coeff[n + offset - t] * signal[x];
            x += 1;
            n += l;
        }
        output[k++] = sum;                                /* dsp.rs:276 */
        t += m;                                           /* dsp.rs:277 */
    }
    *nout = k;
    return output;
}

/* dsp.rs:294-307 */
float *oracle_decimate(const float *x, uint64_t len, uint32_t m, uint64_t *nout) {
    uint64_t count = len / m;
    float *out = (float *)malloc((size_t)(count ? count : 1) * sizeof(float));
    if (!out) { *nout = 0; return NULL; }
    for (uint64_t i = 0; i < count; ++i) out[i] = x[i * m];
    *nout = count;
    return out;
}

/* dsp.rs:350-383 */
int oracle_demodulate(const float *signal, uint64_t len, float carrier_pi_rad, float *output) {
    if (len == 0) return ORACLE_ERR_BAD_ARG;              /* signal[0] panics, dsp.rs:367 */
    float phi = 2.f * oracle_freq_get_rad(carrier_pi_rad); /* dsp.rs:360 */
    float cosphi2 = cosf(phi) * 2.f;                      /* dsp.rs:362 */
    float sinphi = sinf(phi);                             /* dsp.rs:363 */

    output[0] = 0.f;                                      /* vec![0; len], dsp.rs:357 */
    float prev = signal[0];
    float prev_sq = signal[0] * signal[0];
    for (uint64_t i = 1; i < len; ++i) {
        float curr = signal[i];
        float curr_sq = signal[i] * signal[i];
        /* dsp.rs:373: (prev_sq + curr_sq - (prev * curr * cosphi2)).sqrt() / sinphi */
        output[i] = sqrtf(prev_sq + curr_sq - (prev * curr * cosphi2)) / sinphi;
        prev = curr;
        prev_sq = curr_sq;
    }
    return ORACLE_OK;
}

/* dsp.rs:386-410 */
int oracle_filter(const float *signal, uint64_t len, const float *coeff, size_t ncoeff, float *output) {
    for (uint64_t i = 0; i < len; ++i) {
        float sum = 0.f;
        for (size_t j = 0; j < ncoeff; ++j) {
            if (i > j) sum += signal[i - j] * coeff[j];   /* strict i > j, dsp.rs:399 */
        }
        output[i] = sum;
    }
    return ORACLE_OK;
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {         /* gcd crate 2.3.0, dsp.rs:73 */
    while (b) { uint32_t t = a % b; a = b; b = t; }
    return a;
}

/* dsp.rs:62-126 */
int oracle_resample_with_filter(const float *x, uint64_t len, uint32_t in_rate, uint32_t out_rate,
                                int kind, float cutout_pi, float atten, float delta_w_pi,
                                float **out, uint64_t *nout) {
    *out = NULL; *nout = 0;
    if (out_rate == 0) return ORACLE_ERR_RESAMPLE_TO_ZERO; /* dsp.rs:69-71 */
    if (in_rate == 0) return ORACLE_ERR_BAD_ARG;           /* m == 0 -> division panic */

    uint32_t g = gcd_u32(in_rate, out_rate);
    uint32_t l = out_rate / g;                            /* dsp.rs:74 */
    uint32_t m = in_rate / g;                             /* dsp.rs:75 */

    if (l > 1) {
        /* dsp.rs:82-91 */
        uint64_t wide = (uint64_t)in_rate * (uint64_t)l;
        if (wide > 0xFFFFFFFFull) return ORACLE_ERR_RATE_OVERFLOW;
        uint32_t interpolated_rate = (uint32_t)wide;

        /* filt.resample(input_rate, interpolated_rate): filters.rs:90-94 / 134-138 */
        if (kind != ORACLE_FILTER_NONE) {
            float ratio = (float)interpolated_rate / (float)in_rate;
            cutout_pi /= ratio;
            delta_w_pi /= ratio;
        }
        size_t nc = 0;
        float *coeff = oracle_design(kind, cutout_pi, atten, delta_w_pi, &nc); /* dsp.rs:94 */
        if (!coeff) return ORACLE_ERR_NOMEM;
        *out = oracle_fast_resampling(x, len, l, m, coeff, nc, nout);          /* dsp.rs:98 */
        free(coeff);
        return *out ? ORACLE_OK : ORACLE_ERR_NOMEM;
    }

    /* dsp.rs:105-123 */
    size_t nc = 0;
    float *coeff = oracle_design(kind, cutout_pi, atten, delta_w_pi, &nc);
    if (!coeff) return ORACLE_ERR_NOMEM;
    float *filtered = (float *)malloc((size_t)(len ? len : 1) * sizeof(float));
    if (!filtered) { free(coeff); return ORACLE_ERR_NOMEM; }
    oracle_filter(x, len, coeff, nc, filtered);
    free(coeff);
    *out = oracle_decimate(filtered, len, m, nout);
    free(filtered);
    return *out ? ORACLE_OK : ORACLE_ERR_NOMEM;
}

/* dsp.rs:132-162 */
int oracle_resample(const float *x, uint64_t len, uint32_t in_rate, uint32_t out_rate,
                    float atten, float delta_w_pi, float **out, uint64_t *nout) {
    float cutout;
    if (out_rate > in_rate) {
        cutout = oracle_freq_hz((float)in_rate / 2.f, in_rate);   /* dsp.rs:144 */
    } else {
        cutout = oracle_freq_hz((float)out_rate / 2.f, in_rate);  /* dsp.rs:148 */
    }
    return oracle_resample_with_filter(x, len, in_rate, out_rate, ORACLE_FILTER_LOWPASS,
                                       cutout, atten, delta_w_pi, out, nout);
}

/* ---------------------------------------------------------------- decode */

/* decode.rs:171-199 */
int oracle_generate_sync_frame(uint32_t work_rate, int8_t **out, size_t *n) {
    *out = NULL; *n = 0;
    if (work_rate % FINAL_RATE != 0) return ORACLE_ERR_WORK_RATE;
    size_t pixel_width = work_rate / FINAL_RATE;
    size_t sync_pulse_width = pixel_width * 2;
    size_t total = sync_pulse_width + 7 * 2 * sync_pulse_width + 8 * pixel_width;
    int8_t *g = (int8_t *)malloc(total ? total : 1);
    if (!g) return ORACLE_ERR_NOMEM;
    size_t i = 0;
    for (size_t j = 0; j < sync_pulse_width; ++j) g[i++] = -1;
    for (size_t j = 0; j < 7 * 2 * sync_pulse_width; ++j) {
        /* cycle of (-1 x spw, +1 x spw) truncated to 14*spw */
        g[i++] = ((j / sync_pulse_width) % 2 == 0) ? -1 : 1;
    }
    for (size_t j = 0; j < 8 * pixel_width; ++j) g[i++] = -1;
    *out = g; *n = total;
    return ORACLE_OK;
}

typedef struct { uint64_t pos; float val; } peak_t;

/* decode.rs:204-263 */
int oracle_find_sync(const float *signal, uint64_t len, uint32_t work_rate,
                     uint64_t **pos, size_t *npos, float *corr_out) {
    *pos = NULL; *npos = 0;
    int8_t *guard = NULL; size_t glen = 0;
    int rc = oracle_generate_sync_frame(work_rate, &guard, &glen);
    if (rc != ORACLE_OK) return rc;
    if (len < glen) { free(guard); return ORACLE_ERR_BAD_ARG; } /* usize underflow panics, decode.rs:225 */

    size_t cap = 1024, npeaks = 0;
    peak_t *peaks = (peak_t *)malloc(cap * sizeof(peak_t));
    if (!peaks) { free(guard); return ORACLE_ERR_NOMEM; }
    peaks[npeaks].pos = 0; peaks[npeaks].val = 0.f; npeaks++;   /* decode.rs:208-209 */

    uint64_t samples_per_work_row = (uint64_t)PX_PER_ROW * work_rate / FINAL_RATE; /* decode.rs:212 */
    uint64_t min_distance = samples_per_work_row * 8 / 10;                         /* decode.rs:216 */

    for (uint64_t i = 0; i < len - glen; ++i) {            /* decode.rs:225 */
        float corr = 0.f;
        for (size_t j = 0; j < glen; ++j) {                /* decode.rs:227-233 */
            if (guard[j] == 1) corr += signal[i + j];
            else corr -= signal[i + j];
        }
        if (corr_out) corr_out[i] = corr;

        if (i - peaks[npeaks - 1].pos > min_distance) {    /* decode.rs:241 */
            while (i / samples_per_work_row > npeaks) {    /* decode.rs:244 */
                if (npeaks == cap) {
                    cap *= 2;
                    peak_t *np = (peak_t *)realloc(peaks, cap * sizeof(peak_t));
                    if (!np) { free(peaks); free(guard); return ORACLE_ERR_NOMEM; }
                    peaks = np;
                }
                peaks[npeaks].pos = i; peaks[npeaks].val = corr; npeaks++;
            }
        } else if (corr > peaks[npeaks - 1].val) {         /* decode.rs:250 */
            peaks[npeaks - 1].pos = i; peaks[npeaks - 1].val = corr;
        }
    }
    free(guard);

    uint64_t *p = (uint64_t *)malloc(npeaks * sizeof(uint64_t));
    if (!p) { free(peaks); return ORACLE_ERR_NOMEM; }
    for (size_t i = 0; i < npeaks; ++i) p[i] = peaks[i].pos;
    free(peaks);
    *pos = p; *npos = npeaks;
    return ORACLE_OK;
}

void oracle_steps_free(oracle_steps *s) {
    if (!s) return;
    free(s->resampled); free(s->demodulated); free(s->filtered);
    free(s->sync_pos); free(s->aligned);
    memset(s, 0, sizeof(*s));
}

/* decode.rs:43-162 */
int oracle_decode_steps(const float *x, uint64_t len, uint32_t in_rate, const oracle_settings *s,
                        int sync, float **out, uint64_t *nout, oracle_steps *steps) {
    *out = NULL; *nout = 0;
    oracle_steps local; memset(&local, 0, sizeof(local));
    if (steps) memset(steps, 0, sizeof(*steps));
    if (s->work_rate == 0) return ORACLE_ERR_RESAMPLE_TO_ZERO;

    uint32_t samples_per_work_row = PX_PER_ROW * s->work_rate / FINAL_RATE;  /* decode.rs:55 */
    uint32_t work_rate = s->work_rate;
    int rc;

    /* decode.rs:65-77 */
    float cutout = oracle_freq_hz(s->resample_cutout, in_rate);
    float delta_w = oracle_freq_hz(s->resample_delta_freq, in_rate);
    rc = oracle_resample_with_filter(x, len, in_rate, work_rate, ORACLE_FILTER_LOWPASS_DC,
                                     cutout, s->resample_atten, delta_w,
                                     &local.resampled, &local.n_resampled);
    if (rc != ORACLE_OK) goto fail;

    /* decode.rs:79-83 */
    if (local.n_resampled < 10ull * samples_per_work_row) { rc = ORACLE_ERR_TOO_SHORT; goto fail; }

    /* decode.rs:89 */
    local.n_demodulated = local.n_resampled;
    local.demodulated = (float *)malloc((size_t)local.n_demodulated * sizeof(float));
    if (!local.demodulated) { rc = ORACLE_ERR_NOMEM; goto fail; }
    rc = oracle_demodulate(local.resampled, local.n_resampled,
                           oracle_freq_hz((float)CARRIER_FREQ, work_rate), local.demodulated);
    if (rc != ORACLE_OK) goto fail;

    /* decode.rs:95-102 */
    {
        float c = (float)FINAL_RATE / (float)work_rate;   /* Freq::pi_rad(FINAL_RATE as f32 / work_rate as f32) */
        size_t nc = 0;
        float *coeff = oracle_design(ORACLE_FILTER_LOWPASS, c, s->demodulation_atten, c / 5.f, &nc);
        if (!coeff) { rc = ORACLE_ERR_NOMEM; goto fail; }
        local.n_filtered = local.n_demodulated;
        local.filtered = (float *)malloc((size_t)local.n_filtered * sizeof(float));
        if (!local.filtered) { free(coeff); rc = ORACLE_ERR_NOMEM; goto fail; }
        oracle_filter(local.demodulated, local.n_demodulated, coeff, nc, local.filtered);
        free(coeff);
    }

    if (sync) {
        /* decode.rs:110-134 */
        rc = oracle_find_sync(local.filtered, local.n_filtered, work_rate,
                              &local.sync_pos, &local.n_sync_pos, NULL);
        if (rc != ORACLE_OK) goto fail;
        if (local.n_sync_pos < 5) { rc = ORACLE_ERR_FEW_SYNC_FRAMES; goto fail; }

        uint64_t rows = 0;
        for (size_t i = 0; i + 1 < local.n_sync_pos; ++i)
            if (local.sync_pos[i] + samples_per_work_row < local.n_filtered) rows++;
        local.n_aligned = rows * samples_per_work_row;
        local.aligned = (float *)malloc((size_t)(local.n_aligned ? local.n_aligned : 1) * sizeof(float));
        if (!local.aligned) { rc = ORACLE_ERR_NOMEM; goto fail; }
        uint64_t w = 0;
        for (size_t i = 0; i + 1 < local.n_sync_pos; ++i) {
            if (local.sync_pos[i] + samples_per_work_row < local.n_filtered) {   /* decode.rs:127 */
                memcpy(local.aligned + w, local.filtered + local.sync_pos[i],
                       samples_per_work_row * sizeof(float));
                w += samples_per_work_row;
            }
        }
    } else {
        /* decode.rs:141-147 */
        local.n_aligned = local.n_filtered / samples_per_work_row * samples_per_work_row;
        local.aligned = (float *)malloc((size_t)(local.n_aligned ? local.n_aligned : 1) * sizeof(float));
        if (!local.aligned) { rc = ORACLE_ERR_NOMEM; goto fail; }
        memcpy(local.aligned, local.filtered, (size_t)local.n_aligned * sizeof(float));
    }

    /* decode.rs:158-159: resample_with_filter(work_rate -> 4160, NoFilter) */
    rc = oracle_resample_with_filter(local.aligned, local.n_aligned, work_rate, FINAL_RATE,
                                     ORACLE_FILTER_NONE, 0.f, 0.f, 0.f, out, nout);
    if (rc != ORACLE_OK) goto fail;

    if (steps) *steps = local; else oracle_steps_free(&local);
    return ORACLE_OK;

fail:
    oracle_steps_free(&local);
    return rc;
}

int oracle_decode(const float *x, uint64_t len, uint32_t in_rate, const oracle_settings *s,
                  int sync, float **out, uint64_t *nout) {
    return oracle_decode_steps(x, len, in_rate, s, sync, out, nout, NULL);
}

/* ------------------------------------------------- either side of the path */

/* wav.rs:31-40 (hound yields i32, then `*x as f32`) */
void oracle_pcm16_to_f32(const int16_t *in, uint64_t n, float *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = (float)(int32_t)in[i];
}

/* noaa_apt.rs:249-259 */
void oracle_map_signal_u8(const float *x, uint64_t n, float low, float high, uint8_t *out) {
    float range = high - low;
    for (uint64_t i = 0; i < n; ++i) {
        float v = (x[i] - low) / range * 255.f;
        /* .max(0.).min(255.): f32::max/min return the non-NaN operand */
        v = fmaxf(v, 0.f);
        v = fminf(v, 255.f);
        v = roundf(v);                                    /* half away from zero */
        out[i] = (uint8_t)v;                              /* `as u8`, already in range; NaN -> 0 handled by fmaxf */
    }
}

/* wav.rs:71-85: (*sample / max * (i16::MAX as f32)) as i16 with max = dsp::get_max */
int oracle_quantize_i16(const float *x, uint64_t n, int16_t *out) {
    if (n == 0) return ORACLE_ERR_BAD_ARG;                /* get_max errors on empty, dsp.rs:21-25 */
    float max = x[0];
    for (uint64_t i = 0; i < n; ++i) if (x[i] > max) max = x[i];   /* dsp.rs:27-32 */
    for (uint64_t i = 0; i < n; ++i) {
        float v = x[i] / max * 32767.f;
        /* Rust `as i16`: truncate toward zero, saturate, NaN -> 0 */
        int16_t q;
        if (v != v) q = 0;
        else if (v >= 32767.f) q = 32767;
        else if (v <= -32768.f) q = -32768;
        else q = (int16_t)v;
        out[i] = q;
    }
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Next-row (f)3: contrast bounds and telemetry statistics of the decoded image (after the hot path).
 * ---------------------------------------------------------------------------------------------- */

/* dsp.rs:20-54 get_min / get_max: first-wins strict comparisons */
static float sig_min(const float *x, uint64_t n) {
    float m = x[0];
    for (uint64_t i = 0; i < n; ++i) if (x[i] < m) m = x[i];
    return m;
}
static float sig_max(const float *x, uint64_t n) {
    float m = x[0];
    for (uint64_t i = 0; i < n; ++i) if (x[i] > m) m = x[i];
    return m;
}

int oracle_minmax(const float *x, uint64_t n, float *low, float *high) {
    if (n == 0) return ORACLE_ERR_BAD_ARG;                 /* Internal("Can't get maximum of a zero length vector") */
    *low = sig_min(x, n);
    *high = sig_max(x, n);
    return ORACLE_OK;
}

/* misc.rs:119-175 percent(): 1000 buckets between min and max; buckets[trunc((x-min)/range*1000)] clamped */
int oracle_percent_buckets(const float *x, uint64_t n, uint32_t *buckets /* [1000] */, float *mn, float *mx) {
    if (n == 0) return ORACLE_ERR_BAD_ARG;
    const float min = sig_min(x, n), max = sig_max(x, n);
    const float total_range = max - min;
    memset(buckets, 0, 1000 * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) {
        const float t = truncf((x[i] - min) / total_range * 1000.f);
        /* `as usize` saturates: negative / NaN -> 0, huge -> usize::MAX; then .max(0).min(999) */
        uint32_t b;
        if (!(t >= 0.f)) b = 0;
        else if (t >= 999.f) b = 999;
        else b = (uint32_t)t;
        buckets[b] += 1;
    }
    *mn = min;
    *mx = max;
    return ORACLE_OK;
}

int oracle_percent(const float *x, uint64_t n, float percent, float *low, float *high) {
    if (percent < 0.f || percent > 1.f) return ORACLE_ERR_BAD_ARG;   /* Internal("Percent given should be between 0 and 1") */
    uint32_t buckets[1000];
    float min, max;
    int rc = oracle_percent_buckets(x, n, buckets, &min, &max);
    if (rc != ORACLE_OK) return rc;
    const float remainder = (1.f - percent) / 2.f;
    const float total_range = max - min;
    uint32_t accum = 0;
    int low_bucket = -1, high_bucket = -1;
    for (int b = 0; b < 1000; ++b) {
        accum += buckets[b];
        const float frac = (float)accum / (float)n;
        if (low_bucket < 0 && frac > remainder) low_bucket = b;
        else if (high_bucket < 0 && frac > 1.f - remainder) high_bucket = b;
    }
    if (high_bucket < 0) high_bucket = 999;
    if (low_bucket < 0) return ORACLE_ERR_BAD_ARG;          /* low_bucket.unwrap() panics */
    *low = (float)low_bucket / 1000.f * total_range + min;
    *high = (float)high_bucket / 1000.f * total_range + min;
    return ORACLE_OK;
}

/* telemetry.rs:147-170: per image row the means of the two telemetry bands and their pooled variance */
int oracle_telemetry_rows(const float *x, uint64_t n, float *mean_a, float *mean_b, float *variance) {
    const uint64_t rows = n / 2080;
    for (uint64_t r = 0; r < rows; ++r) {
        const float *line = x + r * 2080;
        const float *a = line + 994, *b = line + 2034;
        float sa = 0.f, sb = 0.f;
        for (int i = 0; i < 44; ++i) sa += a[i];
        for (int i = 0; i < 44; ++i) sb += b[i];
        const float ma = sa / 44.f, mb = sb / 44.f;
        float va = 0.f, vb = 0.f;
        for (int i = 0; i < 44; ++i) { const float d = a[i] - ma; va += d * d; }
        for (int i = 0; i < 44; ++i) { const float d = b[i] - mb; vb += d * d; }
        mean_a[r] = ma;
        mean_b[r] = mb;
        variance[r] = (va + vb) / 88.f;
    }
    return ORACLE_OK;
}

/* telemetry.rs:125-243 read_telemetry + :30-66 Telemetry::from_bands: the best frame start and the 16 wedge values of
 * each channel; contrast bounds are low = wedge 9, high = wedge 8 averaged over both channels (noaa_apt.rs:143-149). */
int oracle_read_telemetry(const float *x, uint64_t n, float *wedges_a /* [16] */, float *wedges_b /* [16] */,
                          uint64_t *best_row) {
    static const float pattern[25] = {31.f, 63.f, 95.f, 127.f, 159.f, 191.f, 224.f, 255.f, 0.f,
                                      0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                                      31.f, 63.f, 95.f, 127.f, 159.f, 191.f, 224.f, 255.f, 0.f};
    float sample[200];
    for (int i = 0; i < 200; ++i) sample[i] = pattern[i / 8];
    const uint64_t rows = n / 2080;
    if (rows < 200) return ORACLE_ERR_TOO_SHORT;           /* Internal("Recording too short for telemetry decoding") */
    float *ma = (float *)malloc(rows * sizeof(float)), *mb = (float *)malloc(rows * sizeof(float)),
          *var = (float *)malloc(rows * sizeof(float));
    if (!ma || !mb || !var) { free(ma); free(mb); free(var); return ORACLE_ERR_NOMEM; }
    oracle_telemetry_rows(x, n, ma, mb, var);
    uint64_t best = 0;
    float best_q = 0.f;
    for (uint64_t i = 0; i + 200 < rows; ++i) {
        float sum = 0.f;
        for (int j = 0; j < 200; ++j) {
            sum += sample[j] * ma[i + j];
            sum += sample[j] * mb[i + j];
        }
        float dev = 0.f;
        for (int j = 0; j < 200; ++j) dev += sqrtf(var[i + j]);
        const float q = sum / dev;
        if (q > best_q) { best = i; best_q = q; }
    }
    /* from_bands: means of 8 contiguous rows from `best`, 16 + 9 wedges (fewer if the image ends) */
    float wa[25], wb[25];
    int nw = 0;
    for (; nw < 25 && best + 8ull * (nw + 1) <= rows; ++nw) {
        float sa = 0.f, sb = 0.f;
        for (int k = 0; k < 8; ++k) sa += ma[best + 8 * nw + k];
        for (int k = 0; k < 8; ++k) sb += mb[best + 8 * nw + k];
        wa[nw] = sa / 8.f;
        wb[nw] = sb / 8.f;
    }
    free(ma); free(mb); free(var);
    if (nw < 25) return ORACLE_ERR_BAD_ARG;                /* index out of bounds panic in from_bands */
    for (int w = 1; w <= 16; ++w) {
        wedges_a[w - 1] = w <= 9 ? (wa[w - 1] + wa[w + 15]) / 2.f : wa[w - 1];
        wedges_b[w - 1] = w <= 9 ? (wb[w - 1] + wb[w + 15]) / 2.f : wb[w - 1];
    }
    *best_row = best;
    return ORACLE_OK;
}
