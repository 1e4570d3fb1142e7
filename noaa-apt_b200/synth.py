"""Deterministic synthetic NOAA APT recordings (2.4 kHz AM sub-carrier).

Used by the tests and by bench.py to build the workloads BASELINE.json names
(there are no recordings on the GPU box and no network).  The line layout
follows the constants of the reference (decode.rs:16-35: sync 39 px, space 47,
image 909, telemetry 45, two channels = 2080 px at 4160 px/s) and the sync-A
pattern of `generate_sync_frame` (decode.rs:188-198) at 1-px resolution.

Samples are quantised to int16 and handed out either as int16 (what a WAV
holds) or as the f32 cast `wav::load_wav` performs (wav.rs:37) -- raw integer
values, not normalised.  Gaussian noise from a counter-based generator
(numpy Philox, keyed by the seed) is mandatory: a clean signal has exact
correlation ties that the sync picker resolves by first-wins (decode.rs:250).
"""
import numpy as np

PX_PER_ROW = 2080
FINAL_RATE = 4160
CARRIER_HZ = 2400.0
_WEDGES = np.array([31, 63, 95, 127, 159, 191, 224, 255, 0, 60, 120, 180, 90, 150, 210, 30], dtype=np.float64)


def _sync_a():
    # 2 low, 7 x (2 low, 2 high), 8 low = 38 px, +1 low px to fill PX_SYNC_FRAME = 39
    px = [0, 0]
    for _ in range(7):
        px += [0, 0, 1, 1]
    px += [0] * 8 + [0]
    return np.array(px, dtype=np.float64) * 255.0


def _sync_b():
    # 7 pulses at 832 Hz => 5-px period (3 high, 2 low), padded with low px to 39
    px = [0, 0, 0, 0]
    for _ in range(7):
        px += [1, 1, 1, 0, 0]
    return np.array(px[:39] + [0] * max(0, 39 - len(px)), dtype=np.float64) * 255.0


_SYNC_A = _sync_a()
_SYNC_B = _sync_b()


def _row_template(seed):
    """Column-only part of a row: sync + space; image/telemetry filled per line."""
    t = np.zeros(PX_PER_ROW, dtype=np.float64)
    t[0:39] = _SYNC_A
    t[39:86] = 8.0          # space A (dark)
    t[1040:1079] = _SYNC_B
    t[1079:1126] = 245.0    # space B (bright)
    return t


def pixels(line, col, seed):
    """Pixel value 0..255 for (line, col) arrays of equal shape (vectorised)."""
    line = np.asarray(line)
    col = np.asarray(col)
    tmpl = _row_template(seed)
    out = tmpl[col]
    ph = 0.37 * (seed % 97)
    # image A: cols 86..994, image B: cols 1126..2034
    in_a = (col >= 86) & (col < 995)
    in_b = (col >= 1126) & (col < 2035)
    ca = col - 86
    cb = col - 1126
    img_a = 127.5 + 70.0 * np.sin(ca / 57.0 + ph) * np.cos(line / 83.0 + 0.5 * ph) \
        + 40.0 * np.sin((ca + 2.0 * line) / 19.0)
    img_b = 127.5 + 85.0 * np.cos(cb / 41.0 - ph) * np.sin(line / 61.0 + ph) \
        + 25.0 * np.cos((cb - line) / 13.0)
    out = np.where(in_a, img_a, out)
    out = np.where(in_b, img_b, out)
    # telemetry: 16 wedges of 8 lines each (telemetry.rs:129-133)
    wedge = _WEDGES[(line // 8) % 16]
    in_t = ((col >= 995) & (col < 1040)) | (col >= 2035)
    out = np.where(in_t, wedge, out)
    return np.clip(out, 0.0, 255.0)


def apt_pcm16(rate_hz, seconds=None, seed=0, n_samples=None, start=0, amplitude=20000.0,
              noise_sigma=200.0, chunk=1 << 22):
    """int16 samples [start, start + n) of the synthetic recording `seed` at `rate_hz`."""
    if n_samples is None:
        n_samples = int(round(rate_hz * seconds))
    out = np.empty(n_samples, dtype=np.int16)
    theta = 0.1 + 0.01 * (seed % 50)
    done = 0
    while done < n_samples:
        m = min(chunk - (start + done) % chunk, n_samples - done)
        n = np.arange(start + done, start + done + m, dtype=np.int64)
        pix = (n * FINAL_RATE) // rate_hz
        line = pix // PX_PER_ROW
        col = pix % PX_PER_ROW
        px = pixels(line, col, seed)
        # phase kept small: (n mod rate) keeps the argument exact in f64
        phase = 2.0 * np.pi * CARRIER_HZ * ((n % rate_hz).astype(np.float64) / rate_hz) + theta
        env = amplitude * (0.05 + 0.87 * px / 255.0)
        # counter-based noise: one Philox stream per (seed, chunk index) so any
        # [start, start+n) slice aligned to `chunk` reproduces the same samples
        rng = np.random.Generator(np.random.Philox(key=[0xA97 + seed, (start + done) // chunk]))
        skip = (start + done) % chunk
        noise = rng.standard_normal(skip + m)[skip:] * noise_sigma if noise_sigma > 0 else 0.0
        v = np.rint(env * np.sin(phase) + noise)
        out[done:done + m] = np.clip(v, -32768, 32767).astype(np.int16)
        done += m
    return out


def apt_signal(rate_hz, seconds=None, seed=0, **kw):
    """The same recording as the f32 `Signal` wav::load_wav would return (wav.rs:37)."""
    return apt_pcm16(rate_hz, seconds, seed, **kw).astype(np.float32)
