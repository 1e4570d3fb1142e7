"""SURVEY.md §8 (f)3 on the device: contrast bounds, telemetry statistics and the u8 map (kernels_post.cuh) against the
oracle's restatement of noaa_apt.rs:132-190,249-259, misc.rs:119-175, telemetry.rs:30-66,125-243.  Given the same f32
rows every result must be bit-identical; through decode() the rows themselves differ from the oracle's by <= 1e-5, so the
end-to-end image may differ by one grey level at rounding boundaries."""
import numpy as np
import pytest

import noaa_apt_b200 as na
from noaa_apt_b200 import image, synth
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rows():
    x = synth.apt_signal(11025, 150, seed=12)      # 300 rows: enough for a telemetry frame (200 rows)
    return oracle.decode(x, 11025), x


def test_reference_test_map_vector():
    # noaa_apt.rs:266-281 test_map
    vals = np.array([-10., -5., -1., 0., 1., 2.4, 50., 120., 199.6, 255., 256., 300.], np.float32)
    shifted = (vals * np.float32(123.123) - np.float32(234.234)).astype(np.float32)
    low = np.float32(0.) * np.float32(123.123) - np.float32(234.234)
    high = np.float32(255.) * np.float32(123.123) - np.float32(234.234)
    got = image.map_signal_u8(shifted, float(low), float(high))
    assert got.tolist() == [0, 0, 0, 0, 1, 2, 50, 120, 200, 255, 255, 255]
    assert np.array_equal(got, oracle.map_signal_u8(shifted, float(low), float(high)))


def test_map_signal_u8_bit_exact(rows):
    r, _ = rows
    lo, hi = oracle.minmax(r)
    for low, high in ((lo, hi), (float(np.percentile(r, 2)), float(np.percentile(r, 98))), (0.0, 1.0)):
        assert np.array_equal(image.map_signal_u8(r, low, high), oracle.map_signal_u8(r, low, high))


def test_contrast_minmax_and_percent_bit_exact(rows):
    r, _ = rows
    lo, hi, _ = image.contrast_bounds(r, image.MINMAX)
    assert (lo, hi) == oracle.minmax(r)
    for p in (1.0, 0.98, 0.95, 0.9, 0.5):
        lo, hi, _ = image.contrast_bounds(r, image.PERCENT, p)
        assert (lo, hi) == oracle.percent(r, p), p
    # misc.rs:515-543 test_percent: uniform 0..9999
    u = np.arange(10000, dtype=np.float32)
    for p in (1.0, 0.95, 0.90, 0.80, 0.50):
        lo, hi, _ = image.contrast_bounds(u, image.PERCENT, p)
        assert (lo, hi) == oracle.percent(u, p)
        rem = (1 - p) / 2
        assert rem - 0.005 < lo / 10000 < rem + 0.005 and 1 - (rem + 0.005) < hi / 10000 < 1 - (rem - 0.005)


def test_telemetry_rows_and_frame_bit_exact(rows):
    r, _ = rows
    a, b, v = image.telemetry_rows(r)
    ra, rb, rv = oracle.telemetry_rows(r)
    assert np.array_equal(a, ra) and np.array_equal(b, rb) and np.array_equal(v, rv)
    lo, hi, info = image.contrast_bounds(r, image.TELEMETRY)
    wa, wb, best = oracle.read_telemetry(r)
    assert info["telemetry_row"] == best
    assert np.array_equal(info["wedges_a"], wa) and np.array_equal(info["wedges_b"], wb)
    assert lo == (np.float32(wa[8]) + np.float32(wb[8])) / np.float32(2) and hi == (np.float32(wa[7]) + np.float32(wb[7])) / np.float32(2)
    with pytest.raises(na.err.Internal):        # "Recording too short for telemetry decoding"
        image.contrast_bounds(r[: 2080 * 150], image.TELEMETRY)


@pytest.mark.parametrize("contrast,percent", [(image.MINMAX, 0.0), (image.PERCENT, 0.98), (image.TELEMETRY, 0.0)])
def test_decode_image_u8_end_to_end(rows, contrast, percent):
    ref_rows, x = rows
    img, info = image.decode_image_u8(na.Context(), na.Settings(), x, 11025, True, contrast, percent)
    assert img.shape == (ref_rows.size // 2080, 2080) and info["rows"] == img.shape[0]
    # bit-exact against the oracle applied to the GPU's own f32 rows ...
    gpu_rows = na.decode(na.Context(), na.Settings(), x, 11025, True)
    if contrast == image.MINMAX:
        low, high = oracle.minmax(gpu_rows)
    elif contrast == image.PERCENT:
        low, high = oracle.percent(gpu_rows, percent)
    else:
        wa, wb, _ = oracle.read_telemetry(gpu_rows)
        low = float((np.float32(wa[8]) + np.float32(wb[8])) / np.float32(2))
        high = float((np.float32(wa[7]) + np.float32(wb[7])) / np.float32(2))
    assert (info["low"], info["high"]) == (low, high)
    assert np.array_equal(img.ravel(), oracle.map_signal_u8(gpu_rows, low, high))
    # ... and within one grey level of the all-oracle pipeline (the rows agree to 1e-5)
    if contrast == image.MINMAX:
        rl, rh = oracle.minmax(ref_rows)
    elif contrast == image.PERCENT:
        rl, rh = oracle.percent(ref_rows, percent)
    else:
        wa, wb, _ = oracle.read_telemetry(ref_rows)
        rl = float((np.float32(wa[8]) + np.float32(wb[8])) / np.float32(2))
        rh = float((np.float32(wa[7]) + np.float32(wb[7])) / np.float32(2))
    ref_img = oracle.map_signal_u8(ref_rows, rl, rh)
    diff = np.abs(img.ravel().astype(np.int16) - ref_img.astype(np.int16))
    assert diff.max() <= 1 and np.count_nonzero(diff) < 0.002 * diff.size
    # PCM16 input and the decoder-object form give the same image
    pcm = synth.apt_pcm16(11025, 150, seed=12)
    img16, _ = image.decode_image_u8(na.Context(), na.Settings(), pcm, 11025, True, contrast, percent)
    assert np.array_equal(img16, img)


def test_quantize_i16_and_resample_tool(tmp_path):
    """wav.rs:71-85 on the device, bit-exact; resample.rs:17-71 end to end on a WAV written by the stdlib."""
    import wave
    from noaa_apt_b200 import wav
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(50000) * 3000).astype(np.float32)
    assert np.array_equal(wav.quantize_i16(x), oracle.quantize_i16(x))
    assert np.array_equal(wav.quantize_i16(-np.abs(x) - 1), oracle.quantize_i16(-np.abs(x) - 1))   # negative maximum: saturation
    pcm = synth.apt_pcm16(11025, 3, seed=1)
    src, dst = str(tmp_path / "in.wav"), str(tmp_path / "out.wav")
    with wave.open(src, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(11025); w.writeframes(pcm.tobytes())
    for out_rate in (48000, 6000, 3675):                       # test/test.sh:47-49
        n = wav.resample_wav(src, dst, out_rate, 40.0, 0.1)
        ref = oracle.resample(pcm.astype(np.float32), 11025, out_rate, 40.0, 0.1)
        got, rate = wav.load_wav_pcm16(dst)
        assert rate == out_rate and n == ref.size == got.size
        refq = oracle.quantize_i16(ref)
        assert np.max(np.abs(got.astype(np.int32) - refq.astype(np.int32))) <= 1     # the resampled values agree to 1e-5
