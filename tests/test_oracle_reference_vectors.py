"""Pins the CPU oracle against every vector the reference's own tests hold for
this path.  Each test cites the reference test it re-expresses.
"""
import math

import numpy as np
import pytest

import oracle


def ulps(a, b):
    a = np.float32(a).view(np.int32).astype(np.int64)
    b = np.float32(b).view(np.int32).astype(np.int64)
    return abs(int(a) - int(b))


# decode.rs:270-319 test_sample_sync_frame -- exact golden vectors
def _literal_frame(pw):
    # the rows of the reference's literal vectors: 2 rows of -1, 7 x (+1 row, -1 row), 3 more -1 rows;
    # each row is 2*pw values (decode.rs:275-293 for pw=5, :298-316 for pw=2)
    rows = [-1, -1] + [1, -1] * 7 + [-1, -1, -1]
    return [v for r in rows for v in [r] * (2 * pw)]


def test_sync_frame_golden_x5():
    literal = _literal_frame(5)
    assert len(literal) == 190
    got = oracle.generate_sync_frame(4160 * 5)
    assert got.dtype == np.int8
    assert got.tolist() == literal


def test_sync_frame_golden_x2():
    literal = _literal_frame(2)
    assert len(literal) == 76
    assert oracle.generate_sync_frame(4160 * 2).tolist() == literal


def test_sync_frame_needs_multiple_of_final_rate():
    # decode.rs:172-176
    with pytest.raises(oracle.OracleError) as e:
        oracle.generate_sync_frame(11025)
    assert e.value.code == oracle.ERR_WORK_RATE


# misc.rs:494-513 test_bessel_i0 -- GNU Octave values, max_relative = 1e-3
BESSEL_KAT = [
    (0.0, 1.00000000000000), (0.5, 1.06348337074132), (1.0, 1.26606587775201),
    (1.5, 1.64672318977289), (2.0, 2.27958530233607), (2.5, 3.28983914405012),
    (3.0, 4.88079258586502), (3.5, 7.37820343222548), (4.0, 11.3019219521363),
    (4.5, 17.4811718556093), (5.0, 27.2398718236044), (5.5, 42.6946451518478),
    (6.0, 67.2344069764780), (6.5, 106.292858243996), (7.0, 168.593908510290),
]


@pytest.mark.parametrize("x,expected", BESSEL_KAT)
def test_bessel_i0(x, expected):
    got = oracle.bessel_i0(x)
    assert abs(got - expected) <= 1e-3 * max(abs(got), abs(expected))


# frequency.rs:325-416 test_frequency_conversion -- max_ulps = 10
FREQ_EQUIV = [
    (0.435374149659864, 1.367768230134332, 2400.0, 11025),
    (-0.435374149659864, -1.367768230134332, -2400.0, 11025),
    (0.1, 0.3141592653589793, 100.0, 2000),
    (-0.1, -0.3141592653589793, -100.0, 2000),
    (0.0, 0.0, 0.0, 11025),
    (1.0, math.pi, 5512.5, 11025),
    (-1.0, -math.pi, -5512.5, 11025),
    (2.0, 2 * math.pi, 11025.0, 11025),
    (-2.0, -2 * math.pi, -11025.0, 11025),
    (300.0, 300 * math.pi, 150.0, 1),
    (-300.0, -300 * math.pi, -150.0, 1),
]


@pytest.mark.parametrize("pi_rad,rad,hz,rate", FREQ_EQUIV)
def test_frequency_conversion(pi_rad, rad, hz, rate):
    for f in (np.float32(pi_rad), oracle.freq_rad(np.float32(rad)), oracle.freq_hz(np.float32(hz), rate)):
        assert ulps(f, pi_rad) <= 10
        assert ulps(oracle.freq_get_rad(f), rad) <= 10
        assert ulps(oracle.freq_get_hz(f, rate), hz) <= 10


# filters.rs:243-366 test_lowpass / test_lowpass_dc_removal -- ripple properties
FILTER_PARAMS = [(1 / 4, 20.0, 1 / 10), (1 / 3, 35.0, 1 / 30), (2 / 5, 60.0, 1 / 20)]


@pytest.mark.parametrize("cutout,atten,delta_w", FILTER_PARAMS)
def test_lowpass_ripple(cutout, atten, delta_w):
    ripple = 10.0 ** (-atten / 20.0)
    coeff = oracle.design(oracle.FILTER_LOWPASS, cutout, atten, delta_w)
    assert coeff.size % 2 == 1
    fft = np.abs(np.fft.fft(coeff.astype(np.float64)))
    for i, v in enumerate(fft):
        w = 2.0 * i / fft.size
        if w < cutout - delta_w / 2:
            assert 1 - ripple < v < 1 + ripple
        elif cutout + delta_w / 2 < w < 1.0:
            assert v < ripple


@pytest.mark.parametrize("cutout,atten,delta_w", FILTER_PARAMS)
def test_lowpass_dc_removal_ripple(cutout, atten, delta_w):
    ripple = 10.0 ** (-atten / 20.0)
    coeff = oracle.design(oracle.FILTER_LOWPASS_DC, cutout, atten, delta_w)
    fft = np.abs(np.fft.fft(coeff.astype(np.float64)))
    for i, v in enumerate(fft):
        w = 2.0 * i / fft.size
        if i == 0:
            assert v < 2 * ripple
        if delta_w < w < cutout - delta_w / 2:
            assert 1 - ripple < v < 1 + ripple
        elif cutout + delta_w / 2 < w < 1.0:
            assert v < ripple


# filters.rs:368-372
def test_no_filter():
    assert oracle.design(oracle.FILTER_NONE).tolist() == [1.0]


# filters.rs:377-413: a filter resampled 1000 -> 3000 Hz equals one designed at 3000 Hz
def test_filter_resample_equivalence():
    ratio = np.float32(3000) / np.float32(1000)
    for f_hz in (123.0, 12.0):
        a = np.float32(oracle.freq_hz(f_hz, 1000)) / ratio
        b = np.float32(oracle.freq_hz(f_hz, 3000))
        assert a == b


# dsp.rs:420-434 test_rate_overflow
def test_rate_overflow():
    with pytest.raises(oracle.OracleError) as e:
        oracle.resample_with_filter(np.zeros(1000, np.float32), 99371, 93911, oracle.FILTER_NONE)
    assert e.value.code == oracle.ERR_RATE_OVERFLOW


# dsp.rs:440-468 test_fast_resampling / _short: no overflow/panic on zeros
def test_fast_resampling_zeros():
    out = oracle.fast_resampling(np.zeros(1000, np.float32), 3, 2, np.zeros(100, np.float32))
    assert out.size == math.ceil((1000 * 3 - 49) / 2)
    assert not out.any()


def test_fast_resampling_coeff_longer_than_signal():
    out = oracle.fast_resampling(np.zeros(100, np.float32), 3, 2, np.zeros(1000, np.float32))
    assert out.size == 0  # interpolated_len (300) <= offset (499): the while never runs
    assert oracle.fast_resampling(np.zeros(400, np.float32), 3, 2, np.zeros(1000, np.float32)).size == \
        math.ceil((1200 - 499) / 2)


# dsp.rs:69-71
def test_resample_to_zero_hz():
    with pytest.raises(oracle.OracleError) as e:
        oracle.resample_with_filter(np.zeros(10, np.float32), 11025, 0, oracle.FILTER_NONE)
    assert e.value.code == oracle.ERR_RESAMPLE_TO_ZERO


# noaa_apt.rs:266-281 test_map
def test_map_signal_u8():
    expected = [0, 0, 0, 0, 1, 2, 50, 120, 200, 255, 255, 255]
    vals = np.array([-10., -5., -1., 0., 1., 2.4, 50., 120., 199.6, 255., 256., 300.], np.float32)
    shifted = vals * np.float32(123.123) - np.float32(234.234)
    low = np.float32(0.) * np.float32(123.123) - np.float32(234.234)
    high = np.float32(255.) * np.float32(123.123) - np.float32(234.234)
    assert oracle.map_signal_u8(shifted, low, high).tolist() == expected
