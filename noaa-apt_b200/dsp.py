"""dsp.rs: resample_with_filter, resample, demodulate, filter -- host-buffer stage entry points."""
import ctypes as C

import numpy as np

from . import _lib
from .err import raise_for
from .frequency import Freq, Rate, _rate_hz  # noqa: F401  (re-exported like dsp.rs:8-9)

Signal = np.ndarray  # dsp.rs:16: `type Signal = Vec<f32>`


def _f32(signal):
    return np.ascontiguousarray(signal, dtype=np.float32)


def resample_with_filter(context, signal, input_rate, output_rate, filt):
    """dsp.rs:62-126"""
    lib = _lib.load()
    x = _f32(signal)
    cf = filt.to_c()
    n = C.c_uint64(0)
    raise_for(lib.apt_resample_len(x.size, _rate_hz(input_rate), _rate_hz(output_rate), C.byref(cf), C.byref(n)))
    out = np.empty(n.value, dtype=np.float32)
    raise_for(lib.apt_resample_with_filter(x.ctypes.data, x.size, _rate_hz(input_rate), _rate_hz(output_rate),
                                           C.byref(cf), out.ctypes.data, out.size, C.byref(n)))
    return out[: n.value]


def resample(context, signal, input_rate, output_rate, atten, delta_w):
    """dsp.rs:132-162"""
    lib = _lib.load()
    x = _f32(signal)
    ir, orr = _rate_hz(input_rate), _rate_hz(output_rate)
    dw = delta_w.get_pi_rad() if isinstance(delta_w, Freq) else float(delta_w)
    n = C.c_uint64(0)
    st = lib.apt_resample(x.ctypes.data, x.size, ir, orr, atten, dw, None, 0, C.byref(n))
    if st != _lib.ERR_CAPACITY:
        raise_for(st)
    out = np.empty(n.value, dtype=np.float32)
    if n.value:
        raise_for(lib.apt_resample(x.ctypes.data, x.size, ir, orr, atten, dw, out.ctypes.data, out.size, C.byref(n)))
    return out


def demodulate(context, signal, carrier_freq):
    """dsp.rs:350-383"""
    x = _f32(signal)
    out = np.empty_like(x)
    raise_for(_lib.load().apt_demodulate(x.ctypes.data, x.size, carrier_freq.get_pi_rad(), out.ctypes.data))
    return out


def filter(context, signal, filt):  # noqa: A001  (the reference's name)
    """dsp.rs:386-410"""
    x = _f32(signal)
    out = np.empty_like(x)
    if isinstance(filt, np.ndarray):
        c = _f32(filt)
        raise_for(_lib.load().apt_filter_taps(x.ctypes.data, x.size, c.ctypes.data, c.size, out.ctypes.data))
    else:
        cf = filt.to_c()
        raise_for(_lib.load().apt_filter_signal(x.ctypes.data, x.size, C.byref(cf), out.ctypes.data))
    return out
