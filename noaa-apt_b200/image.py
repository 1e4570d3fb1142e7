"""The front of noaa_apt::process (noaa_apt.rs:132-190) on the device: contrast bounds (MinMax / misc::percent /
telemetry wedges) and map_signal_u8 -- the decoded rows become the u8 image before they leave the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from .config import Settings
from .err import raise_for
from .frequency import _rate_hz

MINMAX, PERCENT, TELEMETRY = _lib.CONTRAST_MINMAX, _lib.CONTRAST_PERCENT, _lib.CONTRAST_TELEMETRY


def _info(ci):
    return {"low": ci.low, "high": ci.high, "rows": int(ci.rows), "telemetry_row": int(ci.telemetry_row),
            "wedges_a": np.array(ci.wedges_a[:], dtype=np.float32), "wedges_b": np.array(ci.wedges_b[:], dtype=np.float32)}


def map_signal_u8(signal, low, high):
    """noaa_apt.rs:249-259"""
    x = np.ascontiguousarray(signal, dtype=np.float32)
    out = np.empty(x.size, dtype=np.uint8)
    raise_for(_lib.load().apt_map_signal_u8(x.ctypes.data, x.size, low, high, out.ctypes.data))
    return out


def contrast_bounds(signal, contrast, percent=0.98):
    """(low, high, info): MinMax (dsp.rs:20-54), misc::percent (misc.rs:119-175) or telemetry wedges 9 / 8."""
    x = np.ascontiguousarray(signal, dtype=np.float32)
    ci = _lib.CImageInfo()
    raise_for(_lib.load().apt_contrast_bounds(x.ctypes.data, x.size, int(contrast), float(percent), C.byref(ci)))
    info = _info(ci)
    return info["low"], info["high"], info


def telemetry_rows(signal):
    """telemetry.rs:147-170 -> (mean_a, mean_b, variance) per image row"""
    x = np.ascontiguousarray(signal, dtype=np.float32)
    rows = x.size // 2080
    a, b, v = (np.empty(rows, dtype=np.float32) for _ in range(3))
    raise_for(_lib.load().apt_telemetry_rows(x.ctypes.data, x.size, a.ctypes.data, b.ctypes.data, v.ctypes.data))
    return a, b, v


def decode_image_u8(context, settings, signal, input_rate, sync=True, contrast=MINMAX, percent=0.98):
    """decode() + contrast + map_signal_u8 in one call -> (u8 image of shape (rows, 2080), info)."""
    lib = _lib.load()
    x = np.ascontiguousarray(signal)
    fmt = _lib.PCM16 if x.dtype == np.int16 else _lib.F32
    if fmt == _lib.F32:
        x = np.ascontiguousarray(x, dtype=np.float32)
    s = (settings or Settings()).to_c()
    rate = _rate_hz(input_rate)
    bound = C.c_uint64(0)
    raise_for(lib.apt_decode_len_bound(x.size, rate, C.byref(s), C.byref(bound)))
    out = np.empty(max(bound.value, 1), dtype=np.uint8)
    n = C.c_uint64(0)
    ci = _lib.CImageInfo()

    def _cb(progress, desc, _user):
        if context is not None:
            context.status(progress, desc.decode())

    cb = _lib.STATUS_CB(_cb)
    raise_for(lib.apt_decode_image_u8(x.ctypes.data, fmt, x.size, rate, C.byref(s), int(bool(sync)), int(contrast),
                                      float(percent), out.ctypes.data, out.size, C.byref(n), C.byref(ci), cb, None))
    return out[: n.value].reshape(-1, 2080).copy(), _info(ci)
