"""ctypes binding of the CPU oracle (oracle/apt_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(noaa-apt_b200/) never imports this module.

Parity status: pinned against the reference's own unit-test vectors; the
outputs of fast_resampling/demodulate/filter/find_sync/decode are
"parity unpinned" (the Rust reference cannot be built here), see apt_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK = 0
ERR_RESAMPLE_TO_ZERO = 1
ERR_TOO_SHORT = 2
ERR_FEW_SYNC_FRAMES = 3
ERR_WORK_RATE = 4
ERR_RATE_OVERFLOW = 5
ERR_BAD_ARG = 7
ERR_NOMEM = 8

FILTER_NONE, FILTER_LOWPASS, FILTER_LOWPASS_DC = 0, 1, 2


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error code {code}")
        self.code = code


class Settings(C.Structure):
    _fields_ = [
        ("work_rate", C.c_uint32),
        ("resample_atten", C.c_float),
        ("resample_delta_freq", C.c_float),
        ("resample_cutout", C.c_float),
        ("demodulation_atten", C.c_float),
    ]


class _Steps(C.Structure):
    _fields_ = [
        ("resampled", C.POINTER(C.c_float)), ("n_resampled", C.c_uint64),
        ("demodulated", C.POINTER(C.c_float)), ("n_demodulated", C.c_uint64),
        ("filtered", C.POINTER(C.c_float)), ("n_filtered", C.c_uint64),
        ("sync_pos", C.POINTER(C.c_uint64)), ("n_sync_pos", C.c_size_t),
        ("aligned", C.POINTER(C.c_float)), ("n_aligned", C.c_uint64),
    ]


def build(force=False):
    """Compile liboracle.so with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "apt_oracle.c")
    hdr = os.path.join(_HERE, "apt_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    L.oracle_free.argtypes = [C.c_void_p]
    L.oracle_free.restype = None
    L.oracle_default_settings.argtypes = [C.POINTER(Settings)]
    L.oracle_default_settings.restype = None
    for name in ("oracle_freq_rad", "oracle_freq_get_rad", "oracle_bessel_i0"):
        getattr(L, name).argtypes = [C.c_float]
        getattr(L, name).restype = C.c_float
    L.oracle_freq_hz.argtypes = [C.c_float, C.c_uint32]
    L.oracle_freq_hz.restype = C.c_float
    L.oracle_freq_get_hz.argtypes = [C.c_float, C.c_uint32]
    L.oracle_freq_get_hz.restype = C.c_float
    L.oracle_kaiser.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_size_t)]
    L.oracle_kaiser.restype = C.c_void_p
    L.oracle_design.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_size_t)]
    L.oracle_design.restype = C.c_void_p
    L.oracle_fast_resampling.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    L.oracle_fast_resampling.restype = C.c_void_p
    L.oracle_decimate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
    L.oracle_decimate.restype = C.c_void_p
    L.oracle_demodulate.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
    L.oracle_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]
    L.oracle_resample_with_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                              C.c_float, C.c_float, C.c_float,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.oracle_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.oracle_generate_sync_frame.argtypes = [C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.oracle_find_sync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t), C.c_void_p]
    L.oracle_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(Settings), C.c_int,
                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.oracle_decode_steps.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(Settings), C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(_Steps)]
    L.oracle_steps_free.argtypes = [C.POINTER(_Steps)]
    L.oracle_steps_free.restype = None
    L.oracle_pcm16_to_f32.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.oracle_pcm16_to_f32.restype = None
    L.oracle_map_signal_u8.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_float, C.c_void_p]
    L.oracle_map_signal_u8.restype = None
    L.oracle_quantize_i16.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.oracle_minmax.argtypes = [C.c_void_p, C.c_uint64, fp, fp]
    L.oracle_percent_buckets.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, fp, fp]
    L.oracle_percent.argtypes = [C.c_void_p, C.c_uint64, C.c_float, fp, fp]
    L.oracle_telemetry_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oracle_read_telemetry.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _take(ptr, n, dtype):
    """Copy n items out of a malloc'd oracle buffer and free it."""
    if not ptr:
        return np.zeros(0, dtype=dtype)
    ctype = np.ctypeslib.as_ctypes_type(dtype)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(int(n), 1),))[: int(n)].copy()
    lib().oracle_free(ptr)
    return arr


def _check(rc):
    if rc != OK:
        raise OracleError(rc)


def default_settings():
    s = Settings()
    lib().oracle_default_settings(C.byref(s))
    return s


def freq_hz(f, rate):
    return lib().oracle_freq_hz(f, rate)


def freq_rad(f):
    return lib().oracle_freq_rad(f)


def freq_get_rad(pi_rad):
    return lib().oracle_freq_get_rad(pi_rad)


def freq_get_hz(pi_rad, rate):
    return lib().oracle_freq_get_hz(pi_rad, rate)


def bessel_i0(x):
    return lib().oracle_bessel_i0(x)


def kaiser(atten, delta_w_pi):
    n = C.c_size_t(0)
    p = lib().oracle_kaiser(atten, delta_w_pi, C.byref(n))
    return _take(p, n.value, np.float32)


def design(kind, cutout_pi=0.0, atten=0.0, delta_w_pi=0.0):
    n = C.c_size_t(0)
    p = lib().oracle_design(kind, cutout_pi, atten, delta_w_pi, C.byref(n))
    return _take(p, n.value, np.float32)


def fast_resampling(x, l, m, coeff):
    x = _f32(x)
    coeff = _f32(coeff)
    n = C.c_uint64(0)
    p = lib().oracle_fast_resampling(x.ctypes.data, x.size, l, m, coeff.ctypes.data, coeff.size, C.byref(n))
    return _take(p, n.value, np.float32)


def decimate(x, m):
    x = _f32(x)
    n = C.c_uint64(0)
    p = lib().oracle_decimate(x.ctypes.data, x.size, m, C.byref(n))
    return _take(p, n.value, np.float32)


def demodulate(x, carrier_pi_rad):
    x = _f32(x)
    out = np.empty_like(x)
    _check(lib().oracle_demodulate(x.ctypes.data, x.size, carrier_pi_rad, out.ctypes.data))
    return out


def filter(x, coeff):  # noqa: A001 - mirrors dsp::filter
    x = _f32(x)
    coeff = _f32(coeff)
    out = np.empty_like(x)
    _check(lib().oracle_filter(x.ctypes.data, x.size, coeff.ctypes.data, coeff.size, out.ctypes.data))
    return out


def resample_with_filter(x, in_rate, out_rate, kind, cutout_pi=0.0, atten=0.0, delta_w_pi=0.0):
    x = _f32(x)
    p = C.c_void_p(None)
    n = C.c_uint64(0)
    _check(lib().oracle_resample_with_filter(x.ctypes.data, x.size, in_rate, out_rate, kind,
                                             cutout_pi, atten, delta_w_pi, C.byref(p), C.byref(n)))
    return _take(p.value, n.value, np.float32)


def resample(x, in_rate, out_rate, atten, delta_w_pi):
    x = _f32(x)
    p = C.c_void_p(None)
    n = C.c_uint64(0)
    _check(lib().oracle_resample(x.ctypes.data, x.size, in_rate, out_rate, atten, delta_w_pi,
                                 C.byref(p), C.byref(n)))
    return _take(p.value, n.value, np.float32)


def generate_sync_frame(work_rate):
    p = C.c_void_p(None)
    n = C.c_size_t(0)
    _check(lib().oracle_generate_sync_frame(work_rate, C.byref(p), C.byref(n)))
    return _take(p.value, n.value, np.int8)


def find_sync(x, work_rate, want_corr=False):
    x = _f32(x)
    p = C.c_void_p(None)
    n = C.c_size_t(0)
    corr = None
    cptr = None
    if want_corr:
        glen = 38 * (work_rate // 4160)
        corr = np.empty(max(x.size - glen, 0), dtype=np.float32)
        cptr = corr.ctypes.data
    _check(lib().oracle_find_sync(x.ctypes.data, x.size, work_rate, C.byref(p), C.byref(n), cptr))
    pos = _take(p.value, n.value, np.uint64)
    return (pos, corr) if want_corr else pos


def decode(x, in_rate, settings=None, sync=True):
    x = _f32(x)
    s = settings if settings is not None else default_settings()
    p = C.c_void_p(None)
    n = C.c_uint64(0)
    _check(lib().oracle_decode(x.ctypes.data, x.size, in_rate, C.byref(s), int(sync), C.byref(p), C.byref(n)))
    return _take(p.value, n.value, np.float32)


def decode_steps(x, in_rate, settings=None, sync=True):
    """decode() plus the intermediate signals Context::step would dump."""
    x = _f32(x)
    s = settings if settings is not None else default_settings()
    p = C.c_void_p(None)
    n = C.c_uint64(0)
    st = _Steps()
    _check(lib().oracle_decode_steps(x.ctypes.data, x.size, in_rate, C.byref(s), int(sync),
                                     C.byref(p), C.byref(n), C.byref(st)))
    out = _take(p.value, n.value, np.float32)

    def grab(ptr, cnt, dtype):
        if not ptr or cnt == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(int(cnt),)).astype(dtype, copy=True)

    steps = {
        "resampled": grab(st.resampled, st.n_resampled, np.float32),
        "demodulated": grab(st.demodulated, st.n_demodulated, np.float32),
        "filtered": grab(st.filtered, st.n_filtered, np.float32),
        "sync_pos": grab(st.sync_pos, st.n_sync_pos, np.uint64),
        "aligned": grab(st.aligned, st.n_aligned, np.float32),
    }
    lib().oracle_steps_free(C.byref(st))
    return out, steps


def pcm16_to_f32(x):
    x = np.ascontiguousarray(x, dtype=np.int16)
    out = np.empty(x.size, dtype=np.float32)
    lib().oracle_pcm16_to_f32(x.ctypes.data, x.size, out.ctypes.data)
    return out


def map_signal_u8(x, low, high):
    x = _f32(x)
    out = np.empty(x.size, dtype=np.uint8)
    lib().oracle_map_signal_u8(x.ctypes.data, x.size, low, high, out.ctypes.data)
    return out


def quantize_i16(x):
    x = _f32(x)
    out = np.empty(x.size, dtype=np.int16)
    _check(lib().oracle_quantize_i16(x.ctypes.data, x.size, out.ctypes.data))
    return out


def minmax(x):
    x = _f32(x)
    lo, hi = C.c_float(0), C.c_float(0)
    _check(lib().oracle_minmax(x.ctypes.data, x.size, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def percent_buckets(x):
    x = _f32(x)
    b = np.zeros(1000, dtype=np.uint32)
    lo, hi = C.c_float(0), C.c_float(0)
    _check(lib().oracle_percent_buckets(x.ctypes.data, x.size, b.ctypes.data, C.byref(lo), C.byref(hi)))
    return b, lo.value, hi.value


def percent(x, p):
    """misc.rs:119-175"""
    x = _f32(x)
    lo, hi = C.c_float(0), C.c_float(0)
    _check(lib().oracle_percent(x.ctypes.data, x.size, p, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def telemetry_rows(x):
    """telemetry.rs:147-170 -> (mean_a, mean_b, variance), one value per image row"""
    x = _f32(x)
    rows = x.size // 2080
    a, b, v = (np.empty(rows, dtype=np.float32) for _ in range(3))
    _check(lib().oracle_telemetry_rows(x.ctypes.data, x.size, a.ctypes.data, b.ctypes.data, v.ctypes.data))
    return a, b, v


def read_telemetry(x):
    """telemetry.rs:125-243 -> (wedges_a[16], wedges_b[16], best_row)"""
    x = _f32(x)
    wa, wb = np.empty(16, dtype=np.float32), np.empty(16, dtype=np.float32)
    best = C.c_uint64(0)
    _check(lib().oracle_read_telemetry(x.ctypes.data, x.size, wa.ctypes.data, wb.ctypes.data, C.byref(best)))
    return wa, wb, best.value
