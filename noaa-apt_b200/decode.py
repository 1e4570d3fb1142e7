"""decode.rs: decode(), find_sync(), generate_sync_frame() -- plus the explicit-state Decoder
(one CUDA stream, device workspaces) and the multi-device batch entry point."""
import ctypes as C

import numpy as np

from . import _lib
from .config import Settings
from .context import Context
from .err import raise_for
from .frequency import _rate_hz

FINAL_RATE = 4160        # decode.rs:14
PX_PER_ROW = 2080        # decode.rs:35
CARRIER_FREQ = 2400      # decode.rs:38


def _fmt_of(arr):
    if arr.dtype == np.int16:
        return _lib.PCM16
    if arr.dtype == np.float32:
        return _lib.F32
    raise TypeError("signal must be float32 (Signal) or int16 (PCM16 of the WAV)")


def generate_sync_frame(work_rate):
    """decode.rs:171-199"""
    lib = _lib.load()
    n = C.c_size_t(0)
    raise_for(lib.apt_generate_sync_frame(_rate_hz(work_rate), None, 0, C.byref(n)))
    out = np.empty(n.value, dtype=np.int8)
    raise_for(lib.apt_generate_sync_frame(_rate_hz(work_rate), out.ctypes.data, out.size, C.byref(n)))
    return out


def find_sync(context, signal, work_rate, want_correlation=False):
    """decode.rs:204-263 -> positions of the sync frames (and the correlation if asked)."""
    lib = _lib.load()
    x = np.ascontiguousarray(signal, dtype=np.float32)
    wr = _rate_hz(work_rate)
    cap = x.size // max(1, PX_PER_ROW * wr // FINAL_RATE) + 4
    pos = np.empty(cap, dtype=np.uint64)
    n = C.c_size_t(0)
    corr = None
    if want_correlation:
        corr = np.empty(max(0, x.size - 38 * (wr // FINAL_RATE)), dtype=np.float32)
    raise_for(lib.apt_find_sync(x.ctypes.data, x.size, wr, pos.ctypes.data, cap, C.byref(n),
                                corr.ctypes.data if corr is not None else None))
    pos = pos[: n.value].copy()
    return (pos, corr) if want_correlation else pos


def decode_len_bound(n, input_rate, settings=None):
    s = (settings or Settings()).to_c()
    b = C.c_uint64(0)
    raise_for(_lib.load().apt_decode_len_bound(int(n), _rate_hz(input_rate), C.byref(s), C.byref(b)))
    return b.value


def decode(context, settings, signal, input_rate, sync=True):
    """noaa_apt::decode (decode.rs:43-162): Signal at input_rate -> rows*2080 f32 pixels.

    `signal` may also be the int16 PCM of the WAV (the `as f32` cast of wav.rs:37 is then done
    on the device)."""
    lib = _lib.load()
    x = np.ascontiguousarray(signal)
    fmt = _fmt_of(x)
    s = (settings or Settings()).to_c()
    rate = _rate_hz(input_rate)
    bound = C.c_uint64(0)
    raise_for(lib.apt_decode_len_bound(x.size, rate, C.byref(s), C.byref(bound)))
    out = np.empty(max(bound.value, 1), dtype=np.float32)
    n = C.c_uint64(0)

    def _cb(progress, desc, _user):
        if context is not None:
            context.status(progress, desc.decode())

    cb = _lib.STATUS_CB(_cb)
    fn = lib.apt_decode_pcm16 if fmt == _lib.PCM16 else lib.apt_decode
    raise_for(fn(x.ctypes.data, x.size, rate, C.byref(s), int(bool(sync)), out.ctypes.data, out.size, C.byref(n),
                 cb, None))
    return out[: n.value].copy()


class Decoder:
    """apt_decoder: plan + device workspaces + one stream; submit()/wait() keep one job in flight."""

    def __init__(self, input_rate, settings=None, max_samples=0, device=0):
        self._lib = _lib.load()
        self.settings = settings or Settings()
        self.input_rate = _rate_hz(input_rate)
        self.device = int(device)
        self.max_samples = int(max_samples)
        s = self.settings.to_c()
        h = C.c_void_p(None)
        raise_for(self._lib.apt_decoder_create(self.device, self.input_rate, C.byref(s), self.max_samples, C.byref(h)))
        self._h = h
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.apt_decoder_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def out_bound(self, n):
        return decode_len_bound(n, self.input_rate, self.settings)

    # --- raw pointer interface (device-resident inputs: torch tensors' data_ptr(), apt_device_alloc) ---
    def submit_device(self, signal_ptr, fmt, n, sync, out_ptr, cap):
        raise_for(self._lib.apt_decoder_submit_device(self._h, signal_ptr, fmt, n, int(bool(sync)), out_ptr, cap))

    def submit_host_ptr(self, signal_ptr, fmt, n, sync, out_ptr, cap):
        raise_for(self._lib.apt_decoder_submit_host(self._h, signal_ptr, fmt, n, int(bool(sync)), out_ptr, cap))

    # --- numpy interface ---
    def submit(self, signal, sync=True, out=None):
        x = np.ascontiguousarray(signal)
        fmt = _fmt_of(x)
        if out is None:
            out = np.empty(max(self.out_bound(x.size), 1), dtype=np.float32)
        self._keep = (x, out)
        self.submit_host_ptr(x.ctypes.data, fmt, x.size, sync, out.ctypes.data, out.size)
        return out

    def wait(self):
        n = C.c_uint64(0)
        st = self._lib.apt_decoder_wait(self._h, C.byref(n))
        keep, self._keep = self._keep, None
        raise_for(st)
        if keep is not None:
            return keep[1][: n.value]
        return n.value

    def decode(self, signal, sync=True):
        self.submit(signal, sync)
        return self.wait().copy()

    # --- introspection ---
    def last_sync(self):
        n = C.c_size_t(0)
        raise_for(self._lib.apt_decoder_last_sync(self._h, None, 0, C.byref(n)))
        pos = np.empty(n.value, dtype=np.uint64)
        if n.value:
            raise_for(self._lib.apt_decoder_last_sync(self._h, pos.ctypes.data, pos.size, C.byref(n)))
        return pos

    def last_counts(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        raise_for(self._lib.apt_decoder_last_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        r = C.c_uint64(0)
        raise_for(self._lib.apt_decoder_last_root_count(self._h, C.byref(r)))
        return {"n_work": a.value, "n_rows": b.value, "n_peaks": c.value, "n_roots": r.value}

    def last_roots(self):
        n = C.c_size_t(0)
        raise_for(self._lib.apt_decoder_last_roots(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint64)
        if n.value:
            raise_for(self._lib.apt_decoder_last_roots(self._h, out.ctypes.data, out.size, C.byref(n)))
        return out

    def read_stage(self, which):
        idx = {"demodulated": 0, "filtered": 1, "correlation": 2}[which] if isinstance(which, str) else int(which)
        n = C.c_uint64(0)
        raise_for(self._lib.apt_decoder_read_stage(self._h, idx, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        if n.value:
            raise_for(self._lib.apt_decoder_read_stage(self._h, idx, out.ctypes.data, out.size, C.byref(n)))
        return out

    def set_profiling(self, enabled):
        raise_for(self._lib.apt_decoder_set_profiling(self._h, int(bool(enabled))))

    def kernel_times_ms(self):
        cnt = self._lib.apt_decoder_kernel_count(self._h)
        ms = (C.c_float * max(cnt, 1))()
        got = C.c_int(0)
        raise_for(self._lib.apt_decoder_kernel_ms(self._h, ms, cnt, C.byref(got)))
        return [(self._lib.apt_decoder_kernel_name(self._h, i).decode(), float(ms[i])) for i in range(got.value)]

    @property
    def stream(self):
        return self._lib.apt_decoder_stream(self._h)

    @property
    def launch_count(self):
        return int(self._lib.apt_decoder_launch_count(self._h))


def decode_batch(signals, input_rate, settings=None, sync=True, devices=None, streams_per_device=4):
    """Independent recordings sharded recording i -> devices[i % G], no collective.
    Returns (list of row arrays or None, list of status codes)."""
    lib = _lib.load()
    s = (settings or Settings()).to_c()
    rate = _rate_hz(input_rate)
    xs = [np.ascontiguousarray(x) for x in signals]
    if not xs:
        return [], []
    fmt = _fmt_of(xs[0])
    if any(_fmt_of(x) != fmt for x in xs):
        raise TypeError("all recordings of a batch must share one sample format")
    count = len(xs)
    outs = [np.empty(max(decode_len_bound(x.size, rate, settings), 1), dtype=np.float32) for x in xs]
    sig_ptrs = (C.c_void_p * count)(*[x.ctypes.data for x in xs])
    out_ptrs = (C.c_void_p * count)(*[o.ctypes.data for o in outs])
    lens = (C.c_uint64 * count)(*[x.size for x in xs])
    caps = (C.c_uint64 * count)(*[o.size for o in outs])
    nouts = (C.c_uint64 * count)()
    statuses = (C.c_int * count)(*([_lib.ERR_CUDA] * count))   # never "ok" unless the library says so
    devs = list(devices) if devices else [0]
    dev_arr = (C.c_int * len(devs))(*devs)
    rc = lib.apt_decode_batch(sig_ptrs, fmt, lens, count, rate, C.byref(s), int(bool(sync)), out_ptrs, caps, nouts,
                              statuses, dev_arr, len(devs), int(streams_per_device))
    if rc != _lib.OK and all(int(v) == rc for v in statuses):
        raise_for(rc)                      # nothing decoded at all (no device, bad settings ...): an error, not a result
    res = [outs[i][: nouts[i]] if statuses[i] == _lib.OK else None for i in range(count)]
    return res, [int(v) for v in statuses]
