"""filters.rs: trait Filter { design(); resample() } and its three implementations.
Design runs in the library's host code (csrc/filters_host.cpp)."""
import ctypes as C

import numpy as np

from . import _lib
from .err import raise_for
from .frequency import Freq, _rate_hz


class Filter:
    kind = _lib.FILTER_NONE

    def to_c(self):
        return _lib.CFilter(self.kind, 0.0, 0.0, 0.0)

    def design(self):
        lib = _lib.load()
        cf = self.to_c()
        n = C.c_size_t(0)
        raise_for(lib.apt_filter_design(C.byref(cf), None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        raise_for(lib.apt_filter_design(C.byref(cf), out.ctypes.data, out.size, C.byref(n)))
        return out

    def resample(self, input_rate, output_rate):
        pass


class NoFilter(Filter):
    """filters.rs:48-54"""

    def __eq__(self, o):
        return isinstance(o, NoFilter)


class _Windowed(Filter):
    def __init__(self, cutout, atten, delta_w):
        self.cutout = cutout if isinstance(cutout, Freq) else Freq(cutout)
        self.atten = float(atten)
        self.delta_w = delta_w if isinstance(delta_w, Freq) else Freq(delta_w)

    def to_c(self):
        return _lib.CFilter(self.kind, self.cutout.get_pi_rad(), self.atten, self.delta_w.get_pi_rad())

    def resample(self, input_rate, output_rate):
        cf = self.to_c()
        _lib.load().apt_filter_resample(C.byref(cf), _rate_hz(input_rate), _rate_hz(output_rate))
        self.cutout = Freq(cf.cutout_pi)
        self.delta_w = Freq(cf.delta_w_pi)

    def __eq__(self, o):
        return type(o) is type(self) and (o.cutout, o.atten, o.delta_w) == (self.cutout, self.atten, self.delta_w)


class Lowpass(_Windowed):
    """filters.rs:56-95"""
    kind = _lib.FILTER_LOWPASS


class LowpassDcRemoval(_Windowed):
    """filters.rs:97-139"""
    kind = _lib.FILTER_LOWPASS_DC
