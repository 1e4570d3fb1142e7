"""Freq / Rate of the reference (frequency.rs:30-117); arithmetic is done by the library in f32."""
import numpy as np

from . import _lib


class Rate:
    """Integer sample rate in Hz (frequency.rs:98-117)."""

    def __init__(self, hz):
        self._hz = int(hz)

    @staticmethod
    def hz(r):
        return Rate(r)

    def get_hz(self):
        return self._hz

    def checked_mul(self, other):
        v = self._hz * int(other)
        return Rate(v) if v <= 0xFFFFFFFF else None

    def __eq__(self, o):
        return isinstance(o, Rate) and o._hz == self._hz

    def __repr__(self):
        return f"Rate({self._hz})"


def _rate_hz(r):
    return r.get_hz() if isinstance(r, Rate) else int(r)


class Freq:
    """Discrete-time frequency stored as a fraction of pi rad/sample (frequency.rs:30-87)."""

    def __init__(self, pi_rad):
        self._pi_rad = np.float32(pi_rad)

    @staticmethod
    def pi_rad(f):
        return Freq(f)

    @staticmethod
    def hz(f, rate):
        return Freq(_lib.load().apt_freq_hz(float(f), _rate_hz(rate)))

    def get_pi_rad(self):
        return float(self._pi_rad)

    def __truediv__(self, d):
        return Freq(np.float32(self._pi_rad) / np.float32(d))

    def __eq__(self, o):
        return isinstance(o, Freq) and np.float32(o._pi_rad) == np.float32(self._pi_rad)

    def __repr__(self):
        return f"Freq(pi_rad={float(self._pi_rad)!r})"
