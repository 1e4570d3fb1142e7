"""err::Error of the reference (err.rs:9-44), restricted to the variants this path can raise."""
from . import _lib


class Error(Exception):
    """Base of every error the decode path reports; `code` is the apt_status."""

    def __init__(self, code, message=""):
        self.code = int(code)
        if not message:
            message = _lib.load().apt_strerror(self.code).decode()
        super().__init__(message)


class Internal(Error):
    """err::Error::Internal -- the reference's four message-only failures on this path."""


class RateOverflow(Error):
    """err::Error::RateOverflow (dsp.rs:82-91)."""


class InvalidInput(Error):
    """Inputs on which the reference panics (empty signal, rate 0, NULL) or a too-small buffer."""


class Io(Error):
    """err::Error::Io / WavOpen (err.rs:11-17): a WAV file cannot be opened, parsed or written."""


class CudaError(Error):
    """CUDA failure or no device.  There is no CPU fallback."""


def raise_for(code):
    if code == _lib.OK:
        return
    msg = _lib.load().apt_last_error().decode() or _lib.load().apt_strerror(code).decode()
    if code in (_lib.ERR_RESAMPLE_TO_ZERO, _lib.ERR_TOO_SHORT, _lib.ERR_FEW_SYNC_FRAMES, _lib.ERR_WORK_RATE,
                _lib.ERR_EMPTY_RESULT):
        raise Internal(code, msg)
    if code == _lib.ERR_RATE_OVERFLOW:
        raise RateOverflow(code, msg)
    if code == _lib.ERR_IO:
        raise Io(code, msg)
    if code in (_lib.ERR_CUDA, _lib.ERR_NOMEM):
        raise CudaError(code, msg)
    raise InvalidInput(code, msg)
