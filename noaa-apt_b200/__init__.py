"""noaa-apt_b200 -- B200-native APT decode path behind the reference's own surface.

Mirrors the module layout of martinber/noaa-apt's hot path:

    noaa_apt_b200.decode(context, settings, signal, input_rate, sync)   # noaa_apt::decode
    noaa_apt_b200.dsp.{resample_with_filter, resample, demodulate, filter, Freq, Rate}
    noaa_apt_b200.filters.{NoFilter, Lowpass, LowpassDcRemoval}
    noaa_apt_b200.Context, noaa_apt_b200.Settings, noaa_apt_b200.err

Every compute call goes through the C ABI of libaptb200.so (include/aptb200.h) into hand-written
sm_100a CUDA kernels.  There is no CPU path: importing works anywhere, computing needs a GPU.
"""
from . import _lib, config, context, dsp, err, filters, frequency, image, wav  # noqa: F401
from . import decode as _decode_mod
from .config import Settings
from .context import Context
from .decode import (CARRIER_FREQ, FINAL_RATE, PX_PER_ROW, Decoder, decode, decode_batch, decode_len_bound,
                     find_sync, generate_sync_frame)
from .frequency import Freq, Rate

__all__ = ["decode", "decode_batch", "decode_len_bound", "find_sync", "generate_sync_frame", "Decoder",
           "Settings", "Context", "Freq", "Rate", "dsp", "filters", "err", "config", "frequency", "image", "wav",
           "FINAL_RATE", "PX_PER_ROW", "CARRIER_FREQ"]


def library_path():
    return _lib.LIB_PATH


def device_count():
    return _lib.load().apt_device_count()
