"""ctypes loader for libaptb200.so -- the C ABI declared in include/aptb200.h.

The library is the product.  If it has not been built this module raises: there
is no Python or CPU fallback for any compute entry point.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("APTB200_LIB") or os.path.join(HERE, "libaptb200.so")   # APTB200_LIB: an experimental build

OK = 0
ERR_RESAMPLE_TO_ZERO = 1
ERR_TOO_SHORT = 2
ERR_FEW_SYNC_FRAMES = 3
ERR_WORK_RATE = 4
ERR_RATE_OVERFLOW = 5
ERR_CUDA = 6
ERR_BAD_ARG = 7
ERR_NOMEM = 8
ERR_CAPACITY = 9
ERR_EMPTY_RESULT = 10
ERR_IO = 11

FILTER_NONE, FILTER_LOWPASS, FILTER_LOWPASS_DC = 0, 1, 2
F32, PCM16 = 0, 1


class CSettings(C.Structure):
    _fields_ = [
        ("work_rate", C.c_uint32),
        ("resample_atten", C.c_float),
        ("resample_delta_freq", C.c_float),
        ("resample_cutout", C.c_float),
        ("demodulation_atten", C.c_float),
    ]


class CFilter(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("cutout_pi", C.c_float),
        ("atten", C.c_float),
        ("delta_w_pi", C.c_float),
    ]


class CTileInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("usable", "groups", "p_out", "p_in", "usteps", "row_len", "rows_per_tile",
                                          "smem_bytes", "slices", "slice_stride", "half_taps", "shift", "iters",
                                          "group_stride", "ctas_per_sm", "pair_pitch", "halves", "rows_per_copy")]


class CUtInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("usable", "l", "m", "np", "q", "rows_per_block", "vec", "back", "chunks",
                                          "slot_floats", "slot_stride", "nslot", "warps", "smem_bytes", "nvec",
                                          "stream_b", "halo_u0", "halo_n", "chunk_len")] + [("cs", C.c_uint32 * 8), ("ce", C.c_uint32 * 8)]


class CImageInfo(C.Structure):
    _fields_ = [("low", C.c_float), ("high", C.c_float), ("rows", C.c_uint64), ("telemetry_row", C.c_uint64),
                ("wedges_a", C.c_float * 16), ("wedges_b", C.c_float * 16)]


CONTRAST_MINMAX, CONTRAST_PERCENT, CONTRAST_TELEMETRY = 0, 1, 2

class CWavInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("is_float", C.c_uint32),
                ("frames", C.c_uint64)]


class CPhInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("usable", "l", "m", "j", "jpad", "pitch", "row_len", "smem_bytes")]


STATUS_CB = C.CFUNCTYPE(None, C.c_float, C.c_char_p, C.c_void_p)

# name -> (restype, argtypes); every symbol include/aptb200.h declares.
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)
SIGNATURES = {
    "apt_strerror": (C.c_char_p, [C.c_int]),
    "apt_last_error": (C.c_char_p, []),
    "apt_abi_version": (C.c_int, []),
    "apt_device_count": (C.c_int, []),
    "apt_default_settings": (None, [C.POINTER(CSettings)]),
    "apt_profile_settings": (C.c_int, [C.c_char_p, C.POINTER(CSettings)]),
    "apt_freq_hz": (C.c_float, [C.c_float, C.c_uint32]),
    "apt_bessel_i0": (C.c_float, [C.c_float]),
    "apt_filter_resample": (None, [C.POINTER(CFilter), C.c_uint32, C.c_uint32]),
    "apt_filter_design": (C.c_int, [C.POINTER(CFilter), C.c_void_p, C.c_size_t, _szp]),
    "apt_resample_len": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(CFilter), _u64p]),
    "apt_resample_with_filter": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(CFilter),
                                           C.c_void_p, C.c_uint64, _u64p]),
    "apt_resample": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                               C.c_void_p, C.c_uint64, _u64p]),
    "apt_demodulate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]),
    "apt_filter_signal": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(CFilter), C.c_void_p]),
    "apt_filter_taps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "apt_generate_sync_frame": (C.c_int, [C.c_uint32, C.c_void_p, C.c_size_t, _szp]),
    "apt_find_sync": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_size_t, _szp, C.c_void_p]),
    "apt_decode_len_bound": (C.c_int, [C.c_uint64, C.c_uint32, C.POINTER(CSettings), _u64p]),
    "apt_decode": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(CSettings), C.c_int,
                             C.c_void_p, C.c_uint64, _u64p, STATUS_CB, C.c_void_p]),
    "apt_decode_pcm16": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(CSettings), C.c_int,
                                   C.c_void_p, C.c_uint64, _u64p, STATUS_CB, C.c_void_p]),
    "apt_wav_info_read": (C.c_int, [C.c_char_p, C.POINTER(CWavInfo)]),
    "apt_wav_load": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, _u64p, C.POINTER(C.c_uint32)]),
    "apt_wav_load_pcm16": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, _u64p, C.POINTER(C.c_uint32)]),
    "apt_wav_write_i16": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint32]),
    "apt_quantize_i16": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "apt_resample_wav": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint32, C.c_float, C.c_float, _u64p]),
    "apt_decode_image_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(CSettings), C.c_int, C.c_int,
                                      C.c_float, C.c_void_p, C.c_uint64, _u64p, C.POINTER(CImageInfo), STATUS_CB, C.c_void_p]),
    "apt_map_signal_u8": (C.c_int, [C.c_void_p, C.c_uint64, C.c_float, C.c_float, C.c_void_p]),
    "apt_contrast_bounds": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.POINTER(CImageInfo)]),
    "apt_telemetry_rows": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "apt_decoder_set_image_mode": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "apt_decoder_image_info": (C.c_int, [C.c_void_p, C.POINTER(CImageInfo)]),
    "apt_decoder_create": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(CSettings), C.c_uint64, C.POINTER(C.c_void_p)]),
    "apt_decoder_destroy": (None, [C.c_void_p]),
    "apt_decoder_submit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]),
    "apt_decoder_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]),
    "apt_decoder_poll": (C.c_int, [C.c_void_p]),
    "apt_cache_clear": (None, []),
    "apt_bind_thread_to_device": (C.c_int, [C.c_int]),
    "apt_decoder_wait": (C.c_int, [C.c_void_p, _u64p]),
    "apt_decoder_last_sync": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, _szp]),
    "apt_decoder_last_counts": (C.c_int, [C.c_void_p, _u64p, _u64p, _u64p]),
    "apt_decoder_last_root_count": (C.c_int, [C.c_void_p, _u64p]),
    "apt_decoder_last_roots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, _szp]),
    "apt_decoder_read_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, _u64p]),
    "apt_decoder_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "apt_decoder_kernel_count": (C.c_int, [C.c_void_p]),
    "apt_decoder_kernel_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "apt_decoder_kernel_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "apt_decoder_stream": (C.c_void_p, [C.c_void_p]),
    "apt_decoder_launch_count": (C.c_uint64, [C.c_void_p]),
    "apt_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "apt_host_free": (None, [C.c_void_p]),
    "apt_device_alloc": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_size_t]),
    "apt_device_free": (None, [C.c_int, C.c_void_p]),
    "apt_memcpy_h2d": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "apt_memcpy_d2h": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "apt_ph_plan": (C.c_int, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(CPhInfo), C.c_void_p, C.c_size_t,
                              C.c_void_p, C.c_size_t]),
    "apt_ut_plan": (C.c_int, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(CUtInfo), C.c_void_p, C.c_size_t]),
    "apt_tile_plan": (C.c_int, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(CTileInfo), C.c_void_p,
                                C.c_size_t, C.c_void_p, C.c_size_t]),
    "apt_decode_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, _u64p, C.c_int, C.c_uint32, C.POINTER(CSettings),
                                   C.c_int, C.POINTER(C.c_void_p), _u64p, _u64p, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.c_int, C.c_int]),
}

_lib = None


def load():
    """dlopen libaptb200.so and attach the prototypes.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} has not been built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means the ABI and the header have diverged
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
