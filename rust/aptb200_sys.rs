//! Raw FFI declarations for libaptb200.so (hand-written equivalent of what `bindgen include/aptb200.h`
//! emits).  SOURCE ONLY: there is no Rust toolchain in the image this repository is built and tested in,
//! so this file has never been compiled; the ABI it binds is exercised by the C++ and Python hosts.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_float, c_int, c_void};

pub const APT_OK: c_int = 0;
pub const APT_ERR_RESAMPLE_TO_ZERO: c_int = 1;
pub const APT_ERR_TOO_SHORT: c_int = 2;
pub const APT_ERR_FEW_SYNC_FRAMES: c_int = 3;
pub const APT_ERR_WORK_RATE: c_int = 4;
pub const APT_ERR_RATE_OVERFLOW: c_int = 5;
pub const APT_ERR_CUDA: c_int = 6;
pub const APT_ERR_BAD_ARG: c_int = 7;
pub const APT_ERR_NOMEM: c_int = 8;
pub const APT_ERR_CAPACITY: c_int = 9;
pub const APT_ERR_EMPTY_RESULT: c_int = 10;

pub const APT_FILTER_NONE: c_int = 0;
pub const APT_FILTER_LOWPASS: c_int = 1;
pub const APT_FILTER_LOWPASS_DC: c_int = 2;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct apt_settings {
    pub work_rate: u32,
    pub resample_atten: c_float,
    pub resample_delta_freq: c_float,
    pub resample_cutout: c_float,
    pub demodulation_atten: c_float,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct apt_filter {
    pub kind: c_int,
    pub cutout_pi: c_float,
    pub atten: c_float,
    pub delta_w_pi: c_float,
}

pub type apt_status_cb = Option<unsafe extern "C" fn(progress: c_float, description: *const c_char, user: *mut c_void)>;

#[link(name = "aptb200")]
extern "C" {
    pub fn apt_strerror(status: c_int) -> *const c_char;
    pub fn apt_last_error() -> *const c_char;
    pub fn apt_filter_resample(f: *mut apt_filter, input_rate: u32, output_rate: u32);
    pub fn apt_filter_design(f: *const apt_filter, out: *mut c_float, cap: usize, n: *mut usize) -> c_int;
    pub fn apt_freq_hz(f_hz: c_float, rate_hz: u32) -> c_float;
    pub fn apt_resample_len(n: u64, input_rate: u32, output_rate: u32, f: *const apt_filter, nout: *mut u64) -> c_int;
    pub fn apt_resample_with_filter(signal: *const c_float, n: u64, input_rate: u32, output_rate: u32,
                                    f: *const apt_filter, out: *mut c_float, cap: u64, nout: *mut u64) -> c_int;
    pub fn apt_demodulate(signal: *const c_float, n: u64, carrier_pi: c_float, out: *mut c_float) -> c_int;
    pub fn apt_filter_signal(signal: *const c_float, n: u64, f: *const apt_filter, out: *mut c_float) -> c_int;
    pub fn apt_find_sync(signal: *const c_float, n: u64, work_rate: u32, positions: *mut u64, cap: usize,
                         npositions: *mut usize, corr: *mut c_float) -> c_int;
    pub fn apt_decode_len_bound(n: u64, input_rate: u32, s: *const apt_settings, bound: *mut u64) -> c_int;
    pub fn apt_decode(signal: *const c_float, n: u64, input_rate: u32, s: *const apt_settings, sync: c_int,
                      out: *mut c_float, cap: u64, nout: *mut u64, cb: apt_status_cb, user: *mut c_void) -> c_int;
    pub fn apt_decode_pcm16(pcm: *const i16, n: u64, input_rate: u32, s: *const apt_settings, sync: c_int,
                            out: *mut c_float, cap: u64, nout: *mut u64, cb: apt_status_cb, user: *mut c_void) -> c_int;
    pub fn apt_cache_clear();

    // WAV files and the resample tool (wav.rs, resample.rs)
    pub fn apt_wav_info_read(path: *const c_char, info: *mut apt_wav_info) -> c_int;
    pub fn apt_wav_load(path: *const c_char, out: *mut c_float, cap: u64, n: *mut u64, sample_rate: *mut u32) -> c_int;
    pub fn apt_wav_load_pcm16(path: *const c_char, out: *mut i16, cap: u64, n: *mut u64, sample_rate: *mut u32) -> c_int;
    pub fn apt_wav_write_i16(path: *const c_char, samples: *const i16, n: u64, sample_rate: u32) -> c_int;
    pub fn apt_quantize_i16(signal: *const c_float, n: u64, out: *mut i16) -> c_int;
    pub fn apt_resample_wav(input_path: *const c_char, output_path: *const c_char, output_rate: u32, atten: c_float,
                            delta_w_pi: c_float, nout: *mut u64) -> c_int;

    // image stage (front of noaa_apt::process)
    pub fn apt_decode_image_u8(signal: *const c_void, format: c_int, n: u64, input_rate: u32, s: *const apt_settings,
                               sync: c_int, contrast: c_int, percent: c_float, out: *mut u8, cap: u64, nout: *mut u64,
                               info: *mut apt_image_info, cb: apt_status_cb, user: *mut c_void) -> c_int;
    pub fn apt_map_signal_u8(signal: *const c_float, n: u64, low: c_float, high: c_float, out: *mut u8) -> c_int;
    pub fn apt_contrast_bounds(signal: *const c_float, n: u64, contrast: c_int, percent: c_float,
                               info: *mut apt_image_info) -> c_int;
    pub fn apt_telemetry_rows(signal: *const c_float, n: u64, mean_a: *mut c_float, mean_b: *mut c_float,
                              variance: *mut c_float) -> c_int;
}

pub const APT_F32: c_int = 0;
pub const APT_PCM16: c_int = 1;
pub const APT_CONTRAST_MINMAX: c_int = 0;
pub const APT_CONTRAST_PERCENT: c_int = 1;
pub const APT_CONTRAST_TELEMETRY: c_int = 2;
pub const APT_ERR_IO: c_int = 11;

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct apt_wav_info { pub sample_rate: u32, pub channels: u32, pub bits_per_sample: u32, pub is_float: u32, pub frames: u64 }

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct apt_image_info { pub low: f32, pub high: f32, pub rows: u64, pub telemetry_row: u64,
                            pub wedges_a: [f32; 16], pub wedges_b: [f32; 16] }
