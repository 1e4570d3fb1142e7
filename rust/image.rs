//! Optional fast path for the front of `noaa_apt::process` (noaa_apt.rs:132-190): decode + contrast bounds +
//! `map_signal_u8` in one library call, so that the rows leave the GPU as u8.  SOURCE ONLY (no Rust toolchain in the
//! image this repository is built in).  `Contrast::Histogram`, false colour, map overlay and rotation stay in Rust.

use std::os::raw::c_void;

use crate::aptb200_sys as sys;
use crate::config;
use crate::context::Context;
use crate::decode::to_error;
use crate::dsp::{Rate, Signal};
use crate::err;
use crate::noaa_apt::Contrast;

/// (GrayImage bytes of width PX_PER_ROW, contrast bounds) for the three contrast modes the library implements.
pub fn decode_to_gray(
    context: &mut Context,
    settings: &config::Settings,
    signal: &Signal,
    input_rate: Rate,
    sync: bool,
    contrast: &Contrast,
) -> err::Result<(Vec<u8>, f32, f32)> {
    let (mode, percent) = match contrast {
        Contrast::MinMax => (sys::APT_CONTRAST_MINMAX, 0.0f32),
        Contrast::Percent(p) => (sys::APT_CONTRAST_PERCENT, *p),
        Contrast::Telemetry => (sys::APT_CONTRAST_TELEMETRY, 0.0f32),
        Contrast::Histogram => return Err(err::Error::FeatureNotAvailable(vec!["device histogram equalisation".into()])),
    };
    let s = sys::apt_settings::from(settings);
    let mut bound = 0u64;
    let st = unsafe { sys::apt_decode_len_bound(signal.len() as u64, input_rate.get_hz(), &s, &mut bound) };
    if st != 0 { return Err(to_error(st)); }
    let mut out = vec![0u8; bound.max(1) as usize];
    let mut n = 0u64;
    let mut info = sys::apt_image_info::default();
    let st = unsafe {
        sys::apt_decode_image_u8(signal.as_ptr() as *const c_void, sys::APT_F32, signal.len() as u64, input_rate.get_hz(), &s,
                                 sync as i32, mode, percent, out.as_mut_ptr(), out.len() as u64, &mut n, &mut info,
                                 Some(crate::decode::status_trampoline), context as *mut _ as *mut c_void)
    };
    if st != 0 { return Err(to_error(st)); }
    out.truncate(n as usize);
    Ok((out, info.low, info.high))
}
