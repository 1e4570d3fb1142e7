//! Replacement bodies for `src/dsp.rs` (resample_with_filter :62, demodulate :350, filter :386) over
//! libaptb200.  `resample` (:132) keeps its Rust body -- it only builds a `Lowpass` and calls
//! `resample_with_filter`.  SOURCE ONLY (never compiled here).

use crate::aptb200_sys as sys;
use crate::context::Context;
use crate::decode::to_error;
use crate::err;
use crate::filters;
pub use crate::frequency::{Freq, Rate};

pub type Signal = Vec<f32>;

/// `filters::Filter` gains one method so that a filter can cross the boundary as a POD:
/// `fn to_c(&self) -> sys::apt_filter` (NoFilter -> kind 0; Lowpass -> kind 1; LowpassDcRemoval -> kind 2,
/// with `cutout.get_pi_rad()`, `atten`, `delta_w.get_pi_rad()`); `design()` may keep its Rust body or call
/// `apt_filter_design` (bit-identical taps: same libm, same expression order).
pub fn resample_with_filter(
    _context: &mut Context,
    signal: &Signal,
    input_rate: Rate,
    output_rate: Rate,
    filt: impl filters::Filter,
) -> err::Result<Signal> {
    let f = filt.to_c();
    let mut n: u64 = 0;
    let st = unsafe { sys::apt_resample_len(signal.len() as u64, input_rate.get_hz(), output_rate.get_hz(), &f, &mut n) };
    if st != sys::APT_OK {
        return Err(to_error(st));
    }
    let mut out: Signal = vec![0_f32; n as usize];
    let st = unsafe {
        sys::apt_resample_with_filter(signal.as_ptr(), signal.len() as u64, input_rate.get_hz(), output_rate.get_hz(),
                                      &f, out.as_mut_ptr(), out.len() as u64, &mut n)
    };
    if st != sys::APT_OK {
        return Err(to_error(st));
    }
    Ok(out)
}

pub fn demodulate(_context: &mut Context, signal: &Signal, carrier_freq: Freq) -> err::Result<Signal> {
    let mut out: Signal = vec![0_f32; signal.len()];
    let st = unsafe { sys::apt_demodulate(signal.as_ptr(), signal.len() as u64, carrier_freq.get_pi_rad(), out.as_mut_ptr()) };
    if st != sys::APT_OK {
        return Err(to_error(st));
    }
    Ok(out)
}

pub fn filter(_context: &mut Context, signal: &Signal, filter: impl filters::Filter) -> err::Result<Signal> {
    let f = filter.to_c();
    let mut out: Signal = vec![0_f32; signal.len()];
    let st = unsafe { sys::apt_filter_signal(signal.as_ptr(), signal.len() as u64, &f, out.as_mut_ptr()) };
    if st != sys::APT_OK {
        return Err(to_error(st));
    }
    Ok(out)
}
