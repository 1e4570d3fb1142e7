//! Optional replacements around `src/wav.rs` / `src/resample.rs` of martinber/noaa-apt.  SOURCE ONLY (never compiled
//! here).  `load_wav_pcm16` hands the file's 16-bit samples to `apt_decode_pcm16` (the `as f32` of wav.rs:37 then runs on
//! the GPU and the upload is half the bytes); `resample` is resample.rs:17-71 with the DSP and the i16 quantisation of
//! wav.rs:71-85 on the GPU (the modification-time copy of resample.rs:30,68 stays here).

use std::ffi::CString;
use std::path::Path;

use crate::aptb200_sys as sys;
use crate::config;
use crate::context::Context;
use crate::decode::to_error;
use crate::err;
use crate::misc;

pub fn load_wav_pcm16(filename: &Path) -> err::Result<(Vec<i16>, u32)> {
    let path = CString::new(filename.to_string_lossy().as_bytes()).map_err(|e| err::Error::Internal(e.to_string()))?;
    let mut info = sys::apt_wav_info::default();
    let st = unsafe { sys::apt_wav_info_read(path.as_ptr(), &mut info) };
    if st != 0 { return Err(to_error(st)); }
    let mut samples = vec![0i16; info.frames.max(1) as usize];
    let (mut n, mut rate) = (0u64, 0u32);
    let st = unsafe { sys::apt_wav_load_pcm16(path.as_ptr(), samples.as_mut_ptr(), samples.len() as u64, &mut n, &mut rate) };
    if st != 0 { return Err(to_error(st)); }
    samples.truncate(n as usize);
    Ok((samples, rate))
}

pub fn resample(
    context: &mut Context,
    settings: config::Settings,
    input_filename: &Path,
    output_filename: &Path,
    output_rate: u32,
) -> err::Result<()> {
    context.status(0.0, "Reading WAV file".to_string());
    let timestamp = misc::read_timestamp(input_filename)?;
    let src = CString::new(input_filename.to_string_lossy().as_bytes()).map_err(|e| err::Error::Internal(e.to_string()))?;
    let dst = CString::new(output_filename.to_string_lossy().as_bytes()).map_err(|e| err::Error::Internal(e.to_string()))?;
    context.status(0.2, format!("Resampling to {}", output_rate));
    let mut n = 0u64;
    let st = unsafe {
        sys::apt_resample_wav(src.as_ptr(), dst.as_ptr(), output_rate, settings.wav_resample_atten,
                              settings.wav_resample_delta_freq, &mut n)
    };
    if st != 0 { return Err(to_error(st)); }
    misc::write_timestamp(timestamp, output_filename)?;
    context.status(1., "Finished".to_string());
    Ok(())
}
