//! `speechsauce::functions` STFT — same signatures (reference functions.rs:86,199).
//! The overlap memory is `cfg.analysis_mem` (a RefCell like the reference's, config.rs:126): read before the first row,
//! left holding the memory after the last row, carried across calls.
use crate::config::{check, SpeechConfig};
use crate::ffi;
use ndarray::{Array2, Array3, ArrayView1, ArrayView2};
use num_complex::Complex32;

/// functions.rs:86 — the overlap memory lives in the config's plan and carries across rows and calls
pub fn stft2(input: ArrayView2<f32>, cfg: &SpeechConfig) -> Array3<Complex32> {
    let x = input.as_standard_layout();
    let (ch, ttd) = x.dim();
    let tp = unsafe { ffi::ss_stft_num_frames(ttd as i64, cfg.frame_size as i64) }.max(0) as usize;
    let mut out = Array3::<Complex32>::zeros((ch, tp, cfg.freq_size));
    // Complex32 is repr(C) {re, im}: the [ch][T'][F][2] f32 layout of the C ABI
    unsafe { check(ffi::ss_stft(cfg.plan, x.as_ptr(), ch as i64, ttd as i64, ttd as i64, out.as_mut_ptr() as *mut f32, cfg.analysis_mem_ptr())) };
    out
}

/// functions.rs:199
pub fn stft1(input: ArrayView1<f32>, cfg: &SpeechConfig) -> Array2<Complex32> {
    let x = input.as_standard_layout();
    let ttd = x.len();
    let tp = unsafe { ffi::ss_stft_num_frames(ttd as i64, cfg.frame_size as i64) }.max(0) as usize;
    let mut out = Array2::<Complex32>::zeros((tp, cfg.freq_size));
    unsafe { check(ffi::ss_stft(cfg.plan, x.as_ptr(), 1, ttd as i64, ttd as i64, out.as_mut_ptr() as *mut f32, cfg.analysis_mem_ptr())) };
    out
}
