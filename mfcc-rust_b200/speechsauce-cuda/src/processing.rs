//! `speechsauce::processing` — same signatures (reference processing.rs:31,65,179,265).
use crate::config::check;
use crate::ffi;
use ndarray::{Array1, Array2, ArrayView1, ArrayView2};
use std::ptr;

/// processing.rs:31 — consumes the signal like the reference
pub fn preemphasis(signal: Array1<f32>, shift: isize, cof: f32) -> Array1<f32> {
    let x = signal.as_standard_layout();
    let mut out = Array1::<f32>::zeros(x.len());
    unsafe { check(ffi::ss_preemphasis(x.as_ptr(), out.as_mut_ptr(), x.len() as i64, shift as i64, cof)) };
    out
}

/// processing.rs:65 — `filter` must be None (the reference's filter path is broken, quirk Q4)
pub fn stack_frames(signal: ArrayView1<f32>, sample_rate: usize, frame_length: f32, frame_stride: f32,
                    filter: Option<fn(usize) -> Array2<f32>>, zero_padding: bool) -> Array2<f32> {
    assert!(filter.is_none(), "stack_frames(filter = Some(..)) is unsupported (shape bug in the reference, processing.rs:122-126)");
    let x = signal.as_standard_layout();
    let l = (sample_rate as f32 * frame_length).round() as i64;
    let s = (sample_rate as f32 * frame_stride).round() as i64;
    let mut t: i64 = 0;
    unsafe { check(ffi::ss_stack_frames(x.as_ptr(), x.len() as i64, l, s, zero_padding as i32, 0, ptr::null_mut(), &mut t)) };
    let mut out = Array2::<f32>::zeros((t as usize, l as usize));
    unsafe { check(ffi::ss_stack_frames(x.as_ptr(), x.len() as i64, l, s, zero_padding as i32, 0, out.as_mut_ptr(), &mut t)) };
    out
}

/// processing.rs:179
pub fn power_spectrum(frames: Array2<f32>, fft_points: usize) -> Array2<f32> {
    let f = frames.as_standard_layout();
    let (rows, l) = f.dim();
    let mut out = Array2::<f32>::zeros((rows, fft_points / 2 + 1));
    unsafe { check(ffi::ss_power_spectrum(f.as_ptr(), rows as i64, l as i64, fft_points as i64, out.as_mut_ptr())) };
    out
}

/// processing.rs:194 (private in the reference)
pub fn log_power_spectrum(frames: Array2<f32>, fft_points: usize, normalize: bool) -> Array2<f32> {
    let f = frames.as_standard_layout();
    let (rows, l) = f.dim();
    let mut out = Array2::<f32>::zeros((rows, fft_points / 2 + 1));
    unsafe {
        check(ffi::ss_log_power_spectrum(f.as_ptr(), rows as i64, l as i64, fft_points as i64, normalize as i32, out.as_mut_ptr()))
    };
    out
}

/// processing.rs:265
pub fn cmvn(vec: ArrayView2<f32>, variance_normalization: bool) -> Array2<f32> {
    let v = vec.as_standard_layout();
    let (rows, cols) = v.dim();
    let mut out = Array2::<f32>::zeros((rows, cols));
    unsafe { check(ffi::ss_cmvn(v.as_ptr(), rows as i64, cols as i64, variance_normalization as i32, out.as_mut_ptr())) };
    out
}

/// processing.rs:222 — delta features along the feature axis (numpy 'edge' padding)
pub fn derivative_extraction(feat: &Array2<f32>, delta_windows: usize) -> Array2<f32> {
    let f = feat.as_standard_layout();
    let (rows, cols) = f.dim();
    let mut out = Array2::<f32>::zeros((rows, cols));
    unsafe { check(ffi::ss_derivative_extraction(f.as_ptr(), rows as i64, cols as i64, delta_windows as i32, out.as_mut_ptr())) };
    out
}

/// processing.rs:315 — sliding-window CMVN (numpy 'symmetric' padding along time)
pub fn cmvnw(vec: Array2<f32>, win_size: usize, variance_normalization: bool) -> Array2<f32> {
    let v = vec.as_standard_layout();
    let (rows, cols) = v.dim();
    let mut out = Array2::<f32>::zeros((rows, cols));
    unsafe { check(ffi::ss_cmvnw(v.as_ptr(), rows as i64, cols as i64, win_size as i64, variance_normalization as i32, out.as_mut_ptr())) };
    out
}
