//! speechsauce-cuda: the public API of `speechsauce` (secretsauceai/mfcc-rust) served by hand-written
//! sm_100a CUDA kernels.  Module names and signatures match the reference crate's `lib.rs:2-6`, so
//! `use speechsauce_cuda as speechsauce;` is a drop-in for the calls on the hot path:
//!   feature::{mfcc, mfe, lmfe, ssc, mel_spectrogram1, mel_spectrogram2},
//!   processing::{preemphasis, stack_frames, power_spectrum, cmvn},
//!   functions::{stft1, stft2}, config::{SpeechConfig, SpeechConfigBuilder}.
//! There is no CPU fallback: without a CUDA device every call panics with the library's message.
pub mod config;
pub mod feature;
pub mod ffi;
pub mod functions;
pub mod processing;

#[cfg(test)]
mod tests {
    //! the reference's own tests (speechsauce/src/lib.rs:41-134) against this crate
    use crate::config::SpeechConfigBuilder;
    use crate::feature::{mfcc, mfe};
    use crate::processing::{cmvn, preemphasis, stack_frames};
    use ndarray::Array;

    fn signal() -> ndarray::Array1<f32> {
        // deterministic stand-in for Normal(0, 0.1): the assertions are shape / NaN only
        Array::from_iter((0..1_000_000u32).map(|i| (((i.wrapping_mul(2654435761)) >> 8) as f32 / 16777216.0 - 0.5) * 0.3))
    }

    #[test]
    fn test_preemphasis() {
        let s = signal();
        assert_eq!(preemphasis(s.clone(), 1, 0.98).shape(), s.shape());
    }

    #[test]
    fn test_stack_frames() {
        let s = signal();
        let frames = stack_frames(s.view(), 16000, 0.02, 0.02, None, true);
        assert_eq!(frames.shape()[0], ((1_000_000f32 - 320.0) / 320.0).ceil() as usize);
    }

    #[test]
    fn test_cmvn() {
        let v = Array::from_shape_fn((50, 100), |(i, j)| ((i * 131 + j * 71) % 97) as f32 / 97.0);
        let y = cmvn(v.view(), true);
        assert_eq!(y.shape(), v.shape());
        for c in y.mean_axis(ndarray::Axis(0)).unwrap().iter() { assert!(c.abs() < 1e-5); }
    }

    #[test]
    fn test_mfcc_mfe() {
        let s = signal();
        let cfg = SpeechConfigBuilder::new(16000).build();
        let m = mfcc(s.view(), &cfg);
        assert_eq!(m.shape(), &[6248, 13]);
        assert!(m.iter().all(|v| !v.is_nan()));
        let (f, e) = mfe(s.view(), &cfg);
        assert_eq!(f.shape(), &[6248, 40]);
        assert_eq!(e.shape(), &[6248]);
    }
}
