//! `speechsauce::feature` — same signatures (reference feature.rs:99,151,163,200,242), device path.
use crate::config::{check, SpeechConfig};
use crate::ffi;
use ndarray::{Array1, Array2, Array3, ArrayView1, ArrayView2};

fn frames(cfg: &SpeechConfig, n: usize) -> usize {
    let t = unsafe { ffi::ss_num_frames(n as i64, cfg.frame_len as i64, cfg.frame_step as i64, 0) };
    if t < 1 { panic!("speechsauce-cuda: signal of {} samples yields no frame (reference underflows, processing.rs:101-105)", n) }
    t as usize
}

/// feature.rs:99 — `mfcc(signal, &config) -> Array2<f32>` of shape (num_frames, num_cepstral)
pub fn mfcc(signal: ArrayView1<f32>, cfg: &SpeechConfig) -> Array2<f32> {
    let x = signal.as_standard_layout();
    let t = frames(cfg, x.len());
    let mut out = Array2::<f32>::zeros((t, cfg.num_cepstral));
    unsafe { check(ffi::ss_mfcc(cfg.plan, x.as_ptr(), 1, x.len() as i64, x.len() as i64, out.as_mut_ptr(), ffi::SS_CMVN_NONE)) };
    out
}

/// batched form: one clip per row, clips sharded over the GPUs; `cmvn_mode` applies
/// processing::cmvn per clip inside the same kernel
pub fn mfcc_batch(signals: ArrayView2<f32>, cfg: &SpeechConfig, cmvn_mode: i32) -> Array3<f32> {
    let x = signals.as_standard_layout();
    let (n, len) = x.dim();
    let t = frames(cfg, len);
    let mut out = Array3::<f32>::zeros((n, t, cfg.num_cepstral));
    unsafe { check(ffi::ss_mfcc(cfg.plan, x.as_ptr(), n as i64, len as i64, len as i64, out.as_mut_ptr(), cmvn_mode)) };
    out
}

/// feature.rs:200 — `mfe(signal, &config) -> (features [T, nf], frame_energies [T])`
pub fn mfe(signal: ArrayView1<f32>, cfg: &SpeechConfig) -> (Array2<f32>, Array1<f32>) {
    let x = signal.as_standard_layout();
    let t = frames(cfg, x.len());
    let mut feat = Array2::<f32>::zeros((t, cfg.num_filters));
    let mut energy = Array1::<f32>::zeros(t);
    unsafe { check(ffi::ss_mfe(cfg.plan, x.as_ptr(), 1, x.len() as i64, x.len() as i64, feat.as_mut_ptr(), energy.as_mut_ptr())) };
    (feat, energy)
}

/// feature.rs:242 (private in the reference, public here: decision D3)
pub fn lmfe(signal: ArrayView1<f32>, cfg: &SpeechConfig) -> Array2<f32> {
    let x = signal.as_standard_layout();
    let t = frames(cfg, x.len());
    let mut out = Array2::<f32>::zeros((t, cfg.num_filters));
    unsafe { check(ffi::ss_lmfe(cfg.plan, x.as_ptr(), 1, x.len() as i64, x.len() as i64, out.as_mut_ptr())) };
    out
}

/// spectral subband centroids (README.md:15 names it; defined in DESIGN.md, decision D4)
pub fn ssc(signal: ArrayView1<f32>, cfg: &SpeechConfig) -> Array2<f32> {
    let x = signal.as_standard_layout();
    let t = frames(cfg, x.len());
    let mut out = Array2::<f32>::zeros((t, cfg.num_filters));
    unsafe { check(ffi::ss_ssc(cfg.plan, x.as_ptr(), 1, x.len() as i64, x.len() as i64, out.as_mut_ptr())) };
    out
}

/// feature.rs:163 — `mel_spectrogram2(signal [ch, ttd], &config) -> Array3<f32>` [ch, nf, T']
pub fn mel_spectrogram2(signal: ArrayView2<f32>, cfg: &SpeechConfig) -> Array3<f32> {
    let x = signal.as_standard_layout();
    let (ch, ttd) = x.dim();
    let tp = unsafe { ffi::ss_stft_num_frames(ttd as i64, cfg.frame_size as i64) }.max(0) as usize;
    let mut out = Array3::<f32>::zeros((ch, cfg.num_filters, tp));
    unsafe { check(ffi::ss_mel_spectrogram(cfg.plan, x.as_ptr(), ch as i64, ttd as i64, ttd as i64, out.as_mut_ptr(), cfg.analysis_mem_ptr(), 0)) };
    out
}

/// librosa.feature.melspectrogram(y, sr, n_fft, hop_length, window="hann", center=True, pad_mode="reflect",
/// power=2.0, n_mels, fmin, fmax, htk=False, norm="slaney") on the rows of `signal`, for a config made by
/// `SpeechConfig::librosa` -> [ch, n_mels, 1 + ttd / hop]; `db`: power_to_db(ref=1.0, amin=1e-10, top_db=None).
/// Not in the reference (its README names librosa compatibility as the goal).
pub fn melspectrogram_librosa(signal: ArrayView2<f32>, cfg: &SpeechConfig, db: bool) -> Array3<f32> {
    let x = signal.as_standard_layout();
    let (ch, ttd) = x.dim();
    let tp = unsafe { ffi::ss_librosa_num_frames(ttd as i64, cfg.frame_size as i64) }.max(0) as usize;
    let mut out = Array3::<f32>::zeros((ch, cfg.num_filters, tp));
    unsafe { check(ffi::ss_librosa_melspectrogram(cfg.plan, x.as_ptr(), ch as i64, ttd as i64, ttd as i64, out.as_mut_ptr(), db as i32)) };
    out
}

/// feature.rs:151 — intended semantics (decision D2): one-row mel_spectrogram2 -> [nf, T']
pub fn mel_spectrogram1(signal: ArrayView1<f32>, cfg: &SpeechConfig) -> Array2<f32> {
    let x = signal.as_standard_layout();
    let ttd = x.len();
    let tp = unsafe { ffi::ss_stft_num_frames(ttd as i64, cfg.frame_size as i64) }.max(0) as usize;
    let mut out = Array2::<f32>::zeros((cfg.num_filters, tp));
    unsafe { check(ffi::ss_mel_spectrogram(cfg.plan, x.as_ptr(), 1, ttd as i64, ttd as i64, out.as_mut_ptr(), cfg.analysis_mem_ptr(), 0)) };
    out
}

/// feature.rs:253 (private in the reference) — [T, C] -> [T, C, 3] = static, delta, delta-delta
pub fn extract_derivative_feature(feature: Array2<f32>) -> Array3<f32> {
    let x = feature.as_standard_layout();
    let (rows, cols) = x.dim();
    let mut out = Array3::<f32>::zeros((rows, cols, 3));
    unsafe { check(ffi::ss_extract_derivative_feature(x.as_ptr(), rows as i64, cols as i64, out.as_mut_ptr())) };
    out
}
