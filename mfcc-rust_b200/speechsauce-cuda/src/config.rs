//! `speechsauce::config` with the same public surface (reference config.rs:10-190).  The FFT plan,
//! window, DCT handler and STFT overlap memory of the reference struct live in the device-side
//! `ss_plan`; the public scalar fields are kept so existing callers compile unchanged.
use crate::ffi;
use std::cell::RefCell;
use std::ptr;

pub(crate) fn check(code: i32) {
    // every error is a panic in the reference; keep that convention
    if code != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ss_last_error()) }.to_string_lossy().into_owned();
        panic!("speechsauce-cuda: [ss_status {}] {}", code, msg);
    }
}

#[derive(Default)]
pub struct SpeechConfigBuilder {
    sample_rate: usize,
    fft_points: usize,
    frame_length: f32,
    frame_stride: f32,
    num_cepstral: usize,
    num_filters: usize,
    low_frequency: f32,
    high_frequency: f32,
    dc_elimination: bool,
}

impl SpeechConfigBuilder {
    /// reference config.rs:35-47
    pub fn new(sample_rate: usize) -> SpeechConfigBuilder {
        SpeechConfigBuilder {
            sample_rate,
            fft_points: 512,
            frame_length: 0.02,
            frame_stride: 0.01,
            num_cepstral: 13,
            num_filters: 40,
            low_frequency: 0.0,
            high_frequency: sample_rate as f32 / 2.0,
            dc_elimination: true,
        }
    }
    pub fn high_freq(mut self, v: f32) -> Self { self.high_frequency = v; self }
    pub fn dc_elimination(mut self, v: bool) -> Self { self.dc_elimination = v; self }
    pub fn low_freq(mut self, v: f32) -> Self { self.low_frequency = v; self }
    pub fn num_cepstral(mut self, v: usize) -> Self { self.num_cepstral = v; self }
    /// not present in the reference builder (quirk Q15); added so 80/128-mel configs are reachable
    pub fn num_filters(mut self, v: usize) -> Self { self.num_filters = v; self }
    pub fn frame_stride(mut self, v: f32) -> Self { self.frame_stride = v; self }
    pub fn frame_length(mut self, v: f32) -> Self { self.frame_length = v; self }
    pub fn fft_points(mut self, v: usize) -> Self { self.fft_points = v; self }
    pub fn build(self) -> SpeechConfig {
        SpeechConfig::new(self.sample_rate, self.fft_points, self.frame_length, self.frame_stride, self.num_cepstral,
                          self.num_filters, self.low_frequency, self.high_frequency, self.dc_elimination)
    }
}

/// Send but not Sync, like the reference (its RefCell fields, config.rs:126,130): the raw plan
/// pointer makes the type !Sync and !Send by default; Send is restored explicitly.
pub struct SpeechConfig {
    pub sample_rate: usize,
    pub window_size: usize,
    pub window_size_half: usize,
    pub frame_length: f32,
    pub frame_stride: f32,
    pub num_cepstral: usize,
    pub num_filters: usize,
    pub low_frequency: f32,
    pub high_frequency: f32,
    pub freq_size: usize,
    pub frame_size: usize,
    pub dc_elimination: bool,
    pub wnorm: f32,
    /// the analysis window (reference config.rs:125,151-160), read back from the plan's tables
    pub window: Vec<f32>,
    /// the STFT overlap memory, `window_size - frame_size` samples (reference config.rs:126,162).  As in the
    /// reference it is the state itself: `stft1/2` and `mel_spectrogram1/2` read it before the first row and leave
    /// the memory after the last row in it; it is never reset behind the caller's back.
    pub analysis_mem: RefCell<Vec<f32>>,
    pub(crate) plan: *mut ffi::ss_plan,
    /// frame length / step as `stack_frames` derives them (round, processing.rs:77-78)
    pub(crate) frame_len: usize,
    pub(crate) frame_step: usize,
    raw: ffi::ss_config,
}

unsafe impl Send for SpeechConfig {}

impl Default for SpeechConfig {
    fn default() -> Self { SpeechConfigBuilder::new(16000).build() }
}

impl Clone for SpeechConfig {
    /// a clone owns its own plan (and its own STFT overlap memory, copied like the RefCell clone)
    fn clone(&self) -> Self {
        let c = SpeechConfig::from_raw(self.raw);
        *c.analysis_mem.borrow_mut() = self.analysis_mem.borrow().clone();
        c
    }
}

impl Drop for SpeechConfig {
    fn drop(&mut self) { unsafe { ffi::ss_plan_destroy(self.plan) } }
}

impl SpeechConfig {
    /// reference config.rs:140-150 — same nine arguments in the same order
    #[allow(clippy::too_many_arguments)]
    pub fn new(sample_rate: usize, fft_points: usize, frame_length: f32, frame_stride: f32, num_cepstral: usize,
               num_filters: usize, low_frequency: f32, high_frequency: f32, dc_elimination: bool) -> Self {
        Self::from_raw(ffi::ss_config {
            sample_rate: sample_rate as i64,
            fft_points: fft_points as i64,
            frame_length,
            frame_stride,
            num_cepstral: num_cepstral as i64,
            num_filters: num_filters as i64,
            low_frequency,
            high_frequency,
            dc_elimination: dc_elimination as i32,
            flags: 0,
            dct2_scale: 0.0,
        })
    }

    /// A plan for `feature::melspectrogram_librosa` (SS_FLAG_LIBROSA = 0x4): periodic Hann window of
    /// `n_fft`, centred reflect-padded frames every `hop_length` samples, Slaney mel bank.  The hop travels
    /// in `frame_length` (the library derives hop = trunc(frame_length * sample_rate) like config.rs:154).
    pub fn librosa(sample_rate: usize, n_fft: usize, hop_length: usize, n_mels: usize, fmin: f32, fmax: f32) -> Self {
        Self::from_raw(ffi::ss_config {
            sample_rate: sample_rate as i64,
            fft_points: n_fft as i64,
            frame_length: (hop_length as f32 + 0.5) / sample_rate as f32,
            frame_stride: hop_length as f32 / sample_rate as f32,
            num_cepstral: n_mels.min(13) as i64,
            num_filters: n_mels as i64,
            low_frequency: fmin,
            high_frequency: fmax,
            dc_elimination: 1,
            flags: ffi::SS_FLAG_LIBROSA,
            dct2_scale: 0.0,
        })
    }

    fn from_raw(raw: ffi::ss_config) -> Self {
        let mut plan: *mut ffi::ss_plan = ptr::null_mut();
        // n_devices = -1: every visible GPU; batches shard over them
        unsafe { check(ffi::ss_plan_create(&raw, ptr::null(), -1, &mut plan)) };
        let mut info = ffi::ss_plan_info::default();
        unsafe { check(ffi::ss_plan_get_info(plan, &mut info)) };
        SpeechConfig {
            sample_rate: raw.sample_rate as usize,
            window_size: raw.fft_points as usize,
            window_size_half: raw.fft_points as usize / 2,
            frame_length: raw.frame_length,
            frame_stride: raw.frame_stride,
            num_cepstral: raw.num_cepstral as usize,
            num_filters: raw.num_filters as usize,
            low_frequency: raw.low_frequency,
            high_frequency: raw.high_frequency,
            freq_size: info.freq_size as usize,
            frame_size: info.hop as usize,
            dc_elimination: raw.dc_elimination != 0,
            wnorm: info.wnorm,
            window: {
                let mut w = vec![0f32; raw.fft_points.max(0) as usize];
                unsafe { check(ffi::ss_debug_window(plan, w.as_mut_ptr())) };
                w
            },
            // config.rs:162: vec![0.0; window_size - frame_size] (empty where the reference's subtraction underflows)
            analysis_mem: RefCell::new(vec![0f32; (raw.fft_points.max(0) as usize).saturating_sub(info.hop.max(0) as usize)]),
            plan,
            frame_len: info.frame_len as usize,
            frame_step: info.frame_step as usize,
            raw,
        }
    }

    pub fn builder() -> SpeechConfigBuilder { SpeechConfigBuilder::default() }

    /// explicit form of the hidden STFT state (reference: `analysis_mem`, never reset)
    pub fn reset_analysis_mem(&self) { self.analysis_mem.borrow_mut().iter_mut().for_each(|v| *v = 0.0) }

    /// pointer handed to the C ABI as `analysis_mem_inout` (null when the STFT is undefined for this config, so that
    /// the library reports the reference's panic condition itself)
    pub(crate) fn analysis_mem_ptr(&self) -> *mut f32 {
        let mut m = self.analysis_mem.borrow_mut();
        let ok = self.frame_size > 0 && 2 * self.frame_size <= self.window_size && m.len() == self.window_size - self.frame_size;
        if ok { m.as_mut_ptr() } else { ptr::null_mut() }
    }
}
