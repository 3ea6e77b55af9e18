//! `extern "C"` declarations of include/speechsauce_cuda.h — one to one, nothing else.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ss_config {
    pub sample_rate: i64,
    pub fft_points: i64,
    pub frame_length: f32,
    pub frame_stride: f32,
    pub num_cepstral: i64,
    pub num_filters: i64,
    pub low_frequency: f32,
    pub high_frequency: f32,
    pub dc_elimination: i32,
    pub flags: u32,
    pub dct2_scale: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ss_plan_info {
    pub frame_len: i64,
    pub frame_step: i64,
    pub hop: i64,
    pub n_fft: i64,
    pub freq_size: i64,
    pub n_pad: i64,
    pub num_filters: i64,
    pub num_cepstral: i64,
    pub wnorm: f32,
    pub n_devices: i32,
    pub fast_path: i32,
    pub fast_stft: i32,
}

#[repr(C)]
pub struct ss_feature_outputs {
    pub mfe: *mut f32,
    pub energy: *mut f32,
    pub lmfe: *mut f32,
    pub ssc: *mut f32,
    pub mfcc: *mut f32,
    pub mfcc_cmvn_mode: i32,
    pub reserved: i32,
}

#[repr(C)]
pub struct ss_plan {
    _private: [u8; 0],
}

pub const SS_FLAG_FRAMING_LITERAL: u32 = 0x1;
pub const SS_FLAG_FORCE_GENERIC: u32 = 0x2;
pub const SS_FLAG_LIBROSA: u32 = 0x4;
pub const SS_CMVN_NONE: c_int = 0;
pub const SS_CMVN_MEAN: c_int = 1;
pub const SS_CMVN_MEAN_VAR: c_int = 2;

extern "C" {
    pub fn ss_device_count() -> c_int;
    pub fn ss_last_error() -> *const c_char;
    pub fn ss_version() -> *const c_char;
    pub fn ss_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn ss_host_free(ptr: *mut c_void) -> c_int;
    pub fn ss_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn ss_host_unregister(ptr: *mut c_void) -> c_int;
    pub fn ss_host_is_pinned(ptr: *const c_void) -> c_int;
    pub fn ss_copy_threads() -> c_int;
    pub fn ss_debug_copy2d(dst: *mut c_void, dst_pitch: usize, src: *const c_void, src_pitch: usize, row_bytes: usize, rows: usize) -> c_int;
    pub fn ss_release_scratch() -> c_int;
    pub fn ss_plan_create(cfg: *const ss_config, devices: *const c_int, n_devices: c_int, out: *mut *mut ss_plan) -> c_int;
    pub fn ss_plan_destroy(plan: *mut ss_plan);
    pub fn ss_plan_get_info(plan: *const ss_plan, info: *mut ss_plan_info) -> c_int;
    pub fn ss_plan_set_tuning(plan: *mut ss_plan, tile_frames: i64, chunk_bytes: i64) -> c_int;
    pub fn ss_plan_get_analysis_mem(plan: *const ss_plan, out: *mut f32, n: i64) -> c_int;
    pub fn ss_plan_set_analysis_mem(plan: *mut ss_plan, input: *const f32, n: i64) -> c_int;
    pub fn ss_num_frames(n: i64, frame_len: i64, frame_step: i64, zero_padding: c_int) -> i64;
    pub fn ss_stft_num_frames(ttd: i64, hop: i64) -> i64;
    pub fn ss_features(plan: *mut ss_plan, pcm: *const f32, n_clips: i64, clip_len: i64, row_stride: i64,
                       outs: *const ss_feature_outputs) -> c_int;
    pub fn ss_mfe(plan: *mut ss_plan, pcm: *const f32, n_clips: i64, clip_len: i64, row_stride: i64,
                  feat: *mut f32, energy: *mut f32) -> c_int;
    pub fn ss_lmfe(plan: *mut ss_plan, pcm: *const f32, n_clips: i64, clip_len: i64, row_stride: i64, out: *mut f32) -> c_int;
    pub fn ss_mfcc(plan: *mut ss_plan, pcm: *const f32, n_clips: i64, clip_len: i64, row_stride: i64,
                   out: *mut f32, cmvn_mode: c_int) -> c_int;
    pub fn ss_ssc(plan: *mut ss_plan, pcm: *const f32, n_clips: i64, clip_len: i64, row_stride: i64, out: *mut f32) -> c_int;
    pub fn ss_stft(plan: *mut ss_plan, pcm: *const f32, ch: i64, ttd: i64, row_stride: i64, out_ri: *mut f32,
                   analysis_mem_inout: *mut f32) -> c_int;
    pub fn ss_mel_spectrogram(plan: *mut ss_plan, pcm: *const f32, ch: i64, ttd: i64, row_stride: i64, out: *mut f32,
                              analysis_mem_inout: *mut f32, log_mode: c_int) -> c_int;
    pub fn ss_preemphasis(input: *const f32, out: *mut f32, n: i64, shift: i64, cof: f32) -> c_int;
    pub fn ss_preemphasis_batch(input: *const f32, out: *mut f32, n_rows: i64, n: i64, row_stride: i64, shift: i64, cof: f32) -> c_int;
    pub fn ss_stack_frames(input: *const f32, n: i64, frame_len: i64, frame_step: i64, zero_padding: c_int, flags: u32,
                           out: *mut f32, numframes: *mut i64) -> c_int;
    pub fn ss_power_spectrum(frames: *const f32, rows: i64, frame_len: i64, n_fft: i64, out: *mut f32) -> c_int;
    pub fn ss_librosa_melspectrogram(plan: *mut ss_plan, pcm: *const f32, ch: i64, ttd: i64, row_stride: i64, out: *mut f32,
                                     db_mode: c_int) -> c_int;
    pub fn ss_librosa_melspectrogram_dev(plan: *mut ss_plan, dev_index: c_int, cuda_stream: *mut c_void, d_pcm: *const f32,
                                         ch: i64, ttd: i64, row_stride: i64, d_out: *mut f32, db_mode: c_int) -> c_int;
    pub fn ss_librosa_num_frames(ttd: i64, hop: i64) -> i64;
    pub fn ss_log_power_spectrum(frames: *const f32, rows: i64, frame_len: i64, n_fft: i64, normalize: c_int, out: *mut f32) -> c_int;
    pub fn ss_log_power_spectrum_dev(cuda_stream: *mut c_void, d_frames: *const f32, rows: i64, frame_len: i64, n_fft: i64,
                                     normalize: c_int, d_out: *mut f32) -> c_int;
    pub fn ss_cmvn(input: *const f32, rows: i64, cols: i64, variance_normalization: c_int, out: *mut f32) -> c_int;
    pub fn ss_cmvn_batch(input: *const f32, n_mats: i64, rows: i64, cols: i64, variance_normalization: c_int, out: *mut f32) -> c_int;
    pub fn ss_features_dev(plan: *mut ss_plan, dev_index: c_int, cuda_stream: *mut c_void, d_pcm: *const f32, n_clips: i64,
                           clip_len: i64, row_stride: i64, d_outs: *const ss_feature_outputs) -> c_int;
    pub fn ss_stft_dev(plan: *mut ss_plan, dev_index: c_int, cuda_stream: *mut c_void, d_pcm: *const f32, ch: i64, ttd: i64,
                       row_stride: i64, d_out_ri: *mut f32, d_mem_in: *const f32, d_mem_out: *mut f32) -> c_int;
    pub fn ss_mel_spectrogram_dev(plan: *mut ss_plan, dev_index: c_int, cuda_stream: *mut c_void, d_pcm: *const f32, ch: i64,
                                  ttd: i64, row_stride: i64, d_out: *mut f32, d_mem_in: *const f32, d_mem_out: *mut f32,
                                  log_mode: c_int) -> c_int;
    pub fn ss_preemphasis_dev(cuda_stream: *mut c_void, d_in: *const f32, d_out: *mut f32, n_rows: i64, n: i64, row_stride: i64,
                              shift: i64, cof: f32) -> c_int;
    pub fn ss_stack_frames_dev(cuda_stream: *mut c_void, d_in: *const f32, n: i64, frame_len: i64, frame_step: i64,
                               zero_padding: c_int, flags: u32, d_out: *mut f32, numframes: i64) -> c_int;
    pub fn ss_power_spectrum_dev(cuda_stream: *mut c_void, d_frames: *const f32, rows: i64, frame_len: i64, n_fft: i64,
                                 d_out: *mut f32) -> c_int;
    pub fn ss_cmvn_dev(cuda_stream: *mut c_void, d_in: *const f32, n_mats: i64, rows: i64, cols: i64,
                       variance_normalization: c_int, d_out: *mut f32) -> c_int;
    pub fn ss_derivative_extraction(input: *const f32, rows: i64, cols: i64, delta_windows: c_int, out: *mut f32) -> c_int;
    pub fn ss_extract_derivative_feature(input: *const f32, rows: i64, cols: i64, out_cube: *mut f32) -> c_int;
    pub fn ss_cmvnw(input: *const f32, rows: i64, cols: i64, win_size: i64, variance_normalization: c_int, out: *mut f32) -> c_int;
    pub fn ss_derivative_extraction_dev(cuda_stream: *mut c_void, d_in: *const f32, rows: i64, cols: i64, delta_windows: c_int,
                                        d_out: *mut f32) -> c_int;
    pub fn ss_extract_derivative_feature_dev(cuda_stream: *mut c_void, d_in: *const f32, rows: i64, cols: i64,
                                             d_out_cube: *mut f32) -> c_int;
    pub fn ss_cmvnw_dev(cuda_stream: *mut c_void, d_in: *const f32, n_mats: i64, rows: i64, cols: i64, win_size: i64,
                        variance_normalization: c_int, d_out: *mut f32) -> c_int;
    pub fn ss_debug_filterbank(plan: *const ss_plan, out_nf_by_f: *mut f32) -> c_int;
    pub fn ss_debug_window(plan: *const ss_plan, out_w: *mut f32) -> c_int;
    pub fn ss_kernel_launch_count() -> i64;
    pub fn ss_last_kernel_name() -> *const c_char;
}
