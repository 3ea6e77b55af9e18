"""Times functions::stft2-shaped output (complex [ch][T'][F]) device-resident: the Whisper-shaped config c4
on a slice (the full batch's output is 39.5 GB), register-FFT kernel vs the generic one."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "mfcc-rust_b200"), str(ROOT / "oracle")]
import torch  # noqa: E402
import speechsauce_b200 as sb  # noqa: E402
from speechsauce_b200 import _lib  # noqa: E402

C4 = dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01, num_cepstral=13, num_filters=80,
          low_frequency=0.0, high_frequency=8000.0)
C3 = dict(sample_rate=22050, fft_points=2048, frame_length=512 / 22050, frame_stride=0.01, num_cepstral=13,
          num_filters=128, low_frequency=0.0, high_frequency=11025.0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for shape, kw, n, ttd, T, F in (("c4", C4, 1024, 480000, 3000, 201), ("c3", C3, 2048, 110250, 216, 1025)):
  x = torch.empty((n, ttd), dtype=torch.float32, device="cuda").normal_(0, 0.1)
  out = torch.empty((n, T, F, 2), dtype=torch.float32, device="cuda")
  for name, flags in (("fast", 0), ("generic", _lib.SS_FLAG_FORCE_GENERIC)):
    cfg = sb.SpeechConfig(**kw, devices=[0], flags=flags)

    def step():
        _lib.check(sb.lib.ss_stft_dev(cfg.plan, 0, sp, C.c_void_p(x.data_ptr()), n, ttd, ttd, C.c_void_p(out.data_ptr()), None, None))
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (x.numel() + out.numel()) * 4 / 1e9
    print(f"stft {shape}-shaped, {n} rows: {name:8s} {sb.lib.ss_last_kernel_name().decode():18s} {ms:8.3f} ms  {gb / ms * 1e3:7.1f} GB/s algorithmic")

# librosa-compatible mel spectrogram (SS_FLAG_LIBROSA) on the same two shapes: n_fft = 400 on the register-FFT
# kernel (full-spectrum tile), everything else on the any-length generic kernel
for shape, sr, n_fft, hop, n_mels, n, ttd in (("c4", 16000, 400, 160, 80, 1024, 480000), ("c3", 22050, 2048, 512, 128, 2048, 110250),
                                              ("n_fft 1024", 22050, 1024, 256, 128, 2048, 110250), ("n_fft 512", 16000, 512, 160, 80, 2048, 160000), ("n_fft 256", 8000, 256, 80, 40, 4096, 80000)):
  x = torch.empty((n, ttd), dtype=torch.float32, device="cuda").normal_(0, 0.1)
  T = 1 + ttd // hop
  out = torch.empty((n, n_mels, T), dtype=torch.float32, device="cuda")
  for name, flags in (("fast", 0), ("generic", _lib.SS_FLAG_FORCE_GENERIC)):
    cfg = sb.feature.librosa_config(sr, n_fft, hop, n_mels, devices=[0], flags=flags)

    def step():
        _lib.check(sb.lib.ss_librosa_melspectrogram_dev(cfg.plan, 0, sp, C.c_void_p(x.data_ptr()), n, ttd, ttd, C.c_void_p(out.data_ptr()), 1))
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"librosa log-mel {shape}-shaped, {n} rows: {name:8s} {sb.lib.ss_last_kernel_name().decode():20s} {ms:8.3f} ms  {n * ttd / sr / ms * 1e3:.3e} audio-s/s")
