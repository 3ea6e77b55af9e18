"""speechsauce_b200 — Python mirror of the speechsauce API over the B200 C-ABI.

Top level = the py-speechsauce package surface (py-speechsauce/speechsauce/__init__.py:37-132):
`mfcc`, `mel_spectrogram`, `cmvn`, `preemphasis` with the same keyword arguments and defaults and
the same lru_cache'd config factory.  Submodules mirror the Rust crate's modules: `config`,
`feature`, `processing`, `functions`."""
from functools import lru_cache

import numpy as np

from . import config, device, feature, functions, hostmem, processing  # noqa: F401
from ._lib import LIB_PATH, SpeechsauceError, lib  # noqa: F401
from .config import SpeechConfig, SpeechConfigBuilder  # noqa: F401
from .processing import cmvn, preemphasis  # noqa: F401

__all__ = ["mfcc", "mel_spectrogram", "preemphasis", "cmvn", "SpeechConfig", "SpeechConfigBuilder"]


@lru_cache(maxsize=32)
def _get_speech_config(sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13,
                       num_filters=40, fft_length=512, low_frequency=0, high_frequency=None,
                       dc_elimination=True):
    # py-speechsauce/speechsauce/__init__.py:8-34 -> _speech_config (py-speechsauce/src/lib.rs:226-254)
    return SpeechConfig(sampling_frequency, fft_length, frame_length, frame_stride, num_cepstral, num_filters,
                        low_frequency, high_frequency, dc_elimination)


def mfcc(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13, num_filters=40,
         fft_length=512, low_frequency=0, high_frequency=None, dc_elimination=True):
    """py-speechsauce/speechsauce/__init__.py:37-83."""
    cfg = _get_speech_config(sampling_frequency, frame_length, frame_stride, num_cepstral, num_filters,
                             fft_length, low_frequency, high_frequency, dc_elimination)
    return feature.mfcc(np.asarray(signal), cfg)


def mel_spectrogram(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13,
                    num_filters=40, fft_length=512, low_frequency=0, high_frequency=None, dc_elimination=True):
    """py-speechsauce/speechsauce/__init__.py:85-132; dispatches on ndim like py-speechsauce/src/lib.rs:179-204."""
    cfg = _get_speech_config(sampling_frequency, frame_length, frame_stride, num_cepstral, num_filters,
                             fft_length, low_frequency, high_frequency, dc_elimination)
    x = np.asarray(signal)
    if x.ndim == 1:
        return feature.mel_spectrogram1(x, cfg)
    if x.ndim == 2:
        return feature.mel_spectrogram2(x, cfg)
    raise ValueError("Input signal must be 1d or 2d")
