"""Host-side sharding helpers for the one-process-per-GPU deployment (torchrun).

The hot path shards trivially: clips / streams are independent, so each rank takes a contiguous
block of rows and there is NO data-path collective.  The one coupling in the reference is the STFT
overlap memory that `stft2` threads from each row into the next (functions.rs:99-160, quirk Q13):
a shard's first row needs the tail of the row before it.  That tail is a pure function of the
*input*, so it is restated on the host as an input halo (`stft_halo`) instead of being exchanged
between devices."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous block split, identical to the library's own split over a plan's devices
    (ss_api.cu run_on_devices): rows [n*rank/world, n*(rank+1)/world)"""
    return n_items * rank // world, n_items * (rank + 1) // world


def stft_halo(pcm: np.ndarray, row0: int, window_size: int, frame_size: int, mem0: np.ndarray | None = None) -> np.ndarray:
    """Overlap memory in front of row `row0`: the last W-H samples of
    mem0 ++ pad(row 0) ++ ... ++ pad(row row0-1), each row zero-padded to ceil(ttd/H)*H samples.
    Pass the result as `analysis_mem` to stft2 / mel_spectrogram2 for the shard starting at row0."""
    W, H = int(window_size), int(frame_size)
    memlen = W - H
    ttd = pcm.shape[1]
    span = int(np.ceil(np.float32(ttd) / np.float32(H))) * H
    out = np.zeros(memlen, np.float32)
    total = memlen + row0 * span
    for i in range(memlen):
        q = total - memlen + i
        if q < memlen:
            out[i] = 0.0 if mem0 is None else mem0[q]
        else:
            q -= memlen
            r, o = divmod(q, span)
            out[i] = pcm[r, o] if o < ttd else 0.0
    return out


def max_over_ranks(value: float, device=None) -> float:
    """max-reduce a scalar over the default process group (bench.py's timing rule)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
