"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic: contiguous sharding with no
collective on the data path, the STFT input halo that replaces the reference's row-to-row overlap
memory, and the max-over-ranks timing reduction bench.py uses.  The oracle stands in for the
device here (this container has no GPU); the same halo is checked against the CUDA path in
tests/test_gpu_parity.py::test_stft_explicit_memory_equals_plan_memory_and_sharding_halo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import speechsauce_oracle as o
from speechsauce_b200 import sharding

C4 = dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01, num_cepstral=13,
          num_filters=80, low_frequency=0.0, high_frequency=8000.0)
C1 = dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01, num_cepstral=13,
          num_filters=40, low_frequency=0.0, high_frequency=8000.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- speechpy path: clips are independent, shards concatenate to the whole
        x = o.synth_signal(7, 16000, 5)
        a, b = sharding.shard_range(x.shape[0], rank, world)
        cfg = o.SpeechConfig(**C1)
        mine = np.stack([o.mfcc(c, cfg) for c in x[a:b]])
        np.save(os.path.join(tmp, f"mfcc_{rank}.npy"), mine)
        # ---- STFT path: rows are coupled through the overlap memory -> host-side halo
        y = o.synth_signal(6, 3200, 9)
        a, b = sharding.shard_range(y.shape[0], rank, world)
        c4 = o.SpeechConfig(**C4)
        c4.analysis_mem = sharding.stft_halo(y, a, c4.window_size, c4.frame_size)
        np.save(os.path.join(tmp, f"stft_{rank}.npy"), o.stft2(y[a:b], c4))
        # ---- timing rule: max over ranks
        t = sharding.max_over_ranks(1.0 + rank)
        assert t == float(world)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x = o.synth_signal(7, 16000, 5)
    whole = np.stack([o.mfcc(c, o.SpeechConfig(**C1)) for c in x])
    got = np.concatenate([np.load(tmp_path / f"mfcc_{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, whole)
    y = o.synth_signal(6, 3200, 9)
    whole = o.stft2(y, o.SpeechConfig(**C4))
    got = np.concatenate([np.load(tmp_path / f"stft_{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, whole)  # the halo reproduces the sequential row-to-row state exactly


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 65536, 1024):
        for world in (1, 2, 4, 8):
            rs = [sharding.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def test_halo_equals_oracle_memory_after_rows():
    y = o.synth_signal(5, 1000, 3)  # ttd not a multiple of H: the tail includes zero padding
    cfg = o.SpeechConfig(**C4)
    for r in range(1, 5):
        c = o.SpeechConfig(**C4)
        o.stft2(y[:r], c)
        np.testing.assert_array_equal(sharding.stft_halo(y, r, cfg.window_size, cfg.frame_size), c.analysis_mem)
    tiny = o.synth_signal(4, 100, 4)  # rows shorter than the memory: the halo spans several rows
    for r in range(1, 4):
        c = o.SpeechConfig(**C4)
        o.stft2(tiny[:r], c)
        np.testing.assert_array_equal(sharding.stft_halo(tiny, r, 400, 160), c.analysis_mem)
