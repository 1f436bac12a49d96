"""CPU, world_size 2, gloo: the split-recording hand-over logic of sdr-j-dab_b200/parallel.py.  The engine needs a
GPU, so a stand-in decoder with the same interface is used: its output depends on everything it has seen (a
running hash), hence the split decode equals the one-shot decode only if ranges, state export/import and the
neighbour send/recv are right.  The real engine goes through the same code in tests/test_multi_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import engine_pkg


class ChainDecoder:
    """stateful stand-in: consumes whole 'frames' of 1000 samples, keeps the remainder as tail"""
    FRAME = 1000

    def __init__(self):
        self.h, self.tail = np.uint64(1469598103934665603), np.zeros(0, np.uint8)

    def decode(self, iq, out):
        data = np.concatenate([self.tail, np.asarray(iq, np.uint8)])
        n = data.size // (2 * self.FRAME)
        for f in range(n):
            blk = data[2 * self.FRAME * f:2 * self.FRAME * (f + 1)]
            self.h = (self.h ^ np.uint64(int(blk.astype(np.uint64).sum()))) * np.uint64(1099511628211)
            out.append(int(self.h))
        self.tail = data[2 * self.FRAME * n:]
        return out

    def export_state(self):
        return np.concatenate([np.frombuffer(np.uint64(self.h).tobytes(), np.uint8), self.tail])

    def import_state(self, blob):
        self.h = np.frombuffer(blob[:8].tobytes(), np.uint64)[0]
        self.tail = blob[8:].copy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    rng = np.random.default_rng(7)
    iq = rng.integers(0, 256, 2 * 10370, dtype=np.uint8)          # 10.37 frames
    a, b = par.shard_ranges(iq.size // 2, world, 777)[rank]       # cut lengths unrelated to the frame length
    res = par.decode_split(ChainDecoder(), iq[2 * a:2 * b], [], rank, world, dist)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_split_decode_equals_one_shot():
    with np.errstate(over="ignore"):
        rng = np.random.default_rng(7)
        iq = rng.integers(0, 256, 2 * 10370, dtype=np.uint8)
        want = ChainDecoder().decode(iq, [])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] + got[1] == want and len(got[0]) > 0 and len(got[1]) > 0


def test_shard_ranges():
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    r = par.shard_ranges(10 * 196608 + 5000, 4, 196608)
    assert r[0][0] == 0 and r[-1][1] == 10 * 196608 + 5000
    assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    assert [(b - a) // 196608 for a, b in r] == [3, 3, 2, 2]
    assert par.shard_ranges(100, 1, 196608) == [(0, 100)]
