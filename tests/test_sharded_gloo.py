"""CPU, gloo, world 2 and 3: the PARALLEL split of one recording (sdr-j-dab_b200/parallel.py: decode_sharded) with a
stand-in engine that has the real engine's contract: a tracking state that is predictable while locked, per-frame
outputs, 16 CIFs of de-interleaver memory with warm-up, state blob.  Checks that (i) a locked stream is decoded in
parallel and the concatenation equals the one-shot decode, (ii) a corrector that moves mid-stream is caught by the
boundary verification and the exact serial chain takes over, (iii) a receiver that is not locked after the lead-in
goes straight to the chain.  The real engine runs the same code in tests/test_multi_gpu.py."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

RATE = 2048000


class FakeEngine:
    frame_len, cifs_per_frame, frame_need = 1000, 4, 1100

    def __init__(self):
        self.abs_base, self.tail = 0, np.zeros(0, np.uint8)
        self.synced, self.f2, self.coarse, self.fine, self.lp = 0, 1, 0, 0, 0
        self.seen, self.hist, self.cifs_seen, self.frames = 0, [0] * 15, 0, 0

    # ---- state
    @staticmethod
    def make_state(**kw):
        return types.SimpleNamespace(**kw)

    def state_get(self):
        return self.make_state(abs_pos=self.abs_base, coarse=self.coarse, fine=self.fine, localPhase=self.lp,
                               f2Correction=self.f2, synced=self.synced, previous_1=0, previous_2=0, frames=self.frames, cifs=0)

    def state_set(self, s):
        self.abs_base, self.tail = s.abs_pos, np.zeros(0, np.uint8)
        self.coarse, self.fine, self.lp, self.f2, self.synced = s.coarse, s.fine, s.localPhase, s.f2Correction, s.synced

    def state_predict(self, s, n):
        assert s.synced == 1 and s.f2Correction == 0
        o = types.SimpleNamespace(**vars(s))
        o.abs_pos = s.abs_pos + n * self.frame_len
        o.localPhase = (s.localPhase - n * self.frame_len * (s.coarse + s.fine)) % RATE
        return o

    def export_state(self):
        meta = np.array([self.abs_base, self.synced, self.f2, self.coarse, self.fine, self.lp, self.seen, self.cifs_seen,
                         self.frames] + self.hist, np.int64)
        return np.concatenate([np.frombuffer(meta.tobytes(), np.uint8), self.tail])

    def import_state(self, blob):
        meta = np.frombuffer(blob[:24 * 8].tobytes(), np.int64)
        (self.abs_base, self.synced, self.f2, self.coarse, self.fine, self.lp, self.seen, self.cifs_seen, self.frames) = [int(x) for x in meta[:9]]
        self.hist = [int(x) for x in meta[9:24]]
        self.tail = blob[24 * 8:].copy()

    # ---- decode
    def decode(self, iq, max_frames):
        data = np.concatenate([self.tail, np.asarray(iq, np.uint8)])
        n, pos = data.size // 2, 0
        r = types.SimpleNamespace(nframes=0, info=[], fic=[], msc=[[]])
        while r.nframes < max_frames and n - pos >= self.frame_need:
            blk = data[2 * pos:2 * (pos + self.frame_len)].astype(np.int64)
            if not self.synced:                               # "acquisition": three frames, then locked
                self.seen += 1
                self.synced = 1
            if self.f2 and self.seen >= 3:
                self.f2 = 0
            self.seen += 1
            r.info.append(types.SimpleNamespace(pos=self.abs_base + pos, coarse=self.coarse, fine=self.fine, phase0=self.lp))
            r.fic.append(int((blk.sum() * 31 + self.fine * 7 + self.lp) % 1000003))
            for c in range(4):
                v = int(blk[500 * c:500 * (c + 1)].sum()) + self.fine
                if self.cifs_seen >= 16:
                    r.msc[0].append(int((sum((i + 1) * h for i, h in enumerate(self.hist)) + 17 * v) % 1000003))
                self.hist = self.hist[1:] + [v]
                self.cifs_seen += 1
            self.lp = (self.lp - self.frame_len * (self.coarse + self.fine)) % RATE
            if blk[0] == 255 and blk[1] == 255:               # the AFC integrator moves on this frame
                self.fine += 3
            pos += self.frame_len
            r.nframes += 1
            self.frames += 1
        self.tail = data[2 * pos:]
        self.abs_base += pos
        return r

    @staticmethod
    def drop_frames(r, n):
        return types.SimpleNamespace(nframes=r.nframes - n, info=r.info[n:], fic=r.fic[n:], msc=r.msc)

    @staticmethod
    def concat_results(a, b):
        return types.SimpleNamespace(nframes=a.nframes + b.nframes, info=a.info + b.info, fic=a.fic + b.fic,
                                     msc=[x + y for x, y in zip(a.msc, b.msc)])


def _recording(kind):
    rng = np.random.default_rng(11)
    iq = rng.integers(0, 255, 2 * 61300, dtype=np.uint8)      # 61.3 frames; 255 never occurs by chance
    if kind == "drift":
        iq[2 * 37000:2 * 37000 + 2] = 255                     # frame 37 moves the corrector
    return iq


def _worker(rank, world, port, kind, lead, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    res, first, mode = par.decode_sharded(FakeEngine(), _recording(kind), lambda n: n, rank, world, dist, "cpu", lead_frames=lead)
    q.put((rank, mode, res.fic, res.msc[0], [i.pos for i in res.info]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,lead,want_mode", [(2, "locked", 8, "parallel"), (3, "locked", 8, "parallel"),
                                                       (3, "drift", 8, "chain"), (2, "locked", 2, "chain")])
def test_sharded_decode_equals_one_shot(world, kind, lead, want_mode):
    one = FakeEngine().decode(_recording(kind), 10 ** 6)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, lead, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g[1] == want_mode for g in got), [g[1] for g in got]
    assert sum((g[2] for g in got), []) == one.fic
    assert sum((g[3] for g in got), []) == one.msc[0]
    assert sum((g[4] for g in got), []) == [i.pos for i in one.info]
    assert all(len(g[2]) > 0 for g in got)


def test_frame_shards():
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    assert par.frame_shards(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert par.frame_shards(9, 4, first_min=4) == [(0, 4), (4, 5), (5, 7), (7, 9)]
    assert par.frame_shards(0, 2) == [(0, 0), (0, 0)]
