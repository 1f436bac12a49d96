"""GPU: state export/import (the multi-GPU hand-over) on one device, and -- when two devices are visible -- the
split-recording chain over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

import dabmod
from util import engine_pkg

pytestmark = pytest.mark.gpu
SUB = [(0, 96, 128, 1, 0o103), (100, 24, 32, 0, 5)]


def _stream(port, nframes=26, seed=31):
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103), (100, 32, 0, 5)], seed)
    return mod.generate(nframes, cfo_hz=-1220.0, snr_db=20.0, lead=3000, tail=6000)["iq"]


def test_state_blob_handover_single_gpu(port):
    """engine A decodes the first part, exports; a fresh engine B imports and continues: equal to one shot"""
    pkg = engine_pkg()
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    iq = _stream(port)
    one = pkg.DabGpu(mode=1); one.set_subchannels(SUB)
    want = one.decode(iq, one.alloc_result(40))
    ranges = par.shard_ranges(iq.size // 2, 3, 196608)
    parts, blob = [], None
    for a, b in ranges:
        e = pkg.DabGpu(mode=1); e.set_subchannels(SUB)
        if blob is not None:
            e.import_state(blob)
        parts.append(e.decode(iq[2 * a:2 * b], e.alloc_result(40)))
        blob = e.export_state()
        e.close()
    assert sum(p.nframes for p in parts) == want.nframes
    assert np.array_equal(np.concatenate([p.soft for p in parts]), want.soft)
    assert np.array_equal(np.concatenate([p.fic_bits for p in parts]), want.fic_bits)
    for i in range(len(SUB)):
        assert np.array_equal(np.concatenate([p.msc[i] for p in parts]), want.msc[i])
    assert [(i.pos, i.fine) for p in parts for i in p.info] == [(i.pos, i.fine) for i in want.info]
    one.close()


def _worker(rank, world, port_no, iq, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = engine_pkg()
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    e = pkg.DabGpu(mode=1, device=rank); e.set_subchannels(SUB)
    a, b = par.shard_ranges(iq.size // 2, world, 196608)[rank]
    r = par.decode_split(e, iq[2 * a:2 * b], e.alloc_result(40), rank, world, dist, torch.device("cuda", rank))
    q.put((rank, r.nframes, r.soft.copy(), [m.copy() for m in r.msc]))
    dist.barrier()
    dist.destroy_process_group()


def test_split_recording_two_gpus_nccl(port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    pkg = engine_pkg()
    iq = _stream(port)
    one = pkg.DabGpu(mode=1); one.set_subchannels(SUB)
    want = one.decode(iq, one.alloc_result(40))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, iq, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert got[0][1] + got[1][1] == want.nframes and got[1][1] > 0
    assert np.array_equal(np.concatenate([got[0][2], got[1][2]]), want.soft)
    for i in range(len(SUB)):
        assert np.array_equal(np.concatenate([got[0][3][i], got[1][3][i]]), want.msc[i])
    one.close()


# ---- one recording, shards decoded in parallel from the predicted tracking state (parallel.decode_sharded) ----
def _sharded_worker(rank, world, port_no, iq, q, backend, lead):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    gpu = rank if backend == "nccl" else 0                    # gloo: all ranks share GPU 0 (the plumbing runs on CPU tensors)
    torch.cuda.set_device(gpu)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", gpu))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = engine_pkg()
    par = __import__("importlib").import_module("sdr-j-dab_b200.parallel")
    e = pkg.DabGpu(mode=1, device=gpu); e.set_subchannels(SUB)
    dev = torch.device("cuda", gpu) if backend == "nccl" else "cpu"
    r, first, mode = par.decode_sharded(e, iq, lambda n: e.alloc_result(n), rank, world, dist, dev, lead_frames=lead)
    q.put((rank, mode, r.nframes, r.soft.copy(), r.fic_bits.copy(), [m.copy() for m in r.msc], [(i.pos, i.fine, i.phase0) for i in r.info]))
    dist.barrier()
    dist.destroy_process_group()
    e.close()


@pytest.mark.parametrize("world,cfo,want_mode", [(2, 8.0, "parallel"), (3, 8.0, "parallel"), (2, -1220.0, "chain")])
def test_sharded_recording_equals_one_shot(port, world, cfo, want_mode):
    """cfo 8 Hz: inside the fine corrector's dead zone, the receiver is locked right after the coarse search -> the
    shards run in parallel; cfo -1220 Hz: the integrator is still moving after a 16-frame lead-in -> the boundary
    check fails and the serial chain takes over.  Either way the output equals the one-GPU decode."""
    import torch
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    pkg = engine_pkg()
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103), (100, 32, 0, 5)], 57)
    iq = mod.generate(46, cfo_hz=cfo, snr_db=20.0, lead=3000, tail=6000)["iq"]
    one = pkg.DabGpu(mode=1); one.set_subchannels(SUB)
    want = one.decode(iq, one.alloc_result(60))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port_no, iq, q, backend, 16)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert [g[1] for g in got] == [want_mode] * world
    assert sum(g[2] for g in got) == want.nframes and all(g[2] > 0 for g in got)
    assert np.array_equal(np.concatenate([g[3] for g in got]), want.soft)
    assert np.array_equal(np.concatenate([g[4] for g in got]), want.fic_bits)
    for i in range(len(SUB)):
        assert np.array_equal(np.concatenate([g[5][i] for g in got]), want.msc[i])
    assert sum((g[6] for g in got), []) == [(i.pos, i.fine, i.phase0) for i in want.info]
    one.close()


def test_buffer_cache_release(port):
    """buffers of closed handles are reused by the next handle and can be handed back to the driver at any time"""
    import torch
    pkg = engine_pkg()
    lib = pkg.load_library()
    mod = dabmod.Modulator(port, 1, [(0, 128, 1, 0o103)], 8)
    iq = mod.generate(12, cfo_hz=300.0, snr_db=25.0, lead=2500, tail=6000)["iq"]
    outs = []
    for k in range(3):
        e = pkg.DabGpu(mode=1); e.set_subchannels([(0, 96, 128, 1, 0o103)])
        r = e.decode(iq, e.alloc_result(16))
        outs.append((r.nframes, r.soft.copy(), r.fic_bits.copy(), r.msc[0].copy()))
        e.close()
        if k == 1:
            free0 = torch.cuda.mem_get_info()[0]
            assert lib.dabgpu_release_cached_memory() == 0
            assert torch.cuda.mem_get_info()[0] > free0            # the cache held device memory and gave it back
    for o in outs[1:]:
        assert o[0] == outs[0][0] and all(np.array_equal(a, b) for a, b in zip(o[1:], outs[0][1:]))
