"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

Two ways the path shards:
  * independent ensembles / recordings: every rank decodes its own streams, no communication (bench.py --gpus N);
  * ONE long recording split into contiguous sample ranges: rank r needs, from rank r-1, the sync/AFC state, the
    unconsumed sample tail and the last 15 CIFs of soft bits (the time de-interleaver halo).  That state travels
    as one blob (`DabGpu.export_state`) with a neighbour send/recv -- NCCL over NVLink between GPUs, gloo in the
    CPU test.  No all-reduce / all-gather anywhere.

The functions are engine-agnostic (duck typed: `decode`, `export_state`, `import_state`) so the hand-over logic
can be tested on CPU with a stand-in decoder.
"""
import numpy as np


def shard_ranges(nsamples, world, frame_len):
    """Contiguous sample ranges, one per rank, cut at multiples of the nominal frame length (the real frame
    boundaries are found by the engine; whatever a shard cannot finish travels on in the state's sample tail)."""
    frames = nsamples // frame_len
    base, extra = divmod(frames, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        last = first + n * frame_len if r < world - 1 else nsamples
        out.append((first, last))
        first = last
    return out


def _send_blob(dist, blob, dst, device):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(blob)).to(device)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=device)
    dist.send(n, dst)
    dist.send(t, dst)


def _recv_blob(dist, src, device):
    import torch
    n = torch.zeros(1, dtype=torch.int64, device=device)
    dist.recv(n, src)
    t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.recv(t, src)
    return t.cpu().numpy()


def decode_split(engine, iq_shard, out, rank, world, dist=None, device="cpu"):
    """Decode this rank's shard of one recording.  Ranks run as a chain: receive the stream state of the
    previous shard, decode, pass the state on.  The concatenation of all ranks' outputs equals the one-GPU
    decode of the whole recording bit for bit."""
    if world > 1 and rank > 0:
        engine.import_state(_recv_blob(dist, rank - 1, device))
    res = engine.decode(iq_shard, out)
    if world > 1 and rank < world - 1:
        _send_blob(dist, engine.export_state(), rank + 1, device)
    return res


# ---------------------------------------------------------------------------------------------------------------
# One recording, shards decoded IN PARALLEL (BASELINE.json configs[4]).
#
# The chain above is exact but serial.  A locked receiver is predictable, though: every frame is T_F samples long and
# the correctors stand still, so the tracking state at any later frame boundary follows in closed form from one
# locked state (`dabgpu_host_state_predict`) -- the same assumption the engine's own frame-parallel pass makes inside
# one GPU.  So: rank 0 decodes a short lead-in (acquisition, AFC convergence) and broadcasts the locked state; every
# rank then decodes its frame range starting from the PREDICTED state, `overlap` frames early so that its time
# de-interleaver (15 CIFs of memory, dab-concurrent.cpp:41-43) is full of real data by the time its own range starts
# (a fresh engine swallows exactly those 16 CIFs as warm-up, dab-concurrent.cpp:172-175).  No soft-bit halo travels.
# Afterwards each rank checks its assumption against the TRUE final state of its left neighbour (one 6-word message);
# if every boundary agrees the concatenated output equals the one-GPU decode bit for bit, otherwise (receiver not
# locked, corrector moved) the ranks fall back to the exact serial chain.
# ---------------------------------------------------------------------------------------------------------------
def frame_shards(nframes, world, first_min=0):
    """contiguous frame ranges [a, b), one per rank; rank 0 gets at least first_min frames"""
    base, extra = divmod(nframes, world)
    cuts, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        cuts.append([a, b])
        a = b
    if world > 1 and cuts[0][1] < first_min:
        cuts[0][1] = min(first_min, nframes)
        for r in range(1, world):
            cuts[r][0] = max(cuts[r][0], cuts[r - 1][1])
            cuts[r][1] = max(cuts[r][1], cuts[r][0])
    return [tuple(c) for c in cuts]


def _state_words(s):
    return [int(s.abs_pos), int(s.coarse), int(s.fine), int(s.localPhase), int(s.f2Correction), int(s.synced)]


def decode_sharded(engine, iq, alloc, rank, world, dist=None, device="cpu", lead_frames=24, timing=None):
    """Decode ONE recording `iq` (interleaved u8 I,Q; every rank can address all of it but touches only its own sample
    range) on `world` ranks in parallel.  `engine` is a fresh handle of this rank, `alloc(max_frames)` returns a result
    buffer.  Returns (result, first_frame, mode): `result` holds this rank's frames (overlap already removed), and
    mode is "parallel" or "chain" (the fallback).  Rank 0's result starts with the lead-in frames.
    timing (optional dict): timing["decoded"] = time.perf_counter() when this rank's frames are decoded, delivered and the
    boundaries verified -- before the host-side assembly of the returned arrays (trimming the overlap / prepending the lead-in)."""
    import os, sys, time, torch
    T_F, cpf, need = engine.frame_len, engine.cifs_per_frame, engine.frame_need
    overlap = -(-16 // cpf)                                   # frames that hold the 16 warm-up CIFs
    nsamp = len(iq) // 2
    trace = os.environ.get("DAB_SHARD_TRACE")
    t_last = [time.perf_counter()]

    def mark(what):
        if trace:
            now = time.perf_counter()
            sys.stderr.write("[shard %d] %-28s %8.2f ms  (launches so far %s)\n" % (rank, what, (now - t_last[0]) * 1e3, engine.launch_count() if hasattr(engine, "launch_count") else "?"))
            t_last[0] = now

    def bcast(words):
        t = torch.tensor(words, dtype=torch.int64, device=device)
        if world > 1:
            dist.broadcast(t, 0)
        return [int(x) for x in t.tolist()]

    def all_min(v):
        t = torch.tensor([v], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    # ---- 1. lead-in on rank 0 -> locked state for everybody
    lead_samples = min(nsamp, lead_frames * T_F)
    lead = None
    words = [0] * 10
    if rank == 0:
        lead = engine.decode(iq[:2 * lead_samples], alloc(lead_frames + 2))
        s0 = engine.state_get()
        words = _state_words(s0) + [int(s0.previous_1), int(s0.previous_2), int(s0.frames), int(s0.cifs)]
    mark("lead-in")
    words = bcast(words)
    mark("state broadcast")
    s0 = engine.make_state(abs_pos=words[0], coarse=words[1], fine=words[2], localPhase=words[3], f2Correction=words[4],
                           synced=words[5], previous_1=words[6], previous_2=words[7], frames=words[8], cifs=words[9])
    locked = s0.synced == 1 and s0.f2Correction == 0
    total = (nsamp - s0.abs_pos - need) // T_F + 1 if nsamp - s0.abs_pos >= need else 0      # whole frames after the lead-in
    plan = frame_shards(max(total, 0), world, first_min=overlap)
    a, b = plan[rank]
    last = rank == world - 1
    end_abs = nsamp if last else min(nsamp, s0.abs_pos + b * T_F + need - T_F + 64)         # enough for frame b-1, not for frame b
    ok = 1 if locked and total >= world * overlap else 0
    res = None
    if all_min(ok):
        # ---- 2. every rank decodes its range from the predicted state
        if rank == 0:
            res = engine.decode(iq[2 * lead_samples:2 * end_abs], alloc(b - a + 2))
            first_state = None
        else:
            sp = engine.state_predict(s0, a - overlap)
            engine.state_set(sp)
            res = engine.decode(iq[2 * sp.abs_pos:2 * end_abs], alloc(b - a + overlap + 2))
            first_state = res.info[overlap] if res.nframes > overlap else None
        fin = _state_words(engine.state_get())
        mark("decode of my range")
        # ---- 3. verify every boundary: my assumed state at frame a == the left neighbour's true final state
        if world > 1:
            t = torch.tensor(fin, dtype=torch.int64, device=device)
            got = torch.zeros(6, dtype=torch.int64, device=device)
            reqs = []
            if rank < world - 1:
                reqs.append(dist.isend(t, rank + 1))
            if rank > 0:
                dist.recv(got, rank - 1)
            for q in reqs:
                q.wait()
            if rank > 0:
                g = [int(x) for x in got.tolist()]
                mine = None if first_state is None else [int(first_state.pos), int(first_state.coarse), int(first_state.fine), int(first_state.phase0), 0, 1]
                ok = 1 if (mine == g and fin[4] == 0 and fin[5] == 1 and res.nframes == b - a + overlap) or (a == b and first_state is None) else 0
            else:
                ok = 1 if (fin[4] == 0 and fin[5] == 1) else 0
        mark("boundary exchange")
        if all_min(ok):
            mark("agreement")
            if timing is not None:
                timing["decoded"] = time.perf_counter()
            if rank > 0:
                res = engine.drop_frames(res, overlap)
            elif lead is not None:
                res = engine.concat_results(lead, res)
            return res, (0 if rank == 0 else a), "parallel"
    # ---- fallback: the exact serial chain over sample ranges (fresh engines for ranks > 0; rank 0 keeps what is valid)
    ranges = shard_ranges(nsamp, world, T_F)
    lo, hi = ranges[rank]
    if rank == 0:
        if res is None:                                       # nothing decoded in parallel mode: continue after the lead-in
            out = engine.decode(iq[2 * lead_samples:2 * hi], alloc((hi - lead_samples) // T_F + 3)) if hi > lead_samples else None
            merged = lead if out is None else engine.concat_results(lead, out)
            fed = max(hi, lead_samples)
        else:
            merged, fed = engine.concat_results(lead, res), end_abs
        edge = torch.tensor([fed], dtype=torch.int64, device=device)
        if world > 1:
            dist.send(edge, 1)
            _send_blob(dist, engine.export_state(), 1, device)
        return merged, 0, "chain"
    edge = torch.zeros(1, dtype=torch.int64, device=device)
    dist.recv(edge, rank - 1)
    engine.import_state(_recv_blob(dist, rank - 1, device))
    lo = int(edge.item())
    hi = max(hi, lo)
    out = engine.decode(iq[2 * lo:2 * hi], alloc((hi - lo) // T_F + 3))
    if rank < world - 1:
        dist.send(torch.tensor([hi], dtype=torch.int64, device=device), rank + 1)
        _send_blob(dist, engine.export_state(), rank + 1, device)
    return out, None, "chain"
