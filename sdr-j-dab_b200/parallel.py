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
