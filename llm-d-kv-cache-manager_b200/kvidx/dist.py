"""One-process-per-GPU plumbing for the replica mode (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The Score() path shards by prompt: every rank holds a full replica of the index (10 M blocks = 1 GB of slots),
scores its own contiguous slice of the batch, and no data-path collective is needed.  What has to be replicated is
the write path: every rank must apply the same KV events, in the same per-pod order.  `broadcast_event_batch` does
that from the ingesting rank; `shard_range` splits a batch; `max_over_ranks` is the timing reduction bench.py uses.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

EVENT_DTYPE = np.dtype([("op", "u1"), ("has_parent", "u1"), ("podtier", "<u2"), ("model", "<u4"),
                        ("parent_hash", "<u8"), ("hash_off", "<u8"), ("tok_off", "<u8"),
                        ("n_hashes", "<u4"), ("n_tokens", "<u4")])


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced split of n prompts: ranks [0, n % world) get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed milliseconds -> max over ranks (the number a multi-GPU step is judged by)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_event_batch(events, hashes, tokens, src: int = 0, device=None):
    """Replicate one decoded event batch (kvidx_event_t array + flat hash / token arrays) from `src` to all ranks.
    Returns numpy arrays on every rank.  Sizes go first, then the three payloads as byte tensors."""
    rank = dist.get_rank()
    if rank == src:
        ev = np.ascontiguousarray(events, EVENT_DTYPE)
        hs = np.ascontiguousarray(hashes, np.uint64)
        tk = np.ascontiguousarray(tokens, np.uint32)
        sizes = torch.tensor([len(ev), len(hs), len(tk)], dtype=torch.int64, device=device)
    else:
        sizes = torch.zeros(3, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src)
    n_ev, n_hs, n_tk = (int(v) for v in sizes.tolist())
    out = []
    for arr, n, dt in ((events, n_ev, EVENT_DTYPE), (hashes, n_hs, np.dtype(np.uint64)), (tokens, n_tk, np.dtype(np.uint32))):
        nbytes = n * dt.itemsize
        if rank == src:
            buf = torch.from_numpy(np.ascontiguousarray(arr, dt).view(np.uint8).reshape(-1).copy())
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8)
        if device is not None:
            buf = buf.to(device)
        if nbytes:
            dist.broadcast(buf, src)
        out.append(buf.cpu().numpy().view(dt) if nbytes else np.zeros(0, dt))
    return tuple(out)


def gather_scores(local_scores: np.ndarray, n_total: int, device=None):
    """Collect the per-rank score slices (shard_range order) on every rank (all_gather of padded rows)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    P = local_scores.shape[1]
    cap = -(-n_total // world)
    pad = np.full((cap, P), -1.0)
    pad[: len(local_scores)] = local_scores
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.empty((n_total, P))
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi] = parts[r].cpu().numpy()[: hi - lo]
    return out


def connect_shards(ix, device=None):
    """Hash-range sharded mode: exchange the CUDA-IPC handle blobs of every rank's shard (one all_gather of 192 bytes
    per rank -- the only collective this mode needs) and map them.  After this call every kernel of `ix` can probe /
    lock slots of any shard through NVLink peer memory."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.frombuffer(bytearray(ix.shard_export()), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    for r in range(world):
        if r != rank:
            ix.shard_import(r, bytes(parts[r].cpu().numpy().tobytes()))
    dist.barrier()


def compact_shards(ix):
    """Drop the tombstones of every rank's shard (kvidx_shard_compact).  Collective: a barrier either side keeps every
    rank off the index while the shards are rewritten in place (peers keep their mappings)."""
    dist.barrier()
    ix.shard_compact()
    dist.barrier()


def events_for_rank(events, rank: int, world: int):
    """Ingest rule of the sharded mode: all events of one pod go through one rank (pod id modulo world), which keeps
    the reference's per-pod ordering (kvevents/pool.go:129-144) without any cross-GPU coordination."""
    pods = events["podtier"] >> 4
    return events[(pods % world) == rank]


def _a2a(out, inp, out_split=None, in_split=None):
    """all_to_all_single on whatever backend the group has (gloo has no CUDA all-to-all: stage through the host there)."""
    if dist.get_backend() == "gloo" and inp.is_cuda:
        o = torch.empty(out.shape, dtype=out.dtype)
        dist.all_to_all_single(o, inp.cpu(), out_split, in_split)
        out.copy_(o)
    else:
        dist.all_to_all_single(out, inp, out_split, in_split)


def score_alltoall(ix, d_tok, d_off, n, d_scores, block_size=16, model0=0, d_has=None):
    """The ROUTED form of a sharded Score(): every key of every prompt is hashed at the origin, sent to the rank
    that owns its hash range (NCCL all-to-all), looked up there, and its 32-byte slot image sent back (second all-to-all); the
    origin then walks and scores.  No early exit, no prefix sharing: all n_blocks keys travel.  Tensors are torch CUDA
    tensors; `ix` is a sharded handle whose stream is the CURRENT torch stream, which must be a real (non-default) stream
    -- ix.set_stream(stream.cuda_stream); handle 0 would mean the library's own stream.  Returns a dict of volumes."""
    world = dist.get_world_size()
    dev = d_tok.device
    lens = torch.div(d_off[1:n + 1] - d_off[:n], block_size, rounding_mode="floor")
    koff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=koff[1:])
    nk = int(koff[-1].item())
    keys = torch.empty(max(nk, 1), dtype=torch.int64, device=dev)
    ix.hash_keys_dev(d_tok.data_ptr(), d_off.data_ptr(), n, koff.data_ptr(), keys.data_ptr())
    keys = keys[:nk]
    owners = torch.empty(max(nk, 1), dtype=torch.uint8, device=dev)
    ix.key_owners_dev(keys.data_ptr(), nk, owners.data_ptr(), model0)
    owners = owners[:nk]
    order = torch.argsort(owners, stable=True)
    counts = torch.bincount(owners, minlength=world).to(torch.int64)
    send = keys[order].contiguous()
    recv_counts = torch.empty_like(counts)
    _a2a(recv_counts, counts)
    in_split, out_split = counts.tolist(), recv_counts.tolist()
    recv = torch.empty(sum(out_split), dtype=torch.int64, device=dev)
    _a2a(recv, send, out_split, in_split)
    slots = torch.empty((max(len(recv), 1), 4), dtype=torch.int64, device=dev)
    ix.probe_slots_dev(recv.data_ptr(), len(recv), slots.data_ptr(), model0)
    slots = slots[: len(recv)]
    back = torch.empty((nk, 4), dtype=torch.int64, device=dev)
    _a2a(back, slots.contiguous(), in_split, out_split)
    slots_pm = torch.empty_like(back)
    slots_pm[order] = back
    ix.score_slots_dev(slots_pm.data_ptr(), koff.data_ptr(), n, d_scores.data_ptr(), d_has_keys=(d_has.data_ptr() if d_has is not None else 0))
    return {"keys": nk, "bytes_out": 8 * (nk - in_split[dist.get_rank()]), "bytes_back": 32 * (nk - in_split[dist.get_rank()])}
