"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the two
collectives the path has -- nothing on the data path itself.

The pair list of /root/reference/apps/sfm/main.cpp:84-92 is a set of independent units, so it shards with no
exchange step: every rank holds the whole descriptor arena (262 MB for 2,000 x 4k ORB -- trivial next to 180 GB of
HBM3e), matches its share of the pair list (whole blocks of the image x image grid, ``shard_owner``) and keeps its results. The collectives are
  * one broadcast of the arena bytes from the rank that owns the host descriptors, and
  * an optional gather of the per-pair results to one rank.
torch is used only to move bytes; the arena memory belongs to libeacham_gpu.so on every rank.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


SHARD_BLOCK = 16      # images per side of a block of the image x image grid; the library hands pairs out in the same blocks


def shard_owner(pairs: np.ndarray, world: int) -> np.ndarray:
    """Rank that owns each pair. Whole 16 x 16 blocks of the image x image grid go to one rank (largest block first, each to the rank with the
    least pairs so far): every rank then works through complete blocks whose ~32 images stay resident in its L2. (A plain ``pairs[rank::world]``
    leaves each rank 1/world of every block, so its ~148 pairs in flight span ``world`` times as many images: at 8 GPUs that cost
    23 % of the per-GPU rate.) Sets with fewer than 8 x world occupied blocks use 8 x 8, 4 x 4, ... blocks instead. Order inside a
    rank is the input order."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if p.shape[0] == 0:
        return np.zeros(0, np.int32)
    side = SHARD_BLOCK
    while True:                                       # small sets: finer blocks until there are enough of them to balance
        nb = int(p.max()) // side + 1
        block = (p[:, 0] // side) * nb + p[:, 1] // side
        uniq, dense, counts = np.unique(block, return_inverse=True, return_counts=True)
        if side == 1 or uniq.shape[0] >= 8 * world:
            break
        side //= 2
    # blocks differ in size (diagonal blocks, window lists): largest first, each to the rank with the least pairs so far
    load = [0] * world
    owner_of_block = np.zeros(uniq.shape[0], np.int32)
    for b in np.argsort(-counts, kind="stable").tolist():
        r = min(range(world), key=load.__getitem__)
        owner_of_block[b] = r
        load[r] += int(counts[b])
    return owner_of_block[dense]


def shard_pairs(pairs: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The pairs rank ``rank`` matches (see shard_owner), in input order."""
    arr = np.asarray(pairs, dtype=np.uint32).reshape(-1, 2)
    return np.ascontiguousarray(arr[shard_owner(arr, world) == rank])


def shard_sizes(pairs: np.ndarray, world: int) -> List[int]:
    return np.bincount(shard_owner(pairs, world), minlength=world).astype(int).tolist()


def unshard(per_rank: Sequence[np.ndarray], pairs: np.ndarray) -> np.ndarray:
    """Inverse of shard_pairs for per-pair arrays: per_rank[r][k] belongs to the k-th pair owned by rank r."""
    owner = shard_owner(pairs, len(per_rank))
    out = np.empty((owner.shape[0],) + per_rank[0].shape[1:], dtype=per_rank[0].dtype)
    for r in range(len(per_rank)):
        out[owner == r] = per_rank[r]
    return out


class _CudaBytes:
    """A raw device range exposed through __cuda_array_interface__ so torch can address library-owned memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def arena_tensor(matcher):
    """uint8 CUDA tensor aliasing the matcher's committed device arena (no copy)."""
    import torch
    ptr, nbytes = matcher.arena()
    return torch.as_tensor(_CudaBytes(ptr, nbytes), device=f"cuda:{matcher.device}")


def upload_and_broadcast(matcher, descriptors: Optional[Sequence[np.ndarray]], src: int = 0, group=None) -> int:
    """Every rank lays out the same arena; ``src`` uploads the bytes (one H2D) and broadcasts them once over NCCL.
    ``descriptors`` is needed on ``src`` only. Returns the arena size in bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if rank == src:
        shapes = [(0 if d.dtype == np.uint8 else 1, int(d.shape[0])) for d in descriptors]
    else:
        shapes = None
    box = [shapes]
    dist.broadcast_object_list(box, src=src, group=group)
    shapes = box[0]
    matcher.Clear()
    for i, (kind, rows) in enumerate(shapes):
        if rank == src:
            matcher.SetDescriptors(i, descriptors[i])
        else:
            matcher.Reserve(i, kind, rows)
    matcher.Commit()
    t = arena_tensor(matcher)
    dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(matcher.device)
    return int(t.numel())


def gather_results(res: np.ndarray, matches: np.ndarray, pairs: np.ndarray, dst: int = 0, group=None,
                   device: Optional[str] = None) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """Gathers each rank's (results, matches) record arrays onto ``dst`` and restores the original pair order.
    Offsets are rebased into the concatenated match buffer. Works on gloo (CPU tensors) and nccl (CUDA tensors).
    Returns (results[n_pairs_total], matches[total]) on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    sizes = torch.tensor([res.shape[0], matches.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    max_r = max(s[0] for s in all_sizes); max_m = max(s[1] for s in all_sizes)

    def padded(arr: np.ndarray, itemsize: int, n_max: int):
        raw = np.zeros(max(n_max, 1) * itemsize, np.uint8)
        b = arr.view(np.uint8).reshape(-1)
        raw[:b.shape[0]] = b
        return torch.from_numpy(raw).to(dev)

    tr = padded(res, np.dtype(L.RESULT_DTYPE).itemsize, max_r)
    tm = padded(matches, np.dtype(L.MATCH_DTYPE).itemsize, max_m)
    if rank == dst:
        gr = [torch.empty_like(tr) for _ in range(world)]
        gm = [torch.empty_like(tm) for _ in range(world)]
    else:
        gr = gm = None
    dist.gather(tr, gr, dst=dst, group=group)
    dist.gather(tm, gm, dst=dst, group=group)
    if rank != dst:
        return None
    per_res, per_m, base = [], [], 0
    for r in range(world):
        nr, nm = all_sizes[r]
        rr = gr[r].cpu().numpy()[: nr * np.dtype(L.RESULT_DTYPE).itemsize].view(L.RESULT_DTYPE).copy()
        mm = gm[r].cpu().numpy()[: nm * np.dtype(L.MATCH_DTYPE).itemsize].view(L.MATCH_DTYPE).copy()
        rr["offset"] += base
        base += nm
        per_res.append(rr); per_m.append(mm)
    return unshard(per_res, pairs), np.concatenate(per_m) if per_m else np.zeros(0, L.MATCH_DTYPE)


_PINNED = {}


def _pinned(tag: str, nbytes: int):
    """Reusable pinned host buffer (cudaHostAlloc of ~1 GB costs more than the copy it serves)."""
    import torch
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=True)
        _PINNED[tag] = buf
    return buf


def gather_results_device(matcher, pairs: np.ndarray, dst: int = 0, group=None) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """Same as gather_results, but straight from the device-resident outputs of the last MatchPairsDevice call:
    shards travel GPU -> GPU over NVLink (NCCL gather) and only ``dst`` does one D2H copy."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = f"cuda:{matcher.device}"
    rptr, mptr, n_res, n_m = matcher.device_results()
    rs, ms = np.dtype(L.RESULT_DTYPE).itemsize, np.dtype(L.MATCH_DTYPE).itemsize
    sizes = torch.tensor([n_res, n_m], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    max_r = max(1, max(x[0] for x in all_sizes)); max_m = max(1, max(x[1] for x in all_sizes))
    tr = torch.zeros(max_r * rs, dtype=torch.uint8, device=dev)
    tm = torch.empty(max_m * ms, dtype=torch.uint8, device=dev)
    if n_res:
        tr[: n_res * rs] = torch.as_tensor(_CudaBytes(rptr, n_res * rs), device=dev)
    if n_m:
        tm[: n_m * ms] = torch.as_tensor(_CudaBytes(mptr, n_m * ms), device=dev)
    gr = [torch.empty_like(tr) for _ in range(world)] if rank == dst else None
    gm = [torch.empty_like(tm) for _ in range(world)] if rank == dst else None
    dist.gather(tr, gr, dst=dst, group=group)
    dist.gather(tm, gm, dst=dst, group=group)
    if rank != dst:
        return None
    # one D2H per array, into pinned host memory, each rank's valid prefix copied straight to its final position
    tot_r = sum(x[0] for x in all_sizes) * rs; tot_m = sum(x[1] for x in all_sizes) * ms
    host_r = _pinned("results", max(tot_r, 1))
    host_m = _pinned("matches", max(tot_m, 1))
    pr = pm = 0
    for r in range(world):
        nr, nm = all_sizes[r][0] * rs, all_sizes[r][1] * ms
        host_r[pr: pr + nr].copy_(gr[r][:nr], non_blocking=True)
        host_m[pm: pm + nm].copy_(gm[r][:nm], non_blocking=True)
        pr += nr; pm += nm
    torch.cuda.synchronize(matcher.device)
    res_all = host_r[:tot_r].numpy().view(L.RESULT_DTYPE).copy()
    m_all = host_m[:tot_m].numpy().view(L.MATCH_DTYPE)          # view of the reusable pinned buffer: valid until the next gather
    per_res, pos, base = [], 0, 0
    for r in range(world):
        nr, nm = all_sizes[r]
        rr = res_all[pos: pos + nr]
        rr["offset"] += base
        per_res.append(rr)
        pos += nr; base += nm
    return unshard(per_res, pairs), m_all
