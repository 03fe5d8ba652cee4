"""CPU, world_size 2, gloo: the host-side multi-rank logic (shard -> per-rank results -> gather -> original order).
The per-pair compute is stood in by the oracle (test infrastructure) so that no GPU is needed."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_images, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from eacham_b200 import synth, distributed as D, _lib as L
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs = synth.orb_image_set(n_images, 300, seed=21, pool=330)
    pairs = synth.exhaustive_pairs(n_images)
    mine = D.shard_pairs(pairs, rank, world)
    assert len(mine) == D.shard_sizes(pairs, world)[rank]
    res = np.zeros(len(mine), dtype=L.RESULT_DTYPE); chunks = []; off = 0
    for k, (i, j) in enumerate(mine.tolist()):
        r = O.c_match_pair(imgs[i], imgs[j])
        m = r["matches"] if r["connected"] else np.zeros((0, 2), np.uint32)
        res[k] = (r["n12"], r["n21"], r["n_mutual"], (1 if r["gated"] else 0) | (2 if r["connected"] else 0), off, len(m))
        rec = np.zeros(len(m), dtype=L.MATCH_DTYPE); rec["query"] = m[:, 0]; rec["train"] = m[:, 1]
        chunks.append(rec); off += len(m)
    buf = np.concatenate(chunks) if chunks else np.zeros(0, L.MATCH_DTYPE)
    got = D.gather_results(res, buf, pairs, dst=0)
    if rank == 0:
        full_res, full_buf = got
        ok = True
        for k, (i, j) in enumerate(pairs.tolist()):
            r = O.c_match_pair(imgs[i], imgs[j])
            fr = full_res[k]
            mm = full_buf[int(fr["offset"]): int(fr["offset"]) + int(fr["count"])]
            want = r["matches"] if r["connected"] else np.zeros((0, 2), np.uint32)
            ok &= (int(fr["n12"]), int(fr["n21"]), int(fr["n_mutual"])) == (r["n12"], r["n21"], r["n_mutual"])
            ok &= np.array_equal(np.stack([mm["query"], mm["train"]], 1), want)
        np.save(out_path, np.array([int(ok), len(pairs), int(full_res["count"].sum())]))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_roundtrip_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, _free_port(), 7, out), nprocs=2, join=True)
    ok, n_pairs, n_matches = np.load(out).tolist()
    assert ok == 1 and n_pairs == 21 and n_matches > 0


def test_shard_unshard_identity():
    from eacham_b200 import distributed as D
    from eacham_b200 import synth
    for pairs in (np.arange(2 * 37, dtype=np.uint32).reshape(37, 2), synth.exhaustive_pairs(70), synth.window_pairs(200, 20)):
        for world in (1, 2, 3, 4, 8):
            parts = [D.shard_pairs(pairs, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == len(pairs) and [len(p) for p in parts] == D.shard_sizes(pairs, world)
            assert np.array_equal(D.unshard(parts, pairs), pairs)
            sizes = D.shard_sizes(pairs, world)
            assert max(sizes) - min(sizes) <= max(4, len(pairs) // 3)      # small sets fall back to finer blocks: nobody is left idle
    big = synth.exhaustive_pairs(2000)                        # BASELINE config 5: shards within 0.2 % of each other ...
    sizes = D.shard_sizes(big, 8)
    assert max(sizes) - min(sizes) < 0.002 * len(big)
    owner = D.shard_owner(big, 8)                             # ... and a 16 x 16 block of the image grid is never split between ranks
    blocks = (big[:, 0].astype(np.int64) // 16) * 1000 + big[:, 1] // 16
    order = np.argsort(blocks, kind="stable")
    starts = np.flatnonzero(np.diff(blocks[order], prepend=-1))
    assert np.array_equal(np.maximum.reduceat(owner[order], starts), np.minimum.reduceat(owner[order], starts))
