"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the sharded path + NCCL broadcast + NVLink gather must
return exactly what one GPU returns for the whole pair list."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path, n_images=9):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import eacham_b200
    from eacham_b200 import synth, distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    imgs = synth.orb_image_set(n_images, 1500, seed=31, pool=2500) if rank == 0 else None
    pairs = synth.exhaustive_pairs(n_images)
    m = eacham_b200.FeatureMatcherGpu(0.8, device=rank)
    D.upload_and_broadcast(m, imgs, src=0)
    m.MatchPairsDevice(D.shard_pairs(pairs, rank, world))
    got = D.gather_results_device(m, pairs, dst=0)
    if rank == 0:
        res, buf = got
        m.Upload(imgs)
        ref_res, ref_buf = m.MatchPairsRaw(pairs)
        ok = True
        for k in range(len(pairs)):
            a, b = res[k], ref_res[k]
            ok &= (a["n12"], a["n21"], a["n_mutual"], a["flags"], a["count"]) == (b["n12"], b["n21"], b["n_mutual"], b["flags"], b["count"])
            ok &= np.array_equal(buf[int(a["offset"]): int(a["offset"] + a["count"])], ref_buf[int(b["offset"]): int(b["offset"] + b["count"])])
        np.save(out_path, np.array([int(ok), int(ref_res["count"].sum())]))
    m.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok, n = np.load(out).tolist()
    assert ok == 1 and n > 0


def test_sharded_with_an_empty_shard(tmp_path):
    """Two images = one pair: one of the two ranks has nothing to match; broadcast and gather must still work."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok1.npy")
    mp.spawn(_worker, args=(2, _free_port(), out, 2), nprocs=2, join=True)
    ok, n = np.load(out).tolist()
    assert ok == 1
