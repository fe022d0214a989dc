"""Multi-rank host logic on CPU: two `gloo` ranks.

The GPU path shards a frame by row bands or by samples and all-reduces the MLP gradients of
each rank's own training records, with the loss normalised by the GLOBAL record count
(hm_mlp_forward_backward's n_total_records), so that the summed gradient equals the
single-GPU one.  These tests run the same contract on the host: the partition arithmetic
comes from the C ABI (hm_band_partition, no device needed), the gradients from the numpy
oracle of the network (test infrastructure), the exchange from torch.distributed/gloo.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

WORLD = 2


def _records(n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 12), np.float32)
    x[:, 0:3] = rng.uniform(-0.5, 0.5, (n, 3))
    for c in (3, 6):
        v = rng.normal(size=(n, 3))
        x[:, c:c + 3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    y = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    return x, y


def _worker(rank, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from hairmsnn_b200 import api
        import mlp_oracle as mo

        # ---- row bands tile the frame; training records are owned exactly once ----------
        W, H, R = 1024, 1024, 16384
        mine = torch.tensor(api.band_partition(W, H, R, rank, WORLD), dtype=torch.int64)
        every = [torch.zeros(5, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(every, mine)
        rows = [(int(t[0]), int(t[1])) for t in every]
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(rows[i][1] == rows[i + 1][0] for i in range(WORLD - 1))
        slots = [(int(t[2]), int(t[3])) for t in every]
        assert slots[0][0] == 0 and sum(s[1] for s in slots) == R
        assert all(slots[i][0] + slots[i][1] == slots[i + 1][0] for i in range(WORLD - 1))
        assert all(int(t[4]) % 128 == 0 and int(t[4]) <= int(t[3]) for t in every)

        # ---- spp sharding: the ranks' sample ids interleave into 0..K*world-1 ------------
        K = 5
        ids = torch.tensor([api.sample_schedule(rank, WORLD, k) for k in range(K)], dtype=torch.int64)
        allids = [torch.zeros(K, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(allids, ids)
        assert sorted(torch.cat(allids).tolist()) == list(range(K * WORLD))

        # ---- data-parallel training step: sum of per-rank gradients == full-batch gradient ----
        cfg = mo.Config(12)
        params = mo.initial_params(cfg)
        n = 256
        x, y = _records(n, seed=7)                 # same records on both ranks (same seed)
        lo, hi = rank * n // WORLD, (rank + 1) * n // WORLD
        _, g_local = mo.backward(cfg, params, x[lo:hi], y[lo:hi], n_total_records=n, half=False)
        g_local = np.asarray(g_local, np.float64)
        t = torch.from_numpy(g_local.copy())
        dist.all_reduce(t)                         # what NCCL does on the GPU path
        _, g_full = mo.backward(cfg, params, x, y, n_total_records=n, half=False)
        g_full = np.asarray(g_full, np.float64)
        err = np.abs(t.numpy() - g_full).max() / max(np.abs(g_full).max(), 1e-30)
        assert err < 1e-5, f"all-reduced gradient differs from the full-batch gradient: {err}"

        # identical Adam step on identical summed gradients keeps the replicas bit-identical
        opt = mo.Adam(cfg, params.copy())
        p_new = np.asarray(opt.step(t.numpy().astype(np.float32), half=False), np.float32)
        digest = torch.tensor([float(np.float64(p_new.astype(np.float64).sum())), float(np.abs(p_new).max())], dtype=torch.float64)
        both = [torch.zeros(2, dtype=torch.float64) for _ in range(WORLD)]
        dist.all_gather(both, digest)
        assert torch.equal(both[0], both[1])

        # ---- acceleration structure: rank 0 builds and publishes, the others restore (bench.py's start-up) ----
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import small_scene_kwargs
        kw = small_scene_kwargs(strands=150, segs=10)
        cache = os.path.join(tmp, "bvh_cache")
        os.makedirs(cache, exist_ok=True)
        if rank == 0:
            sc = api.Scene.from_arrays(**kw)
            sc.save_bvh_cache(cache)
        dist.barrier()
        if rank != 0:
            os.environ["HM_BVH_CACHE"] = cache
            sc = api.Scene.from_arrays(**kw)
            del os.environ["HM_BVH_CACHE"]
        i = sc.info()
        mine = torch.tensor([i.num_wide_nodes, i.num_wide_leaf_refs, i.wide_depth, i.num_bvh_nodes], dtype=torch.int64)
        allinfo = [torch.zeros(4, dtype=torch.int64) for _ in range(WORLD)]
        dist.all_gather(allinfo, mine)
        assert torch.equal(allinfo[0][:3], allinfo[1][:3]) and int(allinfo[0][0]) > 0
        assert int(allinfo[0][3]) > 0 and int(allinfo[1][3]) == 0      # only the builder holds the binary tree
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_and_gradient_exchange(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(WORLD))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("shape", [(1024, 1024), (512, 256), (640, 360), (4096, 4096)])
def test_band_partition_properties(world, shape):
    from hairmsnn_b200 import api
    W, H = shape
    R = 16384
    every_nth = W * H // R
    parts = [api.band_partition(W, H, R, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == H
    owned = set()
    for r, (row0, row1, s0, ns, tn) in enumerate(parts):
        assert row0 <= row1 and (r == 0 or row0 == parts[r - 1][1])
        for s in (s0, s0 + ns - 1):
            if ns:   # every owned group's pixels lie inside the band
                assert row0 * W <= s * every_nth and (s + 1) * every_nth <= max(row1 * W, (s + 1) * every_nth if r == world - 1 else row1 * W)
        rng = set(range(s0, s0 + ns))
        assert not (rng & owned)
        owned |= rng
        assert tn % 128 == 0 and 0 <= tn <= ns
    assert len(owned) <= R
    if (W % every_nth == 0) and all(p[0] * W % every_nth == 0 for p in parts):
        assert len(owned) == R          # aligned bands lose no record


def test_band_partition_rejects_bad_arguments():
    from hairmsnn_b200 import api
    for args in [(0, 4, 16, 0, 1), (4, 4, 16, 1, 1), (4, 4, 16, -1, 2), (4, 4, -1, 0, 1), (4, 4, 16, 0, 0)]:
        with pytest.raises(api.HairMSNNError):
            api.band_partition(*args)
