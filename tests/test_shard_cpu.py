"""CPU (gloo, world size 2) tests of the multi-GPU host logic: contiguous env shards + the single observation all-gather,
and shard invariance of the per-env path (here with the CPU checker standing in for the kernels)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_env_shard_partitions():
    from tacex_b200.shard import env_shard

    for n in (1, 7, 8, 4096, 8192):
        for w in (1, 2, 4, 8):
            blocks = [env_shard(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import canon
    from tacex_b200 import synth
    from tacex_b200.calib import TaximTables
    from tacex_b200.shard import all_gather_obs, env_shard

    H, W = 240, 320
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = TaximTables.load(os.path.join(root, "tests", "golden", "gsmini_tables_320x240.npz"))
    cn = canon.CanonTaxim(H, W, t.poly_grad.numpy(), t.background.numpy(), None, t.params.blur_taps((H, W)))
    hm = synth.height_map_mm(synth.config1(4, seed=5)["depth_m"])
    a, b = env_shard(4, rank, world)
    loc = hm[a:b].numpy()
    rgb = torch.from_numpy(cn.render(loc, cn.indentation_depth(loc), want=("rgb",))["rgb"])
    full = all_gather_obs(rgb)
    if rank == 0:
        ref = cn.render(hm.numpy(), cn.indentation_depth(hm.numpy()), want=("rgb",))["rgb"]
        q.put(bool(np.array_equal(full.numpy(), ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process_bitwise():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
