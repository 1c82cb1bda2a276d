"""Multi-GPU test (needs >= 2 CUDA devices; skipped on a one-GPU box): the observation all-gather by NVLink peer copies into
symmetric memory delivers, on every rank, exactly what a single process renders for all envs (shard invariance, bit for bit)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, mode="copy"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tacex_b200 import synth
    from tacex_b200.calib import TaximTables
    from tacex_b200.engine import TactileEngine
    from tacex_b200.shard import PeerObsGather, env_shard

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = TaximTables.load(os.path.join(root, "tests", "golden", "gsmini_tables_320x240.npz"))
    n = 8
    hm = synth.height_map_mm(synth.config1(n, seed=5)["depth_m"])
    a, b = env_shard(n, rank, world)
    eng = TactileEngine(t, max_envs=n, device=f"cuda:{rank}")
    rects = mode.startswith("rects") or mode == "fused"
    g = PeerObsGather((b - a, 240, 320, 3), torch.float32, torch.device("cuda", rank), n_slots=2, with_rects=rects,
                      multicast=mode in ("rects", "fused"))
    if mode == "fused" and not g.fused_available():
        q.put("skip")
        dist.barrier()
        dist.destroy_process_group()
        return
    if rects and rank == 0:
        print(f"[{mode}] NVSwitch multicast stores: {'yes' if g.mc_rgb[0] else 'no (one store per peer)'}", flush=True)
    ok = True
    if rects:
        for buf in g.bufs:
            buf.fill_(float("nan"))  # the first fill of a slot has to write every pixel outside the rectangles
        torch.cuda.synchronize()
        dist.barrier()
    for step in range(5):  # both slots, slot reuse; in rectangle mode the contacts move between the steps
        slot = step & 1
        if rects:
            hm_s = torch.roll(hm, shifts=(7 * step, -11 * step), dims=(1, 2)) if step else hm
            if step == 3:
                hm_s = hm_s.clone()
                hm_s[::2] = hm_s.max()  # every other env loses contact: its old rectangle must be restored completely
            if mode == "fused":  # the render kernel's epilogue stores the rectangles into every GPU's buffer (multimem.st)
                g.begin_fused(eng, slot)
                eng.render(hm_s[a:b].cuda(rank).contiguous(), None, out=g.local_block(slot))
                eng.set_multicast_output(0, 0)
                eng.set_rect_output(None)
                full = g.finish_fused(eng, slot, torch.cuda.current_stream())
            else:
                eng.set_rect_output(g.local_rects(slot))
                eng.render(hm_s[a:b].cuda(rank).contiguous(), None, out=g.local_block(slot))
                eng.set_rect_output(None)
                full = g.gather_rects(eng, slot, torch.cuda.current_stream())
        else:
            hm_s = hm
            eng.render(hm_s[a:b].cuda(rank).contiguous(), None, out=g.local_block(slot))
            full = g.gather(g.local_block(slot), slot)
        ref = eng.render(hm_s.cuda(rank).contiguous(), None)
        torch.cuda.synchronize()
        ok = ok and torch.equal(full, ref)
    q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["copy", "rects", "rects-unicast", "fused"])
def test_peer_copy_all_gather_matches_single_gpu(mode):
    """copy: whole frames through the copy engines; rects: only the non-flat rectangle of every half frame crosses the link
    (tx_obs_push / tx_obs_fill; NVSwitch multicast stores when the symmetric allocation has a multicast mapping, else -- and
    in the rects-unicast case -- one store per peer), the rest is completed locally from the flat image; fused: the render kernel
    itself stores the rectangles through the multicast mapping (compute + collective in one kernel). All bit-identical to one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + ["copy", "rects", "rects-unicast", "fused"].index(mode)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    oks = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    if "skip" in oks:
        pytest.skip("the symmetric allocation has no NVSwitch multicast mapping on this box")
    assert all(oks) and all(p.exitcode == 0 for p in procs)
