"""Env sharding across GPUs: contiguous env blocks per rank and the single observation all-gather.

The reference makes no collective call at all (Isaac Lab's ``--distributed`` launches one process per GPU and the RL library
all-reduces gradients); envs are fully independent, so the only cross-GPU traffic this engine adds is ONE all-gather of the
final observation tensor per step, over NCCL / NVLink (gloo in the CPU tests)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def env_shard(num_envs: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the contiguous env block of ``rank``; blocks differ by at most one env."""
    base, rem = divmod(num_envs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_obs(local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Gathers equally sized env shards (dim 0) from every rank into one tensor, rank-major = env order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class PeerObsGather:
    """The observation all-gather as direct NVLink peer copies into symmetric buffers.

    Every rank owns ``n_slots`` buffers of the full observation ``[world * E, ...]`` allocated as symmetric memory
    (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every peer process) and PUSHES its shard into the
    ``[rank * E, (rank + 1) * E)`` block of every peer's buffer with ``cudaMemcpyAsync`` (copy engines over NVLink / NVSwitch:
    no SM is taken from the compute kernels, unlike a NCCL all-gather kernel that has to wait for free SMs behind a full grid).
    Two signal-pad barriers per step order the pushes against the peers' consumers: (1) everybody has finished with the slot,
    (2) everybody's pushes have landed. All calls are stream-ordered on the current stream; nothing synchronises the host.
    """

    def __init__(self, local_shape, dtype, device, n_slots: int = 2, with_rects: bool = False, multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.E = int(local_shape[0])
        full = (self.world * self.E,) + tuple(local_shape[1:])
        self.bufs, self.hdls, self.views = [], [], []
        self.rects, self.rect_views, self.prev_rects = [], [], []
        self.mc_rgb, self.mc_rect = [], []  # multicast (NVSwitch) addresses of this rank's block, 0 when unavailable
        frame_bytes = 4
        for d in local_shape[1:]:
            frame_bytes *= int(d)
        for _ in range(n_slots):
            t = symm_mem.empty(full, dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, dist.group.WORLD)
            self.bufs.append(t)
            self.hdls.append(h)
            self.views.append([h.get_buffer(p, full, dtype) for p in range(self.world)])
            if with_rects:  # per half frame: the rectangle outside of which the frame equals the flat image
                rshape = (self.world * self.E, 2, 4)
                r = symm_mem.empty(rshape, dtype=torch.int32, device=device)
                rh = symm_mem.rendezvous(r, dist.group.WORLD)
                r.zero_()
                self.rects.append(r)
                self.rect_views.append([rh.get_buffer(p, rshape, torch.int32) for p in range(self.world)])
                mc_a, mc_b = (int(getattr(h, "multicast_ptr", 0) or 0), int(getattr(rh, "multicast_ptr", 0) or 0)) if multicast else (0, 0)
                ok = mc_a != 0 and mc_b != 0
                self.mc_rgb.append(mc_a + self.rank * self.E * frame_bytes if ok else 0)
                self.mc_rect.append(mc_b + self.rank * self.E * 32 if ok else 0)
                # what the buffer holds after its last fill (local): starts as "anything anywhere" = the whole half frame
                hh, ww = int(local_shape[1]) // 2, int(local_shape[2])
                self.prev_rects.append(torch.tensor([0, hh - 1, 0, ww - 1], dtype=torch.int32, device=device).repeat(rshape[0], 2, 1).contiguous())

    def local_rects(self, slot: int) -> torch.Tensor:
        lo = self.rank * self.E
        return self.rects[slot][lo:lo + self.E]

    def gather_rects(self, engine, slot: int, stream) -> torch.Tensor:
        """The same all-gather with the link traffic reduced to the pixels that carry information: the frames of
        ``local_block(slot)`` were rendered with ``engine.set_rect_output(local_rects(slot))``; their rectangles (and the
        16-byte descriptors) are stored into every peer's buffer by one kernel (NVLink peer-to-peer stores), and after the
        barrier every rank completes the remote envs of its own buffer from its flat image (only where the rectangle of the
        previous fill of this slot is not covered by the new one). Bit-identical to ``gather``. ``stream`` must be the
        current stream (the barriers are enqueued on it)."""
        h, lo = self.hdls[slot], self.rank * self.E
        peers = [(self.rank + k) % self.world for k in range(1, self.world)]
        h.barrier(channel=0)  # every rank is done reading the previous content of this slot
        engine.obs_push(self.local_block(slot), self.local_rects(slot),
                        [self.views[slot][p][lo:lo + self.E] for p in peers],
                        [self.rect_views[slot][p][lo:lo + self.E] for p in peers], stream,
                        mc_rgb=self.mc_rgb[slot], mc_rect=self.mc_rect[slot])
        h.barrier(channel=1)  # every rank's pushes into this slot have landed
        engine.obs_fill(self.bufs[slot], self.rects[slot], self.prev_rects[slot], lo, lo + self.E, stream)
        return self.bufs[slot]

    def fused_available(self, slot: int = 0) -> bool:
        return bool(self.rects) and bool(self.mc_rgb[slot])

    def begin_fused(self, engine, slot: int) -> None:
        """Fused all-gather, step 1 (current stream, BEFORE the render of this slot): every rank is done reading the slot's previous
        content (barrier), then the engine is pointed at the multicast mapping of this rank's block: the render kernel's epilogue
        stores every frame's rectangle + descriptor into ALL GPUs' buffers while it computes."""
        self.hdls[slot].barrier(channel=0)
        engine.set_rect_output(self.local_rects(slot))
        engine.set_multicast_output(self.mc_rgb[slot], self.mc_rect[slot])

    def finish_fused(self, engine, slot: int, stream) -> torch.Tensor:
        """Fused all-gather, step 2 (``stream`` = current stream, after the render): every rank's stores have landed (barrier), then
        the remote frames of this rank's copy are completed from its flat image. Bit-identical to ``gather``."""
        lo = self.rank * self.E
        self.hdls[slot].barrier(channel=1)
        engine.obs_fill(self.bufs[slot], self.rects[slot], self.prev_rects[slot], lo, lo + self.E, stream)
        return self.bufs[slot]

    def local_block(self, slot: int) -> torch.Tensor:
        """This rank's block of its own gathered buffer: render straight into it and the self-copy disappears."""
        lo = self.rank * self.E
        return self.bufs[slot][lo:lo + self.E]

    def gather(self, local: torch.Tensor, slot: int) -> torch.Tensor:
        h, lo = self.hdls[slot], self.rank * self.E
        own = self.local_block(slot)
        in_place = local.data_ptr() == own.data_ptr()
        h.barrier(channel=0)  # every rank is done reading the previous content of this slot
        for k in range(self.world):  # staggered targets: at any time every GPU receives from one peer
            p = (self.rank + k) % self.world
            if p == self.rank and in_place:
                continue
            self.views[slot][p][lo:lo + self.E].copy_(local, non_blocking=True)
        h.barrier(channel=1)  # every rank's pushes into this slot have landed
        return self.bufs[slot]
