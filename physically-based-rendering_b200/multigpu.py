"""Sharding one render job over the GPUs of a box: one process per GPU (torchrun), the scene replicated,
ONE collective on the float4 accumulation buffer per frame (BASELINE.json north_star).

Two partitionings, both pure functions of (rank, world size) so that the host logic can be tested on
CPU with the gloo backend:

  tiles : rank r renders image rows [tile_rows(H, r, R)) of every frame; the frame is completed with
          an all-gather of the row blocks.  Pixels are independent (each work-item of the reference
          kernel touches only its own pixel, pt_rgb.cl:13-20), so the result is bit-identical to one
          GPU.  Strong scaling of one frame.
  spp   : every rank renders whole frames with its own seeds -- frame k of rank r uses the global
          frame index k * R + r in the seed schedule seed_j = 0.0333f * (j + 1) -- and keeps its own
          running average; the displayed frame is the mean over ranks: all-reduce(sum) * (1 / R).
          R times the samples per unit time (weak scaling); float summation order differs from one
          GPU, so parity is within tolerance, not bit-exact.

torch / torch.distributed are used for what they are here: device tensors, streams, NCCL plumbing.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover - torch is part of the image
    torch = None
    dist = None


def tile_rows(height, rank, world, align=4):
    """Rows [y0, y1) of rank `rank`: contiguous blocks, multiples of `align` rows (the kernel walks
    8x4 pixel blocks) except possibly the last block.  Covers [0, height) exactly."""
    blocks = (height + align - 1) // align
    b0 = (blocks * rank) // world
    b1 = (blocks * (rank + 1)) // world
    return min(b0 * align, height), min(b1 * align, height)


def global_frame_index(local_frame, rank, world):
    """Seed-schedule index of local frame k on rank r under spp sharding."""
    return local_frame * world + rank


def frame_seed(global_index):
    return np.float32(np.float32(0.0333) * np.float32(global_index + 1))


def pixel_weight(sample_count):
    return np.float32(np.float32(sample_count) / np.float32(sample_count + 1))


class DeviceImage:
    """torch view (no copy) of a device buffer owned by libpbr_b200.so."""

    def __init__(self, ptr, height, width, device):
        self.__cuda_array_interface__ = {
            "shape": (height, width, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None,
        }
        self.tensor = torch.as_tensor(self, device=device)


def combine_tiles(image, rank, world, group=None):
    """All-gather the row blocks in place: image [H, W, 4] holds this rank's rows on entry and the
    whole frame on return.  One collective per frame."""
    height = image.shape[0]
    rows = [tile_rows(height, r, world) for r in range(world)]
    sizes = {y1 - y0 for y0, y1 in rows}
    if len(sizes) == 1:
        y0, y1 = rows[rank]
        # equal blocks: gather straight into the frame (the local block is already in place)
        dist.all_gather_into_tensor(image.view(-1), image[y0:y1].reshape(-1).clone(), group=group)
    else:
        parts = [torch.empty_like(image[y0:y1]) for y0, y1 in rows]
        y0, y1 = rows[rank]
        dist.all_gather(parts, image[y0:y1].contiguous(), group=group)
        for (a, b), p in zip(rows, parts):
            image[a:b].copy_(p)
    return image


def stripe_rows_for(height, world, want=8):
    """Largest stripe height <= want such that height is a multiple of stripe * world (0 if world does not divide)."""
    if height % world:
        return 0
    local = height // world
    for s in range(min(want, local), 0, -1):
        if local % s == 0:
            return s
    return 0


def combine_stripes(image, rank, world, stripe, group=None):
    """Interleaved row sharding (pbr_set_tile_stripes): image [H, W, 4] holds this rank's stripes on entry and the
    whole frame on return.  One all-gather per frame; the stripes are packed / unpacked with two strided copies."""
    height, width, ch = image.shape
    groups = height // (stripe * world)
    v = image.view(groups, world, stripe, width, ch)
    send = v[:, rank].contiguous()
    recv = torch.empty((world,) + tuple(send.shape), dtype=image.dtype, device=image.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    v.copy_(recv.permute(1, 0, 2, 3, 4))
    return image


def combine_spp(image, world, out=None, group=None):
    """Mean over ranks of the per-rank running averages: out = all_reduce_sum(image) / world.
    `image` is left untouched (it keeps accumulating); one collective per frame."""
    if out is None:
        out = torch.empty_like(image)
    out.copy_(image)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    out.mul_(1.0 / world)
    return out
