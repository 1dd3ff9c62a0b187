"""Work partitioning of the render path across the GPUs of one box (one process per GPU).

The path shards without any data-path exchange: pixels are independent given read-only inputs.
  * frames of an animation (the `t` variable), frame f on rank f mod G  -- weak scaling per GPU
  * one large frame split into row bands: contiguous, with the reference's own formula
    rows [y0 + h*i/n, y0 + h*(i+1)/n) (mathmap_common.c:997-998), or 8-row blocks interleaved
    across ranks when the filter's cost varies over the image (escape-time fractals)
  * input drawables are replicated to every GPU, because samplers read arbitrary coordinates: with one
    broadcast from the rank that holds the image, or -- when every rank can read the host image -- each rank
    uploads one band and one all-gather replicates (NCCL over NVLink on GPUs, gloo in the CPU tests).
torch.distributed is plumbing only.
"""
import numpy as np


def frames_for_rank(num_frames, rank, world):
    """Frame f is rendered by rank f mod world."""
    return list(range(rank, num_frames, world))


def band_for_rank(first_row, last_row, rank, world):
    """The reference's band split: [first + h*i/n, first + h*(i+1)/n)."""
    h = last_row - first_row
    return first_row + h * rank // world, first_row + h * (rank + 1) // world


def interleaved_rows_for_rank(height, rank, world, block=8):
    """Absolute rows of the 8-row blocks b with b % world == rank, in the compact order the kernel stores them."""
    rows = []
    for b in range(rank, (height + block - 1) // block, world):
        rows.extend(range(b * block, min(height, (b + 1) * block)))
    return rows


def assemble_interleaved(parts, height, block=8):
    """Inverse of the interleaved split: parts[r] holds rank r's compact rows ([rows_r, W, C] arrays)."""
    world = len(parts)
    out = np.empty((height,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        rows = interleaved_rows_for_rank(height, r, world, block)
        out[rows] = p[:len(rows)]
    return out


def broadcast_drawable(tensor, src=0):
    """Replicates an input drawable (uint8 [H, W, 4] tensor, already allocated on every rank) from `src`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def replicate_drawable_bands(tensor):
    """Replicates an input drawable of which every rank has uploaded only its own contiguous band (band_for_rank over the
    rows of `tensor`, uint8 [H, W, 4], allocated on every rank): the host->device traffic of the frame is spread over the
    ranks' own PCIe links, and the bands travel between the GPUs once (one in-place all-gather over NVLink when the
    bands are equal, else one broadcast per band)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    world, rank, h = dist.get_world_size(), dist.get_rank(), tensor.shape[0]
    if h % world == 0 and dist.get_backend() == "nccl" and tensor.is_contiguous():
        r0, r1 = band_for_rank(0, h, rank, world)
        dist.all_gather_into_tensor(tensor.view(-1), tensor[r0:r1].view(-1))  # in place: the send buffer is this rank's slot
    else:
        for r in range(world):
            r0, r1 = band_for_rank(0, h, r, world)
            if r1 > r0:
                dist.broadcast(tensor[r0:r1], src=r)
    return tensor


def max_over_ranks(value, device=None):
    """Max of a Python float over all ranks (timing reduction)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
