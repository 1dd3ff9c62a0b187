"""Work partitioning of the render path across the GPUs of one box (one process per GPU).

The path shards without any data-path exchange: pixels are independent given read-only inputs.
  * frames of an animation (the `t` variable), frame f on rank f mod G  -- weak scaling per GPU
  * one large frame split into row bands: contiguous, with the reference's own formula
    rows [y0 + h*i/n, y0 + h*(i+1)/n) (mathmap_common.c:997-998), or 8-row blocks interleaved
    across ranks when the filter's cost varies over the image (escape-time fractals)
  * input drawables are replicated to every GPU with one broadcast (NCCL over NVLink on GPUs,
    gloo in the CPU tests), because samplers read arbitrary coordinates.
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


def max_over_ranks(value, device=None):
    """Max of a Python float over all ranks (timing reduction)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
