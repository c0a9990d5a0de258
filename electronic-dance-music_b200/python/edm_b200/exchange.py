"""Hill exchange between the replicas of one bias (one process per GPU).

Replaces the reference's flush_buffers / check_for_flush broadcast loop (lib/edm_bias.cpp:614-706):
each rank packs the hills it selected into a fixed-capacity block {count, centres...}, ONE all-gather
moves the blocks, and every rank commits the rank-major concatenation, so all replicas deposit the
same hills in the same canonical order and cum_bias_ needs no reduction.

The block layout is the one edm_bias_hills_pack_dev / edm_bias_hills_commit_dev use on the device
(include/edm_b200.h); the numpy helpers below restate it for host-side use and for the CPU tests.
torch.distributed is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np


def block_doubles(dim, cap):
    return 1 + cap * dim


def pack_block(centres, dim, cap):
    """centres: (n, dim) in candidate order -> float64 block of 1 + cap*dim doubles."""
    centres = np.asarray(centres, dtype=np.float64).reshape(-1, dim)
    n = centres.shape[0]
    if n > cap:
        raise ValueError("%d accepted hills exceed the exchange capacity %d" % (n, cap))
    blk = np.zeros(block_doubles(dim, cap))
    blk[0] = n
    blk[1:1 + n * dim] = centres.ravel()
    return blk


def unpack_blocks(blocks, dim, cap):
    """Rank-major concatenation of the centres held in `nblocks` consecutive blocks."""
    blocks = np.asarray(blocks, dtype=np.float64).reshape(-1, block_doubles(dim, cap))
    out = []
    for blk in blocks:
        n = int(blk[0])
        out.append(blk[1:1 + n * dim].reshape(n, dim))
    return np.concatenate(out, axis=0) if out else np.zeros((0, dim))


def all_gather_blocks(block, group=None):
    """block: 1-D torch tensor (CPU for gloo, CUDA for NCCL).  Returns the world_size blocks, rank-major."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(block.numel() * world, dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block, group=group)
    return out
