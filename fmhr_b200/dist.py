"""Host-side helpers for the multi-GPU HAM iteration (views shard across ranks).

The loss of one iteration is a mean over the GLOBAL batch:
    sfs  = sfs_weight  * sum_{valid px, rgb} |tmp_img - img| / (3 * N_valid_global)       (mesh_sfs_optim.py:289)
    mask = mask_weight * sum_{px} (pred_mask - valid_mask)^2 / (n_views_global * H * W)    (mesh_sfs_optim.py:295)
so every rank accumulates UN-NORMALISED sums (gradient accumulators and the three scalars) over its own views, one
all-reduce (sum) merges them, and the normalisation by the global counts happens afterwards, identically on every
rank (fmhr_ham_step_update).  Regularisers (Laplacian, edge, delta) are view-independent and computed redundantly.
"""
import torch


def shard_views(num_views, rank, world_size):
    """Round-robin view ownership: rank r owns views r, r + world, r + 2*world, ..."""
    return list(range(rank, num_views, world_size))


def allreduce_packed(packed, group=None):
    """One sum all-reduce of the packed [12V gradient accumulators | n_valid, abs_sum, mask_sq_sum, pad] buffer."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM, group=group)
    return packed


def normalisation_scales(conf, n_valid_global, n_views_global, H, W):
    """(s_photo, s_mask): factors that turn the un-normalised accumulators into the reference's gradients."""
    s_photo = conf["sfs_weight"] / (3.0 * float(n_valid_global))
    s_mask = 2.0 * conf["mask_weight"] / (float(n_views_global) * H * W)
    return s_photo, s_mask
