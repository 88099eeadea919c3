"""Image-level data parallelism: pairs are independent units, so the batch is sharded
contiguously across ranks with NO collective on the data path; the only exchange is one
all-reduce(SUM) of the statistics vector at the end of a sweep."""
import torch
import torch.distributed as dist

STAT_KEYS = ("bits_r", "bits_d", "se_r", "se_d", "pixels", "pairs")


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def new_stats():
    return {k: 0.0 for k in STAT_KEYS}


def add_pair_stats(stats, r_strings, d_strings, rgb, depth, rec_r, rec_d):
    """Accumulate bits and squared error of a (batch of) pair(s)."""
    stats["bits_r"] += 8.0 * sum(len(s) for grp in r_strings for s in grp)
    stats["bits_d"] += 8.0 * sum(len(s) for grp in d_strings for s in grp)
    stats["se_r"] += float(((rec_r.double() - rgb.double()) ** 2).sum())
    stats["se_d"] += float(((rec_d.double() - depth.double()) ** 2).sum())
    stats["pixels"] += float(rgb.shape[0] * rgb.shape[2] * rgb.shape[3])
    stats["pairs"] += float(rgb.shape[0])
    return stats


def allreduce_stats(stats, device="cpu"):
    """SUM over ranks (NCCL on GPUs, gloo in the CPU tests). Returns a new dict."""
    v = torch.tensor([stats[k] for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return dict(zip(STAT_KEYS, v.tolist()))


def summarize(stats):
    px = max(stats["pixels"], 1.0)
    import math
    out = {"pairs": int(stats["pairs"]), "bpp_r": stats["bits_r"] / px, "bpp_d": stats["bits_d"] / px}
    for m, ch in (("r", 3), ("d", 1)):
        mse = stats[f"se_{m}"] / (px * ch)
        out[f"psnr_{m}"] = 99.0 if mse <= 0 else 10 * math.log10(1.0 / mse)
    return out
