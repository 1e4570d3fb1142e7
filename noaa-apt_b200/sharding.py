"""Host-side sharding of independent recordings across GPUs / ranks.

Recordings never interact, so there is no data-path collective: recording i goes to device i mod G and, on
that device, to stream (i // G) mod S.  Across processes (one per GPU, torchrun) rank r owns the recordings
with i mod world == r.  The only cross-rank communication is for measurement: a barrier and the MAX of the
per-rank elapsed times (bench.py), done with torch.distributed (nccl on GPUs, gloo in the CPU tests).
"""


def assign(i, n_devices, streams_per_device):
    """(device index, stream slot) of recording i -- the rule apt_decode_batch implements in C."""
    return i % n_devices, (i // n_devices) % streams_per_device


def rank_recordings(count, rank, world):
    """Indices of the recordings rank `rank` of `world` decodes."""
    return list(range(rank, count, world))


def aggregate_throughput(samples_local, ms_local, dist=None, device=None):
    """Whole-job Msamples/s: total samples of all ranks / MAX elapsed over ranks."""
    import torch
    total = torch.tensor([float(samples_local)], dtype=torch.float64, device=device)
    worst = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return float(total.item()) / (float(worst.item()) * 1e-3) / 1e6, float(worst.item())
