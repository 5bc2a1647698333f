"""Multi-GPU layout of the hot path: a single sweep does not shard (latency bound,
strictly sequential LM), independent sequences do.  Sequence s lives on rank s mod G; nothing crosses
NVLink per frame; one all_gather of poses + timings at the end of the run."""
import numpy as np


def sequences_of_rank(n_sequences, rank, world_size):
    """Round-robin assignment: sequence s -> rank s mod world_size."""
    return [s for s in range(n_sequences) if s % world_size == rank]


def gather_results(dist, poses, timings, device=None):
    """poses: float64[K,14] of this rank, timings: float64[T].  Returns (poses[world,K,14],
    timings[world,T]) on every rank; with dist None (single process) it just adds the rank axis."""
    import torch
    p = torch.as_tensor(np.ascontiguousarray(poses), dtype=torch.float64)
    t = torch.as_tensor(np.ascontiguousarray(timings), dtype=torch.float64)
    if dist is None or not dist.is_initialized():
        return p[None].numpy(), t[None].numpy()
    if device is not None:
        p, t = p.to(device), t.to(device)
    ws = dist.get_world_size()
    gp = [torch.zeros_like(p) for _ in range(ws)]
    gt = [torch.zeros_like(t) for _ in range(ws)]
    dist.all_gather(gp, p)
    dist.all_gather(gt, t)
    return torch.stack(gp).cpu().numpy(), torch.stack(gt).cpu().numpy()


def aggregate_throughput(timings_ms, frames_per_rank):
    """Whole-job scans/s: all ranks' frames over the slowest rank's time (max over ranks)."""
    timings_ms = np.asarray(timings_ms, float)
    return float(len(timings_ms) * frames_per_rank / (timings_ms.max() * 1e-3))
