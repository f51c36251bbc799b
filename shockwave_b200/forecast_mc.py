"""Monte-Carlo (GBM) remaining-runtime forecast, sharded over GPUs by sample path.

One process per GPU (torch.distributed, NCCL over NVLink): rank r simulates the global paths
[r*P/N, (r+1)*P/N) of every job with swb_gbm_forecast, then ONE all-reduce (SUM) of the [2, J] float64
vector of partial sums gives every rank the mean and variance per job.  Because every
path's random stream is keyed by its GLOBAL id, the result does not depend on N (up to float64
summation order).  `reduce_fn` / `kernel_fn` are injectable so the sharding logic is testable on CPU
with the gloo backend.
"""
from __future__ import annotations

import numpy as np


def path_range(P, rank, world):
    """Contiguous shard of P global paths owned by `rank` (first P % world ranks get one more)."""
    base, extra = divmod(int(P), int(world))
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def analytic_mean(R0, H, mu):
    """E[R] = R0 * mean_{h=1..H} e^{mu h}   (E[exp((mu - s^2/2)h + s W_h)] = e^{mu h})."""
    R0, H, mu = np.asarray(R0, float), np.asarray(H, int), np.asarray(mu, float)
    out = np.array(R0, dtype=float)
    for j in range(len(R0)):
        if H[j] > 0:
            out[j] = R0[j] * np.mean(np.exp(mu[j] * np.arange(1, H[j] + 1)))
    return out


def gbm_forecast(engine, R0, H, mu, sigma, P, seed=0, rank=0, world=1, reduce_fn=None, kernel_fn=None,
                 device_tensor=None):
    """Returns (mean[J], var[J]) of the remaining runtime under the GBM model, using all `world` GPUs.

    device_tensor: optional torch float64 CUDA tensor [2, J]; when given the kernel writes its partial
    sums straight into it and `reduce_fn(device_tensor)` (e.g. dist.all_reduce) runs on the device."""
    lo, n_local = path_range(P, rank, world)
    if kernel_fn is None:
        kernel_fn = engine.gbm_forecast
    if device_tensor is not None:
        try:        # the kernel runs on the engine's own (non-blocking) stream: whatever torch still has queued on
            import torch   # ITS stream for this tensor (a zero-fill, a previous reduce) must be finished first
            torch.cuda.current_stream(device_tensor.device).synchronize()
        except ImportError:
            pass
        kernel_fn(R0, H, mu, sigma, n_local, lo, seed, out_device_ptr=device_tensor.data_ptr())
        if reduce_fn is not None:
            reduce_fn(device_tensor)
        sums = device_tensor.cpu().numpy()
    else:
        sums = kernel_fn(R0, H, mu, sigma, n_local, lo, seed)
        if reduce_fn is not None:
            sums = reduce_fn(sums)
    mean = sums[0] / P
    var = np.maximum(0.0, sums[1] / P - mean * mean)
    return mean, var
