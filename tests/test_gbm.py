"""GBM Monte-Carlo forecast: GPU self-checks (parity unpinned — not in the reference) and
the CPU/gloo test of the path-sharding + all-reduce logic."""
import os
import sys

import numpy as np
import pytest

from shockwave_b200 import forecast_mc as mc


def _params(J, seed=0):
    rng = np.random.default_rng(seed)
    R0 = rng.uniform(100, 1e5, J)
    H = rng.integers(1, 257, J).astype(np.int32)
    mu = rng.uniform(-2e-3, 2e-3, J)
    sigma = rng.uniform(0.0, 0.05, J)
    return R0, H, mu, sigma


@pytest.mark.gpu
def test_sigma_zero_reproduces_deterministic_forecast(engine):
    R0, H, mu, sigma = _params(64)
    mean, var = mc.gbm_forecast(engine, R0, H, 0 * mu, 0 * sigma, P=512)
    assert np.allclose(mean, R0, rtol=2e-6) and np.all(var <= 1e-9 * R0 ** 2)


@pytest.mark.gpu
def test_mean_converges_to_analytic_expectation(engine):
    R0, H, mu, sigma = _params(128, seed=1)
    P = 16384
    mean, var = mc.gbm_forecast(engine, R0, H, mu, sigma, P=P, seed=7)
    want = mc.analytic_mean(R0, H, mu)
    se = np.sqrt(var / P) + 2e-6 * want              # Monte-Carlo standard error + fp32 path arithmetic
    assert np.all(np.abs(mean - want) <= 5.0 * se), np.max(np.abs(mean - want) / se)


@pytest.mark.gpu
def test_result_does_not_depend_on_path_sharding(engine):
    R0, H, mu, sigma = _params(96, seed=2)
    P = 4096
    full = engine.gbm_forecast(R0, H, mu, sigma, P, 0, 11)
    parts = np.zeros_like(full)
    for r in range(3):
        lo, n = mc.path_range(P, r, 3)
        parts += engine.gbm_forecast(R0, H, mu, sigma, n, lo, 11)
    assert np.allclose(parts, full, rtol=1e-12)
    other = engine.gbm_forecast(R0, H, mu, sigma, P, 0, 12)
    assert not np.allclose(other[0], full[0], rtol=1e-9)    # the seed matters


@pytest.mark.gpu
def test_single_paths_match_the_numpy_restatement(engine):
    """The kernel against tests/ref_gbm.py path by path (P_local = 1 at several global path ids): stream keying,
    Box-Muller pairing and the running sum; tolerance = the fast fp32 intrinsics of the kernel."""
    from tests.ref_gbm import path_value
    R0, H, mu, sigma = _params(24, seed=5)
    H[:4] = [1, 2, 3, 256]
    worst = 0.0
    for gp in (0, 1, 7, 4095, 8191, 123456789):
        out = engine.gbm_forecast(R0, H, mu, sigma, 1, gp, 99)
        for j in range(len(R0)):
            want = path_value(99, j, gp, R0[j], int(H[j]), mu[j], sigma[j])
            worst = max(worst, abs(out[0, j] - want) / want)
            assert abs(out[0, j] - want) <= 2e-4 * want, (gp, j, out[0, j], want)
            assert abs(out[1, j] - out[0, j] ** 2) <= 1e-9 * out[0, j] ** 2
    print("gbm kernel vs numpy path restatement: worst relative difference", worst)


def test_path_range_partitions_exactly():
    for P in (1, 7, 8192, 8193):
        for w in (1, 2, 3, 8):
            got = [mc.path_range(P, r, w) for r in range(w)]
            assert sum(n for _, n in got) == P
            assert got[0][0] == 0 and all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(w - 1))


def _fake_kernel(R0, H, mu, sigma, n_local, lo, seed, out_device_ptr=None):
    """Deterministic stand-in with the kernel's contract: per-path values keyed by the GLOBAL path id."""
    J = len(R0)
    out = np.zeros((2, J))
    for p in range(lo, lo + n_local):
        v = np.asarray(R0) * (1.0 + 1e-3 * np.sin(0.37 * p + np.arange(J) + seed))
        out[0] += v
        out[1] += v * v
    return out


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R0, H, mu, sigma = _params(16, seed=3)

    def allreduce(sums):
        t = torch.from_numpy(np.ascontiguousarray(sums))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()
    mean, var = mc.gbm_forecast(None, R0, H, mu, sigma, P=1001, seed=5, rank=rank, world=world,
                                reduce_fn=allreduce, kernel_fn=_fake_kernel)
    q.put((rank, mean, var))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_rank():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    R0, H, mu, sigma = _params(16, seed=3)
    mean1, var1 = mc.gbm_forecast(None, R0, H, mu, sigma, P=1001, seed=5, kernel_fn=_fake_kernel)
    for _, mean, var in res:
        assert np.allclose(mean, mean1, rtol=1e-12) and np.allclose(var, var1, rtol=1e-5, atol=1e-9)  # E[x^2]-mean^2 cancels ~10 digits
