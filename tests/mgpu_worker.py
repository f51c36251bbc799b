"""Worker of tests/test_gpu_multigpu.py: launched with torchrun, one rank per GPU, NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shockwave_b200 import Engine, forecast_mc as mc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    rng = np.random.default_rng(5)
    J, P = 512, 4096
    R0 = rng.uniform(100, 1e5, J); H = rng.integers(1, 257, J).astype(np.int32)
    mu = rng.uniform(-1e-3, 1e-3, J); sg = rng.uniform(0, 0.05, J)
    buf = torch.zeros((2, J), dtype=torch.float64, device=f"cuda:{local}")
    mean, var = mc.gbm_forecast(eng, R0, H, mu, sg, P, seed=3, rank=rank, world=world,
                                reduce_fn=lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM), device_tensor=buf)
    if rank == 0:
        mean1, var1 = mc.gbm_forecast(eng, R0, H, mu, sg, P, seed=3)        # all paths on one GPU
        ok = np.allclose(mean, mean1, rtol=1e-12) and np.allclose(var, var1, rtol=1e-6, atol=1e-9)
        print("MGPU_OK" if ok else "MGPU_MISMATCH", float(np.abs(mean / mean1 - 1).max()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
