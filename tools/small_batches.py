"""Stepwise vs fused rollout time for small batches (where the 50 launches, not the kernels, set the pace)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

for n in (1024, 4096, 16384, 65536, 262144):
    env = TarokEnv(n, seed=1); env.set_materialise(False)
    out = []
    for fused in (False, True):
        for _ in range(3): env.rollout(16, 0, fused=fused)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(20): env.rollout(16, i * n, fused=fused)
        b.record(); torch.cuda.synchronize()
        out.append((a.elapsed_time(b) / 20 * 1e3, (time.perf_counter() - t0) / 20 * 1e6))
    print("games %7d: stepwise %7.1f us (host %7.1f us)   fused %7.1f us" % (n, out[0][0], out[0][1], out[1][0]), flush=True)
    env.close()
