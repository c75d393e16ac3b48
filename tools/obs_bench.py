"""CUDA-event timing of obs_expand / select_action at config-4 size (65,536 envs) for each history length."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = TarokEnv(n, seed=3, history=True)
env.setup_synth(2)            # all Dve: one (net, T) bucket per step
for t in range(48):
    kinds, rows = env.obs_shape()
    T = int(rows[0].item())
    if t % 8 in (0, 7):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            arrs, ok = env.obs_expand(1, T)      # warm the caching allocator (two generations of outputs alive)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            arrs, ok = env.obs_expand(1, T)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 5 * 1e3
        byts = sum(x.numel() * 4 for x in arrs)
        print("play %2d T %2d: obs_expand %8.1f us  %7.1f MB written  %7.1f GB/s (incl. torch.empty allocs)" % (t, T, us, byts / 1e6, byts / us / 1e3), flush=True)
        assert float(ok.float().mean()) > 0.999
    env.step_random(1)
