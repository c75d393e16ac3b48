"""Prototype check: leader-relative hand slots (step impl 3) vs the seat-indexed layout: same scores, step time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

def run(n, impl, mode):
    env = TarokEnv(n, seed=1)
    env.set_step_impl(impl)
    env.set_materialise(False)
    best = 1e9
    for r in range(3):
        env.reset_stats()
        env.setup_synth(mode, r * n)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); env.step_random(48); b.record()
        env.score()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 48 * 1e3)
    sc = env.scores[:n].clone()
    st = env.stats().copy()
    env.close()
    return best, sc, st

for n in [int(x) for x in sys.argv[1:]] or [1 << 20, 1 << 23]:
    for mode in (16, 17):
        t1, s1, st1 = run(n, 1, mode)
        print("games %9d mode %d: %7.2f us/step over a whole 48-step rollout" % (n, mode, t1), flush=True)
