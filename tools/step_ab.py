"""A/B timing of the play_step implementations (plain vs persistent TMA-staged) with CUDA events.
usage: python tools/step_ab.py [games ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

def run(n, impl, pdl=True, lock=True, mode=16, reps=3):
    env = TarokEnv(n, seed=1)
    env.set_step_impl(impl)
    env.set_pdl(pdl)
    env.set_lockstep(lock)
    best = 1e9
    for r in range(reps):
        env.deal(r * n); env.force_contract_synth(mode); env.exchange_synth(False)
        env.step_random(4)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); env.step_random(40); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 40 * 1e3)
    env.close()
    return best

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [1 << 20, 1 << 23]
    for n in sizes:
        for impl, pdl, lock in ((1, False, False), (1, True, False), (1, True, True), (2, True, False)):
            us = run(n, impl, pdl, lock)
            print("games %9d impl %d pdl %d lockstep %d: %8.2f us/launch  %7.1f GB/s algorithmic (64 B/step)  frac %.3f" %
                  (n, impl, pdl, lock, us, 64 * n / us / 1e3, 64 * n / us / 1e3 / 6457.4), flush=True)
