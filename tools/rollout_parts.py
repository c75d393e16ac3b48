"""CUDA-event timing of the rollout parts (setup / 48 steps / score / fused) at a given batch size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    env = TarokEnv(n, seed=1)
    print("games", n, "mode", mode)
    print("setup_synth   %8.1f us" % timed(lambda: env.setup_synth(mode, 0)))
    print("deal          %8.1f us" % timed(lambda: env.deal(0)))
    def sep():
        env.deal(0); env.force_contract_synth(mode) if mode not in (17, 18) else env.auction_synth(mode); env.exchange_synth(mode == 17)
    print("deal+begin+ex %8.1f us" % timed(sep))
    def steps():
        env.setup_synth(mode, 0); env.step_random(48)
    print("setup+48steps %8.1f us" % timed(steps))
    env.setup_synth(mode, 0); env.step_random(48)
    print("score         %8.1f us" % timed(lambda: env.score()))
    env.set_materialise(False)
    print("score (no mat)%8.1f us" % timed(lambda: env.score()))
    env.set_materialise(True)
    print("fused rollout %8.1f us" % timed(lambda: env.rollout(mode, 0, fused=True)))
    print("stepwise roll %8.1f us" % timed(lambda: env.rollout(mode, 0, fused=False)))
