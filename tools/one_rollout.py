"""One stepwise rollout (setup -> 48 x play_step -> score) of config 2, for profiler captures:
   ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 8 --launch-count 4 -o out python tools/one_rollout.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 16
env = TarokEnv(n, seed=0x5EED7A20C0001)
env.set_materialise(False)
env.setup_synth(mode, 0)
env.step_random(48)
env.score()
torch.cuda.synchronize()
print(env.stats()[18:21])
env.close()
