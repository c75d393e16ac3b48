"""One stepwise rollout driven by a device-resident action buffer (k_step<false>: externally supplied cards), for profiler
captures.  The actions are the cards of a recorded random rollout of the same deals (teacher forcing):
   ncu --set full --clock-control none --import-source on -k regex:k_step -s 56 -c 4 -o out python tools/one_rollout_forced.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rec = TarokEnv(n, seed=0x5EED7A20C0001, history=True)
rec.setup_synth(mode, 0)
rec.step_random(48)                                    # launches 0..47: k_step<true, POS>
hist = rec.hist
acts = torch.where(hist == 0xFF, hist, hist & 63).contiguous()
del hist
rec.close()
env = TarokEnv(n, seed=0x5EED7A20C0001)
env.set_materialise(False)
env.setup_synth(mode, 0)
for t in range(48):                                    # launches 48..95: k_step<false, POS>
    env.step(acts[t])
env.score()
torch.cuda.synchronize()
print(env.stats()[18:21])
env.close()
