"""One neural self-play rollout at BASELINE config 4's size (65,536 envs, four players), for profiler captures of the
device-side bucketing / observation expansion / action selection kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.samoigra import Samoigra

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
s = Samoigra(n, seed=0x5EED7A20C0001, random_card=0.05, igralci=4)
st, _ = s.odigraj(0)
torch.cuda.synchronize()
print(st[18:21], sum(s.zadnji_koraki))
s.zapri()
