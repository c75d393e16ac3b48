"""CPU enqueue cost vs GPU time of one stepwise rollout (setup -> 48 x play_step -> score), and the same work replayed as a
CUDA graph: tells whether a box is launch-bound.   python tools/launch_cost.py [games]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = TarokEnv(n, seed=1)
env.set_materialise(False)
reps = 20
for _ in range(3):
    env.rollout(16, 0, fused=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for i in range(reps):
    env.rollout(16, i * n, fused=False)
t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
out = {"games": n, "cpu_enqueue_us_per_rollout": (t1 - t0) / reps * 1e6, "gpu_us_per_rollout": a.elapsed_time(b) / reps * 1e3}
s = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    env.rollout(16, 0, fused=False); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        env.rollout(16, 0, fused=False)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
t0 = time.perf_counter(); a.record()
for i in range(reps):
    g.replay()
t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
out.update({"graph_cpu_us_per_replay": (t1 - t0) / reps * 1e6, "graph_gpu_us_per_replay": a.elapsed_time(b) / reps * 1e3})
print(json.dumps({k: round(v, 1) if isinstance(v, float) else v for k, v in out.items()}))
env.close()
