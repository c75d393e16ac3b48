"""Graph-backed vs plain-launch stepwise rollout on the same box: GPU time (CUDA events) and host enqueue time per rollout.
   python tools/ab_graph.py [games]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tarok_b200.env import TarokEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = 20
out = {"games": n}
for rep in range(2):
    for graph in (True, False):
        env = TarokEnv(n, seed=1)
        env.set_materialise(False)
        env.set_graph(graph)
        for i in range(3):
            env.rollout(16, i * n)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for i in range(reps):
            env.rollout(16, (3 + i) * n)
        t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
        k = "graph" if graph else "plain"
        out["%s_gpu_us_%d" % (k, rep)] = round(a.elapsed_time(b) / reps * 1e3, 1)
        out["%s_host_us_%d" % (k, rep)] = round((t1 - t0) / reps * 1e6, 1)
        env.close()
print(json.dumps(out))
