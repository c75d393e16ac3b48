"""Host-buffer entries: time per rollout vs pipeline depth (rows = 57 B/deal, records = 24 B/deal)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tarok_b200.env import TarokEnv, pack_records

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = TarokEnv(n, seed=1)
env.deal(0)
perm = torch.empty((n, 54), dtype=torch.uint8).pin_memory(); perm.copy_(env.export_perm())
rng = np.random.default_rng(0)
c = torch.from_numpy(rng.integers(1, 4, n, dtype=np.uint8)).pin_memory()
d = torch.from_numpy(rng.integers(0, 4, n, dtype=np.uint8)).pin_memory()
k = torch.from_numpy(rng.integers(0, 4, n, dtype=np.uint8)).pin_memory()
rec, _ = pack_records(perm, c, d, k)
sc = torch.empty((n, 4), dtype=torch.int16).pin_memory(); st = torch.zeros(32, dtype=torch.int64).pin_memory()
for chunks in (4, 8, 12, 16, 24, 32):
    env.set_chunks(chunks)
    out = []
    for records in (False, True):
        def once():
            if records: env.rollout_records(rec, sc, st)
            else: env.rollout_host(perm, c, d, k, sc, st, fused=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3): once()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): once()
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / 20)
    print("chunks %2d: rows %.3f ms  records %.3f ms" % (chunks, out[0], out[1]), flush=True)
env.close()
