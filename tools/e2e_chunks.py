"""Host-buffer entries: wall time per rollout vs pipeline depth (rows = 57 B/deal, records = 20 B/deal, rows packed in the call).
   python tools/e2e_chunks.py [games] [pack threads]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tarok_b200.env import TarokEnv, pack_records

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
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
    for kind in ("rows", "records", "packed"):
        def once():
            if kind == "records": env.rollout_records(rec, sc, st)
            elif kind == "packed": env.rollout_host_packed(perm, c, d, k, sc, st, threads=threads)
            else: env.rollout_host(perm, c, d, k, sc, st, fused=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3): once()
        import time
        t0 = time.perf_counter()
        for _ in range(20): once()
        out.append((time.perf_counter() - t0) / 20 * 1e3)
    print("chunks %2d: rows %.3f ms  records (prepacked) %.3f ms  rows packed in the call by %d threads %.3f ms" % (chunks, out[0], out[1], threads, out[2]), flush=True)
env.close()
