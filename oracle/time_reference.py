"""Times the REAL reference (imported from /root/reference through oracle/ref_harness.py) on its own public path:
``Tarok(4 x Bot_igralec, n).paralel_start()`` (Tarok.py:30-62) -- the batched rollout this repository replaces -- and on the
forced-Klop workload of BASELINE config 1.  Build container only (the reference cannot travel to the GPU box); the result
is committed as profiles/r01/reference_cpu_container.json and quoted in DESIGN.md next to the C port's numbers.

    python oracle/time_reference.py [deals_per_worker] [workers]
"""
import json
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _bots(ref):
    return [ref.Igralec.Bot_igralec() for _ in range(4)]


def paralel_start(n, seed):
    """n deals through Tarok.paralel_start with four Bot_igralec (random bids Klop/Tri/Dve/Ena, random legal cards)."""
    from oracle import ref_harness as H
    ref = H.load_reference()
    random.seed(seed)
    import numpy as np
    np.random.seed(seed & 0x7FFFFFFF)
    t = ref.Tarok.Tarok(_bots(ref), n)
    t.izpis = False
    t0 = time.perf_counter()
    t.paralel_start()
    return n, 48 * n, time.perf_counter() - t0          # Bot bids never reach Berac: 48 card plays per deal


def klop(n, seed):
    """n forced-Klop deals (config 1): Igra.razdeli + Klop(...).start() with random legal cards."""
    from oracle import ref_harness as H
    ref = H.load_reference()
    random.seed(seed)
    t0 = time.perf_counter()
    for _ in range(n):
        players = _bots(ref)
        talon = ref.Igra.Igra(players).razdeli()
        list(ref.Klop.Klop(players, talon, 0).start())
    return n, 48 * n, time.perf_counter() - t0


def nevronski(n, seed):
    """n deals through Tarok.paralel_start with four of the reference's own Nevronski_igralec (BASELINE config 4's path) on
    the restated networks (tarok_b200/compat/torch_models.py fills the module missing upstream); CPU forward passes."""
    import importlib.util
    import tempfile
    import numpy as np
    import torch
    from oracle import ref_harness as H
    spec = importlib.util.spec_from_file_location("torch_models", os.path.join(ROOT, "tarok_b200", "compat", "torch_models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["torch_models"] = mod
    ref = H.load_reference()
    random.seed(seed); np.random.seed(seed & 0x7FFFFFFF); torch.manual_seed(seed)
    os.chdir(tempfile.mkdtemp())                      # the player creates its save directory relative to the cwd
    players = [ref.Igralec.Nevronski_igralec() for _ in range(4)]
    t = ref.Tarok.Tarok(players, n)
    t.izpis = False
    t0 = time.perf_counter()
    t.paralel_start()
    dt = time.perf_counter() - t0
    return n, 48 * n, dt                              # upper bound: Berac games stop early (a few per cent of the deals)


def _run(args):
    fn, n, seed = args
    with open(os.devnull, "w") as devnull:            # the reference prints progress lines
        old = sys.stdout
        sys.stdout = devnull
        try:
            return globals()[fn](n, seed)
        finally:
            sys.stdout = old


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    out = {"host": "build container", "cpu_count": os.cpu_count(), "python": sys.version.split()[0], "results": []}
    for fn in ("klop", "paralel_start", "nevronski"):
        for p in ((1,) if fn == "nevronski" else (1, workers)):
            t0 = time.perf_counter()
            if fn == "nevronski":                      # own process: Igralec must be imported with the real network module
                with mp.get_context("spawn").Pool(1) as pool:
                    res = pool.map(_run, [(fn, max(8, per // 8), 1)])
            elif p == 1:
                res = [_run((fn, per, 1))]
            else:
                with mp.Pool(p) as pool:
                    res = pool.map(_run, [(fn, per, 100 + i) for i in range(p)])
            wall = time.perf_counter() - t0
            deals, steps = sum(r[0] for r in res), sum(r[1] for r in res)
            busy = max(r[2] for r in res)
            out["results"].append({"workload": fn, "processes": p, "deals": deals, "env_steps": steps, "seconds": busy,
                                   "wall_seconds_incl_pool_start": wall, "deals_per_sec": deals / busy,
                                   "env_steps_per_sec": steps / busy})
            print(out["results"][-1], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
