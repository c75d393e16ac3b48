"""TEST INFRASTRUCTURE ONLY -- times the REAL reference (the unmodified Python engine, imported through
oracle/ref_harness.py from ``baseline/_ref`` or ``/root/reference``, whichever exists on the box) on its own public paths:

* ``klop``            config 1: forced Klop, ``Igra.razdeli`` + ``Klop(players, talon, 0).start()`` (Klop.py:16-45);
* ``tri/dve/ena``     forced Navadna igra with talon exchange, ``Navadna_igra(players, tip, king, players[d], talon, 0)``
                      (Navadna_igra.py:15, uniform declarer and king like bench.py's config 2);
* ``solo``            forced Solo_tri/dve/ena; ``berac`` forced Berac (Berac.py:5-44, early stop);
* ``paralel_start``   ``Tarok(4 x Bot_igralec, n).paralel_start()`` (Tarok.py:30-62): full bidding + play, the batched rollout
                      this repository replaces;
* ``nevronski``       the same with four of the reference's own ``Nevronski_igralec`` on the restated nets (config 4's path).

All with four ``Bot_igralec`` (uniform-random legal moves, Igralec.py:142-171), one worker process per core
(``multiprocessing.Pool``), an own ``random.seed`` per worker.  The reference cannot travel to the GPU
box, so there bench.py reports "absent on this box"; in the build container the result is committed under profiles/.

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


class _Stevec:
    """Counts the card plays of a contract generator: one 'Pripravljen igrat karto' yield per env-step."""

    @staticmethod
    def run(gen):
        steps = 0
        for item in gen:
            if item == "Pripravljen igrat karto":
                steps += 1
        return steps


def _forced(n, seed, tips, berac=False):
    from oracle import ref_harness as H
    ref = H.load_reference()
    random.seed(seed)
    T, B = ref.Tip_igre.Tip_igre, ref.Karta.Barva
    suits = [B.KARA, B.SRCE, B.PIK, B.KRIZ]
    steps, done = 0, 0
    t0 = time.perf_counter()
    for i in range(n):
        players = _bots(ref)
        talon = ref.Igra.Igra(players).razdeli()
        d = random.randrange(4)
        try:
            if berac:
                g = ref.Berac.Berac(players, players[d], talon, False, 0)
            else:
                tip = getattr(T, tips[i % len(tips)])
                king = random.choice(suits) if tip in (T.Tri, T.Dve, T.Ena) else None
                g = ref.Navadna_igra.Navadna_igra(players, tip, king, players[d], talon, 0)
            steps += _Stevec.run(g.start())
            done += 1
        except ValueError:                                  # fewer than k discardable cards (Igralec.py:166, Q19)
            pass
    return done, steps, time.perf_counter() - t0


def tri(n, seed):
    return _forced(n, seed, ["Tri"])


def dve(n, seed):
    return _forced(n, seed, ["Dve"])


def ena(n, seed):
    return _forced(n, seed, ["Ena"])


def navadna_mix(n, seed):
    """bench.py's config 2 mix: Tri/Dve/Ena in turn, uniform declarer and king."""
    return _forced(n, seed, ["Tri", "Dve", "Ena"])


def solo(n, seed):
    return _forced(n, seed, ["Solo_tri", "Solo_dve", "Solo_ena"])


def berac(n, seed):
    return _forced(n, seed, [], berac=True)


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


def run_workload(fn, total_deals, workers, seed0=100):
    """`total_deals` of workload `fn` spread over `workers` processes (seeds seed0 + i); returns the result row."""
    per = max(1, total_deals // workers)
    t0 = time.perf_counter()
    if workers == 1:
        res = [_run((fn, per, seed0 - 99))]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_run, [(fn, per, seed0 + i) for i in range(workers)])
    wall = time.perf_counter() - t0
    deals, steps = sum(r[0] for r in res), sum(r[1] for r in res)
    busy = max(r[2] for r in res)
    return {"workload": fn, "processes": workers, "deals": deals, "env_steps": steps, "seconds": busy,
            "wall_seconds_incl_pool_start": wall, "deals_per_sec": deals / busy, "env_steps_per_sec": steps / busy}


def bounded(workers, budget_seconds=20.0):
    """bench.py's leg: BASELINE.md section 3's workloads on `workers` processes, sized from a short probe so that the whole
    leg takes about `budget_seconds`: config 1 exactly (10,000 Klop deals) when it fits its share, 2,000 deals of each other
    contract (fewer on a slow box -- the sample is stated per row)."""
    from oracle import ref_harness as H
    H.load_reference()                                                   # once, before the pools fork: the workers inherit it
    rows = []
    run_workload("klop", 10 * workers, workers)                          # warm-up (imports, page faults)
    probe = run_workload("klop", 150 * workers, workers)
    rate = probe["deals_per_sec"]                                        # deals/s with all workers
    plan = [("klop", 10000, 0.35), ("navadna_mix", 2000, 0.1), ("tri", 2000, 0.1), ("solo", 2000, 0.1), ("berac", 2000, 0.1),
            ("paralel_start", 2000, 0.1)]
    for fn, want, share in plan:
        n = int(min(want, max(workers * 20, rate * budget_seconds * share)))
        rows.append(run_workload(fn, n, workers))
        rows[-1]["config_sized"] = n == want
    one = run_workload("klop", int(max(50, min(2000, rate / workers * budget_seconds * 0.1))), 1)      # P = 1 beside P = all
    rows.append(one)
    return rows


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    out = {"host": "build container", "cpu_count": os.cpu_count(), "python": sys.version.split()[0], "results": []}
    for fn in ("klop", "navadna_mix", "tri", "dve", "ena", "solo", "berac", "paralel_start", "nevronski"):
        for p in ((1,) if fn == "nevronski" else (1, workers)):
            if fn == "nevronski":                      # own process: Igralec must be imported with the real network module
                t0 = time.perf_counter()
                with mp.get_context("spawn").Pool(1) as pool:
                    res = pool.map(_run, [(fn, max(8, per // 8), 1)])
                wall = time.perf_counter() - t0
                deals, steps, busy = res[0][0], res[0][1], res[0][2]
                row = {"workload": fn, "processes": 1, "deals": deals, "env_steps": steps, "seconds": busy,
                       "wall_seconds_incl_pool_start": wall, "deals_per_sec": deals / busy, "env_steps_per_sec": steps / busy}
            else:
                row = run_workload(fn, per * p if fn == "klop" else max(p * 50, per * p // 5), p)
            out["results"].append(row)
            print(row, file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
