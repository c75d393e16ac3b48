"""CPU, world_size 2 over gloo: the host-side sharding + the statistics all-reduce.

Each rank plays its shard of global game ids with the CPU oracle (the CUDA env needs a GPU); the summed
statistics must equal a single-process run over the whole range -- the invariance the multi-GPU path
relies on (global-game-id keyed Philox, Tarok.py:34 rotation)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, seed, mode, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tarok_b200.dist import allreduce_stats, shard
    first, count = shard(total, rank, world)
    r = O.rollout(seed, first, count, mode, full=False)
    st = np.zeros(32, np.int64)
    st[0:4], st[4:8], st[19], st[20] = r["stats"][0:4], r["stats"][4:8], r["stats"][8], r["stats"][9]
    ok = r["err"] == 0
    st[8:18] = np.bincount(r["contract"][ok], minlength=10)
    st[18] = ok.sum()
    t = torch.from_numpy(st)
    allreduce_stats(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_shard_ranges_tile_the_batch():
    from tarok_b200.dist import shard
    for total in (1, 7, 8, 1000, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, cnt = shard(total, r, world)
                assert first == nxt
                nxt += cnt
            assert nxt == total


def test_two_rank_allreduce_equals_single_process(tmp_path, oracle):
    total, seed, mode = 30001, 77, 18
    out = str(tmp_path / "stats.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, seed, mode, out), nprocs=2, join=True)
    got = np.load(out)
    r = oracle.rollout(seed, 0, total, mode, full=False)
    ok = r["err"] == 0
    assert (got[0:4] == r["stats"][0:4]).all() and (got[4:8] == r["stats"][4:8]).all()
    assert got[19] == r["stats"][8] and got[18] == ok.sum()
    assert (got[8:18] == np.bincount(r["contract"][ok], minlength=10)).all()


def test_reference_arm_prints_the_contract_line(tmp_path):
    """`bench.py --impl reference` needs no GPU: it must print exactly one JSON line with the contract's keys -- the C port arm
    (`--no-pyref`) and, where the reference tree is installed, the arm that times the unmodified Python engine."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in (["--no-pyref"], []):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                              "--games", "65536"] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-500:]
        lines = [l for l in out.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, out.stdout[-500:]
        d = json.loads(lines[0])
        for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                  "cpu_baseline", "e2e"):
            assert k in d, k
        assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["value"] > 0
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        if extra:
            assert d["cpu_baseline"]["kind"] == "port"
