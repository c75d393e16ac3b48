"""GPU, 2 ranks: game shards + the one NCCL collective (tarok_allreduce_stats) == a single-GPU run of the same games."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from tarok_b200.env import TarokEnv
from tarok_b200.dist import NcclComm, shard, allreduce_stats
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
total, seed, mode = 300001, 99, 17
first, count = shard(total, rank, world)
env = TarokEnv(count, seed=seed, device=rank)
env.rollout(mode, first_game_id=first)
comm = NcclComm(rank)
got = comm.allreduce_stats(env).cpu().numpy()            # raw NCCL through the C ABI
via_torch = allreduce_stats(env.stats_dev.clone()).cpu().numpy()
assert (got == via_torch).all()
if rank == 0:
    ref = TarokEnv(total, seed=seed, device=0)
    ref.rollout(mode, first_game_id=0)
    want = ref.stats()
    assert (got[:21] == want[:21]).all(), (got[:21], want[:21])
    print("MULTI_OK", int(got[19]))
comm.close()
dist.destroy_process_group()
'''


def test_two_gpu_shards_and_nccl_allreduce_equal_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert "MULTI_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
