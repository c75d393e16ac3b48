"""Pure pinned host->device copies on every rank at once: where does the multi-GPU end-to-end path lose its bandwidth?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_contention.py

Every rank copies the e2e input of one bench step (1,048,576 deals x 57 B = 59.8 MB from pinned memory) to its own GPU,
whole and in 8 chunks, alone (ranks take turns) and all at once, and the 8.4 MB of scores back at the same time.
Prints one JSON line: per-rank GB/s alone / concurrent and the aggregate -- if the concurrent aggregate stops growing with N
the wall is the host (memory / root complex), not the GPUs' links."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20
    src = torch.randint(0, 54, (n * 57,), dtype=torch.uint8).pin_memory()
    dst = torch.empty(n * 57, dtype=torch.uint8, device=dev)
    back_d = torch.zeros(n * 8, dtype=torch.uint8, device=dev)
    back_h = torch.empty(n * 8, dtype=torch.uint8).pin_memory()
    side = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(chunks, with_d2h, reps=20):
        step = (n * 57 + chunks - 1) // chunks
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            for c in range(chunks):
                dst[c * step:(c + 1) * step].copy_(src[c * step:(c + 1) * step], non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(side):
                    back_h.copy_(back_d, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        return n * 57 * reps / (a.elapsed_time(b) * 1e-3) / 1e9

    out = {}
    for name, chunks, d2h in (("whole", 1, False), ("chunks8", 8, False), ("chunks8_d2h", 8, True)):
        alone = 0.0
        for r in range(world):                       # ranks take turns
            barrier()
            if r == rank:
                alone = run(chunks, d2h)
        barrier()
        together = run(chunks, d2h)                  # everybody at once
        barrier()
        t = torch.tensor([alone, together], dtype=torch.float64, device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
        else:
            g = [t]
        out[name] = {"alone_gbs_per_rank": [round(float(x[0]), 2) for x in g], "concurrent_gbs_per_rank": [round(float(x[1]), 2) for x in g],
                     "concurrent_aggregate_gbs": round(sum(float(x[1]) for x in g), 2)}
    if rank == 0:
        try:
            topo = os.popen("nvidia-smi topo -m 2>/dev/null | head -12").read()
        except Exception:
            topo = ""
        print(json.dumps({"n_gpus": world, "bytes_per_copy": n * 57, "host_cpus": os.cpu_count(), "results": out, "topo": topo}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
