#!/usr/bin/env python
"""bench.py -- env-steps/s and deals/s of the batched Tarok environment (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--mode M] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: deal -> contract -> talon exchange ->
48 x play_step (uniform-random legal-move players) -> score, for G concurrent deals per GPU
(default: BASELINE config 2 = 1,048,576 Navadna deals on one B200).  Prints ONE JSON line.

* value      whole-job env-steps/s, inputs resident in HBM (Philox deals generated on device); each of the K steps
             is bracketed by its own CUDA-event pair, a 160 MiB L2-flush write runs between the pairs
* e2e        same metric through the host-buffer C-ABI entry (tarok_rollout_host): pinned host deals
             + contracts uploaded, scores + stats downloaded, inside the timed region
* roofline   the dominant kernel k_step<random>: 64 B/env-step x live games per
             launch / its CUDA-event duration over the timed region, against MEASURED_PEAKS.json
* cpu_baseline  the C port of the reference rules (oracle/, OpenMP over all host cores) on a bounded
             sample of the same workload -- a reported baseline, not the target
* --impl reference  times that same CPU port as the reference arm (the Python reference cannot
             travel to the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5EED7A20C0001
BYTES_PER_ENV_STEP = 64          # play_step with the mask fused in
MODE_NAMES = {16: "Navadna igra (Tri/Dve/Ena forced, uniform declarer+king), talon exchange",
              17: "uniform index2igra bids (all contracts incl. Berac), random talon group",
              18: "Bot_igralec bidding (Klop/Tri/Dve/Ena)", 0: "Klop forced"}


def workload_name(mode, games):
    return ("config 2: %s; 4 uniform-random legal-move players; %d concurrent deals per GPU; one step = deal, contract, "
            "talon exchange (one fused setup launch), 48 x play_step, score" % (MODE_NAMES.get(mode, str(mode)), games))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None
        self.nvml, self.samples, self.stop_flag = None, [], False

    def _nvml_loop(self):
        import pynvml as N
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        self.mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((sm, self.mx, rs))
            except Exception:
                break
            time.sleep(0.001)

    def _nvml_result(self):
        import pynvml as N
        self.stop_flag = True
        self.t.join(timeout=2)
        names = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(s[1] for s in self.samples) if sm else None,
                "samples": len(sm), "reasons": reasons, "source": "NVML polling thread, timed region only"}

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            self.nvml = N
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            return self._nvml_result()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v == "Active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_rate(mode, target_seconds, threads_note=True):
    """Times the C port of the reference engine (oracle/synth.c, OpenMP) on a bounded sample."""
    from oracle import oracle as O
    O.build()
    cores = O.use_all_threads()
    t = time.perf_counter()
    st = O.rollout_stats_only(SEED, 0, 20000, mode)
    dt = max(time.perf_counter() - t, 1e-4)
    n = int(min(max(20000 / dt * target_seconds, 20000), 40_000_000))
    t = time.perf_counter()
    st = O.rollout_stats_only(SEED, 0, n, mode)
    dt = time.perf_counter() - t
    out = {"steps_per_s": float(st[8]) / dt, "deals_per_s": n / dt, "cores": cores, "deals": n, "seconds": dt}
    O.set_threads(1)                                   # P = 1 beside P = all
    n1 = max(int(n / dt / cores * 2.0), 10000)
    t = time.perf_counter()
    st1 = O.rollout_stats_only(SEED, 0, n1, mode)
    dt1 = time.perf_counter() - t
    O.use_all_threads()
    out["single_core_steps_per_s"] = float(st1[8]) / dt1
    return out


def run_reference(args, rank):
    """Reference arm: the reference's CPU rule engine (C port, all host threads) on the same workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = O.use_all_threads()          # torchrun exports OMP_NUM_THREADS=1; the arm uses every host core
    t = time.perf_counter()
    O.rollout_stats_only(SEED, 0, 20000, args.mode)
    rate = 20000 / max(time.perf_counter() - t, 1e-4)
    per_step = int(min(max(rate * 1.0, 20000), args.games))          # ~1 s of CPU work per step
    for i in range(args.warmup):
        O.rollout_stats_only(SEED, i * per_step, per_step, args.mode)
    t0 = time.perf_counter()
    steps = 0
    for i in range(args.steps):
        st = O.rollout_stats_only(SEED, (args.warmup + i) * per_step, per_step, args.mode)
        steps += int(st[8])
    dt = time.perf_counter() - t0
    v = steps / dt
    sample = "%d deals per step (of the %d-deal workload), %d steps" % (per_step, args.games, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.mode, args.games), "games_per_gpu": args.games, "mode": args.mode,
                   "seed": hex(SEED)},
        "deals_per_sec": per_step * args.steps / dt,
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of the reference rule engine (oracle/tarok_oracle.c + synth.c, OpenMP); the Python reference "
                "itself cannot travel to the GPU box (probe in the build container: ~0.3k deals/s/core, BASELINE.md)",
    }))


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tarok_b200.env import TarokEnv, pack_records, MODE_AUCTION_UNIFORM, S_STEPS, S_FINISHED, S_ERRORS

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from tarok_b200.dist import bind_to_gpu_numa
        bind_to_gpu_numa(local_rank)              # pinned e2e buffers on the GPU's NUMA node
        dist.init_process_group("nccl", device_id=dev)
    n, mode, total = args.games, args.mode, args.games * world
    env = TarokEnv(n, seed=SEED, device=local_rank)
    env.set_materialise(False)      # the rollout's outputs are scores + statistics (Tarok.rezultati); piles stay in the trick log
    auction = mode in (17, 18)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    stats_ring = torch.zeros((args.warmup + args.steps + 1, 32), dtype=torch.int64, device=dev)
    pending = []
    step_events = []

    iter_events = []

    def rollout(i, timed, last=False):
        gid0 = i * total + rank * n
        if not args.no_flush:
            flush.zero_()                                                  # L2 flush between iterations (not timed)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        env.setup_synth(mode, gid0)                                        # deal + contract + exchange, one launch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step_random(48)
        b.record()
        env.score()
        stats_ring[i].copy_(env.stats_dev)
        if world > 1:
            # the one collective: returns/statistics, 256 B over NCCL/NVLink; asynchronous so that the next
            # rollout's deal overlaps it (waited for before the timed region closes)
            pending.append(dist.all_reduce(stats_ring[i], async_op=True))
        if last:
            for w in pending:                                              # every all-reduce lands inside a timed interval
                w.wait()
        i1.record()
        if timed:
            step_events.append((a, b))
            iter_events.append((i0, i1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        rollout(i, False)
    barrier()
    env.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = env.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(args.steps):
        rollout(args.warmup + i, True, last=(i == args.steps - 1))
    t1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    # K steps, each timed by its own CUDA-event pair on the launching stream (the L2-flush writes run between the pairs)
    region_ms = t0.elapsed_time(t1)
    ms = torch.tensor([sum(x.elapsed_time(y) for x, y in iter_events)], dtype=torch.float64, device=dev)
    step_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in step_events) / (48 * len(step_events))],
                           dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    ms, step_ms = float(ms.item()), float(step_ms.item())
    st = stats_ring[args.warmup + args.steps - 1].cpu().numpy()
    launches = env.launches - launches0
    env_steps, deals, errors = int(st[S_STEPS]), int(st[S_FINISHED]), int(st[S_ERRORS])
    value = env_steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_step<random>) over the timed region
    peak, peak_src = hbm_peak()
    live_per_launch = env_steps / world / (48 * args.steps)               # live games one launch advances
    achieved = BYTES_PER_ENV_STEP * live_per_launch / (step_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "step_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if int(tj.get("games", 0)) == n:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "k_step<random>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * live_per_launch,
                "avg_launch_us": step_ms * 1e3,
                "step_kernel_share_of_rollout": step_ms * 48 * args.steps / ms,
                "note": "achieved = SURVEY 8(d)'s 64 B/env-step x live games / event-timed launch; the kernel itself moves fewer DRAM "
                        "bytes than that figure (`traffic`, ncu) and at this batch size its 50 MB working set stays in the 126 MB L2 "
                        "between launches, so frac can exceed 1; roofline_large is the same kernel on a state 8x larger than L2"}

    # ---- fused rollout (state in registers; not HBM-bound) -- informational
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.rollout(mode, first_game_id=0, fused=True)
    torch.cuda.synchronize()
    fa.record()
    for i in range(5):
        env.rollout(mode, first_game_id=(i + 1) * total + rank * n, fused=True)
    fb.record()
    torch.cuda.synchronize()
    fused_ms = fa.elapsed_time(fb) / 5

    # ---- end to end through the host-buffer C-ABI entry
    env.deal(rank * n)
    perm_h = torch.empty((n, 54), dtype=torch.uint8).pin_memory()
    perm_h.copy_(env.export_perm())
    rng = np.random.default_rng(1234 + rank)
    c_h = torch.from_numpy(rng.integers(1, 4, n, dtype=np.uint8)).pin_memory()
    d_h = torch.from_numpy(rng.integers(0, 4, n, dtype=np.uint8)).pin_memory()
    k_h = torch.from_numpy(rng.integers(0, 4, n, dtype=np.uint8)).pin_memory()
    sc_h = torch.empty((n, 4), dtype=torch.int16).pin_memory()
    st_h = torch.zeros(32, dtype=torch.int64).pin_memory()
    rec_h, _ = pack_records(perm_h, c_h, d_h, k_h)           # the same deals + contracts as 24-byte records (pinned)
    def e2e_run(fused, records=False):
        def once(i):
            if records:
                env.rollout_records(rec_h, sc_h, st_h, first_game_id=i * total + rank * n)
            else:
                env.rollout_host(perm_h, c_h, d_h, k_h, sc_h, st_h, first_game_id=i * total + rank * n, fused=fused)
            torch.cuda.current_stream().synchronize()                     # the host reads the result
            return int(st_h[S_STEPS])
        for i in range(args.warmup):
            once(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cnt = 0
        e0.record()
        for i in range(args.steps):
            cnt += once(args.warmup + i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        c = torch.tensor([cnt], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c)
        return float(c.item()) / (float(t.item()) * 1e-3), float(t.item()) / args.steps

    e2e_value, e2e_ms = e2e_run(True)          # fused kernel behind an 8-chunk upload/compute/download pipeline
    e2e_sw_value, e2e_sw_ms = e2e_run(False)   # stepwise kernels, serial upload -> 52 launches -> download
    e2e_rec_value, e2e_rec_ms = e2e_run(True, records=True)   # same pipeline, 24-byte deal records instead of 57-byte rows

    # ---- HBM-bound regime: the same step kernel on a state 8x larger than L2 (informational)
    big = None
    if rank == 0 and not args.no_large:
        try:
            nb = 8 << 20
            eb = TarokEnv(nb, seed=SEED, device=local_rank)
            eb.deal(0); eb.force_contract_synth(mode) if not auction else eb.auction_synth(mode)
            eb.exchange_synth(mode == MODE_AUCTION_UNIFORM)
            eb.step_random(8)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); eb.step_random(36); b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 36 * 1e3
            live = int(eb.live().sum().item())
            gbs = BYTES_PER_ENV_STEP * live / (us * 1e-6) / 1e9
            big = {"games": nb, "state_bytes": nb * 104, "avg_launch_us": us, "achieved": gbs, "frac": gbs / peak, "unit": "GB/s"}
            eb.close()
        except Exception as ex:    # out of memory on a shared box etc.
            big = {"error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_port_rate(mode, 12.0)
        cpu = {"value": r["steps_per_s"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
               "sample": "%d deals of the same workload (%.1f s, OpenMP over %d threads)" % (r["deals"], r["seconds"], r["cores"]),
               "deals_per_sec": r["deals_per_s"], "single_core_value": r["single_core_steps_per_s"]}
        try:    # the Python reference itself, timed where it exists (build container; static file, not measured here)
            rj = json.load(open(os.path.join(ROOT, "profiles", "r01", "reference_cpu_container.json")))
            cpu["python_reference_in_build_container"] = [
                {k: x[k] for k in ("workload", "processes", "deals", "deals_per_sec", "env_steps_per_sec")} for x in rj["results"]]
        except Exception:
            pass

    if rank == 0:
        out = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(mode, n),
                       "games_per_gpu": n, "mode": mode, "seed": hex(SEED),
                       "outputs": "per-deal scores (int16 x4) + the all-reduced statistics vector = Tarok.rezultati; won-card piles stay "
                                  "in the 4-byte-per-trick log (TAROK_OPT_MATERIALISE=0), as Tarok.paralel_start never returns them",
                       "l2": ("no flush: every iteration regenerates and revisits its own %d MB state (> 126 MB L2)" % (n * 152 >> 20))
                             if args.no_flush else
                             ("160 MiB flush write between iterations, outside the per-iteration event pairs (and every iteration "
                              "regenerates its own %d MB state, larger than L2); within one iteration the 48 play_steps revisit "
                              "that state as the workload prescribes" % (n * 152 >> 20)),
                       "timing": "each of the K steps bracketed by its own CUDA-event pair; ms_per_step = their mean, max over ranks; "
                                 "region_ms_incl_flush = first event to last event including the flush writes"},
            "region_ms_incl_flush": region_ms,
            "deals_per_sec": deals / (ms * 1e-3), "env_steps": env_steps, "deals": deals, "error_games": errors,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": world * n * 57,
                    "d2h_bytes_per_step": world * (n * 8 + 256), "ms_per_step": e2e_ms,
                    "api": "tarok_rollout_host (TarokEnv.rollout_host): pinned host deals+contracts in, scores+stats out; fused "
                           "kernel behind an 8-chunk upload/compute/download pipeline",
                    "stepwise_kernels": {"value": e2e_sw_value, "ms_per_step": e2e_sw_ms},
                    "deal_records": {"value": e2e_rec_value, "ms_per_step": e2e_rec_ms, "h2d_bytes_per_step": world * n * 24,
                                     "api": "tarok_rollout_records: the same deals and contracts serialised as 24-byte records "
                                            "(packed on the host before the timed region, like the rows are dealt before it)"}},
            "gpu_launches": launches * world,
            "roofline": roofline, "roofline_large": big,
            "fused_rollout": {"ms_per_rollout": fused_ms, "env_steps_per_sec_per_gpu": env_steps / args.steps / world / (fused_ms * 1e-3),
                              "note": "one kernel, state in registers; ALU-bound, no HBM fraction claimed"},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    env.close()
    if world > 1:
        dist.destroy_process_group()


def run_config4(args, local_rank):
    """BASELINE config 4: policy-net forward + GPU env step, 65,536 envs per GPU (informational; one JSON line)."""
    import torch
    from tarok_b200.samoigra import Samoigra
    torch.cuda.set_device(local_rank)
    n = args.games if args.games != (1 << 20) else 65536
    s = Samoigra(n, device=local_rank, seed=SEED, random_card=0.05)
    for i in range(2):
        s.odigraj(i * n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps, parts = 0, {}
    for i in range(args.steps):
        st, ms = s.odigraj((2 + i) * n, meri=True)
        steps += int(st[19])
        for k, v in ms.items():
            parts[k] = parts.get(k, 0.0) + v
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "env_steps_per_sec", "value": steps / dt, "unit": "env-steps/s", "n_gpus": 1, "steps": args.steps,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "dtype": "u64 env / fp32 nets", "data": "synthetic",
                      "config": {"workload": "config 4: self-play, restated policy nets (random init) forward + GPU env step, %d envs" % n},
                      "device_ms_per_rollout": {k: v / args.steps for k, v in parts.items()},
                      "note": "env = deal/auction/exchange/step/score kernels; obs = obs_shape/obs_expand/bucketing; forward = the "
                              "reference-architecture nets (cuDNN LSTM, out of scope to accelerate); select = action-selection kernels"}))
    s.zapri()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--games", type=int, default=1 << 20, help="concurrent deals per GPU")
    ap.add_argument("--mode", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-large", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2-flush write between iterations (the 159 MB state is larger than L2)")
    ap.add_argument("--config", type=int, default=2, help="2 (default, the headline workload) or 4 (neural self-play)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config == 4:
        if rank == 0:
            args.steps = min(args.steps, 5) if args.steps == 100 else args.steps
            run_config4(args, local_rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (one process per GPU)" % args.gpus, file=sys.stderr)
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
