#!/usr/bin/env python
"""bench.py -- env steps/s (including observation generation) of the batched Twoarmy hot path.

    python bench.py --gpus 1 --steps 2000 --warmup 200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

Workload (BASELINE.json configs[2]): MiniGrid-twoarmy-17x17-v4, 65536 envs per GPU,
agent_view_size 17 (the reference default), uniform actions over {0,1,2,3,6}, Philox draws,
autoreset.  A "step" is one fused step+gen_obs launch over all of a rank's envs.  To keep the
timed launches out of L2, each rank rotates over `--batches` independent env batches (state
~21 MB each) and writes each step's observations into its own slot of a ring (57 MB per slot):
the working set (batches x 78 MB) exceeds the 126 MB L2.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic bytes per env-step, fixed by SURVEY.md section 8(d)
ALG_BYTES = {17: 1259, 7: 539}
METRIC = "env_steps_per_sec_incl_obs"
UNIT = "env-steps/s"


def alg_bytes(view):
    # action 1 + state 2*32 + grid 289 + dirty cells 32 + obs 3*V*V + reward 4 + term 1 + trunc 1
    return ALG_BYTES.get(view, 1 + 64 + 289 + 32 + 3 * view * view + 6)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(version, view, envs, steps, warmup, threads=None):
    """The CPU arm: the oracle's port of the path, all host threads, bounded sample."""
    from oracle import oracle as O
    cores = threads or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    # size a step so that warmup+steps stay within a couple of minutes: probe first
    probe_steps, probe_s = O.bench_rollout(version, envs, 4, view)
    rate = probe_steps / max(probe_s, 1e-9)
    per_step_T = max(1, int(rate * 0.05 / envs))  # about 50 ms of CPU work per "step"
    total = 0
    t_all = 0.0
    for i in range(warmup + steps):
        n, s = O.bench_rollout(version, envs, per_step_T, view, seed=9981 + i)
        if i >= warmup:
            total += n
            t_all += s
    return total / t_all, cores, per_step_T, t_all


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--view", type=int, default=17)
    ap.add_argument("--version", type=int, default=4)
    ap.add_argument("--batches", type=int, default=8, help="independent env batches rotated per rank (L2 busting)")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the rollout / 7x7-view secondary measurements")
    ap.add_argument("--rollout-T", type=int, default=16)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"MiniGrid-twoarmy-17x17-v{args.version} batched step+gen_obs, {args.envs} envs per GPU, "
                f"view {args.view}x{args.view}x3, random actions, Philox draws, autoreset (BASELINE configs[2])")
    config = {"workload": workload, "envs_per_gpu": args.envs, "view": args.view, "env_version": args.version,
              "parallelism": f"env-sharded x{world}, no data-path collective",
              "l2": f"{args.batches} rotating env batches, each with its own obs slot: working set "
                    f"{args.batches * args.envs * (3 * args.view * args.view + 230) / 1e6:.0f} MB > 126 MB L2",
              "launch": "ta_step launches replayed from a CUDA graph (one kernel node per batch)"}

    if args.impl == "reference":
        if rank != 0:
            return
        envs = 4096
        v, cores, per_step_T, secs = cpu_reference_run(args.version, args.view, envs, min(args.steps, 200), min(args.warmup, 5))
        sample = (f"{envs} envs x {per_step_T} steps per bench step, C port of the reference path "
                  f"(oracle/twoarmy_oracle.c, OpenMP over envs); the Python reference itself is not shippable to the box")
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": min(args.steps, 200),
                "warmup": min(args.warmup, 5), "ms_per_step": 1e3 * secs / max(1, min(args.steps, 200)),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import twoarmy_b200 as pkg

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    n, V, B = args.envs, args.view, args.batches
    amap = torch.tensor([0, 1, 2, 3, 6], dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    class Workload:
        """B independent env batches of n envs (rotated so that the timed launches miss L2), one
        obs/reward/flag slot per batch, pre-sampled uniform actions.  step(i) = ONE fused
        step+gen_obs launch (ta_step) on batch i % B."""

        def __init__(self, view, batches, id_base):
            self.view, self.B = view, batches
            self.envs = [pkg.TwoarmyVecEnv(args.version, n, view, device=dev, seed=9981,
                                           env_id0=id_base + (rank * batches + b) * n) for b in range(batches)]
            for e in self.envs:
                e.reset()
            self.actions = amap[torch.randint(0, 5, (batches, n), generator=g, device=dev)].contiguous()
            self.outs = [dict(obs=torch.empty((n, view, view, 3), dtype=torch.uint8, device=dev),
                              reward=torch.empty(n, dtype=torch.float32, device=dev),
                              terminated=torch.empty(n, dtype=torch.uint8, device=dev),
                              truncated=torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(batches)]
            self.graph = None

        def step(self, i):
            b = i % self.B
            self.envs[b].step(self.actions[b], out=self.outs[b])

        def capture(self):
            """The launch-bound inner loop (B launches, one per batch) as one CUDA graph."""
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(self.B):
                    self.step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                for i in range(self.B):
                    self.step(i)

        def run(self, k):
            """k launches: whole graph replays, the remainder as single launches. Returns the
            number of step kernels enqueued."""
            reps, rem = divmod(k, self.B)
            for _ in range(reps):
                self.graph.replay()
            for i in range(rem):
                self.step(i)
            return k

        def timed(self, steps, warmup):
            self.run(warmup)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            launches = self.run(steps)
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
            if dist:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, launches

        def close(self):
            self.graph = None
            for e in self.envs:
                e.close()

    wl = Workload(V, B, 0)
    wl.capture()
    wl.run(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = wl.timed(args.steps, args.warmup)
    # a short run gives nvidia-smi nothing to sample: keep the GPU under the same load a bit longer
    clocks = None
    if rank == 0:
        t_end = time.time() + max(0.0, 0.7 - ms / 1e3)
        while time.time() < t_end:
            wl.run(25 * B)
            torch.cuda.synchronize()
        clocks = sampler.stop()
    total_steps = args.steps * n * world
    value = total_steps / (ms / 1e3)

    # ---- secondary measurements (same timing rules; reported under "extra") ----------------
    extra = {}
    if not args.no_extra:
        # (a) T-step rollouts in ONE launch (ta_rollout: pre-sampled actions, state stays on chip)
        T = args.rollout_T
        racts = amap[torch.randint(0, 5, (T, n), generator=g, device=dev)].contiguous()
        robs = [torch.empty((T, n, V, V, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        rrew = torch.empty((T, n), dtype=torch.float32, device=dev)
        rte = torch.empty((T, n), dtype=torch.uint8, device=dev)
        rtr = torch.empty((T, n), dtype=torch.uint8, device=dev)
        import ctypes as C
        L = pkg._capi.lib()

        def rollout(i):
            e = wl.envs[i % B]
            pkg._capi.check(L.ta_rollout(e._h, C.c_void_p(racts.data_ptr()), 1, T, C.c_void_p(robs[i % 2].data_ptr()),
                                         C.c_void_p(rrew.data_ptr()), C.c_void_p(rte.data_ptr()), C.c_void_p(rtr.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ta_rollout")

        for i in range(3):
            rollout(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        R = max(4, args.steps // (4 * T))
        ev0.record()
        for i in range(R):
            rollout(i)
        ev1.record()
        barrier()
        rms = ev0.elapsed_time(ev1)
        if dist:
            t = torch.tensor([rms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rms = float(t.item())
        extra["rollout"] = {"call": "ta_rollout", "T": T, "launches": R, "us_per_env_step_batch": rms * 1e3 / (R * T),
                            "value": R * T * n * world / (rms / 1e3), "unit": UNIT,
                            "hbm_bytes_per_env_step": 3 * V * V + 1 + 6 + (2 * 112) / T}
        del robs
        # (b) the north-star 7x7x3 view (agent_view_size=7), single-step launches
        if V != 7:
            wl7 = Workload(7, B, 1 << 40)
            wl7.capture()
            ms7, _ = wl7.timed(args.steps, args.warmup)
            extra["view7"] = {"us_per_launch": ms7 * 1e3 / args.steps, "value": args.steps * n * world / (ms7 / 1e3),
                              "unit": UNIT, "alg_bytes_per_env_step": alg_bytes(7),
                              "achieved_gbs": alg_bytes(7) * n / (ms7 / args.steps / 1e3) / 1e9}
            wl7.close()

    # ---- end to end through the public host-buffer call ---------------------------------
    e2e_env = wl.envs[0]
    actions = wl.actions
    h_act = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_act.copy_(actions[0].cpu())
    h_obs = torch.empty((n, V, V, 3), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_te = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_tr = torch.empty(n, dtype=torch.uint8).pin_memory()
    na, no, nr, nte, ntr = (x.numpy() for x in (h_act, h_obs, h_rew, h_te, h_tr))
    for _ in range(3):
        e2e_env.step_host(na, no, nr, nte, ntr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_env.step_host(na, no, nr, nte, ntr)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = args.e2e_steps * n * world / e2e_s
    launches_e2e = args.e2e_steps

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_launch_ms = ms / args.steps
    achieved = alg_bytes(V) * n / (per_launch_ms / 1e3) / 1e9
    # bytes this layout really has to move per env-step (2-bit packed grid, 80 B each way)
    layout_bytes = 1 + 2 * (80 + 32) + 3 * V * V + 6
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"step_obs_v{V}_n{n}")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": f"step_obs_kernel<{V}>", "alg_bytes_per_env_step": alg_bytes(V),
                     "peak_source": peak_src, "launch_us": per_launch_ms * 1e3,
                     "layout_bytes_per_env_step": layout_bytes,
                     "achieved_layout_bytes": layout_bytes * n / (per_launch_ms / 1e3) / 1e9,
                     "timing": "CUDA events on the launching stream around K graph-replayed launches"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n),
                "d2h_bytes_per_step": int(n * (3 * V * V + 4 + 1 + 1)), "steps": args.e2e_steps,
                "call": "TwoarmyVecEnv.step_host -> ta_step_host (pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks, "extra": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        cn = 4096
        ps, pt = O.bench_rollout(args.version, cn, 8, V)
        T = max(8, int(ps / pt * args.cpu_seconds / cn))
        cs, ct = O.bench_rollout(args.version, cn, T, V)
        line["cpu_baseline"] = {"value": cs / ct, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cn} envs x {T} steps ({ct:.1f} s) of the same workload through "
                                          f"oracle/twoarmy_oracle.c (C port of the reference path, OpenMP over envs)"}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
