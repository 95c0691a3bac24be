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
    # size a step so that warmup+steps stay within a couple of minutes: probe first
    probe_steps, probe_s = O.bench_rollout(version, envs, 4, view, threads=cores)
    cores = O.max_threads()
    rate = probe_steps / max(probe_s, 1e-9)
    per_step_T = max(1, int(rate * 0.05 / envs))  # about 50 ms of CPU work per "step"
    total = 0
    t_all = 0.0
    for i in range(warmup + steps):
        n, s = O.bench_rollout(version, envs, per_step_T, view, seed=9981 + i, threads=cores)
        if i >= warmup:
            total += n
            t_all += s
    return total / t_all, cores, per_step_T, t_all


def bench_ppo(args, rank, world, local_rank):
    """--workload ppo (BASELINE configs[3]): full PPO rollout-update loop on Twoarmy-17x17-v4,
    `--ppo-envs` envs GLOBAL x `--ppo-horizon` steps per update, GAE/advantages on the GPU, NCCL
    gradient all-reduce.  A step = one rollout + one PPO.update (K_epochs x minibatches).  Metric:
    env frames consumed per second by the whole loop."""
    import importlib
    metric, unit = "ppo_frames_per_sec", "frames/s"
    n_global, T = args.ppo_envs, args.ppo_horizon
    config = {"workload": f"PPO rollout+update loop, MiniGrid-twoarmy-17x17-v{args.version}, {n_global} envs x {T}-step horizon "
                          f"(BASELINE configs[3])", "envs_global": n_global, "horizon": T, "K_epochs": args.ppo_epochs,
              "minibatch_per_rank": args.ppo_minibatch, "nets": "Net_PPO_actor/critic (TINet), bf16 autocast, channels_last",
              "parallelism": f"env-sharded dp{world}, NCCL gradient all-reduce"}
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's loop shape on the host cores: ONE env, select_action at B=1, buffer 2048,
        # update K_epochs x minibatches of 128 (soa/train_ppo.py:99-160) -- env via the C port, nets
        # via this repo's PyTorch mirror of all_net.py on the CPU (the Python reference cannot travel)
        import numpy as np
        import torch
        from oracle import oracle as O
        sys.path.insert(0, ROOT)
        import twoarmy_b200 as pkg
        P = importlib.import_module(pkg.__name__ + ".ppo")
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        torch.manual_seed(0)
        agent = P.PPO(device="cpu", autocast=False)
        cap = args.ppo_ref_frames
        ora = O.OracleBatch(args.version, 1, 17, seed=9981)
        ora.reset()
        lut = np.array(P.MATRIX_LUT, np.float32)
        t0 = time.perf_counter()
        f = ora.matrix().astype(np.float32)
        sm = np.repeat(f[:, None], 5, 1)
        ss = np.tile(np.array([[15.0, 3.0]], np.float32), (1, 5, 1))
        g = torch.tensor([[2.0, 14.0]])
        rec = {k: [] for k in ("s", "p", "a", "r", "a_logp")}
        amap = np.array([0, 1, 2, 3, 6], np.int32)
        for i in range(cap):
            a, lp = agent.select_action(torch.from_numpy(sm), torch.from_numpy(ss), g)
            out = ora.step(amap[a.numpy()], None, autoreset=False)
            e = ora.envs[0]
            sm = np.concatenate([sm[:, 1:], ora.matrix().astype(np.float32)[:, None]], 1)
            ss = np.concatenate([ss[:, 1:], np.array([[[float(e["ay"]), float(e["ax"])]]], np.float32)], 1)
            rec["s"].append(sm[0].copy()); rec["p"].append(ss[0].copy()); rec["a"].append([int(a)])
            rec["r"].append([float(out["reward"][0])]); rec["a_logp"].append([float(lp)])
            if out["terminated"][0] or out["truncated"][0]:
                ora.reset()
                f = ora.matrix().astype(np.float32)
                sm = np.repeat(f[:, None], 5, 1)
                ss = np.tile(np.array([[15.0, 3.0]], np.float32), (1, 5, 1))
        buf = {k: torch.tensor(np.array(v)) for k, v in rec.items()}
        buf["g"] = g.repeat(cap, 1)
        agent.update(buf)
        secs = time.perf_counter() - t0
        v = cap / secs
        line = {"metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                 "sample": f"one buffer of {cap} frames: single env (C port) + select_action at B=1 + "
                                           f"PPO.update K=10 x minibatch 128 on {cores} CPU threads"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import twoarmy_b200 as pkg
    P = importlib.import_module(pkg.__name__ + ".ppo")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    n_local = n_global // world
    torch.manual_seed(9981)
    if args.ppo_predictor:
        agent = importlib.import_module(pkg.__name__ + ".predictor").ppo_predictor(device=dev)
        config["workload"] = config["workload"].replace("PPO rollout+update loop", "PPO + frame-predictor rollout+update loop").replace("configs[3]", "configs[4]")
        config["nets"] = "Net_PPO_Predictor_actor/critic (8-channel TINet) + frozen Net_Encoder / LSTM(1024x3) / Net_Decoder, bf16 autocast"
    else:
        agent = P.PPO(device=dev)
    agent.broadcast_parameters()
    torch.manual_seed(9981 + 1000 * (rank + 1))
    env = pkg.TwoarmyVecEnv(args.version, n_local, 17, device=dev, seed=9981, env_id0=rank * n_local, autoreset=False)
    roll = P.VecRollout(env, agent, T)

    def one_step():
        buf = roll.collect()
        return agent.update(buf.flat(), minibatch=args.ppo_minibatch, epochs=args.ppo_epochs)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.ppo_steps, args.ppo_warmup
    for _ in range(warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pkg.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        losses = one_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pkg.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # phase split (one more step, timed per phase)
    t0 = time.perf_counter(); buf = roll.collect(); torch.cuda.synchronize(); t1 = time.perf_counter()
    agent.update(buf.flat(), minibatch=args.ppo_minibatch, epochs=args.ppo_epochs); torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0:
        frames = steps * T * n_local * world
        fwd_flop = 47.53e6  # per sample per net
        upd_flop = 3 * 2 * fwd_flop * T * n_local * args.ppo_epochs + 2 * 2 * fwd_flop * T * n_local  # update + the two critic passes
        roll_flop = fwd_flop * T * n_local  # actor forward per frame
        tf = 1353.2
        try:
            tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            pass
        achieved_tf = (upd_flop + roll_flop) / ((ms / steps) / 1e3) / 1e12
        line = {"metric": metric, "value": frames / (ms / 1e3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config,
                "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": tf, "unit": "TFLOP/s", "frac": achieved_tf / tf,
                             "traffic": None, "kernel": "cuDNN/cuBLAS conv+GEMM of TINet (library calls, per north_star)",
                             "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"},
                "phases": {"rollout_s": t1 - t0, "update_s": t2 - t1},
                "losses": {"action": losses[0], "value": losses[1]},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


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
    ap.add_argument("--workload", default="step", choices=["step", "ppo", "aux"])
    ap.add_argument("--ppo-envs", type=int, default=16384, help="GLOBAL env count of the PPO workload")
    ap.add_argument("--ppo-horizon", type=int, default=128)
    ap.add_argument("--ppo-minibatch", type=int, default=4096)
    ap.add_argument("--ppo-epochs", type=int, default=10)
    ap.add_argument("--ppo-steps", type=int, default=2)
    ap.add_argument("--ppo-warmup", type=int, default=1)
    ap.add_argument("--ppo-ref-frames", type=int, default=2048)
    ap.add_argument("--ppo-predictor", action="store_true",
                    help="BASELINE configs[4]: actor / critic also see 4 frames predicted by the frozen Encoder-LSTM-Decoder")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "ppo":
        return bench_ppo(args, rank, world, local_rank)
    if args.workload == "aux":  # featuriser / advantage kernels against the HBM roofline (rank 0, one GPU)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import aux_kernels_bench
            aux_kernels_bench.main()
        return
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

        def __init__(self, view, batches, id_base, n=n):
            self.view, self.B, self.n = view, batches, n
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

        def capture(self, streams=1):
            """The launch-bound inner loop (B launches, one per batch) as one CUDA graph.  With
            streams=2 the batches alternate between two captured branches, so launches on
            independent batches may overlap (secondary measurement only)."""
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(self.B):
                    self.step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                if streams == 1:
                    for i in range(self.B):
                        self.step(i)
                else:
                    main = torch.cuda.current_stream()
                    branch = torch.cuda.Stream(device=dev)
                    branch.wait_stream(main)
                    for i in range(self.B):
                        with torch.cuda.stream(branch if i % 2 else main):
                            self.step(i)
                    main.wait_stream(branch)

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
        # (a2) single-step launches again, independent batches alternating between two streams
        wl.capture(streams=2)
        ms2, _ = wl.timed(args.steps, args.warmup)
        extra["two_streams"] = {"us_per_launch": ms2 * 1e3 / args.steps, "value": args.steps * n * world / (ms2 / 1e3), "unit": UNIT,
                                "note": "same launches, the 8 independent batches alternate between 2 streams so the serial "
                                        "head/tail of one launch overlaps the observation stores of the other"}
        wl.capture(streams=1)
        # (b) the north-star 7x7x3 view (agent_view_size=7), single-step launches
        if V != 7:
            wl7 = Workload(7, B, 1 << 40)
            wl7.capture()
            ms7, _ = wl7.timed(args.steps, args.warmup)
            extra["view7"] = {"us_per_launch": ms7 * 1e3 / args.steps, "value": args.steps * n * world / (ms7 / 1e3),
                              "unit": UNIT, "alg_bytes_per_env_step": alg_bytes(7),
                              "achieved_gbs": alg_bytes(7) * n / (ms7 / args.steps / 1e3) / 1e9}
            wl7.close()
        # (c) the same single-step launches at larger launch sizes: the 65536-env launch of configs[2] is 15 us of
        #     which several are the serial head / tail of one wave; more envs per launch amortise it
        sweep = []
        for view_s, mult in ((V, 4), (7, 16)):
            wls = Workload(view_s, 2, 1 << 41, n=n * mult)
            wls.capture()
            k = max(16, args.steps // (2 * mult))
            mss, _ = wls.timed(k, max(3, args.warmup // mult))
            sweep.append({"view": view_s, "envs_per_launch": n * mult, "us_per_launch": mss * 1e3 / k,
                          "value": k * n * mult * world / (mss / 1e3), "unit": UNIT,
                          "achieved_gbs": alg_bytes(view_s) * n * mult / (mss / k / 1e3) / 1e9})
            wls.close()
        extra["launch_size"] = sweep

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
        cn = 4096
        ps, pt = O.bench_rollout(args.version, cn, 8, V, threads=cores)
        cores = O.max_threads()
        T = max(8, int(ps / pt * args.cpu_seconds / cn))
        cs, ct = O.bench_rollout(args.version, cn, T, V, threads=cores)
        line["cpu_baseline"] = {"value": cs / ct, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cn} envs x {T} steps ({ct:.1f} s) of the same workload through "
                                          f"oracle/twoarmy_oracle.c (C port of the reference path, OpenMP over envs)"}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
