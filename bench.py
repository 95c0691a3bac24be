#!/usr/bin/env python
"""bench.py -- env steps/s (including observation generation) of the batched Twoarmy hot path,
plus the PPO rollout+update loop's frames/s (BASELINE.json's two-part metric) under `extra.ppo`.

    python bench.py --gpus 1 --steps 2000 --warmup 200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm: the Python reference itself on the host cores
    python bench.py --workload ppo ...        # only the PPO loop (same numbers as extra.ppo, as the headline)

Workload (BASELINE.json configs[2]): MiniGrid-twoarmy-17x17-v4, 65536 envs per GPU,
agent_view_size 17 (the reference default), uniform actions over {0,1,2,3,6}, Philox draws,
autoreset.  A "step" is one fused step+gen_obs launch over all of a rank's envs.  To keep the
timed launches out of L2, each rank rotates over `--batches` independent env batches (state
~21 MB each) and writes each step's observations into its own slot of a ring (57 MB per slot):
the working set (batches x 78 MB) exceeds the 126 MB L2.  The K timed launches are ONE CUDA graph
of exactly K kernel nodes (no eager launch inside the timed region).

extra.ppo (BASELINE.json configs[3]): 16384 envs GLOBAL x 128-step horizon, the reference's
K_epochs = 10, 4096-sample minibatches per rank, GAE/advantages on the GPU, NCCL gradient all-reduce
when N > 1; two cheap warm-up iterations (K_epochs = 1), one timed iteration.

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


def measured_peak(key="hbm_gbs", fallback=6650.0):
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))[key]), f"measured (MEASURED_PEAKS.json {key})"
        except Exception:
            pass
    return fallback, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------
# CPU legs: the Python reference itself (baseline/ref_arm.py) and the C port of the path (oracle/)
# ---------------------------------------------------------------------------------------------
def ref_arm(mode, **kw):
    """baseline/ref_arm.py (the UNMODIFIED Python reference on the host cores) in a child process;
    returns its JSON object, or {"unavailable": why}."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_arm.py"), mode]
    for k, v in kw.items():
        cmd += ["--" + k.replace("_", "-"), str(v)]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        lines = [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
        if res.returncode != 0 or not lines:
            return {"unavailable": (res.stderr.strip().splitlines() or ["no output"])[-1][:300]}
        return json.loads(lines[-1])
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def port_rate(version, view, seconds, envs=4096):
    """The C/OpenMP port of the path (oracle/twoarmy_oracle.c) on all host threads, ~`seconds` of work."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    ps, pt = O.bench_rollout(version, envs, 8, view, threads=cores)
    cores = O.max_threads()
    T = max(8, int(ps / pt * seconds / envs))
    cs, ct = O.bench_rollout(version, envs, T, view, threads=cores)
    return {"value": cs / ct, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{envs} envs x {T} steps ({ct:.1f} s) through oracle/twoarmy_oracle.c (C port of the reference path, "
                      f"OpenMP over envs)"}


def step_cpu_baseline(version, view, rounds, warmup, port_seconds):
    """cpu_baseline of the step workload: the Python reference (P processes x one gym env each, C1 of
    BASELINE.md section 3) when its sources are on this machine, and the C port as a second, labelled number."""
    cores = os.cpu_count() or 1
    port = port_rate(version, view, port_seconds)
    r = ref_arm("c1", version=version, view=view, procs=cores, rounds=rounds, warmup=warmup)
    if "unavailable" in r:
        port["reference_unavailable"] = r["unavailable"]
        return port, None
    cb = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference",
          "sample": f"{cores} processes x one unmodified reference env each (gym.make('MiniGrid-twoarmy-17x17-v{version}', "
                    f"agent_view_size={view}).step incl. gen_obs, reset on done), {r['rounds']} rounds of {r['round_steps']} steps "
                    f"per process ({r['seconds']:.1f} s); {r['per_process_steps_per_s']:.0f} steps/s per process",
          "port": port}
    return cb, r


def reference_arm_step(args):
    """--impl reference, step workload (rank 0 only)."""
    steps, warmup = min(args.steps, 40), min(args.warmup, 5)
    cb, r = step_cpu_baseline(args.version, args.view, steps, warmup, port_seconds=3.0)
    workload = (f"MiniGrid-twoarmy-17x17-v{args.version} batched step+gen_obs, {args.envs} envs per GPU, "
                f"view {args.view}x{args.view}x3, random actions, Philox draws, autoreset (BASELINE configs[2])")
    config = {"workload": workload, "envs_per_gpu": args.envs, "view": args.view, "env_version": args.version}
    if r is not None:
        config["ran"] = (f"CPU arm: {r['procs']} processes x 1 reference env (the reference has no vector env), "
                         f"{r['round_steps']} env.step calls per process per bench step, same env id / view / action distribution")
        value, ms = r["value"], r["ms_per_round"]
    else:
        config["ran"] = "CPU arm: C port of the path, 4096 envs x T-step OpenMP rollouts (reference sources not on this machine)"
        value, ms = cb["value"], None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def reference_arm_ppo(args):
    """--impl reference --workload ppo: soa/train_ppo.py's loop with the reference's own PPO / Env_transact /
    Buffer_gridworld (baseline/ref_arm.py ppo) on the host cores."""
    cores = os.cpu_count() or 1
    r = ref_arm("ppo", version=args.version, frames=args.ppo_ref_frames, threads=cores)
    config = {"workload": f"PPO rollout+update loop, MiniGrid-twoarmy-17x17-v{args.version} (BASELINE configs[0] shape: ONE env, "
                          f"buffer of {args.ppo_ref_frames} frames, K_epochs 10 x minibatches of 128)"}
    if "unavailable" in r:
        print(json.dumps({"impl": "reference", "unavailable": r["unavailable"], "metric": "ppo_frames_per_sec", "config": config}))
        return
    v = r["value"]
    print(json.dumps({"metric": "ppo_frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                      "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": config, "impl": "reference",
                      "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                                       "sample": f"soa/train_ppo.py:99-160 verbatim with the reference's PPO / Env_transact / "
                                                 f"Buffer_gridworld: {r['frames']} frames + one PPO.update, {cores} torch threads "
                                                 f"({r['seconds']:.1f} s: env+render {r['rollout_env_s']:.1f}, select_action "
                                                 f"{r['select_action_s']:.1f}, update {r['update_s']:.1f})"},
                      "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------
# PPO rollout+update loop (BASELINE configs[3]; configs[4] with --ppo-predictor)
# ---------------------------------------------------------------------------------------------
def ppo_measure(args, rank, world, dev, dist, clock_index=None):
    """Two cheap warm-up iterations (K_epochs = 1) and `--ppo-steps` timed iterations of
    VecRollout.collect + PPO.update.  Returns the extra.ppo object on rank 0 (None elsewhere)."""
    import importlib
    import torch
    import twoarmy_b200 as pkg
    P = importlib.import_module(pkg.__name__ + ".ppo")
    n_global, T = args.ppo_envs, args.ppo_horizon
    n_local = n_global // world
    torch.manual_seed(9981)
    if args.ppo_predictor:
        agent = importlib.import_module(pkg.__name__ + ".predictor").ppo_predictor(device=dev)
    else:
        agent = P.PPO(device=dev)
    agent.broadcast_parameters()
    torch.manual_seed(9981 + 1000 * (rank + 1))
    env = pkg.TwoarmyVecEnv(args.version, n_local, 17, device=dev, seed=9981, env_id0=rank * n_local, autoreset=False)
    roll = P.VecRollout(env, agent, T)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(epochs):
        buf = roll.collect()
        return agent.update(buf.flat(), minibatch=args.ppo_minibatch, epochs=epochs)

    one_step(1)   # warm-up: allocator, cuDNN / cuBLAS plans, NCCL channels
    one_step(1)   # second warm-up: VecRollout captures its T-step loop into a CUDA graph on its second collect
    barrier()
    sampler = ClockSampler(clock_index) if (rank == 0 and clock_index is not None) else None
    if sampler:
        sampler.start()
    l0 = pkg.launch_count()
    steps = args.ppo_steps
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    evs[0].record()
    for i in range(steps):
        buf = roll.collect()
        evs[2 * i + 1].record()
        losses = agent.update(buf.flat(), minibatch=args.ppo_minibatch, epochs=args.ppo_epochs)
        evs[2 * i + 2].record()
    barrier()
    ms = evs[0].elapsed_time(evs[-1])
    roll_ms = sum(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(steps))
    upd_ms = sum(evs[2 * i + 1].elapsed_time(evs[2 * i + 2]) for i in range(steps))
    own_launches = pkg.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if dist:
        t = torch.tensor([ms, roll_ms, upd_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, roll_ms, upd_ms = (float(x) for x in t.tolist())
    probe = agent.probe_step(buf.flat(), minibatch=args.ppo_minibatch) if hasattr(agent, "probe_step") else {}
    env.close()
    if rank != 0:
        return None
    B_local = T * n_local
    opt_steps = args.ppo_epochs * ((B_local + args.ppo_minibatch - 1) // args.ppo_minibatch)
    frames = steps * T * n_local * world
    fwd_flop = 47.53e6  # per sample per net
    # per rank: update fwd+bwd of two nets over K epochs, the two critic passes, the actor forward per frame
    flop_rank = (3 * 2 * fwd_flop * args.ppo_epochs + 2 * fwd_flop + fwd_flop) * B_local
    tf, tf_src = measured_peak("bf16_tflops_sustained", 1353.2)
    achieved_tf = flop_rank * world / ((ms / steps) / 1e3) / 1e12
    out = {"metric": "ppo_frames_per_sec", "value": frames / (ms / 1e3), "unit": "frames/s", "n_gpus": world,
           "iterations_timed": steps, "ms_per_iteration": ms / steps, "scaling": "strong",
           "config": {"workload": f"PPO rollout+update loop, MiniGrid-twoarmy-17x17-v{args.version}, {n_global} envs GLOBAL x {T}-step "
                                  f"horizon (BASELINE configs[{4 if args.ppo_predictor else 3}])",
                      "envs_global": n_global, "envs_per_rank": n_local, "horizon": T, "K_epochs": args.ppo_epochs,
                      "minibatch_per_rank": args.ppo_minibatch, "minibatch_global": args.ppo_minibatch * world,
                      "optimizer_steps_per_iteration": opt_steps,
                      "parallelism": f"env-sharded dp{world}; per optimiser step one NCCL all-reduce per network, issued on that "
                                     f"network's stream right after its backward (overlaps the other network's backward)"},
           "phases_ms": {"rollout": roll_ms / steps, "update": upd_ms / steps},
           "ms_per_optimizer_step": upd_ms / steps / opt_steps,
           "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": tf * world, "unit": "TFLOP/s", "frac": achieved_tf / (tf * world),
                        "note": "nominal FLOPs of the reference network (47.53 MFLOP forward per sample per net)", "peak_source": tf_src},
           "losses": {"action": losses[0], "value": losses[1]}, "own_kernel_launches": int(own_launches), "clocks": clocks}
    out.update(probe)
    if getattr(agent, "last_replay_stats", None):
        out["graph_replayed_optimizer_steps"] = agent.last_replay_stats   # CUDA events around the replays of the last update()
    if getattr(agent, "last_update_phases", None):
        out["update_phases_ms"] = agent.last_update_phases   # TA_PPO_TIMING=1 only (synchronising marks: a diagnosis run, not a bench value)
    out["fused_step"] = bool(getattr(agent, "_fused", None))
    return out


# ---------------------------------------------------------------------------------------------
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
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the rollout / 7x7-view secondary measurements")
    ap.add_argument("--no-ppo", action="store_true", help="skip extra.ppo (the PPO rollout+update loop)")
    ap.add_argument("--no-ppo-predictor", action="store_true", help="skip extra.ppo_predictor (BASELINE configs[4] at 8192 envs x 32 steps)")
    ap.add_argument("--no-aux", action="store_true", help="skip extra.aux (featuriser / advantage kernels against the HBM roofline)")
    ap.add_argument("--rollout-T", type=int, default=16)
    ap.add_argument("--workload", default="step", choices=["step", "ppo", "aux"])
    ap.add_argument("--ppo-envs", type=int, default=16384, help="GLOBAL env count of the PPO workload")
    ap.add_argument("--ppo-horizon", type=int, default=128)
    ap.add_argument("--ppo-minibatch", type=int, default=4096, help="per rank")
    ap.add_argument("--ppo-epochs", type=int, default=10)
    ap.add_argument("--ppo-steps", type=int, default=1)
    ap.add_argument("--ppo-ref-frames", type=int, default=256)
    ap.add_argument("--ppo-predictor", action="store_true",
                    help="BASELINE configs[4]: actor / critic also see 4 frames predicted by the frozen Encoder-LSTM-Decoder")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            (reference_arm_ppo if args.workload == "ppo" else reference_arm_step)(args)
        return
    if args.workload == "aux":  # featuriser / advantage kernels against the HBM roofline (rank 0, one GPU)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import aux_kernels_bench
            aux_kernels_bench.main()
        return

    import torch
    import twoarmy_b200 as pkg

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL prints its version banner (NCCL_DEBUG=VERSION and up) on stdout: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "ppo":
        out = ppo_measure(args, rank, world, dev, dist, clock_index=local_rank)
        if rank == 0:
            line = {"metric": out["metric"], "value": out["value"], "unit": out["unit"], "n_gpus": world, "steps": args.ppo_steps,
                    "warmup": 1, "ms_per_step": out["ms_per_iteration"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": out["config"], "roofline": out["roofline"],
                    "gpu_launches": out["own_kernel_launches"], "clocks": out["clocks"], "extra": {"ppo": out}}
            print(json.dumps(line))
        if dist:
            dist.destroy_process_group()
        return

    workload = (f"MiniGrid-twoarmy-17x17-v{args.version} batched step+gen_obs, {args.envs} envs per GPU, "
                f"view {args.view}x{args.view}x3, random actions, Philox draws, autoreset (BASELINE configs[2])")
    config = {"workload": workload, "envs_per_gpu": args.envs, "view": args.view, "env_version": args.version,
              "parallelism": f"env-sharded x{world}, no data-path collective",
              "l2": f"{args.batches} rotating env batches, each with its own obs slot: working set "
                    f"{args.batches * args.envs * (3 * args.view * args.view + 230) / 1e6:.0f} MB > 126 MB L2",
              "launch": "the K timed ta_step launches are one CUDA graph of K kernel nodes (batch i % B)"}

    n, V, B = args.envs, args.view, args.batches
    amap = torch.tensor([0, 1, 2, 3, 6], dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    GRAPH_MAX = 2048   # kernel nodes per captured graph

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
            self.graphs, self.streams = {}, 1
            side = torch.cuda.Stream(device=dev)   # one eager pass (allocations, function attributes) before any capture
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(self.B):
                    self.step(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

        def step(self, i):
            b = i % self.B
            self.envs[b].step(self.actions[b], out=self.outs[b])

        def graph_for(self, k):
            """A CUDA graph of exactly k launches (batch i % B for launch i).  With streams == 2 the launches
            alternate between two captured branches, so launches on independent batches may overlap
            (secondary measurement only)."""
            key = (k, self.streams)
            if key not in self.graphs:
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    if self.streams == 1:
                        for i in range(k):
                            self.step(i)
                    else:
                        main = torch.cuda.current_stream()
                        branch = torch.cuda.Stream(device=dev)
                        branch.wait_stream(main)
                        for i in range(k):
                            with torch.cuda.stream(branch if i % 2 else main):
                                self.step(i)
                        main.wait_stream(branch)
                self.graphs[key] = gr
            return self.graphs[key]

        def plan(self, k):
            """The graph replays that make up k launches -- never an eager launch."""
            reps, rem = divmod(k, GRAPH_MAX)
            return [self.graph_for(GRAPH_MAX)] * reps + ([self.graph_for(rem)] if rem else [])

        def run(self, k):
            for gr in self.plan(k):
                gr.replay()
            return k

        def timed(self, steps, warmup):
            plan = self.plan(steps)           # captured before the timed region
            self.run(warmup)
            for gr in plan[-1:]:
                gr.replay()                   # the timed graph itself has run once (upload) before it is timed
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for gr in plan:
                gr.replay()
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
            if dist:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, steps

        def close(self):
            self.graphs = {}
            for e in self.envs:
                e.close()

    wl = Workload(V, B, 0)
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
        wl.streams = 2
        ms2, _ = wl.timed(args.steps, args.warmup)
        extra["two_streams"] = {"us_per_launch": ms2 * 1e3 / args.steps, "value": args.steps * n * world / (ms2 / 1e3), "unit": UNIT,
                                "note": "same launches, the 8 independent batches alternate between 2 streams so the serial "
                                        "head/tail of one launch overlaps the observation stores of the other"}
        wl.streams = 1
        # (b) the north-star 7x7x3 view (agent_view_size=7), single-step launches
        if V != 7:
            wl7 = Workload(7, B, 1 << 40)
            ms7, _ = wl7.timed(args.steps, args.warmup)
            p7, _ = measured_peak()
            extra["view7"] = {"us_per_launch": ms7 * 1e3 / args.steps, "value": args.steps * n * world / (ms7 / 1e3),
                              "unit": UNIT, "alg_bytes_per_env_step": alg_bytes(7),
                              "achieved_gbs": alg_bytes(7) * n / (ms7 / args.steps / 1e3) / 1e9,
                              "frac": alg_bytes(7) * n / (ms7 / args.steps / 1e3) / 1e9 / p7}
            wl7.close()
        # (c) the same single-step launches at larger launch sizes: the 65536-env launch of configs[2] is 15 us of
        #     which several are the serial head / tail of one wave; more envs per launch amortise it
        sweep = []
        for view_s, mult in ((V, 4), (7, 16)):
            wls = Workload(view_s, 2, 1 << 41, n=n * mult)
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
    def e2e_run(dma, steps):
        for _ in range(3):
            e2e_env.step_host(na, no, nr, nte, ntr, dma=dma)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_env.step_host(na, no, nr, nte, ntr, dma=dma)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        if dist:
            t = torch.tensor([secs], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return steps * n * world / secs, int(e2e_env.host_d2h_bytes())

    # the default transfer (2-bit codes over PCIe + host-thread decode into the caller's array) is the headline; the plain
    # DMA of the expanded bytes into the pinned array is reported beside it
    e2e_value, e2e_d2h = e2e_run(False, args.e2e_steps)
    dma_value, dma_d2h = e2e_run(True, max(8, args.e2e_steps // 2))
    extra["e2e_dma"] = {"value": dma_value, "unit": UNIT, "d2h_bytes_per_step": dma_d2h,
                        "call": "ta_step_host(TA_STEP_HOST_DMA): expanded observations copied by the DMA engine into the pinned array"}
    del h_obs, no
    wl.close()
    del wl
    torch.cuda.empty_cache()

    # ---- the PPO rollout+update loop (second half of BASELINE.json's metric) --------------
    if not args.no_ppo:
        extra["ppo"] = ppo_measure(args, rank, world, dev, dist)
    # ---- PPO + frozen frame predictor (BASELINE configs[4]) under the same world: 8192 envs GLOBAL x 32 steps, K_epochs 2
    if not args.no_ppo and not args.no_ppo_predictor:
        import copy
        pa = copy.copy(args)
        pa.ppo_predictor, pa.ppo_envs, pa.ppo_horizon, pa.ppo_epochs, pa.ppo_steps = True, 8192, 32, 2, 2
        extra["ppo_predictor"] = ppo_measure(pa, rank, world, dev, dist)
    # ---- featuriser / advantage kernels against the HBM roofline (one GPU: the kernels do not communicate) ----
    if world == 1 and not args.no_aux:
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import aux_kernels_bench
        extra["aux"] = aux_kernels_bench.measure()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    per_launch_ms = ms / args.steps
    achieved = alg_bytes(V) * n / (per_launch_ms / 1e3) / 1e9
    # bytes this layout really has to move per env-step (2-bit packed grid, 80 B each way)
    layout_bytes = 1 + 2 * (80 + 32) + 3 * V * V + 6
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(f"step_obs_v{V}_n{n}")
            traffic_src = tj.get("source", "committed ncu capture (profiles/traffic.json), not measured in this run")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": f"step_obs_kernel<{V}>",
                     "alg_bytes_per_env_step": alg_bytes(V), "peak_source": peak_src, "launch_us": per_launch_ms * 1e3,
                     "layout_bytes_per_env_step": layout_bytes,
                     "achieved_layout_bytes": layout_bytes * n / (per_launch_ms / 1e3) / 1e9,
                     "timing": "CUDA events on the launching stream around ONE replay of a CUDA graph of exactly K step launches"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": e2e_d2h,
                "steps": args.e2e_steps, "host_threads": int(os.environ.get("TA_HOST_THREADS", 0)) or max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", "1"))),
                "call": ("TwoarmyVecEnv.step_host -> ta_step_host: H2D actions from pinned memory, fused kernel, D2H of the packed "
                         "observations (2-bit cell codes) + status bytes in 8 pieces, expanded by host threads into the caller's "
                         "uint8 [n,V,V,3] / float32 / uint8 arrays; synchronous per step") if e2e_d2h < n * 3 * V * V else
                        ("TwoarmyVecEnv.step_host -> ta_step_host: H2D actions from pinned memory, fused kernel, expanded observations "
                         "copied by the DMA engine into the pinned array (the library's own choice: a single rank with at most 4 host "
                         "threads) + status bytes; synchronous per step")},
        "gpu_launches": int(launches), "clocks": clocks, "extra": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = step_cpu_baseline(args.version, V, rounds=12, warmup=2, port_seconds=args.cpu_seconds)
        line["cpu_baseline"] = cb
        if not args.no_ppo:   # C3 of BASELINE.md section 3: the reference's own PPO loop on the host cores, bounded sample
            cores = os.cpu_count() or 1
            r = ref_arm("ppo", version=args.version, frames=args.ppo_ref_frames, threads=cores)
            if "unavailable" not in r:
                extra["ppo"]["cpu_baseline"] = {
                    "value": r["value"], "unit": "frames/s", "cores": cores, "kind": "reference",
                    "sample": f"soa/train_ppo.py:99-160 verbatim (reference PPO / Env_transact / Buffer_gridworld, one env, "
                              f"select_action at B=1, get_full_render every step): {r['frames']} frames + PPO.update K_epochs "
                              f"{r['K_epochs']} x minibatches of {r['minibatch']}, {cores} torch threads, {r['seconds']:.1f} s",
                    "as_run_env_steps_per_s_one_process": r["as_run_env_steps_per_s_one_process"]}
            else:
                extra["ppo"]["cpu_baseline"] = {"unavailable": r["unavailable"]}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
