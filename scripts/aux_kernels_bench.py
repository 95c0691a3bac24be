"""Device-timed throughput of the featuriser and advantage kernels against the HBM roofline
(582 B per env-step for the code-form state matrix, 17 B per (t, env) for
GAE, +8 B for the normalisation pass).  Prints one JSON object; used by bench.py --workload aux."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg


def timeit(fn, reps=48, warm=8):
    """Device time per call, CPU launch overhead removed: 8 consecutive calls (fn rotates over
    its own buffers) are captured in a CUDA graph and replayed."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 8):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps // 8 * 8) * 1e-3  # seconds


def measure():
    """The table as a dict (bench.py puts it under extra.aux of its default line)."""
    adv_mod = importlib.import_module(pkg.__name__ + ".advantage")
    dev = torch.device("cuda:0")
    peak = 6536.7
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"peak_gbs": peak}
    n = 65536
    B = 8  # rotate batches: the timed launches miss L2 (each batch's stacks are 95-380 MB)
    envs = [pkg.TwoarmyVecEnv(4, n, 17, device=dev, seed=1, env_id0=b * n) for b in range(B)]
    for e in envs:
        e.reset()
    codes = [torch.empty((n, 289), dtype=torch.uint8, device=dev) for _ in range(B)]
    mats = [torch.empty((n, 289), dtype=torch.float32, device=dev) for _ in range(B)]
    place = torch.empty((n, 2), dtype=torch.float32, device=dev)
    import ctypes as C
    L = pkg._capi.lib()
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    k = [0]

    def sm_codes():
        b = k[0] % B; k[0] += 1
        pkg._capi.check(L.ta_state_matrix(envs[b]._h, p(codes[b]), None, p(place), st()))

    def sm_float():
        b = k[0] % B; k[0] += 1
        pkg._capi.check(L.ta_state_matrix(envs[b]._h, None, p(mats[b]), p(place), st()))

    s = timeit(sm_codes)
    out["state_matrix_codes"] = {"us": s * 1e6, "alg_bytes": 582 * n, "gbs": 582 * n / s / 1e9, "frac": 582 * n / s / 1e9 / peak}
    s = timeit(sm_float)
    byt = (80 + 16 + 289 * 4 + 8) * n
    out["state_matrix_float"] = {"us": s * 1e6, "alg_bytes": byt, "gbs": byt / s / 1e9, "frac": byt / s / 1e9 / peak}
    sc = [torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev) for _ in range(B)]
    so = [torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev) for _ in range(B)]
    po = torch.zeros((n, 5, 2), dtype=torch.float32, device=dev)
    ps = [torch.zeros((n, 5, 2), dtype=torch.float32, device=dev) for _ in range(B)]

    def roll_codes():
        b = k[0] % B; k[0] += 1
        envs[b].stack_roll_codes(sc[b], ps[b])

    s = timeit(roll_codes)
    byt = (80 + 16 + 4 * 289 + 5 * 289 + 2 * 40) * n  # read record+agent, read 4 frames, write 5 frames, p stack r/w
    out["stack_roll_codes"] = {"us": s * 1e6, "alg_bytes": byt, "gbs": byt / s / 1e9, "frac": byt / s / 1e9 / peak}
    def push_codes():
        b = k[0] % B; k[0] += 1
        envs[b].stack_push(sc[b], so[b], ps[b], po)

    s = timeit(push_codes)
    byt = (80 + 16 + 4 * 289 + 5 * 289 + 2 * 40) * n
    out["stack_push_codes"] = {"us": s * 1e6, "alg_bytes": byt, "gbs": byt / s / 1e9, "frac": byt / s / 1e9 / peak}
    del sc, so, codes, mats
    # GAE: BASELINE configs[3] shape (T=128, N=16384) and a larger one that leaves L2
    for (T, N, tag) in ((128, 16384, "cfg4_128x16384"), (128, 262144, "128x262144")):
        R = 8
        r = [torch.randn(T, N, device=dev) for _ in range(R)]
        v = [torch.randn(T, N, device=dev) for _ in range(R)]
        d = [(torch.rand(T, N, device=dev) < 0.02).to(torch.uint8) for _ in range(R)]
        lv = torch.randn(N, device=dev)

        def gae():
            b = k[0] % R; k[0] += 1
            adv_mod.gae(r[b], v[b], d[b], 0.99, 0.95, True, last_value=lv)

        s = timeit(gae)
        byt = 17 * T * N + 4 * N
        out[f"gae_{tag}"] = {"us": s * 1e6, "alg_bytes": byt, "gbs": byt / s / 1e9, "frac": byt / s / 1e9 / peak}

        def gae_norm():
            b = k[0] % R; k[0] += 1
            adv_mod.gae(r[b], v[b], d[b], 0.99, 0.95, True, last_value=lv, normalize=True)

        s = timeit(gae_norm)
        byt = 25 * T * N + 4 * N
        out[f"gae_norm_{tag}"] = {"us": s * 1e6, "alg_bytes": byt, "gbs": byt / s / 1e9, "frac": byt / s / 1e9 / peak}
        del r, v, d
    for e in envs:
        e.close()
    torch.cuda.empty_cache()
    return out


def main():
    print(json.dumps(measure()))


if __name__ == "__main__":
    main()
