"""What the aux timing harness itself costs: a tiny kernel, a 36 MB device copy and the GAE launch at the
BASELINE configs[3] size (128 x 16384), each as a CUDA graph of 8 / 64 kernel nodes rotating over 8 buffer
sets.  Prints one JSON object (us per launch)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg


def timeit(fn, nodes, reps=6):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(8):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(nodes):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nodes) * 1e3  # us


def main():
    adv_mod = importlib.import_module(pkg.__name__ + ".advantage")
    dev = torch.device("cuda:0")
    T, N, R = 128, 16384, 8
    r = [torch.randn(T, N, device=dev) for _ in range(R)]
    v = [torch.randn(T, N, device=dev) for _ in range(R)]
    d = [(torch.rand(T, N, device=dev) < 0.02).to(torch.uint8) for _ in range(R)]
    adv = [torch.empty(T, N, device=dev) for _ in range(R)]
    ret = [torch.empty(T, N, device=dev) for _ in range(R)]
    lv = torch.randn(N, device=dev)
    tiny = torch.zeros(32, device=dev)
    src = [torch.randn(T * N * 17 // 8, device=dev) for _ in range(R)]   # 17.9 MB read + 17.9 MB written = 35.7 MB
    dst = [torch.empty_like(s) for s in src]
    k = [0]

    def f_tiny():
        tiny.add_(1.0)

    def f_copy():
        b = k[0] % R; k[0] += 1
        dst[b].copy_(src[b])

    def f_gae():
        b = k[0] % R; k[0] += 1
        adv_mod.gae(r[b], v[b], d[b], 0.99, 0.95, True, last_value=lv, out=(adv[b], ret[b]))

    def f_gae_norm():
        b = k[0] % R; k[0] += 1
        adv_mod.gae(r[b], v[b], d[b], 0.99, 0.95, True, last_value=lv, normalize=True, out=(adv[b], ret[b]))

    out = {}
    for name, fn in (("tiny", f_tiny), ("copy_36MB", f_copy), ("gae", f_gae), ("gae_norm", f_gae_norm)):
        out[name] = {f"graph{n}": round(timeit(fn, n), 2) for n in (8, 64)}
    out["TA_GAE_SMALL"] = os.environ.get("TA_GAE_SMALL", "")   # read once per process by the library
    print(json.dumps(out))


if __name__ == "__main__":
    main()
