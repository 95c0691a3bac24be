"""Development probe: conv2_dgrad_planes_ws_kernel alone at B = 4096 -- time and where CTA 0's roles wait."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
L = pkg._capi.lib()
L.ta_debug_dgrad_profile.argtypes = [C.c_void_p]
B = int(os.environ.get("B", "4096"))
g = torch.Generator(device="cuda").manual_seed(1)
w = (torch.randn((64, 64, 3, 3), generator=g, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
dz = torch.randn((B, 16, 16, 64), generator=g, device="cuda").to(torch.bfloat16)
wimg = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device="cuda")
vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
pkg._capi.check(L.ta_conv2_dgrad_prep(vp(w), w.stride(0), w.stride(1), w.stride(2), w.stride(3), vp(wimg), st))
planes = torch.empty((4, B * 289, 64), dtype=torch.bfloat16, device="cuda")
mask = torch.randint(0, 2 ** 31 - 1, (B * 289 * 8,), generator=g, device="cuda", dtype=torch.int32)
prof = torch.zeros(16 + 3 * 160, dtype=torch.int64, device="cuda")
for use_mask in (False, True):
    for cm in (1, 0):
        L.ta_debug_dgrad_profile(vp(prof) if cm else None)
        for _ in range(3):
            pkg._capi.check(L.ta_conv2_dgrad_planes(vp(dz), vp(wimg), vp(mask) if use_mask else None, B, cm, vp(planes), st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pkg._capi.check(L.ta_conv2_dgrad_planes(vp(dz), vp(wimg), vp(mask) if use_mask else None, B, cm, vp(planes), st))
        e1.record(); torch.cuda.synchronize()
        print(f"class_major={cm} mask={use_mask}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us; fail {L.ta_debug_conv1_tc_failed()}")
        if cm:
            p = prof.cpu().tolist()
            tiles = (B * 289 + 127) // 128 / 148
            print(f"   CTA 0 ({tiles:.1f} tiles): producer wait a_empty {p[0]} of {p[1]} cycles | MMA wait a_full {p[2]}, acc_empty {p[3]} of {p[4]} | "
                  f"epilogue warp 0 wait acc_full {p[5]}, staging {p[6]} of {p[7]}")

# the fused kernel (data gradient + conv1's weight gradient) against the two kernels it replaces
xc = torch.randint(0, 3, (B, 5, 289), generator=g, device="cuda", dtype=torch.uint8)
dw4 = torch.empty((256, 16), device="cuda"); db4 = torch.empty(256, device="cuda")
pl = torch.empty((4, B * 289, 64), dtype=torch.bfloat16, device="cuda")
def fused():
    pkg._capi.check(L.ta_conv2_dgrad_conv1_bwd(vp(dz), vp(wimg), vp(mask), vp(xc), 1, xc.stride(0), B, vp(dw4), vp(db4), st))
def two():
    pkg._capi.check(L.ta_conv2_dgrad_planes(vp(dz), vp(wimg), None, B, 1, vp(pl), st))
    pkg._capi.check(L.ta_conv1_bwd_planes(vp(xc), 1, xc.stride(0), None, vp(mask), vp(pl), 1, B, vp(dw4), vp(db4), st))
L.ta_debug_dgrad_profile(None)
for name, fn in (("fused dgrad + conv1 wgrad", fused), ("dgrad planes, then conv1 wgrad", two)):
    if fn is fused:
        prof.zero_(); L.ta_debug_dgrad_profile(vp(prof)); fused(); torch.cuda.synchronize(); q = prof.cpu().tolist(); L.ta_debug_dgrad_profile(None)
        print(f"   CTA 0 lifetime {q[11] / 1e3:.1f} us; longest CTA {q[12]} cycles, {q[13] / 1e3:.1f} us")
        per = sorted((q[17 + 3 * i] / 1e3, q[16 + 3 * i], q[18 + 3 * i], i) for i in range(148))
        print("   per-CTA lifetime us (cycles, smid, cta): fastest", per[:4], "median", per[74], "slowest", per[-6:])
        import collections
        by = collections.defaultdict(list)
        for us, cyc, sm, i in per: by[sm // 2 % 4 if False else sm % 2].append(us)
        print("   mean by smid parity", {k: sum(v) / len(v) for k, v in by.items()}, "MHz", [round(c / u) for u, c, _, _ in per[::37]])
        print(f"   fused, CTA 0: producer waits ring {q[0]} of {q[2]} | decoder waits p_empty {q[1]} of {q[10]} | MMA waits slot_full {q[3]}, acc_empty {q[4]}, a2_full/p_full {q[5]} of {q[6]} | "
              f"epilogue warp 0 waits acc_full {q[7]}, a2_empty {q[8]} of {q[9]}")
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us; fail {L.ta_debug_conv1_tc_failed()}")
