"""GPU: advantage / returns kernel against the reference lines (bit-exact in reference mode)
and against the float64 oracle (1e-5 relative, the tolerance north_star states for fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RTOL = 1e-5  # north_star: "GAE/returns within 1e-5 relative in fp32"


def _adv():
    import twoarmy_b200  # noqa: F401
    import importlib
    return importlib.import_module(twoarmy_b200.__name__ + ".advantage")


def test_reference_mode_is_bit_exact(golden):
    fx = golden("td_adv.npz")
    A = _adv()
    dev = "cuda:0"
    adv, tv = A.reference_mode(torch.tensor(fx["r"], device=dev), torch.tensor(fx["v"], device=dev),
                               torch.tensor(fx["v_next"], device=dev), float(fx["gamma"]))
    assert np.array_equal(tv.cpu().numpy(), fx["target_v"])
    assert np.array_equal(adv.cpu().numpy(), fx["adv"])


@pytest.mark.parametrize("T,N", [(128, 16384), (1, 7), (33, 100), (300, 257), (50, 1), (300, 260), (129, 4)])
@pytest.mark.parametrize("lam,use_mask", [(0.95, True), (1.0, True), (0.0, False), (0.9, False)])
def test_gae_matches_float64_oracle(T, N, lam, use_mask):
    from oracle import oracle as O
    A = _adv()
    rng = np.random.default_rng(T * 1000 + N)
    lut = np.array([-0.01, -0.1, -0.9, 0.2, 0.9], np.float32)
    r = lut[rng.choice(5, size=(T, N), p=[0.9, 0.04, 0.02, 0.02, 0.02])]
    v = rng.standard_normal((T, N)).astype(np.float32)
    last = rng.standard_normal(N).astype(np.float32)
    done = (rng.random((T, N)) < 0.03).astype(np.uint8)
    want_adv, want_ret = O.gae(r, v, done, 0.99, lam, use_mask=use_mask, last_v=last)
    dev = "cuda:0"
    adv, ret = A.gae(torch.tensor(r, device=dev), torch.tensor(v, device=dev), torch.tensor(done, device=dev), 0.99,
                     lam, use_mask=use_mask, last_value=torch.tensor(last, device=dev))
    scale = np.maximum(np.abs(want_adv), 1.0)
    assert np.max(np.abs(adv.cpu().numpy() - want_adv) / scale) < RTOL
    scale = np.maximum(np.abs(want_ret), 1.0)
    assert np.max(np.abs(ret.cpu().numpy() - want_ret) / scale) < RTOL


@pytest.mark.parametrize("T,N", [(128, 4096), (24, 65536), (40, 1003), (128, 16384), (16, 64), (130, 4096), (8, 18944), (8, 18976)])
def test_gae_with_per_sample_v_next_and_normalisation(T, N):
    """normalize=True (ta_gae_normalized): rollouts whose grid is resident at once are normalised INSIDE the GAE launch
    (grid barrier on a ticket: 128 x 4096, 128 x 16384 = BASELINE configs[3], 16 x 64 with CTAs of 16 threads, 8 x 18944 =
    the last env count that fits); the others take the moments from the GAE launch plus one pass over adv (130 steps = two
    passes over T, 18976 envs = one CTA too many, 65536 envs = the 128-env-CTA kernel, N % 4 != 0 = the scalar kernel)."""
    from oracle import oracle as O
    A = _adv()
    rng = np.random.default_rng(3)
    r = rng.standard_normal((T, N)).astype(np.float32) * 0.1
    v = rng.standard_normal((T, N)).astype(np.float32)
    vn = rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.02).astype(np.uint8)
    want_adv, _ = O.gae(r, v, done, 0.99, 0.95, use_mask=True, v_next=vn)
    dev = "cuda:0"
    adv, _ = A.gae(torch.tensor(r, device=dev), torch.tensor(v, device=dev), torch.tensor(done, device=dev), 0.99, 0.95,
                   use_mask=True, v_next=torch.tensor(vn, device=dev), normalize=True)
    want = (want_adv - want_adv.mean()) / (want_adv.std(ddof=1) + 1e-8)  # PPO.py:115 (torch.std is unbiased)
    assert np.max(np.abs(adv.cpu().numpy() - want)) < 1e-4
