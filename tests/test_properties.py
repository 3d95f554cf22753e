"""Property tests (hypothesis): size-independent laws of the path on randomly drawn shapes and transforms.
CPU part: the oracle against itself (laws the reference's algorithm obeys).  GPU part (-m gpu): the CUDA path through
the C ABI against the oracle on shapes nobody hand-picked -- odd widths, W % 4 != 0 (no bulk copy, no 128-bit path),
tiny and non-square maps, arbitrary affine transforms."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import ubpl_oracle as O

# derandomize: the same examples in every run (a parity suite must not depend on the day's random seed); database=None:
# nothing is written next to the tests
SET = dict(deadline=None, derandomize=True, database=None,
           suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _theta(rng, n):
    ang = rng.uniform(-0.7, 0.7, n)
    sc = rng.uniform(0.6, 1.4, n)
    th = np.zeros((n, 2, 3), np.float32)
    th[:, 0, 0] = sc * np.cos(ang); th[:, 0, 1] = -sc * np.sin(ang); th[:, 0, 2] = rng.uniform(-0.3, 0.3, n)
    th[:, 1, 0] = sc * np.sin(ang); th[:, 1, 1] = sc * np.cos(ang); th[:, 1, 2] = rng.uniform(-0.3, 0.3, n)
    return th


# ------------------------------------------------------------------------------------------------------------------
# oracle laws (CPU)
# ------------------------------------------------------------------------------------------------------------------
@settings(max_examples=25, **SET)
@given(st.integers(2, 24), st.integers(2, 24), st.integers(0, 2 ** 31 - 1))
def test_identity_warp_is_identity_and_flip_is_an_involution(H, W, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((3, 2, H, W)).astype(np.float32)
    ident = np.tile(np.array([[1, 0, 0], [0, 1, 0]], np.float32), (3, 1, 1))
    back = O.affine_back2(x, ident, np.zeros(3, np.uint8))
    # the identity grid samples every texel at weight 1 up to the rounding of the base grid
    np.testing.assert_allclose(back, x, rtol=0, atol=1e-5 * max(1.0, float(np.abs(x).max())))
    assert np.array_equal(O.fliplr_back_tensor(O.fliplr_back_tensor(x)), x)
    flipped = O.affine_back2(x, ident, np.ones(3, np.uint8))
    np.testing.assert_allclose(flipped, x[..., ::-1], rtol=0, atol=1e-5 * max(1.0, float(np.abs(x).max())))


@settings(max_examples=25, **SET)
@given(st.integers(2, 20), st.integers(2, 20), st.integers(0, 2 ** 31 - 1))
def test_warp_is_linear_and_bounded_by_the_map(H, W, seed):
    """Bilinear sampling with zero padding is linear in the map, and a convex combination of its texels and zero."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((2, 2, H, W)).astype(np.float32)
    th = _theta(rng, 2)
    fl = (rng.random(2) < 0.5).astype(np.uint8)
    wa = O.affine_back2(a, th, fl)
    np.testing.assert_allclose(O.affine_back2(2.0 * a, th, fl), 2.0 * wa, rtol=1e-6, atol=1e-6)
    assert wa.max() <= max(float(a.max()), 0.0) + 1e-5 and wa.min() >= min(float(a.min()), 0.0) - 1e-5


@settings(max_examples=30, **SET)
@given(st.integers(1, 6), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_argmax_first_and_get_preds(J, HW, seed):
    rng = np.random.default_rng(seed)
    H = max(1, HW // 5); W = max(1, HW - H)
    m = rng.integers(-3, 4, (2, J, H, W)).astype(np.float32)          # many ties
    val, idx = O.argmax_first(m)
    flat = m.reshape(2, J, -1)
    assert np.array_equal(val, flat.max(-1))
    assert np.array_equal(idx, flat.argmax(-1))                       # numpy's argmax is the first maximum too
    preds = O.get_preds(m)
    keep = val > 0
    assert np.array_equal(preds[..., 0][keep], (idx % W + 1)[keep].astype(np.float32))
    assert np.array_equal(preds[..., 1][keep], (idx // W + 1)[keep].astype(np.float32))
    assert np.all(preds[~keep] == 0)


@settings(max_examples=30, **SET)
@given(st.integers(1, 200), st.floats(0.0, 0.9999), st.integers(0, 2 ** 31 - 1))
def test_ema_is_a_convex_step(n, alpha, seed):
    rng = np.random.default_rng(seed)
    e = rng.standard_normal(n).astype(np.float32); p = rng.standard_normal(n).astype(np.float32)
    out = O.ema_update(e, p, alpha)
    lo, hi = np.minimum(e, p), np.maximum(e, p)
    tol = 1e-6 * (np.abs(e) + np.abs(p) + 1)
    assert np.all(out >= lo - tol) and np.all(out <= hi + tol)
    np.testing.assert_allclose(O.ema_update(e, e, alpha), e, rtol=2e-7, atol=1e-30)   # a fixed point up to rounding
    assert np.array_equal(O.ema_update(e, p, 0.0), p)                                  # alpha = 0 (epoch 0): a copy


@settings(max_examples=20, **SET)
@given(st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_selected_set_grows_with_the_percentage(n, seed):
    rng = np.random.default_rng(seed)
    ext = np.round(rng.uniform(0, 6, n) * 4) / 4                      # the quarter-pixel grid: ties are the rule
    legal = (rng.random(n) < 0.9).astype(np.float64)
    prev = None
    for pct in (0.1, 0.3, 0.5, 0.8, 1.0):
        rel, thr, mask = O.filter_dual(ext.copy(), legal.copy(), 0.0, pct, 0.0)
        mask = np.asarray(mask, bool)
        assert np.array_equal(mask, rel > thr)
        if prev is not None:
            assert np.all(mask[prev]), "a pseudo-label selected at a lower percentage must stay selected"
        prev = mask


# ------------------------------------------------------------------------------------------------------------------
# CUDA path vs oracle on random shapes (GPU)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ubpl_b200  # noqa: F401
    from ubpl_b200 import ops as _ops
    return _ops


@pytest.mark.gpu
@settings(max_examples=30, **SET)
@given(st.integers(2, 40), st.integers(2, 40), st.integers(1, 3), st.integers(1, 5), st.integers(1, 4), st.integers(0, 2 ** 31 - 1),
       st.sampled_from(["blob", "noise", "negative", "flat"]))
def test_k1_random_shapes_vs_oracle(ops, H, W, V, B, J, seed, kind):
    import torch
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    maps = np.empty((V, B, J, H, W), np.float32)
    for idx in np.ndindex(V, B, J):
        if kind == "blob":
            cx, cy = rng.uniform(0, W), rng.uniform(0, H)
            maps[idx] = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0) + 0.01 * rng.standard_normal((H, W))
        elif kind == "noise":
            maps[idx] = rng.standard_normal((H, W))
        elif kind == "negative":
            maps[idx] = -np.abs(rng.standard_normal((H, W))) - 1e-3
        else:
            maps[idx] = float(rng.integers(-1, 2))
    th = _theta(rng, V * B).reshape(V, B, 2, 3)
    fl = (rng.random((V, B)) < 0.5).astype(np.uint8)
    center = np.full((B, 2), 100.0, np.float32); scale = np.full((B,), 1.0, np.float32)
    back = np.stack([O.affine_back2(maps[v], th[v], fl[v]) for v in range(V)])
    val, idx = O.argmax_first(back)
    xy = np.stack([O.final_preds(back[v], center, scale, [H, W], "f32") for v in range(V)])
    dec = ops.decode_coeffs(torch.as_tensor(center), torch.as_tensor(scale), [H, W]).cuda()
    r = ops.warp_decode(torch.as_tensor(maps).cuda(), torch.as_tensor(th).cuda(), torch.as_tensor(fl).cuda(), dec)
    assert np.array_equal(r["idx"].cpu().numpy().astype(np.int64), idx)
    assert np.array_equal(r["max"].cpu().numpy(), val)
    assert np.array_equal(r["xy"].cpu().numpy(), xy)
    wm = ops.warp_materialize(torch.as_tensor(maps[0]).cuda(), torch.as_tensor(th[0]).cuda(), torch.as_tensor(fl[0]).cuda())
    assert np.array_equal(wm.cpu().numpy(), back[0])


@pytest.mark.gpu
@settings(max_examples=20, **SET)
@given(st.lists(st.integers(1, 20000), min_size=1, max_size=6), st.floats(0.0, 0.9999), st.integers(0, 3), st.integers(0, 2 ** 31 - 1))
def test_ema_random_tensor_lists_vs_oracle(ops, sizes, alpha, off, seed):
    """Standalone and inside K1's launch: every element updated exactly once, bit-identical to the oracle."""
    import torch
    from ubpl_b200 import synth
    rng = np.random.default_rng(seed)
    base = torch.as_tensor(rng.standard_normal(sum(sizes) + 8).astype(np.float32)).cuda()
    params, pos = [], off
    for n in sizes:
        params.append(base[pos:pos + n]); pos += n
    e0 = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    want = [O.ema_update(e, p.cpu().numpy(), alpha) for e, p in zip(e0, params)]
    emas = [torch.as_tensor(e).cuda() for e in e0]
    ops.EmaPlan(params, emas).step(alpha)
    for e, w in zip(emas, want):
        assert np.array_equal(e.cpu().numpy(), w)
    emas = [torch.as_tensor(e).cuda() for e in e0]
    d = synth.make_batch(B=4, K=2, J=3, M=1, S=1, seed=seed % 1000, device="cuda")
    dec = ops.decode_coeffs(d["center"], d["scale"], [64, 64])
    ops.warp_decode_k2(d["teacher"][0], d["theta"], d["flip"], dec, 1, ema=ops.EmaPlan(params, emas), alpha=alpha)
    torch.cuda.synchronize()
    for e, w in zip(emas, want):
        assert np.array_equal(e.cpu().numpy(), w)
