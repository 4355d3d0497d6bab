"""Size-independent properties of the CPU oracle (hypothesis, CPU only): the same invariants the full-size GPU tests
check on the CUDA path (tests/test_gpu_fullsize.py), here on random small inputs."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import onerf_oracle as O


def _rays(rng, n):
    o = rng.uniform(-1, 1, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    near = rng.uniform(0.05, 0.5, size=(n, 1))
    far = near + rng.uniform(0.5, 3.0, size=(n, 1))
    return torch.from_numpy(np.concatenate([o, d, near, far], 1).astype(np.float32))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 9), s=st.integers(2, 40), disp=st.booleans(), jitter=st.booleans())
def test_stratified_depths_are_sorted_and_inside_near_far(seed, n, s, disp, jitter):
    rng = np.random.default_rng(seed)
    rays = _rays(rng, n)
    j = torch.from_numpy(rng.uniform(0, 1, size=(n, s)).astype(np.float32)) if jitter else None
    z = O.stratified_z(rays, s, use_disp=disp, perturb=1.0 if jitter else 0.0, jitter=j)
    assert z.shape == (n, s)
    assert (z[:, 1:] >= z[:, :-1]).all()
    eps = 1e-5 * rays[:, 7:8]
    assert (z >= rays[:, 6:7] - eps).all() and (z <= rays[:, 7:8] + eps).all()


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 7), nb=st.integers(3, 33), k=st.integers(1, 40), det=st.booleans())
def test_sample_pdf_draws_stay_inside_the_bins_and_merge_is_sorted(seed, n, nb, k, det):
    rng = np.random.default_rng(seed)
    bins = torch.from_numpy(np.sort(rng.uniform(0.1, 4.0, size=(n, nb)), axis=1).astype(np.float32))
    w = torch.from_numpy(rng.uniform(0, 1, size=(n, nb - 1)).astype(np.float32))
    w[rng.uniform(size=(n, nb - 1)) < 0.3] = 0.0                      # empty bins: the eps rule
    u = None if det else torch.from_numpy(rng.uniform(0, 1, size=(n, k)).astype(np.float32))
    z = O.sample_pdf(bins, w, k, det=det, u=u)
    assert z.shape == (n, k) and torch.isfinite(z).all()
    assert (z >= bins[:, :1] - 1e-6).all() and (z <= bins[:, -1:] + 1e-6).all()
    if det:
        assert (z[:, 1:] >= z[:, :-1] - 1e-6).all()                   # inverse CDF of sorted u is monotone
    merged = O.merge_sorted(bins, z)
    assert merged.shape == (n, nb + k) and (merged[:, 1:] >= merged[:, :-1]).all()


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 7), s=st.integers(2, 48), last=st.sampled_from([0.0, 1e10]),
       white=st.booleans())
def test_compositing_weights_form_a_sub_probability_and_maps_are_bounded(seed, n, s, last, white):
    rng = np.random.default_rng(seed)
    # sorted depths at least 0.01 apart: "opaque" below means sigma * delta >> 1, not sigma alone
    z = torch.from_numpy((np.sort(rng.uniform(0.1, 5.0, size=(n, s)), axis=1) + 0.01 * np.arange(s)).astype(np.float32))
    sigma = torch.from_numpy(rng.normal(0, 20, size=(n, s)).astype(np.float32))
    rgb = torch.from_numpy(rng.uniform(0, 1, size=(n, s, 3)).astype(np.float32))
    _, w = O.alpha_weights(sigma, z, last)
    assert (w >= 0).all() and (w.sum(1) <= 1 + 1e-4).all()
    # front-to-back: everything behind an opaque sample is occluded
    dense = sigma.clone()
    dense[:, 0] = 1e4
    _, w2 = O.alpha_weights(dense, z, last)
    if s > 2:
        assert (w2[:, 1:].sum(1) <= 1e-3).all()
    opacity, img, depth = O.composite(w, rgb, z, white)
    assert (opacity <= 1 + 1e-4).all() and (img >= -1e-5).all() and (img <= 1 + 1e-4).all()
    assert (depth <= z[:, -1] * (1 + 1e-4)).all()


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 200))
def test_slab_test_agrees_with_point_sampling(seed, n):
    """bbox_intersection: for a hit, the points at near and far lie on the box surface and the midpoint inside it;
    origins inside the box are misses (datasets/geo_utils.py:158-160)."""
    rng = np.random.default_rng(seed)
    bounds = np.stack([rng.uniform(-1, -0.2, 3), rng.uniform(0.2, 1, 3)])
    for _ in range(n):
        o = rng.uniform(-3, 3, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        hit, t0, t1 = O.bbox_intersection(bounds, o, d)
        inside = bool(((o >= bounds[0]) & (o <= bounds[1])).all())
        if inside:
            assert not hit
        if hit:
            assert 0 <= t0 <= t1
            mid = o + d * (t0 + t1) / 2
            assert ((mid >= bounds[0] - 1e-9) & (mid <= bounds[1] + 1e-9)).all()
            for t in (t0, t1):
                p = o + d * t
                assert ((p >= bounds[0] - 1e-7) & (p <= bounds[1] + 1e-7)).all()
                assert np.min(np.minimum(np.abs(p - bounds[0]), np.abs(p - bounds[1]))) <= 1e-7
