"""The CPU oracle against the reference's own known answers, and against an independent float64 restatement.

Reference golden vectors for this path: the physics doc example
(docs/tutorials/minimal-env/index.rst:140-145) and the two ragged tests (megastep/ragged.py:77-103, in
test_host.py). Nothing in the reference pins render/bake, so those are cross-checked here against a second,
independently written numpy/float64 implementation of the same geometry (not sharing code with the C oracle).
"""
import numpy as np
import pytest

import common
from oracle import oracle


def test_physics_known_answer_from_reference_docs():
    # box(5) scene, agent at (3, 3), velocity (1000, 0), fps 10 -> positions (5.8649, 3.0000)
    _, arrays = common.toy_scene('box', 2, 1)
    st = dict(angles=np.zeros((2, 1), np.float32), positions=np.full((2, 1, 2), 3., np.float32),
              angvelocity=np.zeros((2, 1), np.float32), velocity=np.tile(np.float32([1000., 0.]), (2, 1, 1)))
    progress = oracle.physics(arrays, st, fps=10.)
    np.testing.assert_allclose(st['positions'], np.tile(np.float32([5.8649, 3.]), (2, 1, 1)), atol=5e-5)
    assert (progress < 1).all()
    # a collision kills all momentum (kernels.cu:225,227)
    assert (st['velocity'] == 0).all() and (st['angvelocity'] == 0).all()


def test_physics_free_motion_and_angle_wrap():
    _, arrays = common.toy_scene('box', 1, 1)
    st = dict(angles=np.float32([[170.]]), positions=np.float32([[[3., 3.]]]),
              angvelocity=np.float32([[300.]]), velocity=np.float32([[[1., -2.]]]))
    progress = oracle.physics(arrays, st, fps=10.)
    assert progress[0, 0] == 1.
    np.testing.assert_allclose(st['positions'][0, 0], [3.1, 2.8], atol=1e-6)
    np.testing.assert_allclose(st['angles'][0, 0], -160., atol=1e-4)   # 170 + 30 wraps into [-180, 180)
    np.testing.assert_allclose(st['velocity'][0, 0], [1., -2.])


def test_agents_collide_with_each_other():
    _, arrays = common.toy_scene('box', 1, 2)
    st = dict(angles=np.zeros((1, 2), np.float32), positions=np.float32([[[2.5, 3.], [3.5, 3.]]]),
              angvelocity=np.zeros((1, 2), np.float32), velocity=np.float32([[[10., 0.], [0., 0.]]]))
    progress = oracle.physics(arrays, st, fps=10.)
    assert progress[0, 0] < 1 and st['positions'][0, 0, 0] < 3.5 - 2 * common.AGENT_RADIUS + 1e-3
    # the stationary agent "collides" too: relative motion is symmetric (kernels.cu:198)
    assert progress[0, 1] < 1


def _raycast_f64(lines, p, angle, res, fov, radius):
    """Independent float64 brute-force nearest hit, written from the geometry rather than from kernels.cu."""
    lines = lines.astype(np.float64)
    a, b = lines[:, 0], lines[:, 1]
    h = np.tan(np.deg2rad(fov) / 2)
    th = np.deg2rad(angle)
    fwd, left = np.array([np.cos(th), np.sin(th)]), np.array([-np.sin(th), np.cos(th)])
    out_idx, out_dist = [], []
    for r in range(res):
        y = (res - 2 * r - 1) * h / res
        u = fwd + y * left
        v = b - a
        den = u[0] * v[:, 1] - u[1] * v[:, 0]
        with np.errstate(divide='ignore', invalid='ignore'):
            pq = a - p
            s = (pq[:, 0] * v[:, 1] - pq[:, 1] * v[:, 0]) / den
            t = (pq[:, 0] * u[1] - pq[:, 1] * u[0]) / den
        ok = (np.abs(den) >= 1e-3) & (t >= 0) & (t <= 1) & (s > radius / np.linalg.norm(u))
        if ok.any():
            k = np.flatnonzero(ok)[np.argmin(s[ok])]
            out_idx.append(k)
            out_dist.append(s[k] * np.linalg.norm(u))
        else:
            out_idx.append(-1)
            out_dist.append(np.inf)
    return np.array(out_idx), np.array(out_dist)


@pytest.mark.parametrize('fov,res', [(130., 64), (70., 128)])
def test_render_matches_independent_float64_raycast(fov, res):
    gs, arrays = common.synthetic_scene(3, 2, seed=3)
    st = common.random_state(gs, 2, seed=4)
    out = oracle.render(arrays, st, res=res, fov=fov)
    starts = oracle.starts(arrays['line_widths'])
    bad = 0
    for n in range(3):
        ln = out['lines'][starts[n]:starts[n] + arrays['line_widths'][n]]
        for a in range(2):
            idx, dist = _raycast_f64(ln, st['positions'][n, a].astype(np.float64), float(st['angles'][n, a]), res, fov,
                                     common.AGENT_RADIUS)
            got_d = out['distances'][n, a]
            hit = idx >= 0
            assert ((out['indices'][n, a] >= 0) == hit).mean() > .98
            close = np.abs(got_d[hit] - dist[hit]) < 2e-3     # ties between abutting segments resolve either way
            bad += (~close).sum()
    assert bad <= 2


def test_render_outputs_are_well_formed():
    gs, arrays = common.synthetic_scene(2, 4, seed=5)
    st = common.random_state(gs, 4, seed=6)
    out = oracle.render(arrays, st, res=128, fov=70.)
    idx = out['indices']
    assert idx.min() >= -1 and (idx < arrays['line_widths'][:, None, None]).all()
    miss = idx < 0
    assert np.isinf(out['distances'][miss]).all() and np.isnan(out['locations'][miss]).all()
    assert (out['screen'][miss] == 0).all()
    hit = ~miss
    assert ((out['locations'][hit] >= 0) & (out['locations'][hit] <= 1)).all()
    assert (np.abs(out['dots'][hit]) <= 1 + 1e-5).all()
    assert (out['screen'] >= 0).all() and (out['screen'] <= 1 + 1e-5).all()
    # near plane: nothing closer than the agent's own radius
    assert (out['distances'][hit] > common.AGENT_RADIUS * .999).all()


def test_draw_moves_model_lines():
    gs, arrays = common.toy_scene('box', 1, 1)
    st = dict(angles=np.float32([[90.]]), positions=np.float32([[[3., 3.]]]),
              angvelocity=np.zeros((1, 1), np.float32), velocity=np.zeros((1, 1, 2), np.float32))
    out = oracle.render(arrays, st, res=8, fov=90.)
    model = arrays['model']
    # rotating by 90 degrees maps (x, y) -> (-y, x)
    expect = np.stack([-model[..., 1], model[..., 0]], -1) + 3.
    np.testing.assert_allclose(out['lines'][:8], expect, atol=1e-6)
    np.testing.assert_array_equal(out['lines'][8:], arrays['lines'][8:])


def test_bake_light_falloff_and_occlusion():
    _, arrays = common.toy_scene('box', 1, 1)
    baked = arrays['baked']
    tw = arrays['tex_widths']
    ts = oracle.starts(tw, np.int64)
    # agent lines sit at the origin, outside the box: only ambient light reaches them
    np.testing.assert_allclose(baked[:ts[8]], .1, atol=1e-6)
    # wall texels are lit by the single central light; the middle of a wall is brighter than its ends
    w0 = baked[ts[8]:ts[8] + tw[8]]
    assert w0.min() > .1 and w0[len(w0) // 2] > w0[0] and w0[len(w0) // 2] > w0[-1]
    assert baked.max() <= 1.


def test_half_screen_matches_reference_formula():
    for fov in (60., 70., 90., 130.):
        assert abs(oracle.half_screen(fov) - np.tan(np.deg2rad(fov) / 2)) < 1e-6


# ----------------------------------------------------------------------------------------------------------------------
# properties the domain offers, on random inputs (hypothesis)
# ----------------------------------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as hst  # noqa: E402


@settings(max_examples=15, deadline=None)
@given(seed=hst.integers(0, 10 ** 6), speed=hst.floats(0., 40.), n_agents=hst.integers(1, 5))
def test_physics_properties_on_random_states(seed, speed, n_agents):
    """progress in [0, 1]; momentum is kept exactly when nothing was hit and zeroed otherwise (kernels.cu:223-227);
    positions advance by progress * velocity / fps; angles stay in [-180, 180)."""
    gs, arrays = common.synthetic_scene(3, n_agents, seed=seed % 7, bake=False)
    st = common.random_state(gs, n_agents, seed=seed, speed=speed, angspeed=400.)
    before = common.copy_state(st)
    progress = oracle.physics(arrays, st, fps=10.)
    assert ((progress >= 0) & (progress <= 1)).all()
    hit = progress < 1
    assert (st['velocity'][hit] == 0).all() and (st['angvelocity'][hit] == 0).all()
    assert np.array_equal(st['velocity'][~hit], before['velocity'][~hit])
    assert np.array_equal(st['angvelocity'][~hit], before['angvelocity'][~hit])
    np.testing.assert_allclose(st['positions'], before['positions'] + progress[..., None] * before['velocity'] / 10., atol=1e-5)
    assert ((st['angles'] >= -180) & (st['angles'] < 180 + 1e-4)).all()


@settings(max_examples=10, deadline=None)
@given(seed=hst.integers(0, 10 ** 6), res=hst.sampled_from([7, 32, 48, 64]), fov=hst.floats(30., 170.))
def test_render_properties_on_random_states(seed, res, fov):
    """Every hit lies beyond the near plane (the agent's radius) and on its line; misses are -1 / NaN / +inf / black;
    rendering twice changes nothing; the screen is within [0, 1]; the agents' lines are drawn at their poses."""
    n_agents = 2
    gs, arrays = common.synthetic_scene(2, n_agents, seed=seed % 5, bake=True)
    st = common.random_state(gs, n_agents, seed=seed)
    a1 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in arrays.items()}
    r1 = oracle.render(a1, st, res=res, fov=fov)
    r2 = oracle.render(a1, st, res=res, fov=fov)
    for k in r1:
        assert np.array_equal(r1[k], r2[k], equal_nan=True), k
    assert np.array_equal(a1['lines'], arrays['lines'])
    hit = r1['indices'] >= 0
    assert (r1['distances'][hit] > common.AGENT_RADIUS * .999).all() and np.isinf(r1['distances'][~hit]).all()
    assert ((r1['locations'][hit] >= 0) & (r1['locations'][hit] <= 1)).all() and np.isnan(r1['locations'][~hit]).all()
    assert (r1['screen'] >= 0).all() and (r1['screen'] <= 1 + 1e-4).all() and (r1['screen'][~hit] == 0).all()
    widths = np.asarray(arrays['line_widths'])[:, None, None]
    assert (r1['indices'] < widths).all()
    # draw (kernels.cu:297-318): agent a's model lines sit within the model's radius of its position
    lines = r1['lines'].reshape(-1, 2, 2)                     # the oracle returns the drawn lines, the input is untouched
    ls = oracle.starts(arrays['line_widths'])
    F = len(arrays['model'])
    rad = np.abs(np.asarray(arrays['model'])).reshape(-1, 2)
    rad = np.hypot(rad[:, 0], rad[:, 1]).max()
    for n in range(2):
        for a in range(n_agents):
            seg = lines[ls[n] + a * F: ls[n] + (a + 1) * F].reshape(-1, 2)
            assert (np.hypot(*(seg - st['positions'][n, a]).T) <= rad + 1e-4).all()
