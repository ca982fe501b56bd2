"""Golden vectors: outputs of the reference's OWN CUDA build (oracle/_ref) on small seeded scenes, captured on a B200
by tests/golden/make_golden.py and committed as tests/golden/reference_*.npz.

* CPU (always): the C oracle must reproduce them — floats to a few ulp (its 1/x and sqrt are correctly rounded where
  the GPU's MUFU.RCP/SQRT are approximate), hit indices and collision flags exactly except at near-ties. This is what
  pins the oracle to the real reference rather than to our reading of it.
* GPU: the sm_100a kernels must reproduce them bit-for-bit.
"""
import os
import sys

import numpy as np
import pytest

import common
from oracle import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_golden  # noqa: E402

CASES = sorted(make_golden.CASES)


def load(name):
    kind, N, A, res, fov, seed = make_golden.CASES[name]
    gs, arrays, st = make_golden.inputs(kind, N, A, seed)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', f'reference_{name}.npz'))
    return (kind, N, A, res, fov, seed), gs, arrays, st, gold


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_bake(name):
    _, gs, arrays, st, gold = load(name)
    got = oracle.bake(arrays)
    bad = np.abs(got - gold['baked']) > 1e-5
    # a texel whose light ray grazes a segment end can flip between lit and unlit under a 1-ulp change of 1/x
    assert bad.mean() < 2e-3, f'{bad.mean():.3%} of baked texels differ from the reference build'


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_render(name):
    (kind, N, A, res, fov, seed), gs, arrays, st, gold = load(name)
    arrays['baked'] = gold['baked']
    out = oracle.render(arrays, st, res=res, fov=fov)
    differ, really = common.index_agreement(out['indices'], gold['render_indices'], out['distances'], gold['render_distances'])
    assert really == 0. and differ < 2e-3
    same = out['indices'] == gold['render_indices']
    hit = same & (out['indices'] >= 0)
    np.testing.assert_allclose(out['distances'][hit], gold['render_distances'][hit], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(out['locations'][hit], gold['render_locations'][hit], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['dots'][hit], gold['render_dots'][hit], rtol=0, atol=2e-5)
    # screen: compare where no dynamically lit (agent-hit) ray could have flipped a shadow test
    np.testing.assert_allclose(out['screen'][same], gold['render_screen'][same], rtol=0, atol=1e-3)
    close = np.abs(out['screen'][same] - gold['render_screen'][same]) < 1e-5
    assert close.mean() > .995
    np.testing.assert_allclose(out['lines'], gold['render_lines'], rtol=0, atol=1e-6)
    assert np.isinf(gold['render_distances'][gold['render_indices'] < 0]).all()


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_physics(name):
    (kind, N, A, res, fov, seed), gs, arrays, st, gold = load(name)
    st = common.copy_state(st)
    for t, kick in enumerate(make_golden.kicks(st['velocity'].shape, seed + 2)):
        st['velocity'] += kick
        progress = oracle.physics(arrays, st, fps=10.)
        flags = progress < 1
        want = gold[f'physics{t}_progress'] < 1
        assert (flags == want).mean() > .99
        ok = flags == want
        np.testing.assert_allclose(progress[ok], gold[f'physics{t}_progress'][ok], rtol=0, atol=2e-5)
        np.testing.assert_allclose(st['positions'][ok], gold[f'physics{t}_positions'][ok], rtol=0, atol=2e-5)
        d = (st['angles'] - gold[f'physics{t}_angles'] + 180.) % 360. - 180.
        assert np.abs(d[ok]).max() < 1e-3
        # continue from the reference's state so a near-tie cannot compound
        for k in ('angles', 'positions', 'angvelocity', 'velocity'):
            st[k] = gold[f'physics{t}_{k}'].copy()


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_kernels_reproduce_reference_bit_for_bit(name):
    import torch
    from megastep_b200 import cuda, scene
    (kind, N, A, res, fov, seed), gs, arrays, st, gold = load(name)
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, res, fov, 10.))
    assert np.array_equal(s.baked.vals.cpu().numpy(), gold['baked'])
    arrays['baked'] = gold['baked']
    c = common.to_device(arrays, st, res, fov)
    r = c.render()
    for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
        got = getattr(r, k).cpu().numpy()
        assert np.array_equal(got, gold[f'render_{k}'], equal_nan=True), f'{name}: render.{k} differs from the reference build'
    assert np.array_equal(c.scenery.lines.vals.cpu().numpy(), gold['render_lines'])
    for t, kick in enumerate(make_golden.kicks(st['velocity'].shape, seed + 2)):
        c.agents.velocity.add_(torch.as_tensor(kick).cuda())
        p = c.physics()
        assert np.array_equal(p.progress.cpu().numpy(), gold[f'physics{t}_progress'])
        for k in ('angles', 'positions', 'angvelocity', 'velocity'):
            assert np.array_equal(getattr(c.agents, k).cpu().numpy(), gold[f'physics{t}_{k}']), f'{name}: tick {t} {k}'
