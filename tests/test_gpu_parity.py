"""Parity of the sm_100a kernels — called through the C ABI — on a real GPU.

Three checkers, strongest first:
  1. oracle/_ref: the reference's OWN kernels.cu/wrappers.cpp compiled unmodified for sm_100a. Bar: hit indices and
     collision flags bit-exact, floats bit-exact too (the kernels restate its arithmetic op-for-op); the asserted
     tolerance for positions/depth is the north-star's 1e-5 abs, and the exact-match fractions are printed.
  2. oracle/megastep_oracle.c: the CPU restatement (always available). MUFU.RCP/SQRT are not reproducible on a
     CPU, so floats are compared to 1e-4 relative and indices must agree except at near-ties.
  3. size-independent properties at the benchmark's full size (BASELINE.json configs[2]).
"""
import numpy as np
import pytest
import torch

import common
from oracle import oracle

pytestmark = pytest.mark.gpu

CASES = [
    # name, scene kind, n_envs, n_agents, res, fov
    ('box-explorer', 'box', 3, 1, 64, 130.),
    ('column', 'column', 2, 2, 32, 90.),
    ('synthetic-explorer', 'synthetic', 24, 1, 64, 130.),
    ('synthetic-deathmatch', 'synthetic', 16, 4, 128, 70.),
    ('synthetic-ragged-res', 'synthetic', 5, 3, 48, 100.),     # res not a multiple of 32
    ('synthetic-wide', 'synthetic', 4, 2, 512, 70.),
    ('synthetic-six-agents', 'synthetic', 6, 6, 96, 90.),       # more agents than warps in a CTA
]


def make(kind, n_envs, n_agents, seed=11):
    if kind == 'synthetic':
        gs, arrays = common.synthetic_scene(n_envs, n_agents, seed=seed)
    else:
        gs, arrays = common.toy_scene(kind, n_envs, n_agents, seed=seed)
    return gs, arrays, common.random_state(gs, n_agents, seed=seed + 1)


@pytest.fixture(scope='module')
def ref():
    return common.reference_module()


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------------------------
# 2. against the CPU oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,kind,N,A,res,fov', CASES)
def test_render_against_oracle(name, kind, N, A, res, fov):
    gs, arrays, st = make(kind, N, A)
    c = common.to_device(arrays, st, res, fov)
    r = c.render()
    want = oracle.render(arrays, st, res=res, fov=fov)
    torch.cuda.synchronize()
    idx, dist = _np(r.indices), _np(r.distances)
    differ, really = common.index_agreement(idx, want['indices'], dist, want['distances'])
    assert really == 0., f'{name}: {really:.2%} of rays hit something at a different depth than the oracle'
    assert differ < 2e-3, f'{name}: {differ:.2%} of hit indices differ from the oracle'
    same = idx == want['indices']
    hit = same & (idx >= 0)
    np.testing.assert_allclose(dist[hit], want['distances'][hit], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(_np(r.locations)[hit], want['locations'][hit], rtol=0, atol=2e-4)
    np.testing.assert_allclose(_np(r.dots)[hit], want['dots'][hit], rtol=0, atol=1e-4)
    np.testing.assert_allclose(_np(r.screen)[same], want['screen'][same], rtol=0, atol=2e-3)
    miss = same & (idx < 0)
    assert np.isinf(dist[miss]).all() and np.isnan(_np(r.locations)[miss]).all() and (_np(r.screen)[miss] == 0).all()
    # draw side effect: the agents' model lines now sit at the agents' poses
    np.testing.assert_allclose(_np(c.scenery.lines.vals), want['lines'], rtol=0, atol=1e-5)


@pytest.mark.parametrize('name,kind,N,A,res,fov', CASES[:4])
def test_physics_against_oracle(name, kind, N, A, res, fov):
    gs, arrays, st = make(kind, N, A)
    c = common.to_device(arrays, st, res, fov)
    want_st = common.copy_state(st)
    for _ in range(3):     # a few ticks so collisions, stops and re-accelerations all occur
        p = c.physics()
        want_p = oracle.physics(arrays, want_st, fps=10.)
        got = common.read_state(c)
        collided = _np(p.progress) < 1
        assert (collided == (want_p < 1)).mean() > .995
        agree = collided == (want_p < 1)
        np.testing.assert_allclose(_np(p.progress)[agree], want_p[agree], rtol=0, atol=2e-4)
        for k in ('positions', 'velocity'):
            np.testing.assert_allclose(got[k][agree], want_st[k][agree], rtol=0, atol=2e-4, err_msg=k)
        # angles wrap at +-180: compare on the circle
        d = (got['angles'] - want_st['angles'] + 180.) % 360. - 180.
        assert np.abs(d[agree]).max() < 1e-2
        # resync so that one near-tie does not compound over the following ticks
        common.load_state(c, want_st)
        st_rand = np.random.RandomState(5)
        want_st['velocity'] += st_rand.normal(size=want_st['velocity'].shape).astype(np.float32)
        common.load_state(c, want_st)


def test_bake_against_oracle():
    gs, arrays = common.synthetic_scene(6, 2, seed=21, bake=True)
    from megastep_b200 import cuda, scene
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 64, 130., 10.))
    got = _np(s.baked.vals)
    # a texel whose light ray grazes a wall end can flip between lit and unlit: allow a handful
    bad = np.abs(got - arrays['baked']) > 1e-4
    assert bad.mean() < 2e-3, f'{bad.mean():.3%} of baked texels differ from the oracle'


def test_physics_known_answer_from_reference_docs():
    # docs/tutorials/minimal-env/index.rst:140-145
    gs, arrays = common.toy_scene('box', 4, 1)
    st = dict(angles=np.zeros((4, 1), np.float32), positions=np.full((4, 1, 2), 3., np.float32),
              angvelocity=np.zeros((4, 1), np.float32), velocity=np.tile(np.float32([1000., 0.]), (4, 1, 1)))
    c = common.to_device(arrays, st, 64, 130.)
    p = c.physics()
    np.testing.assert_allclose(_np(c.agents.positions), np.tile(np.float32([5.8649, 3.]), (4, 1, 1)), atol=5e-5)
    assert (_np(p.progress) < 1).all() and (_np(c.agents.velocity) == 0).all()


# ------------------------------------------------------------------------------------------------------------------
# 1. against the reference's own CUDA build
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,kind,N,A,res,fov', CASES)
def test_render_bit_exact_against_reference_build(ref, name, kind, N, A, res, fov):
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    gs, arrays, st = make(kind, N, A)
    c = common.to_device(arrays, st, res, fov)
    r = c.render()
    ref.initialize(common.AGENT_RADIUS, res, fov, 10.)
    rs, ra = common.reference_scenery(ref, arrays), common.reference_agents(ref, st)
    rr = ref.render(rs, ra)
    torch.cuda.synchronize()
    assert torch.equal(r.indices, rr.indices), f'{name}: {(r.indices != rr.indices).float().mean():.3%} hit indices differ'
    for k in ('locations', 'dots', 'distances', 'screen'):
        a, b = getattr(r, k), getattr(rr, k)
        same = (a == b) | (a.isnan() & b.isnan())
        print(f'{name}: {k} bit-exact on {same.float().mean():.6%}')
        torch.testing.assert_close(a, b, rtol=0, atol=1e-5, equal_nan=True)
        assert same.all(), f'{name}: {k} is within 1e-5 but not bit-exact on {(~same).float().mean():.4%}'
    assert torch.equal(c.scenery.lines.vals, rs.lines.vals)


@pytest.mark.parametrize('name,kind,N,A,res,fov', CASES)
def test_physics_bit_exact_against_reference_build(ref, name, kind, N, A, res, fov):
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    gs, arrays, st = make(kind, N, A)
    c = common.to_device(arrays, st, res, fov)
    ref.initialize(common.AGENT_RADIUS, res, fov, 10.)
    rs, ra = common.reference_scenery(ref, arrays), common.reference_agents(ref, st)
    rng = np.random.RandomState(9)
    for tick in range(4):
        p, rp = c.physics(), ref.physics(rs, ra)
        assert torch.equal(p.progress < 1, rp.progress < 1), f'{name}: collision flags differ at tick {tick}'
        for k, a, b in [('progress', p.progress, rp.progress)] + [(k, getattr(c.agents, k), getattr(ra, k)) for k in
                                                                   ('positions', 'angles', 'velocity', 'angvelocity')]:
            torch.testing.assert_close(a, b, rtol=0, atol=1e-5, msg=lambda m: f'{name} tick {tick} {k}: {m}')
            assert torch.equal(a, b), f'{name} tick {tick}: {k} within 1e-5 but not bit-exact'
        kick = torch.as_tensor(rng.normal(size=st['velocity'].shape).astype(np.float32) * 2).cuda()
        c.agents.velocity.add_(kick)
        ra.velocity.add_(kick)


def test_more_lights_than_lanes_bit_exact_against_reference_build(ref):
    """An env with 40 lights: dynamic lighting of agent-hit rays leaves the 32-lights-per-warp fast path (dyn_kernel's
    and the inline fallback's), baking sums all of them. Agents packed close so that many rays hit agents."""
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    from megastep_b200 import cuda, scene, toys
    from megastep_b200.arrdict import arrdict
    g = toys.box()
    rng = np.random.RandomState(5)
    g = arrdict(walls=g.walls, lights=rng.uniform(1.5, 5.5, (40, 2)), masks=g.masks, res=g.res)
    gs = [g] * 3
    arrays = scene.scene_arrays(gs, 4, np.random.RandomState(6))
    arrays['baked'] = oracle.bake(arrays)
    st = common.random_state(gs, 4, seed=7)
    st['positions'] = (3.5 + rng.uniform(-.6, .6, st['positions'].shape)).astype(np.float32)
    ref.initialize(common.AGENT_RADIUS, 128, 100., 10.)
    rs, ra = common.reference_scenery(ref, arrays), common.reference_agents(ref, st)
    rr = ref.render(rs, ra)
    for use_ws in (True, False):
        cuda.USE_WORKSPACE = use_ws
        try:
            c = common.to_device(arrays, st, 128, 100.)
            r = c.render()
            torch.cuda.synchronize()
        finally:
            cuda.USE_WORKSPACE = True
        assert int(((r.indices >= 0) & (r.indices < 32)).sum()) > 50
        assert torch.equal(r.indices, rr.indices)
        assert _same(r.screen, rr.screen), f'workspace={use_ws}: {(r.screen != rr.screen).float().mean():.4%} of screen differs'
    # and the baked light map with 40 lights
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 128, 100., 10.))
    rs2 = common.reference_scenery(ref, arrays)
    ref.bake(rs2)
    torch.cuda.synchronize()
    assert torch.equal(s.baked.vals, rs2.baked.vals)


def test_bake_over_the_spatial_table_equals_brute_force():
    """msb_bake's two kernels — every static line per (texel, light) like the reference, or only the runs of the spatial
    table the light ray passes — must agree bit for bit (the cull is conservative), ragged envs and 40 lights included."""
    from megastep_b200 import cuda, scene
    gs, arrays = common.synthetic_scene(9, 3, seed=23, bake=False)
    rng = np.random.RandomState(5)
    extra = rng.uniform(2., 9., (40 - int(arrays['light_widths'][0]), 3)).astype(np.float32)      # env 0 gets 40 lights
    arrays['lights'] = np.concatenate([arrays['lights'][:arrays['light_widths'][0]], extra, arrays['lights'][arrays['light_widths'][0]:]])
    arrays['light_widths'][0] = 40
    outs = []
    for brute in (1, 0):
        cuda.set_option('bake_brute', brute)
        try:
            s = scene.upload(arrays)
            cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 64, 130., 10.))
            torch.cuda.synchronize()
        finally:
            cuda.set_option('bake_brute', 0)
        outs.append(s.baked.vals.clone())
    assert torch.equal(outs[0], outs[1]), f'{(outs[0] != outs[1]).float().mean():.4%} of baked texels differ'
    assert float(outs[0].min()) >= .1 and float(outs[0].max()) <= 1. and float(outs[0].std()) > .05


def test_bake_bit_exact_against_reference_build(ref):
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    gs, arrays = common.synthetic_scene(6, 2, seed=21, bake=False)
    from megastep_b200 import cuda, scene
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 64, 130., 10.))
    ref.initialize(common.AGENT_RADIUS, 64, 130., 10.)
    rs = common.reference_scenery(ref, arrays)
    ref.bake(rs)
    torch.cuda.synchronize()
    same = s.baked.vals == rs.baked.vals
    assert same.all(), f'{(~same).float().mean():.4%} of baked texels differ from the reference build'


# ------------------------------------------------------------------------------------------------------------------
# fused paths against the unfused ones
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('A,res,fov,sub', [(4, 128, 70., 1), (4, 512, 70., 4), (1, 256, 130., 4), (3, 96, 100., 2)])
def test_fused_step_equals_modules_pipeline(A, res, fov, sub):
    """FusedStep (what bench.py times) against this package's unfused modules — MomentumMovement -> render -> RGB / Depth /
    IMU, PyTorch ops around cuda.physics / cuda.render as in the reference — over 6 ticks from the same state with NO
    resynchronisation: everything bit for bit (in-kernel movement restates torch's op order with rounding-explicit
    intrinsics; the pooled heads add in ATen's order). The same comparison against the reference's OWN modules.py runs
    in tests/test_gpu_reference_python.py."""
    from megastep_b200 import modules
    from megastep_b200.arrdict import arrdict
    gs, arrays, st = make('synthetic', 12, A, seed=31)
    c1, c2 = common.to_device(arrays, st, res, fov), common.to_device(arrays, st, res, fov)
    mover, rgb, depth, imu = modules.MomentumMovement(c1), modules.RGB(c1, subsample=sub), modules.Depth(c1, subsample=sub), modules.IMU(c1)
    fused = modules.FusedStep(c2, subsample=sub, raw=True)
    rng = np.random.RandomState(3)
    for tick in range(6):
        actions = torch.as_tensor(rng.randint(0, 7, (12, A))).int().cuda()
        p = mover(arrdict(actions=actions))
        r = modules.render(c1)
        want = arrdict(rgb=rgb(r), d=depth(r), imu=imu())
        out = fused(actions)
        torch.cuda.synchronize()
        assert torch.equal(out.progress, p.progress), f'tick {tick}: progress'
        for k in ('positions', 'angles', 'velocity', 'angvelocity'):
            assert torch.equal(getattr(c2.agents, k), getattr(c1.agents, k)), f'tick {tick}: {k}'
        assert torch.equal(out.render.indices, r.indices.squeeze(2)), f'tick {tick}: indices'
        for k in ('locations', 'dots', 'distances'):
            assert _same(getattr(out.render, k), r[k].squeeze(2)), f'tick {tick}: {k}'
        assert torch.equal(out.obs.imu, want.imu) and torch.equal(out.obs.d, want.d) and torch.equal(out.obs.rgb, want.rgb), f'tick {tick}: obs'


def test_random_spawns_respawn_only_the_flagged_agents_on_the_device():
    """modules.RandomSpawns (reference modules.py:295-326): flagged agents land on one of their pre-computed free
    spawn points with zeroed velocities, the others are untouched — without the reference's host sync."""
    from megastep_b200 import modules
    gs, arrays, st = make('box', 6, 3, seed=41)               # toy geometries carry the free-space masks spawns are drawn from
    c = common.to_device(arrays, st, 64, 90.)
    spawner = modules.RandomSpawns(gs, c, n_spawns=20)
    before = common.read_state(c)
    reset = torch.as_tensor(np.random.RandomState(1).rand(6, 3) < .5).cuda()
    spawner(reset)
    after = common.read_state(c)
    m = _np(reset)
    for k in before:
        assert np.array_equal(before[k][~m], after[k][~m]), k
    assert (after['velocity'][m] == 0).all() and (after['angvelocity'][m] == 0).all()
    spawns = _np(spawner._spawns.positions)                  # (N, A, n_spawns, 2)
    for n, a in zip(*m.nonzero()):
        assert (np.abs(spawns[n, a] - after['positions'][n, a]).sum(-1) == 0).any()
        assert -180 <= after['angles'][n, a] <= 180


def test_graph_replay_with_host_io_equals_plain_launches():
    """FusedStep as plain launches, as a CUDA-graph replay, and as ONE graph launch holding the host copies too
    (step_host) must walk the same trajectory: the programmatic-dependent launches and the streaming queue between
    view_kernel and dyn_kernel are captured as such."""
    from megastep_b200 import modules
    gs, arrays, st = make('box', 6, 4, seed=91)
    st['positions'] = (3.5 + np.random.RandomState(0).uniform(-.8, .8, st['positions'].shape)).astype(np.float32)
    acts = torch.as_tensor(np.random.RandomState(5).randint(0, 7, (6, 6, 4)).astype(np.int32))
    recs = []
    from megastep_b200 import cuda
    pinned = acts.pin_memory()
    for mode in ('plain', 'graph', 'host', 'idx64', 'native', 'native-pinned'):   # idx64: plain launches, 64-bit output indexing
        c = common.to_device(arrays, st, 128, 100.)
        step = modules.FusedStep(c, subsample=2, raw=True, graph=mode == 'graph')
        if mode == 'host':
            step._capture(host_io=True)
        if mode.startswith('native'):                          # the graph captured inside the library, one call per tick
            step.enable_host_graph()
        cuda.set_option('idx64', int(mode == 'idx64'))
        rec = []
        try:
            for t in range(6):
                if mode in ('host', 'native'):
                    out = step.step_host(acts[t])
                elif mode == 'native-pinned':
                    out = step.step_host(pinned[t])
                else:
                    out = step(acts[t].cuda())
                torch.cuda.synchronize()
                rec.append([torch.as_tensor(out.progress).cpu().clone(), out.obs.rgb.cpu().clone(), out.obs.d.cpu().clone(),
                            out.obs.imu.cpu().clone(), out.render.screen.cpu().clone(), out.render.indices.cpu().clone()])
        finally:
            cuda.set_option('idx64', 0)
        recs.append(rec)
    n_dyn = int(((recs[0][0][5] >= 0) & (recs[0][0][5] < 32)).sum())
    assert n_dyn > 20
    for other in recs[1:]:
        for a, b in zip(recs[0], other):
            assert all(_same(x, y) for x, y in zip(a, b))


def downsampled_all(mask, sub):
    return mask.reshape(*mask.shape[:-1], mask.shape[-1] // sub, sub).all(-1)


def test_rgbd_head_equals_rgb_and_depth_modules():
    from megastep_b200 import modules
    gs, arrays, st = make('synthetic', 8, 4, seed=41)
    c = common.to_device(arrays, st, 512, 70.)
    r = modules.render(c)
    want_rgb, want_d, want_imu = modules.RGB(c, subsample=4)(r), modules.Depth(c, subsample=4)(r), modules.IMU(c)()
    obs = modules.RGBD(c, subsample=4)()
    torch.cuda.synchronize()
    assert obs.rgb.shape == want_rgb.shape and obs.d.shape == want_d.shape
    assert torch.equal(obs.rgb, want_rgb) and torch.equal(obs.d, want_d) and torch.equal(obs.imu, want_imu)


OPTIONS = ('nch', 'threads', 'stage_rec', 'idx64', 'persist', 'merge_dyn', 'dyn_groups', 'stages', 'no_sched', 'dyn_warps')


def _reset_options():
    from megastep_b200 import cuda
    for name in OPTIONS:
        cuda.set_option(name, 0)


@pytest.mark.parametrize('res', [48, 64, 128, 512])
def test_every_kernel_variant_gives_identical_results(res):
    """One CTA per env (view_kernel + dyn_kernel) or the persistent grid (tick_kernel, second pass merged in or not); the
    chunking / thread-count / staging options: they change which boxes a warp opens, which segments it skips and who
    runs what, never results."""
    from megastep_b200 import cuda
    gs, arrays, st = make('synthetic', 10, 4, seed=51)
    c = common.to_device(arrays, st, res, 70.)
    cuda.set_option('persist', 2)
    base = c.render()

    def check(what):
        r = c.render()
        for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
            a, b = getattr(r, k), getattr(base, k)
            assert ((a == b) | (a != a) & (b != b)).all(), f'{what}: {k} differs'

    try:
        for stage_rec in (1, 2):                               # the rows' records staged in shared memory / left in global
            for nch in (1, 2, 4):
                for threads in (32, 64, 128, 256):
                    for name, v in (('persist', 2), ('stage_rec', stage_rec), ('nch', nch), ('threads', threads)):
                        cuda.set_option(name, v)
                    check(f'view_kernel stage_rec={stage_rec} nch={nch} threads={threads}')
                for threads in (64, 256):
                    for stages in (2, 3):
                        for name, v in (('persist', 1), ('stage_rec', stage_rec), ('nch', nch), ('threads', threads), ('stages', stages)):
                            cuda.set_option(name, v)
                        check(f'tick_kernel stage_rec={stage_rec} nch={nch} threads={threads} stages={stages}')
        _reset_options()
        for opts in ({'persist': 1, 'merge_dyn': 2}, {'persist': 1, 'dyn_groups': 1}, {'persist': 1, 'dyn_groups': 2},
                     {'persist': 1, 'no_sched': 1}, {'persist': 1, 'stages': 4, 'threads': 128}, {'persist': 1, 'idx64': 1}, {'idx64': 1}, {}):
            _reset_options()
            for name, v in opts.items():
                cuda.set_option(name, v)
            for _ in range(2):                                 # twice: the counters must re-arm themselves
                check(str(opts))
    finally:
        _reset_options()


def _same(a, b):
    return bool(((a == b) | (a != a) & (b != b)).all())


def test_step_with_physics_inside_the_render_kernel_agrees():
    """msb_step with physics as its own launch (default) and fused into view_kernel (option fused_step): same
    progress, same agent state, same observations, over several ticks with slow, stationary and fast agents."""
    from megastep_b200 import cuda, modules
    gs, arrays, st = make('synthetic', 12, 4, seed=71)
    st['velocity'][:3] = 0.                                   # exactly stationary
    st['velocity'][3:6] *= 1e-4                                # slow but moving: the no-cull path
    outs = []
    for fused in (0, 1):
        cuda.set_option('fused_step', fused)
        try:
            c = common.to_device(arrays, st, 64, 90.)
            step = modules.FusedStep(c, subsample=2, raw=True)
            acts = torch.as_tensor(np.random.RandomState(5).randint(0, 7, (3, 12, 4)).astype(np.int32)).cuda()
            rec = []
            for t in range(3):
                out = step(acts[t])
                rec.append((out.progress.clone(), out.obs.rgb.clone(), out.obs.d.clone(), out.obs.imu.clone(), out.render.indices.clone()))
            outs.append((rec, common.read_state(c)))
        finally:
            cuda.set_option('fused_step', 0)
    for a, b in zip(outs[0][0], outs[1][0]):
        assert all(_same(x, y) for x, y in zip(a, b))
    for k in outs[0][1]:
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k


@pytest.mark.parametrize('sub', [1, 4])
def test_second_pass_inline_and_overflow_paths_agree(sub):
    """Agent-hit rays are lit by the persistent kernel itself (tickets: 4, 2 or 1 warps per queue entry), by dyn_kernel
    behind it or behind view_kernel (its warps share each entry's lights 2, 1 or 4 ways), inline by the first pass (no
    workspace), or by a mix (workspace too small): all must give identical screens and observations. Agents are packed
    close so many rays hit agents."""
    from megastep_b200 import cuda
    gs, arrays, st = make('box', 6, 4, seed=61)
    rng = np.random.RandomState(0)
    st['positions'] = (3.5 + rng.uniform(-.6, .6, st['positions'].shape)).astype(np.float32)
    res = 128
    outs = []
    modes = {'view+dyn': {}, 'view-inline': {}, 'view-overflow': {}, 'view+dyn-1warp': {'dyn_warps': 1}, 'view+dyn-4warps': {'dyn_warps': 4},
             'tick': {'persist': 1}, 'tick-2-per-entry': {'persist': 1, 'dyn_groups': 2}, 'tick-1-per-entry': {'persist': 1, 'dyn_groups': 1},
             'tick+dyn_kernel': {'persist': 1, 'merge_dyn': 2}, 'tick-inline': {'persist': 1}, 'tick-overflow': {'persist': 1},
             'tick+dyn_kernel-overflow': {'persist': 1, 'merge_dyn': 2}}
    for mode, opts in modes.items():
        cuda.USE_WORKSPACE = 'inline' not in mode
        _reset_options()
        for name, v in opts.items():
            cuda.set_option(name, v)
        try:
            c = common.to_device(arrays, st, res, 100.)
            plan = cuda.StepPlan(c.scenery, c.agents, c.params, actions=None, raw=True, subsample=sub)
            if 'overflow' in mode:
                small = 32 + 6 * 4 * 32 * 4 + 3 * (48 + 32 * max(4, sub))   # counters + occluder cache + room for three pixel windows only
                plan._wsbuf = torch.zeros(small, dtype=torch.uint8, device='cuda')
                plan._ws = cuda._Workspace(plan._wsbuf.data_ptr(), small)
            for _ in range(3):                                          # several times: the queue must re-arm itself
                plan.render_only()
            torch.cuda.synchronize()
        finally:
            cuda.USE_WORKSPACE = True
            _reset_options()
        outs.append((mode, plan.render.screen.clone(), plan.rgb.clone(), plan.render.indices.clone()))
    n_dyn = int(((outs[0][3] >= 0) & (outs[0][3] < 32)).sum())
    assert n_dyn > 50, 'the scene should have plenty of agent-hit rays'
    for mode, screen, rgb, _ in outs[1:]:
        assert _same(outs[0][1], screen) and _same(outs[0][2], rgb), mode


def test_static_geometry_edited_in_place_is_seen_after_invalidate():
    """The side tables (spatial table, visibility grid) are built once per Scenery: after moving a wall in place,
    Scenery.invalidate() makes the kernels see it — same results as a scenery built from the edited arrays."""
    gs, arrays, st = make('synthetic', 5, 2, seed=81)
    c = common.to_device(arrays, st, 64, 100.)
    before = c.render()
    AF = 2 * 8
    lo, hi = int(c.scenery.lines.starts[0]) + AF, int(c.scenery.lines.ends[0])
    c.scenery.lines.vals[lo:hi] += .37                           # shift every wall of env 0
    c.scenery.invalidate()
    fresh = c.render()
    edited = dict(arrays)
    edited['lines'] = arrays['lines'].copy()
    edited['lines'][lo:hi] += np.float32(.37)
    c2 = common.to_device(edited, st, 64, 100.)
    want = c2.render()
    torch.cuda.synchronize()
    assert not _same(fresh.distances[0], before.distances[0]) and _same(fresh.distances[1:], before.distances[1:])
    for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
        assert _same(getattr(fresh, k), getattr(want, k)), k


def test_native_table_builder_equals_the_torch_restatement():
    """msb_build_table (sort-tile-recursive packing by two shared-memory sorts) against cuda._occluder_table, bit for
    bit: rows, records, run boxes, per-env summary; ragged envs including ones with no static line at all."""
    from megastep_b200 import cuda, scene
    gs, arrays = common.synthetic_scene(9, 3, seed=77, bake=False)
    s = scene.upload(arrays)
    # an env with nothing but the agents' lines, and one with a single static line
    lw = s.lines.widths.clone()
    AF = 3 * s.model.size(0)
    keep = torch.ones(s.lines.vals.size(0), dtype=torch.bool, device='cuda')
    for n, left in ((2, 0), (5, 1)):
        lo, hi = int(s.lines.starts[n]) + AF + left, int(s.lines.ends[n])
        keep[lo:hi] = False
        lw[n] = AF + left
    lines = cuda.Ragged3D(s.lines.vals[keep].contiguous(), lw)
    tex = cuda.Ragged2D(s.textures.vals[keep[s.textures.inverse.long()]].contiguous(), s.textures.widths[keep].contiguous())
    s2 = cuda.Scenery(3, s.lights, lines, tex, s.model)
    got = s2._struct() and s2._occ
    want = cuda._occluder_table(lines, AF, 16, tex.widths, tex._long_starts())
    names = ('occ_lines', 'occ_starts', 'occ_boxes', 'box_starts', 'occ_meta', 'occ_rec')
    for name, a, b in zip(names, got, want):
        assert a.shape == b.shape and torch.equal(a, b), name


def test_visibility_grid_is_conservative_and_changes_nothing():
    """The light-visibility grid (msb_scenery::vis) may only vouch for lights that no static segment comes near, and
    using it may not change a single output bit. Checked on a box scene with packed agents (lit by the room's light)
    and on synthetic floorplans; the vouched-for (cell, light) pairs are re-derived in float64 at sampled points."""
    from megastep_b200 import cuda
    for kind, N, seed in (('box', 4, 61), ('synthetic', 12, 62)):
        gs, arrays, st = make(kind, N, 4, seed=seed)
        if kind == 'box':
            st['positions'] = (3.5 + np.random.RandomState(0).uniform(-.6, .6, st['positions'].shape)).astype(np.float32)
        outs = []
        for no_vis in (0, 1):
            cuda.set_option('no_vis', no_vis)
            try:
                c = common.to_device(arrays, st, 128, 100.)
                plan = cuda.StepPlan(c.scenery, c.agents, c.params, actions=None, raw=True, subsample=1)
                plan.render_only()
                torch.cuda.synchronize()
            finally:
                cuda.set_option('no_vis', 0)
            outs.append((plan.render.screen.clone(), plan.rgb.clone(), plan.render.indices.clone()))
        n_dyn = int(((outs[0][2] >= 0) & (outs[0][2] < 32)).sum())
        assert n_dyn > (50 if kind == 'box' else 0)
        assert _same(outs[0][0], outs[1][0]) and _same(outs[0][1], outs[1][1])
        # the table itself
        sc = c.scenery
        vis, starts, meta = (_np(t) for t in sc._vis)
        lines = _np(sc.lines.vals).reshape(-1, 4).astype(np.float64)
        lstarts, lwidths = _np(sc.lines.starts), _np(sc.lines.widths)
        lights, istarts, iwidths = _np(sc.lights.vals).astype(np.float64), _np(sc.lights.starts), _np(sc.lights.widths)
        rng = np.random.RandomState(3)
        vouched = total = 0
        for n in range(N):
            x0, y0, gx, gy = meta[n]
            gx, gy = int(gx), int(gy)
            words = vis[starts[n]:starts[n] + gx * gy].view(np.uint32)
            seg = lines[lstarts[n] + 32:lstarts[n] + lwidths[n]]
            I = lights[istarts[n]:istarts[n] + iwidths[n]]
            total += gx * gy * min(len(I), 32)
            cells = rng.choice(gx * gy, size=min(gx * gy, 300), replace=False)
            for cell in cells:
                iy, ix = divmod(int(cell), gx)
                for i in range(min(len(I), 32)):
                    if not (words[cell] >> i) & 1:
                        continue
                    vouched += 1
                    pts = np.array([x0, y0]) + (np.array([ix, iy]) + rng.uniform(0, 1, (4, 2))) * cuda.VIS_CELL
                    for P in pts:
                        U = P - I[i, :2]
                        V = seg[:, 2:] - seg[:, :2]
                        PQ = seg[:, :2] - I[i, :2]
                        den = U[0] * V[:, 1] - U[1] * V[:, 0]
                        with np.errstate(divide='ignore', invalid='ignore'):
                            sp = (PQ[:, 0] * V[:, 1] - PQ[:, 1] * V[:, 0]) / den
                            tp = (PQ[:, 0] * U[1] - PQ[:, 1] * U[0]) / den
                        blocked = (np.abs(den) > 1e-9) & (sp > -1e-3) & (sp < 1 + 1e-3) & (tp > -1e-3) & (tp < 1 + 1e-3)
                        assert not blocked.any(), (kind, n, cell, i)
        assert vouched > 0 and total > 0
        # it has to be worth its memory: a good share of the (cell, light) pairs of a box room is vouched for
        if kind == 'box':
            bits = sum(bin(int(w)).count('1') for w in vis.view(np.uint32))
            assert bits > .3 * total, (bits, total)


def test_more_than_2_to_31_texels_copies_of_an_env_render_and_move_alike():
    """64-bit texel offsets in their true regime (the reference's 32-bit accessors stop at 2^31 elements, common.h:30,43):
    a scenery of > 2^31 texels — 256 floorplans repeated on the device, scene.tiled_scenery — in which every copy of a
    floorplan holds its agent in the same pose: the last copies must render and move exactly like the first."""
    from megastep_b200 import core as core_, scene, synthetic
    if torch.cuda.mem_get_info()[1] < 120e9:
        pytest.skip('needs ~80 GB of device memory')
    gs = synthetic.sample(256, seed=5)
    base = scene.scene_arrays(gs, 1, np.random.RandomState(5))
    reps = int(np.ceil((2 ** 31 + 2 ** 27) / len(base['textures'])))
    N = 256 * reps
    s = scene.tiled_scenery(base, N)
    assert s.textures.vals.size(0) > 2 ** 31 and int(s.textures._long_starts()[-1]) > 2 ** 31
    c = core_.Core(s, res=32, fov=130., fps=10.)
    pos, ang = synthetic.spawns(gs, 1, np.random.RandomState(6))
    c.agents.positions.copy_(torch.as_tensor(pos).cuda().repeat(reps, 1, 1))
    c.agents.angles.copy_(torch.as_tensor(ang).cuda().repeat(reps, 1))
    c.agents.velocity.copy_(torch.as_tensor((2 * np.random.RandomState(7).normal(size=(256, 1, 2))).astype(np.float32)).cuda().repeat(reps, 1, 1))
    for tick in range(2):
        p = c.physics()
        r = c.render()
        torch.cuda.synchronize()
        assert torch.equal(p.progress[:256], p.progress[-256:]) and torch.equal(c.agents.positions[:256], c.agents.positions[-256:])
        for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
            a, b = getattr(r, k)[:256], getattr(r, k)[-256:]
            assert _same(a, b), f'tick {tick}: {k} of the last copies differs from the first'
        assert (r.indices >= 0).float().mean() > .9 and float(r.screen[-256:].sum()) > 0
    del c, s, r
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------------
# 3. properties at the benchmark's full size (Deathmatch 4096 x 4 x 128)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full():
    from megastep_b200 import scene, synthetic, cuda, core as core_
    N, A = 4096, 4
    gs = synthetic.sample(N, seed=1, n_unique=256)
    arrays = synthetic.tile_arrays(scene.scene_arrays(gs[:256], A, np.random.RandomState(1)), N)
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 128, 70., 10.))
    c = core_.Core(s, res=128, fov=70., fps=10.)
    pos, ang = synthetic.spawns(gs, A, np.random.RandomState(2))
    c.agents.positions.copy_(torch.as_tensor(pos))
    c.agents.angles.copy_(torch.as_tensor(ang))
    return c, gs, arrays


def test_full_size_render_properties(full):
    c, gs, arrays = full
    r1, r2 = c.render(), c.render()
    torch.cuda.synchronize()
    # idempotent and deterministic
    for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
        a, b = getattr(r1, k), getattr(r2, k)
        assert ((a == b) | (a != a) & (b != b)).all()
    idx = r1.indices
    widths = c.scenery.lines.widths[:, None, None]
    assert (idx >= -1).all() and (idx < widths).all()
    hit = idx >= 0
    assert hit.float().mean() > .95                                        # indoors, nearly every ray hits a wall
    assert (r1.distances[hit] > common.AGENT_RADIUS * .999).all() and r1.distances[~hit].isinf().all()
    assert ((r1.locations[hit] >= 0) & (r1.locations[hit] <= 1)).all()
    assert (r1.screen >= 0).all() and (r1.screen <= 1.0001).all() and (r1.screen[~hit] == 0).all()
    # a sampled slice of envs against the CPU oracle (a checksum of the whole against a checksum of the sample's kin)
    pick = np.arange(0, 4096, 512)
    st = common.read_state(c)
    for n in pick:
        sub = {k: v[n:n + 1] for k, v in st.items()}
        lo, hi = int(c.scenery.lines.starts[n]), int(c.scenery.lines.ends[n])
        tl, th = int(c.scenery.textures._long_starts()[lo]), int(c.scenery.textures._long_starts()[hi - 1] + c.scenery.textures.widths[hi - 1])
        one = dict(n_agents=4, model=arrays['model'], lines=_np(c.scenery.lines.vals[lo:hi]), line_widths=np.int32([hi - lo]),
                   lights=_np(c.scenery.lights[int(n)]), light_widths=np.int32([len(c.scenery.lights[int(n)])]),
                   textures=_np(c.scenery.textures.vals[tl:th]), tex_widths=_np(c.scenery.textures.widths[lo:hi]),
                   baked=_np(c.scenery.baked.vals[tl:th]))
        want = oracle.render(one, sub, res=128, fov=70.)
        differ, really = common.index_agreement(_np(r1.indices[n:n + 1]), want['indices'], _np(r1.distances[n:n + 1]), want['distances'])
        assert really == 0. and differ < 5e-3


def test_full_size_trajectory_bit_exact_against_reference_build(ref, full):
    """BASELINE.json's configuration (4096 envs x 4 agents x 128 rays), several ticks of kick -> physics -> render through
    both libraries from the same start: every output of every tick bit-identical, agents' states included."""
    if ref is None:
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    c, gs, arrays = full
    st0 = common.read_state(c)
    try:
        ref.initialize(common.AGENT_RADIUS, 128, 70., 10.)
        rs = common.reference_scenery(ref, arrays)
        rs.baked.vals.copy_(c.scenery.baked.vals)
        ra = common.reference_agents(ref, st0)
        rng = np.random.RandomState(4)
        for tick in range(4):
            kick = torch.as_tensor((2.5 * rng.normal(size=st0['velocity'].shape)).astype(np.float32)).cuda()
            spin = torch.as_tensor((150 * rng.normal(size=st0['angles'].shape)).astype(np.float32)).cuda()
            for agents in (c.agents, ra):
                agents.velocity.add_(kick)
                agents.angvelocity.add_(spin)
            p, rp = c.physics(), ref.physics(rs, ra)
            r, rr = c.render(), ref.render(rs, ra)
            torch.cuda.synchronize()
            assert torch.equal(p.progress, rp.progress), f'tick {tick}: progress'
            for k in ('positions', 'angles', 'velocity', 'angvelocity'):
                assert torch.equal(getattr(c.agents, k), getattr(ra, k)), f'tick {tick}: {k}'
            assert torch.equal(r.indices, rr.indices), f'tick {tick}: {(r.indices != rr.indices).float().mean():.4%} hit indices differ'
            for k in ('locations', 'dots', 'distances', 'screen'):
                assert _same(getattr(r, k), getattr(rr, k)), f'tick {tick}: {k}'
            assert torch.equal(c.scenery.lines.vals, rs.lines.vals), f'tick {tick}: drawn lines'
        assert float((p.progress < 1).float().mean()) > .02, 'the kicks should make some agents collide'
    finally:
        common.load_state(c, st0)


def test_full_size_physics_properties(full):
    c, gs, arrays = full
    before = common.read_state(c)
    # at rest nothing moves and nothing collides
    c.agents.velocity.zero_()
    c.agents.angvelocity.zero_()
    p = c.physics()
    assert (p.progress == 1).all()
    assert torch.equal(c.agents.positions.cpu(), torch.as_tensor(before['positions']))
    # with random velocities: progress in [0, 1]; collided agents are stopped; nobody tunnels out of the building
    rng = np.random.RandomState(3)
    c.agents.velocity.copy_(torch.as_tensor((4 * rng.normal(size=before['velocity'].shape)).astype(np.float32)))
    c.agents.angvelocity.copy_(torch.as_tensor(rng.uniform(-300, 300, before['angvelocity'].shape).astype(np.float32)))
    v0 = c.agents.velocity.clone()
    for _ in range(10):
        p = c.physics()
        assert ((p.progress >= 0) & (p.progress <= 1)).all()
        stopped = p.progress < 1
        assert (c.agents.velocity[stopped] == 0).all() and (c.agents.angvelocity[stopped] == 0).all()
        assert ((c.agents.angles >= -180) & (c.agents.angles < 180)).all()
        c.agents.velocity.copy_(v0)
    pos = c.agents.positions.cpu().numpy()
    hi = np.array([[g.rooms[:, 2].max(), g.rooms[:, 3].max()] for g in gs])[:, None]
    assert (pos > .5).all() and (pos < hi + .5).all()
    common.load_state(c, before)
