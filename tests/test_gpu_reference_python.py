"""Parity against the reference's OWN Python + its OWN CUDA build (tests/common.py::reference_package).

The checker here is the unmodified megastep/{core,modules,scene}.py and megastep/demo/envs/{explorer,deathmatch}.py of
the reference (pip-installed by oracle/build_ref.sh into baseline/_ref), driving the reference's own extension
(oracle/_ref/megastepcuda*.so). Compared with it, from identical state and with no resynchronisation in between:

  * `modules.FusedStep` — the path bench.py times — at BASELINE.json's full sizes (Deathmatch 4096 x 4 x 128 and
    Explorer 4096 x 1 x 64) over 8 ticks of random actions: `progress`, the agents' state, the five Render tensors and
    the RGB / Depth / IMU observations — pooled ones (subsample 2 ... 16) included — bit for bit (the north star asks
    bit-exact indices / collision flags and 1e-5 abs on positions / depth);
  * this package's unfused `modules.*` (MomentumMovement, render, RGB, Depth, IMU) the same way;
  * `envs.Explorer` / `envs.Deathmatch` — rules only: both sides respawn from the same fixed table.
"""
import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pkg():
    p = common.reference_package()
    if p is None:
        pytest.skip('oracle/_ref (extension + site) not built: needs /root/reference at build time')
    return p


def _same(a, b):
    return bool(((a == b) | (a != a) & (b != b)).all())


def _frac(a, b):
    return float(((a == b) | (a != a) & (b != b)).float().mean())


def _scene(n_envs, n_agents, unique=256, seed=1):
    from megastep_b200 import scene, synthetic
    gs = synthetic.sample(n_envs, seed=seed, n_unique=unique)
    arrays = synthetic.tile_arrays(scene.scene_arrays(gs[:min(unique, n_envs)], n_agents, np.random.RandomState(seed)), n_envs)
    pos, ang = synthetic.spawns(gs, n_agents, np.random.RandomState(seed + 1))
    N = n_envs
    st = dict(angles=ang, positions=pos, angvelocity=np.zeros((N, n_agents), np.float32), velocity=np.zeros((N, n_agents, 2), np.float32))
    return gs, arrays, st


def _pair(pkg, arrays, st, res, fov):
    """(our core, the reference's core) over the same scene, same baked light map, same agents."""
    from megastep_b200 import cuda, core as core_, scene
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, res, fov, 10.))      # bit-exact to ref.bake: test_gpu_parity
    c = core_.Core(s, res=res, fov=fov, fps=10.)
    common.load_state(c, st)
    rc = common.reference_core(pkg, arrays, st, res, fov)
    rc.scenery.baked.vals.copy_(s.baked.vals)
    return c, rc


CONFIGS = [
    # name, envs, agents, res, fov, subsample, ticks
    ('deathmatch-full', 4096, 4, 128, 70., 1, 8),          # BASELINE.json configs[2]: what bench.py times
    ('explorer-full', 4096, 1, 64, 130., 1, 8),            # BASELINE.json configs[1]
    ('deathmatch-demo', 256, 4, 512, 70., 4, 8),           # demo-faithful: render at 4x, pool by 4 (deathmatch.py:26-28)
    ('explorer-demo', 256, 1, 256, 130., 4, 8),            # explorer.py:13-15
    ('three-agents-ragged-res', 64, 3, 48, 100., 1, 8),
    ('six-agents', 64, 6, 96, 90., 2, 8),
    ('pool-by-8', 32, 2, 256, 70., 8, 4),
    ('pool-by-16', 32, 2, 256, 70., 16, 4),
]


@pytest.mark.parametrize('name,N,A,res,fov,sub,ticks', CONFIGS)
def test_fused_step_bit_exact_against_reference_python(pkg, name, N, A, res, fov, sub, ticks):
    from megastep_b200 import modules
    gs, arrays, st = _scene(N, A)
    c, rc = _pair(pkg, arrays, st, res, fov)
    fused = modules.FusedStep(c, subsample=sub, raw=True)
    mover, rgb, depth, imu = (pkg.modules.MomentumMovement(rc), pkg.modules.RGB(rc, subsample=sub),
                              pkg.modules.Depth(rc, subsample=sub), pkg.modules.IMU(rc))
    rng = np.random.RandomState(3)
    collided = 0.
    for tick in range(ticks):
        actions = torch.as_tensor(rng.randint(0, 7, (N, A)).astype(np.int32)).cuda()
        p = mover(pkg.arrdict.arrdict(actions=actions))                      # reference: modules.py:106-118 -> cuda.physics
        r = pkg.modules.render(rc)                                           # modules.py:126-136 -> cuda.render
        want = dict(rgb=rgb(r), d=depth(r), imu=imu())                       # modules.py:211-224, 170-184, 263-270
        out = fused(actions)
        torch.cuda.synchronize()
        where = f'{name} tick {tick}'
        assert torch.equal(out.progress < 1, p.progress < 1), f'{where}: collision flags differ'
        assert torch.equal(out.progress, p.progress), f'{where}: progress'
        for k in ('positions', 'angles', 'velocity', 'angvelocity'):
            a, b = getattr(c.agents, k), getattr(rc.agents, k)
            torch.testing.assert_close(a, b, rtol=0, atol=1e-5, msg=lambda m: f'{where} {k}: {m}')
            assert torch.equal(a, b), f'{where}: {k} within 1e-5 but bit-exact on only {_frac(a, b):.6%}'
        assert torch.equal(out.render.indices, r.indices.squeeze(2)), \
            f'{where}: {(out.render.indices != r.indices.squeeze(2)).float().mean():.4%} of hit indices differ'
        for k in ('locations', 'dots', 'distances'):
            assert _same(getattr(out.render, k), r[k].squeeze(2)), f'{where}: {k}'
        assert _same(out.render.screen, r.screen.squeeze(3).permute(0, 1, 3, 2)), f'{where}: screen'
        assert torch.equal(c.scenery.lines.vals, rc.scenery.lines.vals), f'{where}: drawn lines'
        assert out.obs.rgb.shape == want['rgb'].shape and out.obs.d.shape == want['d'].shape and out.obs.imu.shape == want['imu'].shape
        assert torch.equal(out.obs.imu, want['imu']), f'{where}: imu bit-exact on {_frac(out.obs.imu, want["imu"]):.6%}'
        # (pooled heads, sub > 1: the kernel adds the `sub` pixels of a pooled pixel in the order ATen's mean() does)
        torch.testing.assert_close(out.obs.d, want['d'], rtol=0, atol=1e-6)
        torch.testing.assert_close(out.obs.rgb, want['rgb'], rtol=0, atol=1e-6)
        assert torch.equal(out.obs.d, want['d']), f'{where}: depth (subsample {sub}) bit-exact on {_frac(out.obs.d, want["d"]):.6%}'
        assert torch.equal(out.obs.rgb, want['rgb']), f'{where}: rgb (subsample {sub}) bit-exact on {_frac(out.obs.rgb, want["rgb"]):.6%}'
        collided = max(collided, float((p.progress < 1).float().mean()))
    assert collided > 0 or N * A * ticks < 2000, 'random actions should make some agents run into something'


@pytest.mark.parametrize('name,N,A,res,fov,sub,ticks', [CONFIGS[2], CONFIGS[4]])
def test_unfused_modules_bit_exact_against_reference_python(pkg, name, N, A, res, fov, sub, ticks):
    """This package's own modules.py (the API the north star keeps) next to the reference's, tick for tick."""
    from megastep_b200 import modules
    from megastep_b200.arrdict import arrdict
    gs, arrays, st = _scene(N, A, seed=5)
    c, rc = _pair(pkg, arrays, st, res, fov)
    ours = (modules.MomentumMovement(c), modules.RGB(c, subsample=sub), modules.Depth(c, subsample=sub), modules.IMU(c))
    theirs = (pkg.modules.MomentumMovement(rc), pkg.modules.RGB(rc, subsample=sub), pkg.modules.Depth(rc, subsample=sub), pkg.modules.IMU(rc))
    rng = np.random.RandomState(4)
    for tick in range(ticks):
        actions = torch.as_tensor(rng.randint(0, 7, (N, A)).astype(np.int32)).cuda()
        p, rp = ours[0](arrdict(actions=actions)), theirs[0](pkg.arrdict.arrdict(actions=actions))
        r, rr = modules.render(c), pkg.modules.render(rc)
        torch.cuda.synchronize()
        assert torch.equal(p.progress, rp.progress), f'{name} tick {tick}: progress'
        for k in ('positions', 'angles', 'velocity', 'angvelocity'):
            assert torch.equal(getattr(c.agents, k), getattr(rc.agents, k)), f'{name} tick {tick}: {k}'
        assert sorted(r.keys()) == sorted(rr.keys())
        for k in r:
            assert r[k].shape == rr[k].shape and _same(r[k], rr[k]), f'{name} tick {tick}: {k}'
        assert torch.equal(ours[1](r), theirs[1](rr)) and torch.equal(ours[2](r), theirs[2](rr)) and torch.equal(ours[3](), theirs[3]())


# ----------------------------------------------------------------------------------------------------------------------
# the two demo environments, rules only
# ----------------------------------------------------------------------------------------------------------------------
class FixedSpawns:
    """Stands in for RandomSpawns (modules.py:295-326) on both sides: flagged agents go to the next entry of a table
    drawn once by the test, velocities zeroed — the reference's semantics with the random draw taken out."""

    def __init__(self, core, positions, angles):
        self.core, self.positions, self.angles, self.calls = core, positions, angles, 0

    def __call__(self, reset):
        k = self.calls % self.positions.shape[2]
        self.calls += 1
        ag = self.core.agents
        ag.angles[reset] = self.angles[:, :, k][reset]
        ag.positions[reset] = self.positions[:, :, k][reset]
        ag.velocity[reset] = 0.
        ag.angvelocity[reset] = 0.


def _geometries(n, with_masks=True):
    from megastep_b200 import synthetic
    gs = synthetic.sample(n, seed=9, n_unique=n, with_masks=with_masks)
    for g in gs:
        g['res'] = np.array(g['res'])            # a 0-d array, as np.load gives the reference: it multiplies a shape TUPLE by it (deathmatch.py:44)
    return gs


def _spawn_table(gs, A, k=4, seed=3):
    from megastep_b200 import synthetic
    rng = np.random.RandomState(seed)
    pos, ang = zip(*[synthetic.spawns(gs, A, rng) for _ in range(k)])
    return torch.as_tensor(np.stack(pos, 2)).cuda(), torch.as_tensor(np.stack(ang, 2)).cuda()


def test_explorer_env_against_the_reference_env(pkg):
    """envs.Explorer (explorer.py:8-107): same scene (both builders consume the same seeded numpy stream), same spawns,
    same actions -> same observations, rewards and resets over 12 steps."""
    from megastep_b200 import envs
    from megastep_b200.arrdict import arrdict
    N = 24
    gs = _geometries(N)
    np.random.seed(11)
    ours = envs.Explorer(gs)
    np.random.seed(11)
    pkg.explorer.cubicasa.sample = lambda n: gs
    theirs = pkg.explorer.Explorer(N)
    assert torch.equal(ours.core.scenery.lines.vals, theirs.core.scenery.lines.vals)
    assert torch.equal(ours.core.scenery.textures.vals, theirs.core.scenery.textures.vals)
    assert torch.equal(ours.core.scenery.baked.vals, theirs.core.scenery.baked.vals)
    pos, ang = _spawn_table(gs, 1)
    ours._respawner, theirs._respawner = FixedSpawns(ours.core, pos, ang), FixedSpawns(theirs.core, pos, ang)
    rng = np.random.RandomState(2)
    a, b = ours.reset(), theirs.reset()
    for tick in range(12):
        torch.cuda.synchronize()
        for k in ('rgb', 'd', 'imu'):
            assert torch.equal(a.obs[k], b.obs[k]), f'tick {tick}: obs.{k}'
        assert torch.equal(a.reset, b.reset), f'tick {tick}: reset'
        # the reference indexes its `seen` table with the -1 of rays that hit nothing, i.e. marks the batch's LAST texel
        # (explorer.py:45-47); this package does not — the last env's reward may differ by that one texel
        miss = bool((theirs.core.scenery.lines.widths.sum() > 0) and (pkg.modules.render(theirs.core).indices < 0).any())
        upto = N - 1 if miss else N
        assert torch.equal(a.reward[:upto], b.reward[:upto]), f'tick {tick}: reward'
        assert torch.equal(ours._ledger.potential[:upto], theirs._potential[:upto])
        actions = torch.as_tensor(rng.randint(0, 7, (N, 1))).cuda()
        if tick == 6:                                                        # force a few resets through the rule itself
            ours._lengths[:5] += 1000
            theirs._lengths[:5] += 1000
        a, b = ours.step(arrdict(actions=actions)), theirs.step(pkg.arrdict.arrdict(actions=actions))
    assert float(a.reward.abs().sum()) > 0


def test_deathmatch_env_against_the_reference_env(pkg):
    """envs.Deathmatch (deathmatch.py:20-119): agents packed into view of each other so that shots land."""
    from megastep_b200 import envs
    from megastep_b200.arrdict import arrdict
    N, A = 16, 4
    gs = _geometries(N)
    np.random.seed(12)
    ours = envs.Deathmatch(gs, A)
    np.random.seed(12)
    pkg.deathmatch.cubicasa.sample = lambda n: gs
    theirs = pkg.deathmatch.Deathmatch(4 * N, A)
    assert torch.equal(ours.core.scenery.textures.vals, theirs.core.scenery.textures.vals)
    pos, ang = _spawn_table(gs, A)
    # everybody in one room, looking at the room's middle
    rooms = np.stack([g.rooms[np.argmax((g.rooms[:, 2] - g.rooms[:, 0]) * (g.rooms[:, 3] - g.rooms[:, 1]))] for g in gs])
    mid = torch.as_tensor(np.stack([(rooms[:, 0] + rooms[:, 2]) / 2, (rooms[:, 1] + rooms[:, 3]) / 2], -1)).float().cuda()
    off = torch.as_tensor(np.random.RandomState(1).uniform(-.8, .8, tuple(pos.shape))).float().cuda()
    pos = mid[:, None, None, :] + off
    d = mid[:, None, None, :] - pos
    ang = torch.rad2deg(torch.atan2(d[..., 1], d[..., 0]))
    ours._spawner, theirs._spawner = FixedSpawns(ours.core, pos, ang), FixedSpawns(theirs.core, pos, ang)
    rng = np.random.RandomState(2)
    a, b = ours.reset(), theirs.reset()
    hits = 0.
    for tick in range(12):
        torch.cuda.synchronize()
        for k in ('rgb', 'd', 'imu', 'health'):
            assert a.obs[k].shape == b.obs[k].shape and _same(a.obs[k], b.obs[k]), f'tick {tick}: obs.{k}'
        assert torch.equal(a.reset, b.reset) and torch.equal(a.reward, b.reward), f'tick {tick}'
        assert _same(ours._health, theirs._health) and _same(ours._damage, theirs._damage)
        assert torch.equal(ours.matchings, theirs.matchings)
        hits += float(a.reward.sum())
        if tick == 5:                                                        # a few deaths, through the rule itself
            ours._health[:3] = -1.
            theirs._health[:3] = -1.
        actions = torch.as_tensor(rng.randint(0, 7, (N * A, 1))).cuda()
        a, b = ours.step(arrdict(actions=actions)), theirs.step(pkg.arrdict.arrdict(actions=actions))
    assert hits > 0, 'the packed agents should land some shots'


# ----------------------------------------------------------------------------------------------------------------------
# the envs' fused paths (device kernels for the rules) against their op-for-op restatements
# ----------------------------------------------------------------------------------------------------------------------
def test_fused_explorer_equals_the_unfused_one():
    """envs.Explorer(fused=True) — msb_move, RGBD heads, the bit ledger kernels — against fused=False (itself checked
    against the reference's env above): observations, rewards, resets, potentials and the set of seen texels."""
    from megastep_b200 import envs
    from megastep_b200.arrdict import arrdict
    N = 24
    gs = _geometries(N)
    np.random.seed(21)
    a_env = envs.Explorer(gs, fused=True)
    np.random.seed(21)
    b_env = envs.Explorer(gs, fused=False)
    assert a_env.fused and not b_env.fused
    pos, ang = _spawn_table(gs, 1)
    a_env._respawner, b_env._respawner = FixedSpawns(a_env.core, pos, ang), FixedSpawns(b_env.core, pos, ang)
    rng = np.random.RandomState(2)
    a, b = a_env.reset(), b_env.reset()
    for tick in range(14):
        torch.cuda.synchronize()
        for k in ('rgb', 'd', 'imu'):
            assert torch.equal(a.obs[k], b.obs[k]), f'tick {tick}: obs.{k}'
        assert torch.equal(a.reset, b.reset) and torch.equal(a.reward, b.reward), f'tick {tick}'
        assert torch.equal(a_env._ledger.potential, b_env._ledger.potential)
        assert torch.equal(a_env._ledger.seen, b_env._ledger.seen), f'tick {tick}: seen texels'
        actions = torch.as_tensor(rng.randint(0, 7, (N, 1))).cuda()
        if tick in (5, 9):                                                   # resets through the rule itself, twice
            a_env._lengths[3:9] += 1000
            b_env._lengths[3:9] += 1000
        a, b = a_env.step(arrdict(actions=actions)), b_env.step(arrdict(actions=actions))
    assert float(a.reward.abs().sum()) > 0 and int(a_env._ledger.seen.sum()) > 100
    st = a_env.state(2)
    assert st.seen.shape == b_env.state(2).seen.shape


def test_fused_deathmatch_equals_the_unfused_one():
    from megastep_b200 import envs
    from megastep_b200.arrdict import arrdict
    N, A = 16, 4
    gs = _geometries(N)
    np.random.seed(22)
    a_env = envs.Deathmatch(gs, A, fused=True)
    np.random.seed(22)
    b_env = envs.Deathmatch(gs, A, fused=False)
    pos, ang = _spawn_table(gs, A)
    rooms = np.stack([g.rooms[np.argmax((g.rooms[:, 2] - g.rooms[:, 0]) * (g.rooms[:, 3] - g.rooms[:, 1]))] for g in gs])
    mid = torch.as_tensor(np.stack([(rooms[:, 0] + rooms[:, 2]) / 2, (rooms[:, 1] + rooms[:, 3]) / 2], -1)).float().cuda()
    pos = mid[:, None, None, :] + torch.as_tensor(np.random.RandomState(1).uniform(-.8, .8, tuple(pos.shape))).float().cuda()
    d = mid[:, None, None, :] - pos
    ang = torch.rad2deg(torch.atan2(d[..., 1], d[..., 0]))
    a_env._spawner, b_env._spawner = FixedSpawns(a_env.core, pos, ang), FixedSpawns(b_env.core, pos, ang)
    rng = np.random.RandomState(2)
    a, b = a_env.reset(), b_env.reset()
    hits = 0.
    for tick in range(12):
        torch.cuda.synchronize()
        for k in ('rgb', 'd', 'imu', 'health'):
            assert _same(a.obs[k], b.obs[k]), f'tick {tick}: obs.{k}'
        assert torch.equal(a.reset, b.reset) and torch.equal(a.reward, b.reward), f'tick {tick}'
        assert _same(a_env._health, b_env._health) and _same(a_env._damage, b_env._damage)
        assert torch.equal(a_env.matchings, b_env.matchings)
        hits += float(a.reward.sum())
        if tick == 5:
            a_env._health[:3] = -1.
            b_env._health[:3] = -1.
        if tick == 7:                                                        # somebody wanders out of the building
            a_env.core.agents.positions[5, 1] = -3.
            b_env.core.agents.positions[5, 1] = -3.
        actions = torch.as_tensor(rng.randint(0, 7, (N * A, 1))).cuda()
        a, b = a_env.step(arrdict(actions=actions)), b_env.step(arrdict(actions=actions))
    assert hits > 0


def test_device_respawns_move_only_the_flagged_agents():
    """cuda.env_respawn (RandomSpawns, modules.py:312-326, with the draw on the device): flagged agents land on one of
    THEIR spawn points with zeroed velocities, the others are untouched; with explicit choices it equals the gather."""
    from megastep_b200 import cuda, modules, scene
    gs = _geometries(6)
    arrays = scene.scene_arrays(gs, 3, np.random.RandomState(1))
    st = common.random_state(gs, 3, seed=4)
    c = common.to_device(arrays, st, 64, 90.)
    spawner = modules.RandomSpawns(gs, c, n_spawns=20, fused=True, seed=5)
    before = common.read_state(c)
    reset = torch.as_tensor(np.random.RandomState(1).rand(6, 3) < .5).cuda()
    spawner(reset)
    after = common.read_state(c)
    m = reset.cpu().numpy()
    for k in before:
        assert np.array_equal(before[k][~m], after[k][~m]), k
    assert (after['velocity'][m] == 0).all() and (after['angvelocity'][m] == 0).all()
    spawns = spawner._spawns.positions.float().cpu().numpy()
    picked = set()
    for n, a in zip(*m.nonzero()):
        hit = (np.abs(spawns[n, a] - after['positions'][n, a]).sum(-1) == 0).nonzero()[0]
        assert len(hit) > 0
        picked.add(int(hit[0]))
    assert len(picked) > 2, 'the draws should not all pick the same spawn'
    # explicit choices
    choices = torch.as_tensor(np.random.RandomState(2).randint(0, 20, (6, 3)).astype(np.int32)).cuda()
    cuda.env_respawn(c.scenery, c.agents, torch.ones_like(reset), spawner._flat[0], spawner._flat[1], 0, 0, choices)
    want = spawner._flat[0].gather(2, choices.long()[..., None, None].expand(-1, -1, 1, 2)).squeeze(2)
    assert torch.equal(c.agents.positions, want)
    assert torch.equal(c.agents.angles, spawner._flat[1].gather(2, choices.long()[..., None]).squeeze(2))
