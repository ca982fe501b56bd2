"""Host-side logic, on CPU: ragged containers, dict containers, scene construction, geometry, the module-level
maths, the C-ABI library's exported surface. No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import common
from megastep_b200 import arrdict as ad
from megastep_b200 import cuda, dotdict as dd, geometry, modules, ragged, scene, spaces, synthetic, toys


# ---- the reference's own ragged tests (megastep/ragged.py:77-103), restated against this package -------------
def test_ragged():
    vals = torch.as_tensor([0, 1, 2, 3, 4, 5]).float()
    widths = torch.as_tensor([3, 1, 2]).int()
    r = ragged.Ragged(vals, widths)
    assert isinstance(r, cuda.Ragged1D)
    torch.testing.assert_close(r[1], torch.tensor([3.]))
    torch.testing.assert_close(r[-1], torch.tensor([4., 5.]))
    torch.testing.assert_close(r[:2].vals, torch.tensor([0., 1., 2., 3.]))
    torch.testing.assert_close(r[:2].widths, torch.tensor([3, 1], dtype=torch.int32))
    torch.testing.assert_close(r[1:].vals, torch.tensor([3., 4., 5.]))
    torch.testing.assert_close(r[1:].widths, torch.tensor([1, 2], dtype=torch.int32))


def test_ragged_numpy():
    r = ragged.RaggedNumpy(np.array([0, 1, 2, 3, 4, 5]), np.array([3, 1, 2]))
    np.testing.assert_allclose(r[1], [3])
    np.testing.assert_allclose(r[-1], [4, 5])
    np.testing.assert_allclose(r[:2].vals, [0, 1, 2, 3])
    np.testing.assert_allclose(r[:2].widths, [3, 1])
    np.testing.assert_allclose(r[1:].vals, [3, 4, 5])
    np.testing.assert_allclose(r[1:].widths, [1, 2])


def test_ragged_metadata_and_roundtrip():
    vals = torch.arange(12.).reshape(6, 2)
    r = ragged.Ragged(vals, torch.tensor([2, 0, 3, 1], dtype=torch.int32))
    assert isinstance(r, cuda.Ragged2D)
    assert r.starts.dtype == r.ends.dtype == r.inverse.dtype == torch.int32
    assert r.starts.tolist() == [0, 2, 2, 5] and r.ends.tolist() == [2, 2, 5, 6]
    assert r._long_starts().dtype == torch.int64 and r._long_starts().tolist() == [0, 2, 2, 5]
    n = r.numpyify()
    assert isinstance(n, ragged.RaggedNumpy) and n.starts.tolist() == [0, 2, 2, 5]
    back = n.torchify()
    assert isinstance(back, cuda.Ragged2D) and torch.equal(back.vals, vals)
    c = r.clone()
    c.vals[0, 0] = 99.
    assert r.vals[0, 0] == 0.


def test_ragged_validation_errors():
    with pytest.raises(RuntimeError):
        cuda.Ragged1D(torch.zeros(5), torch.tensor([3, 1], dtype=torch.int32))       # widths do not sum to len
    with pytest.raises(RuntimeError):
        cuda.Ragged1D(torch.zeros(4), torch.tensor([3, 1]))                           # int64 widths
    with pytest.raises(RuntimeError):
        cuda.Ragged2D(torch.zeros(4), torch.tensor([3, 1], dtype=torch.int32))       # wrong ndim
    with pytest.raises(RuntimeError):
        cuda.Ragged1D(torch.zeros(8)[::2], torch.tensor([3, 1], dtype=torch.int32))  # non-contiguous


def test_cpu_tensors_are_refused_by_the_kernels():
    # the reference's TensorProxy insists on CUDA tensors (common.h:12-14,33-37); there is no CPU path here either
    z = torch.zeros
    with pytest.raises(RuntimeError, match='CUDA'):
        cuda.Agents(z(2, 1), z(2, 1, 2), z(2, 1), z(2, 1, 2))
    arrays = scene.scene_arrays([toys.box()], 1, np.random.RandomState(0))
    with pytest.raises(RuntimeError, match='CUDA'):
        scene.upload(arrays, 'cpu')


def test_render_result_exposes_exactly_the_reference_attributes():
    r = cuda.Render(*[torch.zeros(1)] * 5)
    assert sorted(k for k in dir(r) if not k.startswith('_')) == ['distances', 'dots', 'indices', 'locations', 'screen']
    assert sorted(modules.unpack(r).keys()) == ['distances', 'dots', 'indices', 'locations', 'screen']
    assert [k for k in dir(cuda.Physics(torch.zeros(1))) if not k.startswith('_')] == ['progress']


# ---- C ABI -------------------------------------------------------------------------------------------------------
def test_library_exports_every_symbol_the_header_declares():
    header = open(os.path.join(common.ROOT, 'include', 'megastep_b200.h')).read()
    declared = set(re.findall(r'\b(msb_[a-z_]+)\s*\(', header))
    assert {'msb_physics', 'msb_render', 'msb_step', 'msb_bake', 'msb_build_table', 'msb_build_visibility', 'msb_params_init'} <= declared
    lib = ctypes.CDLL(cuda.library_path())
    for name in declared:
        assert hasattr(lib, name), f'{name} is declared in the header but not exported'
    assert lib.msb_abi_version() == 5


def test_params_init_matches_reference_formula_and_validates():
    from oracle import oracle
    for fov in (60., 70., 130.):
        p = cuda.make_params(common.AGENT_RADIUS, 128, fov, 10.)
        assert p.half_screen == pytest.approx(oracle.half_screen(fov), abs=0)   # same host expression
        assert p.res == 128 and p.fps == 10.
    with pytest.raises(RuntimeError, match='fov'):
        cuda.make_params(.1, 64, 180., 10.)
    with pytest.raises(RuntimeError, match='res'):
        cuda.make_params(.1, 0, 90., 10.)


def test_struct_layouts_match_the_header(tmp_path):
    """Compile the header with the C compiler and compare every struct's size and field offsets with the ctypes
    mirrors in megastep_b200/cuda.py."""
    import subprocess
    structs = {'msb_params': cuda.Params, 'msb_scenery': cuda._Scenery, 'msb_agents': cuda._Agents,
               'msb_render_out': cuda._RenderOut, 'msb_obs_out': cuda._ObsOut, 'msb_movement': cuda._Movement,
               'msb_workspace': cuda._Workspace}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "megastep_b200.h"', 'int main(void) {']
    for cname, mirror in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, *_ in mirror._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(common.ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in structs.items():
        assert int(got[cname]) == ctypes.sizeof(mirror), cname
        for fname, *_ in mirror._fields_:
            assert int(got[f'{cname}.{fname}']) == getattr(mirror, fname).offset, f'{cname}.{fname}'


# ---- containers --------------------------------------------------------------------------------------------------
def test_dotdict_and_arrdict_behaviour():
    d = dd.dotdict(a=1, b=dd.dotdict(c=2))
    assert d.a == 1 and d.b.c == 2 and 'a' in dir(d)
    assert dd.mapping(lambda x: x + 1)(d).b.c == 3
    assert dd.leaves(d) == [1, 2]
    with pytest.raises(AttributeError):
        d.nope
    x = ad.arrdict(p=torch.arange(6.).reshape(3, 2), q=torch.arange(3.))
    assert x[1].p.tolist() == [2., 3.] and x[1].q.item() == 1.
    assert (x + 1).q.tolist() == [1., 2., 3.] and (x * x).q.tolist() == [0., 1., 4.]
    assert x.shape == ad.arrdict(p=torch.Size([3, 2]), q=torch.Size([3]))
    y = x.clone()
    y[0] = ad.arrdict(p=torch.tensor([9., 9.]), q=torch.tensor(9.))
    assert y.p[0].tolist() == [9., 9.] and x.p[0].tolist() == [0., 1.]
    with pytest.raises(ValueError):
        x.z = 1
    s = ad.stack([x, x])
    assert s.p.shape == (2, 3, 2)
    c = ad.cat([x, x])
    assert c.q.shape == (6,)
    n = ad.numpyify(x)
    assert isinstance(n.p, np.ndarray)
    t = ad.torchify(ad.arrdict(i=np.arange(3), f=np.ones(2), b=np.array([True])))
    assert t.i.dtype == torch.int32 and t.f.dtype == torch.float32 and t.b.dtype == torch.bool
    assert 'p' in str(x)


def test_spaces():
    assert spaces.MultiImage(2, 3, 1, 64).shape == (2, 3, 1, 64)
    assert spaces.MultiVector(2, 3).shape == (2, 3)
    assert spaces.MultiDiscrete(4, 7).shape == (4, 7)


# ---- scene construction -------------------------------------------------------------------------------------------
def test_agent_model_and_colors():
    m = scene.agent_model()
    assert m.shape == (8, 2, 2)
    np.testing.assert_allclose(m[:, 1], np.roll(m[:, 0], -1, 0))           # a closed outline
    assert np.abs(m).max() == pytest.approx(.075)
    assert np.linalg.norm(m.reshape(-1, 2), axis=1).max() < common.AGENT_RADIUS   # inside the near plane
    c = scene.agent_colors()
    assert c.shape == (8, 3) and c[1].tolist() == [0., .5, 0.] and c[3].tolist() == [1., 0., 0.]
    assert scene.to_rgb('#c185ae') == pytest.approx((0xc1 / 255, 0x85 / 255, 0xae / 255))


def test_scene_arrays_layout():
    gs = [toys.box(), toys.column()]
    a = scene.scene_arrays(gs, 2, np.random.RandomState(0))
    assert a['line_widths'].tolist() == [16 + 4, 16 + 4] and a['light_widths'].tolist() == [1, 4]
    assert a['lines'].shape == (40, 2, 2) and a['lights'].shape == (5, 3)
    assert a['tex_widths'].sum() == len(a['textures'])
    # agent lines are 2 texels each; a 5 m wall is 100 texels (TEXTURE_RES = 5 cm)
    assert (a['tex_widths'][:16] == 2).all() and (a['tex_widths'][16:20] == 100).all()
    # agent texels keep their flat colours (pattern forced to 1), in linear space
    np.testing.assert_allclose(a['textures'][2:4], np.tile(np.array([0., .5, 0.]) ** 2.2, (2, 1)), atol=1e-6)
    assert ((a['lights'][:, 2] >= .5) & (a['lights'][:, 2] < 2.)).all()
    b = scene.scene_arrays(gs, 2, np.random.RandomState(0))
    np.testing.assert_array_equal(a['textures'], b['textures'])            # deterministic in the RandomState


def test_box_geometry_matches_reference_docs():
    g = toys.box(5)
    # corners at 1..6 (toys.py:7-9 with MARGIN = 1), light in the centre
    assert sorted(set(np.round(g.walls.reshape(-1), 6))) == [1., 6.]
    assert g.lights.tolist() == [[3.5, 3.5]]
    m = g.masks
    assert m.shape == (36, 36) and set(np.unique(m)) == {-1, 0, 1}
    i, j = geometry.indices(np.array([3.5, 3.5]), m.shape, g.res)
    assert m[i, j] == 1                                                     # room interior
    i, j = geometry.indices(np.array([6., 3.5]), m.shape, g.res)
    assert m[i, j] == -1                                                    # the right wall
    assert m[0, 0] == 0                                                     # outside
    xy = geometry.centers(np.array([[i, j]]), m.shape, g.res)
    assert np.abs(xy - [6., 3.5]).max() <= g.res


def test_synthetic_floorplans_are_cubicasa_shaped_and_deterministic():
    gs = synthetic.sample(64, seed=7)
    W = np.array([len(g.walls) for g in gs])
    I = np.array([len(g.lights) for g in gs])
    assert 230 < W.mean() < 320 and W.min() >= 60 and W.max() <= 600
    assert 17 < I.mean() < 25
    assert min(g.walls.min() for g in gs) > 0                               # inside the +quadrant, with margin
    again = synthetic.sample(64, seed=7)
    np.testing.assert_array_equal(gs[13].walls, again[13].walls)
    pos, ang = synthetic.spawns(gs, 4, np.random.RandomState(0))
    assert pos.shape == (64, 4, 2) and ang.shape == (64, 4)
    for g, p in zip(gs, pos):
        inside = ((p[:, None, 0] > g.rooms[None, :, 0]) & (p[:, None, 0] < g.rooms[None, :, 2]) &
                  (p[:, None, 1] > g.rooms[None, :, 1]) & (p[:, None, 1] < g.rooms[None, :, 3])).any(1)
        assert inside.all()


def test_tile_arrays_repeats_envs():
    gs = synthetic.sample(3, seed=2)
    a = scene.scene_arrays(gs, 1, np.random.RandomState(0))
    b = synthetic.tile_arrays(a, 7)
    assert b['line_widths'].tolist() == a['line_widths'][[0, 1, 2, 0, 1, 2, 0]].tolist()
    assert b['tex_widths'].sum() == len(b['textures'])
    L0 = a['line_widths'][0]
    np.testing.assert_array_equal(b['lines'][-L0:], a['lines'][:L0])


# ---- module maths ----------------------------------------------------------------------------------------------
def test_frames_are_inverse_rotations():
    ang = torch.tensor([[0., 90., -45.]])
    p = torch.randn(1, 3, 2)
    torch.testing.assert_close(modules.to_local_frame(ang, modules.to_global_frame(ang, p)), p, atol=1e-6, rtol=0)
    g = modules.to_global_frame(torch.tensor([90.]), torch.tensor([[1., 0.]]))
    torch.testing.assert_close(g, torch.tensor([[0., 1.]]), atol=1e-6, rtol=0)


def test_downsample_and_observation_heads_on_a_fake_render():
    class FakeCore:
        res, n_agents, agent_radius = 8, 1, common.AGENT_RADIUS
    r = ad.arrdict(distances=torch.tensor([.5, 1., 2., 4., 8., 16., float('inf'), common.AGENT_RADIUS]).reshape(1, 1, 1, 8),
                   screen=torch.arange(24.).reshape(1, 1, 3, 1, 8))
    assert modules.downsample(r.screen, 4).shape == (1, 1, 3, 1, 2, 4)
    d = modules.Depth(FakeCore(), subsample=2)(r)
    assert d.shape == (1, 1, 1, 1, 4)
    full = 1 - ((r.distances - common.AGENT_RADIUS) / 10).clamp(0, 1)
    torch.testing.assert_close(d.reshape(-1), full.reshape(4, 2).mean(-1))
    assert full.reshape(-1)[6] == 0 and full.reshape(-1)[7] == 1              # infinity -> 0, near plane -> 1
    rgb = modules.RGB(FakeCore(), subsample=4)(r)
    assert rgb.shape == (1, 1, 3, 1, 2)
    torch.testing.assert_close(rgb[0, 0, 0, 0], torch.tensor([1.5, 5.5]))


def test_spatial_table_is_a_padded_permutation_with_tight_boxes():
    """cuda._occluder_table on CPU tensors: every env's static lines appear exactly once (in runs of 16, padded per
    env), each run's box bounds its segments, and every row carries its line's texel offset / count and index."""
    from megastep_b200 import scene, synthetic
    A = 2
    gs = synthetic.sample(5, seed=3)
    arrays = scene.scene_arrays(gs, A, np.random.RandomState(3))
    lines = cuda.Ragged3D(torch.as_tensor(arrays['lines']), torch.as_tensor(arrays['line_widths']))
    tw = torch.as_tensor(arrays['tex_widths'])
    ts = tw.long().cumsum(0) - tw.long()
    AF = A * len(arrays['model'])
    occ, occ_starts, boxes, box_starts, meta, rec = cuda._occluder_table(lines, AF, 16, tw, ts)
    assert occ.shape[0] == rec.shape[0] == 16 * boxes.shape[0]
    for n in range(5):
        L = int(lines.widths[n]); W = L - AF; nb = (W + 15) // 16
        assert int(occ_starts[n]) == 16 * int(box_starts[n])
        rows = slice(int(occ_starts[n]), int(occ_starts[n]) + 16 * nb)
        ids = rec[rows, 3].numpy()
        assert sorted(ids[:W]) == list(range(AF, L)) and (ids[W:] == -1).all()
        mine = lines[n].reshape(-1, 4)
        assert torch.equal(occ[rows][:W], mine[ids[:W]])
        g = int(lines.starts[n]) + ids[:W]
        got_ts = (rec[rows, 1][:W].long() << 32) | (rec[rows, 0][:W].long() & 0xffffffff)
        assert torch.equal(got_ts, ts[g]) and torch.equal(rec[rows, 2][:W], tw[g])
        for b in range(nb):
            seg = occ[rows][16 * b:min(16 * b + 16, W)]
            bx = boxes[int(box_starts[n]) + b]
            assert bx[0] == seg[:, [0, 2]].min() and bx[2] == seg[:, [0, 2]].max()
            assert bx[1] == seg[:, [1, 3]].min() and bx[3] == seg[:, [1, 3]].max()


def test_caller_side_allocations_of_the_side_tables():
    """What the caller owes msb_build_table / msb_build_visibility (cuda._empty_table, cuda._visibility_grid): array
    sizes, the exclusive prefix sums, a grid that covers each env's static geometry in 0.25 m cells; envs without a
    static line get no boxes and no cells."""
    widths = torch.tensor([16 + 40, 16, 16 + 1, 16 + 33], dtype=torch.int32)      # 16 = the agents' lines
    occ, occ_starts, boxes, box_starts, meta, rec = cuda._empty_table(widths, 16)
    assert box_starts.tolist() == [0, 3, 3, 4] and boxes.shape == (7, 4)
    assert occ.shape == (7 * 16, 4) and rec.shape == (7 * 16, 4) and rec.dtype == torch.int32
    assert occ_starts.shape == (4,) and meta.shape == (4, 2)
    boxes = torch.tensor([[1., 1., 3., 2.], [2., 0., 6.1, 2.], [0., 0., 1., 1.],        # env 0: x 0..6.1, y 0..2
                          [5., 5., 5., 9.],                                                # env 2: a vertical sliver
                          [0., 0., 1., 1.], [1., 1., 2., 2.], [2., 2., 3., 3.]])          # env 3
    vis, starts, vmeta = cuda._visibility_grid(boxes, box_starts, widths, 16)
    gx, gy = vmeta[:, 2].long(), vmeta[:, 3].long()
    assert gx.tolist() == [25, 0, 1, 12] and gy.tolist() == [8, 0, 16, 12]
    assert vmeta[0, :2].tolist() == [0., 0.] and vmeta[2, :2].tolist() == [5., 5.]
    assert starts.tolist() == [0, 200, 200, 216] and vis.numel() == 200 + 16 + 144 and vis.dtype == torch.int32
    assert float(vmeta[0, 0] + gx[0] * cuda.VIS_CELL) >= 6.1 and float(vmeta[3, 1] + gy[3] * cuda.VIS_CELL) >= 3.   # reaches the far corner


def test_geometry_cache_round_trip_and_reference_format(tmp_path):
    """megastep_b200.cubicasa against the reference's cache format (cubicasa.py:149-174: gzip(np.savez(flat)) with
    members "<id>/walls|lights|masks|res"): a file written the reference's way loads here, a file written here loads
    the reference's way (np.load on the unzipped bytes), and sample() splits and cycles as the reference does."""
    import gzip
    from io import BytesIO
    from megastep_b200 import cubicasa, toys, synthetic
    gs = {str(100 + i): g for i, g in enumerate([toys.box(), toys.column()] + synthetic.sample(18, seed=5))}
    for g in gs.values():
        if 'masks' not in g:
            g['masks'] = np.zeros((3, 4), np.int16)
    # written the reference's way
    flat = {f'{k}/{f}': np.asarray(g[f]) for k, g in gs.items() for f in cubicasa.FIELDS}
    bs = BytesIO()
    np.savez(bs, **flat)
    theirs = tmp_path / 'theirs.npz.gz'
    theirs.write_bytes(gzip.compress(bs.getvalue()))
    got = cubicasa.load_geometries(theirs)
    assert sorted(got) == sorted(gs)
    for k, g in gs.items():
        for f in cubicasa.FIELDS:
            assert np.array_equal(got[k][f], np.asarray(g[f])) and got[k][f].dtype == np.asarray(g[f]).dtype
    # written here, read the reference's way
    ours = cubicasa.save_geometries(gs, tmp_path / 'ours.npz.gz')
    back = np.load(BytesIO(gzip.decompress(ours.read_bytes())))
    assert sorted(back.files) == sorted(flat)
    assert all(np.array_equal(back[name], flat[name]) for name in flat)
    # flatten / unflatten
    tree = {'a': {'b': 1, 'c': {'d': 2}}, 'e': 3}
    assert cubicasa.flatten(tree) == {'a/b': 1, 'a/c/d': 2, 'e': 3} and cubicasa.unflatten(cubicasa.flatten(tree)) == tree
    # sample(): sorted ids shuffled by the seed, first 90 % training, cycled
    order = np.random.RandomState(7).permutation(sorted(gs))
    train = cubicasa.sample(40, 'training', seed=7, path=ours)
    assert [g.id for g in train] == [order[:18][i % 18] for i in range(40)]
    assert [g.id for g in cubicasa.sample(3, 'test', seed=7, path=ours)] == [order[18], order[19], order[18]]
    assert len(cubicasa.sample(5, 'all', seed=7, path=ours)) == 5
    with pytest.raises(ValueError):
        cubicasa.sample(1, 'validation', path=ours)
    with pytest.raises(FileNotFoundError):
        cubicasa.load_geometries(tmp_path / 'missing.npz.gz')
    # and the geometries feed the scene builder unchanged
    from megastep_b200 import scene
    arrays = scene.scene_arrays(train[:3], 2, np.random.RandomState(0))
    assert len(arrays['line_widths']) == 3


def test_culling_model_packs_runs_like_the_table_builder():
    """scripts/cull_sim.py (the CPU model used to plan kernel changes) must pack runs exactly as the spatial table does,
    or its batch / test counts say nothing about the kernel."""
    import importlib.util
    import os
    from megastep_b200 import scene, synthetic
    spec = importlib.util.spec_from_file_location('cull_sim', os.path.join(os.path.dirname(os.path.dirname(__file__)), 'scripts', 'cull_sim.py'))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    gs = synthetic.sample(3, seed=9)
    arrays = scene.scene_arrays(gs, 2, np.random.RandomState(9))
    lines = cuda.Ragged3D(torch.as_tensor(arrays['lines']), torch.as_tensor(arrays['line_widths']))
    AF = 2 * len(arrays['model'])
    for order in ('str',):                                 # (the model's Morton cells start at each env's own corner)
        old = cuda.TABLE_ORDER
        cuda.TABLE_ORDER = order
        try:
            occ, occ_starts, boxes, box_starts, meta, rec = cuda._occluder_table(lines, AF, 16)
        finally:
            cuda.TABLE_ORDER = old
        for n in range(3):
            walls = lines[n].reshape(-1, 4)[AF:].numpy().astype(np.float64)
            rows, bx = sim.pack(walls, 16, order)
            W = len(walls)
            lo = int(occ_starts[n])
            assert np.array_equal(rows.astype(np.float32), occ[lo:lo + W].numpy()), (order, n)
            nb = (W + 15) // 16
            assert np.allclose(bx, boxes[int(box_starts[n]):int(box_starts[n]) + nb].numpy())


def test_demo_env_rules_on_plain_tensors():
    """megastep_b200.envs: the game rules of the reference's Explorer (explorer.py:34-58) and Deathmatch
    (deathmatch.py:54-80) on CPU tensors, against hand-worked cases."""
    from megastep_b200 import envs
    # --- Explorer: two envs; env 0 has lines of 3 and 2 texels, env 1 one line of 4 texels
    line_starts = torch.tensor([0, 2], dtype=torch.int32)
    tex_widths = torch.tensor([3, 2, 4], dtype=torch.int32)
    tex_starts = torch.tensor([0, 3, 5], dtype=torch.int32)
    indices = torch.tensor([[[0, 1, -1, 0]], [[0, 0, 0, -1]]], dtype=torch.int32)               # (N=2, A=1, R=4)
    locations = torch.tensor([[[0., .99, float('nan'), 1.]], [[.3, .5, 1., float('nan')]]])
    tex = envs.texel_indices(line_starts, tex_starts, tex_widths, indices, locations)
    #   env 0: line 0 @0 -> texel 0; line 1 @.99 -> 3 + floor(1.98) = 4; miss; line 0 @1 -> floor(3) clamped to 2
    #   env 1: global line 2 @.3 -> 5 + 1; @.5 -> 5 + 2; @1 -> clamped to 5 + 3; miss
    assert tex.tolist() == [[[0, 4, -1, 2]], [[6, 7, 8, -1]]]
    texel_env = torch.tensor([0, 0, 0, 0, 0, 1, 1, 1, 1])
    led = envs.ExplorationLedger(texel_env, n_envs=2, rays_per_obs=4)
    r = led.reward(tex, reset=torch.tensor([False, False]))
    assert r.tolist() == [3 / 4, 3 / 4] and led.potential.tolist() == [3., 3.]
    r = led.reward(tex, reset=torch.tensor([False, False]))                                      # nothing new
    assert r.tolist() == [0., 0.]
    more = torch.tensor([[[1, 4, -1, 2]], [[5, 7, 8, -1]]])
    r = led.reward(more, reset=torch.tensor([False, True]))                                      # env 1's reward is void
    assert r.tolist() == [1 / 4, 0.] and led.potential.tolist() == [4., 4.]
    led.forget(torch.tensor([True, False]))
    assert led.potential.tolist() == [0., 4.] and led.seen.tolist() == [False] * 5 + [True] * 4
    # --- Deathmatch: 1 env, 3 agents, 8 rays pooled by 2 -> 4 pixels; the two middle pixels are the crosshairs
    F = 8
    lines = torch.full((1, 3, 1, 8), 40, dtype=torch.int32)                                      # a wall (line 40 >= 3 * F)
    lines[0, 0, 0, 2:6] = 1 * F + 3            # agent 0 has agent 1 across the middle of its view
    lines[0, 1, 0, 0:2] = 2 * F                # agent 1 sees agent 2 at the edge only
    lines[0, 2, 0, 4:6] = 0 * F + 7            # agent 2 has agent 0 in the right middle pixel
    lines[0, 2, 0, 7] = -1
    opp = envs.seen_agents(lines, n_model=F, n_agents=3, subsample=2)
    assert opp.shape == (1, 3, 1, 4)
    assert opp[0, :, 0].tolist() == [[-1, 1, 1, -1], [2, -1, -1, -1], [-1, -1, 0, -1]]
    matchings, hits, wounds = envs.shots(opp, 3)
    assert matchings[0].tolist() == [[False, True, False], [False, False, False], [True, False, False]]
    assert hits[0].tolist() == [1., 0., 1.] and wounds[0].tolist() == [1., 1., 0.]


def test_demo_envs_wiring_on_a_fake_core(monkeypatch):
    """Explorer / Deathmatch end to end on CPU: the Core, the scenery upload, physics and render are faked (they need a
    GPU), everything else — observation modules, spawns, ledger, shooting, the reset/step protocol — is the real thing."""
    from megastep_b200 import envs, modules, scene, toys
    from megastep_b200.arrdict import arrdict

    class FakeScenery:
        def __init__(self, geometries, n_agents):
            a = scene.scene_arrays(geometries, n_agents, np.random.RandomState(0))
            self.n_agents = n_agents
            self.lines = cuda.Ragged3D(torch.as_tensor(a['lines']), torch.as_tensor(a['line_widths']))
            self.textures = cuda.Ragged2D(torch.as_tensor(a['textures']), torch.as_tensor(a['tex_widths']))
            self.model = torch.as_tensor(a['model'])

    class FakeCore:
        def __init__(self, scenery, res=64, fov=130, fps=10):
            self.scenery, self.res, self.fov, self.fps = scenery, res, fov, fps
            self.n_envs, self.n_agents = len(scenery.lines), scenery.n_agents
            self.device = torch.device('cpu')
            self.agent_radius = .1
            self.random = np.random.RandomState(1)
            z = lambda *s: torch.zeros((self.n_envs, self.n_agents, *s))
            self.agents = arrdict(angles=z(), positions=z(2), angvelocity=z(), velocity=z(2))

        def env_full(self, x):
            return torch.full((self.n_envs,), x)

        def agent_full(self, x):
            return torch.full((self.n_envs, self.n_agents), x)

        def state(self, e):
            return arrdict(n_envs=torch.tensor(self.n_envs))

    rng = np.random.RandomState(3)

    def fake_render(core):
        N, A, R = core.n_envs, core.n_agents, core.res
        L = int(core.scenery.lines.widths.min())
        idx = torch.as_tensor(rng.randint(-1, L, (N, A, 1, R)).astype(np.int32))
        loc = torch.as_tensor(rng.rand(N, A, 1, R).astype(np.float32))
        loc[idx < 0] = float('nan')
        dist = torch.as_tensor(rng.uniform(.2, 12., (N, A, 1, R)).astype(np.float32))
        return arrdict(indices=idx, locations=loc, dots=torch.zeros(N, A, 1, R), distances=dist,
                       screen=torch.as_tensor(rng.rand(N, A, 3, 1, R).astype(np.float32)))

    monkeypatch.setattr(envs.scene, 'scenery', FakeScenery)
    monkeypatch.setattr(envs.core_, 'Core', FakeCore)
    monkeypatch.setattr(modules, '_physics', lambda core: None)
    monkeypatch.setattr(modules, 'render', fake_render)
    gs = [toys.box()] * 3

    env = envs.Explorer(gs)
    out = env.reset()
    assert out.obs.rgb.shape == (3, 1, 3, 1, 64) and out.obs.d.shape == (3, 1, 1, 1, 64) and out.obs.imu.shape == (3, 1, 3)
    assert out.reset.all() and (out.reward == 0).all()
    total = torch.zeros(3)
    for _ in range(4):
        out = env.step(arrdict(actions=torch.as_tensor(rng.randint(0, 7, (3, 1)))))
        assert out.reward.shape == (3,) and (out.reward >= 0).all() and not out.reset.any()
        total += out.reward
    assert (total > 0).all() and (total * 64 <= env._ledger.potential + 1e-3).all()      # rewards are new texels / 64 pooled rays
    assert env.state(1).seen.dtype == torch.bool and int(env.state(1).length) == 4

    dm = envs.Deathmatch(gs, 2)
    out = dm.reset()
    assert dm.n_envs == 6 and out.obs.rgb.shape == (6, 1, 3, 1, 128) and out.obs.health.shape == (6, 1, 1)
    assert out.reset.shape == (6,) and out.reset.all() and out.reward.shape == (6,)
    for _ in range(3):
        out = dm.step(arrdict(actions=torch.as_tensor(rng.randint(0, 7, (6, 1)))))
        assert out.obs.imu.shape == (6, 1, 3) and out.reward.shape == (6,) and torch.isfinite(out.obs.health).all()
    assert (dm._health <= 1).all() and dm.matchings.shape == (3, 2, 2)
    # health only ever drops between respawns: by 0.001 per step plus 0.05 per wound
    assert (dm._health < 1).all()


# ----------------------------------------------------------------------------------------------------------------------
# round 2 host logic
# ----------------------------------------------------------------------------------------------------------------------
def test_tile_arrays_cycles_through_the_envs():
    import numpy as np
    from megastep_b200 import scene, sharding, synthetic
    gs = synthetic.sample(5, seed=3)
    base = scene.scene_arrays(gs, 2, np.random.RandomState(0))
    base['baked'] = np.arange(len(base['textures']), dtype=np.float32)
    for n in (5, 7, 10, 13):
        tiled = synthetic.tile_arrays(base, n)
        assert len(tiled['line_widths']) == n
        for i in range(n):
            one, ref = sharding.shard_arrays(tiled, i, i + 1), sharding.shard_arrays(base, i % 5, i % 5 + 1)
            for k in ('lines', 'line_widths', 'lights', 'light_widths', 'textures', 'tex_widths', 'baked'):
                np.testing.assert_array_equal(one[k], ref[k])


def test_spawns_land_inside_rooms_whether_or_not_envs_share_a_floorplan():
    import numpy as np
    from megastep_b200 import synthetic
    for n_unique in (6, 2):
        gs = synthetic.sample(6, seed=4, n_unique=n_unique)
        pos, ang = synthetic.spawns(gs, 3, np.random.RandomState(1))
        assert pos.shape == (6, 3, 2) and ang.shape == (6, 3) and (ang >= -180).all() and (ang < 180).all()
        for n, g in enumerate(gs):
            r = g['rooms']
            inside = ((pos[n, :, None, 0] >= r[None, :, 0]) & (pos[n, :, None, 0] <= r[None, :, 2]) &
                      (pos[n, :, None, 1] >= r[None, :, 1]) & (pos[n, :, None, 1] <= r[None, :, 3])).any(1)
            assert inside.all()
        if n_unique == 2:                                       # copies of a floorplan still get their own poses
            assert not np.array_equal(pos[0], pos[2])


def test_scene_building_does_not_load_the_native_library():
    """bench.py's reference arm builds its scenes with this package: that must not map libmegastep_b200.so."""
    import subprocess
    import sys
    code = ("import sys, numpy as np; from megastep_b200 import scene, synthetic, sharding, geometry, toys; "
            "scene.scene_arrays(synthetic.sample(2, seed=1), 2, np.random.RandomState(0)); "
            "assert 'megastep_b200.cuda' not in sys.modules and 'megastep_b200.core' not in sys.modules; print('clean')")
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=common.ROOT)
    assert out.returncode == 0 and 'clean' in out.stdout, out.stderr[-500:]


def test_package_exposes_its_submodules_lazily():
    import megastep_b200
    for name in ('envs', 'cubicasa', 'sharding', 'synthetic', 'constants'):
        assert getattr(megastep_b200, name).__name__ == f'megastep_b200.{name}'


def test_spawn_points_without_masks_fall_back_to_the_rooms():
    import numpy as np
    from megastep_b200 import modules, synthetic
    gs = synthetic.sample(3, seed=2)                            # no masks
    pts = modules.random_empty_positions(gs, 2, 7, random=np.random.RandomState(0))
    assert pts.shape == (3, 2, 7, 2)
    for n, g in enumerate(gs):
        r = g['rooms']
        p = pts[n].reshape(-1, 2)
        assert ((p[:, None, 0] >= r[None, :, 0]) & (p[:, None, 0] <= r[None, :, 2]) & (p[:, None, 1] >= r[None, :, 1]) & (p[:, None, 1] <= r[None, :, 3])).any(1).all()


def test_library_contains_the_blackwell_instructions_it_claims():
    """The built library's SASS (sm_100a) for view_kernel: bulk (TMA) copies, mbarrier transactions, warp reductions and the
    programmatic-dependent-launch pair; physics_kernel: the bulk L2 prefetch. (profiles/r02_sass_excerpt.txt is the same
    listing, committed.)"""
    import shutil
    import subprocess
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('cuobjdump not available')

    def sass(fn):
        return subprocess.run([tool, '-sass', '-fun', fn, cuda.library_path()], capture_output=True, text=True).stdout

    view = sass('_Z11view_kernelILi2ELb0ELb0EEv5KArgs')
    assert 'sm_100' in view
    for mnemonic in ('UBLKCP', 'SYNCS.ARRIVE.TRANS64', 'SYNCS.PHASECHK.TRANS64.TRYWAIT', 'REDUX', 'ACQBULK', 'PREEXIT'):
        assert mnemonic in view, mnemonic
    assert 'UBLKPF.L2' in sass('_Z14physics_kernel5KArgs')
