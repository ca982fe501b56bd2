"""Shared scene/state builders for the tests (seeded, deterministic)."""
import glob
import importlib.machinery
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AGENT_RADIUS = .15 / 2 ** .5


def synthetic_scene(n_envs, n_agents, seed=1, bake=True):
    """(geometries, scene arrays incl. oracle-baked lighting)"""
    from megastep_b200 import scene, synthetic
    from oracle import oracle
    gs = synthetic.sample(n_envs, seed=seed)
    arrays = scene.scene_arrays(gs, n_agents, np.random.RandomState(seed))
    if bake:
        arrays['baked'] = oracle.bake(arrays)
    return gs, arrays


def toy_scene(kind, n_envs, n_agents, seed=1):
    from megastep_b200 import scene, toys
    from oracle import oracle
    gs = [getattr(toys, kind)()] * n_envs
    arrays = scene.scene_arrays(gs, n_agents, np.random.RandomState(seed))
    arrays['baked'] = oracle.bake(arrays)
    return gs, arrays


def random_state(gs, n_agents, seed=2, speed=3., angspeed=200.):
    """Agents at random free poses with random velocities (m/s, deg/s)."""
    from megastep_b200 import synthetic
    rng = np.random.RandomState(seed)
    if 'rooms' in gs[0]:
        pos, ang = synthetic.spawns(gs, n_agents, rng)
    else:
        pos = rng.uniform(1.6, 5.4, (len(gs), n_agents, 2)).astype(np.float32)
        ang = rng.uniform(-180, 180, (len(gs), n_agents)).astype(np.float32)
    N = len(gs)
    return dict(angles=ang, positions=pos,
                angvelocity=rng.uniform(-angspeed, angspeed, (N, n_agents)).astype(np.float32),
                velocity=(speed * rng.normal(size=(N, n_agents, 2))).astype(np.float32))


def copy_state(st):
    return {k: v.copy() for k, v in st.items()}


def to_device(arrays, st, res, fov, fps=10., device='cuda'):
    """Scene arrays + state -> (core, cuda module); the Core's agents are loaded with `st` and the scenery's baked
    light map with arrays['baked'] (so lighting parity is tested separately from render parity)."""
    import torch
    from megastep_b200 import core as core_, scene
    s = scene.upload(arrays, device)
    if 'baked' in arrays:
        s.baked.vals.copy_(torch.as_tensor(arrays['baked']))
    c = core_.Core(s, res=res, fov=fov, fps=fps)
    load_state(c, st)
    return c


def load_state(c, st):
    import torch
    for k in ('angles', 'positions', 'angvelocity', 'velocity'):
        getattr(c.agents, k).copy_(torch.as_tensor(st[k]))


def read_state(c):
    return {k: getattr(c.agents, k).cpu().numpy() for k in ('angles', 'positions', 'angvelocity', 'velocity')}


def reference_module():
    """The reference's own extension (oracle/_ref, built from its unmodified sources), or None when not built."""
    hits = glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'megastepcuda*.so'))
    if not hits:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    loader = importlib.machinery.ExtensionFileLoader('megastepcuda', hits[0])
    spec = importlib.util.spec_from_loader('megastepcuda', loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def reference_scenery(ref, arrays, device='cuda'):
    import torch
    t = lambda k, dtype: torch.as_tensor(arrays[k], dtype=dtype).contiguous().to(device)
    s = ref.Scenery(n_agents=arrays['n_agents'],
                    lights=ref.Ragged2D(t('lights', torch.float32), t('light_widths', torch.int32)),
                    lines=ref.Ragged3D(t('lines', torch.float32), t('line_widths', torch.int32)),
                    textures=ref.Ragged2D(t('textures', torch.float32), t('tex_widths', torch.int32)),
                    model=t('model', torch.float32))
    if 'baked' in arrays:
        s.baked.vals.copy_(torch.as_tensor(arrays['baked']))
    return s


def reference_scenery_tiled(ref, arrays, n_envs, baked=None, device='cuda'):
    """A reference Scenery of `n_envs` envs cycling through the envs of `arrays`, repeated on the device (what
    megastep_b200.scene.tiled_scenery does for ours). `baked`: the distinct envs' baked light map, tiled the same way."""
    import torch
    t = lambda k, dtype: torch.as_tensor(arrays[k], dtype=dtype).contiguous().to(device)
    lw, iw, tw = t('line_widths', torch.int32), t('light_widths', torch.int32), t('tex_widths', torch.int32)
    u = lw.numel()
    reps, rest = divmod(n_envs, u)
    l_rest, i_rest = int(lw[:rest].sum()), int(iw[:rest].sum())
    t_rest = int(tw[:l_rest].long().sum())
    cyc = lambda x, n_rest: torch.cat([x.repeat(reps, *([1] * (x.dim() - 1))), x[:n_rest]]).contiguous()
    s = ref.Scenery(n_agents=arrays['n_agents'],
                    lights=ref.Ragged2D(cyc(t('lights', torch.float32), i_rest), cyc(iw, rest)),
                    lines=ref.Ragged3D(cyc(t('lines', torch.float32), l_rest), cyc(lw, rest)),
                    textures=ref.Ragged2D(cyc(t('textures', torch.float32), t_rest), cyc(tw, l_rest)),
                    model=t('model', torch.float32))
    if baked is not None:
        s.baked.vals.copy_(cyc(torch.as_tensor(baked).to(device), t_rest))
    return s


def reference_agents(ref, st, device='cuda'):
    import torch
    return ref.Agents(**{k: torch.as_tensor(st[k]).contiguous().to(device) for k in ('angles', 'positions', 'angvelocity', 'velocity')})


def index_agreement(a, b, dist_a, dist_b, tol=2e-3):
    """Fraction of rays whose hit index differs AND whose hit distance differs by more than `tol` (i.e. real
    disagreements rather than ties between coincident/abutting segments)."""
    a, b = np.asarray(a), np.asarray(b)
    diff = a != b
    da, db = np.asarray(dist_a), np.asarray(dist_b)
    with np.errstate(invalid='ignore'):
        far = np.abs(da - db) > tol * np.maximum(1., np.minimum(np.abs(da), np.abs(db)))
    far |= np.isinf(da) != np.isinf(db)
    return float(diff.mean()), float((diff & far).mean())


# ----------------------------------------------------------------------------------------------------------------------
# The reference's own PYTHON package as a checker: megastep/{core,modules,scene,ragged}.py and the demo
# envs, unmodified, as pip-installed by oracle/build_ref.sh into baseline/_ref (git-ignored, travels to the GPU box),
# running on the reference's own extension build (oracle/_ref/megastepcuda*.so). The package's __init__ (a JIT build
# with -std=c++14 that torch >= 2 rejects) is bypassed; matplotlib / rasterio / shapely / bs4, absent from this image
# and used by none of the code the checker runs, are stubbed.
# ----------------------------------------------------------------------------------------------------------------------
_REFPKG = None


def _stub_absent_modules():
    import sys
    import types
    from unittest import mock

    def absent(name):
        if name in sys.modules:
            return False
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError):
            return True

    if absent('matplotlib'):
        from megastep_b200.scene import to_rgb               # '#rrggbb', 'g' / 'r', '.25': all reference scene.py asks of it
        mpl = types.ModuleType('matplotlib')
        mpl.__path__ = []
        mpl.colors = types.ModuleType('matplotlib.colors')
        mpl.colors.to_rgb = to_rgb
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.colors'] = mpl.colors
        for sub in ('pyplot', 'tight_bbox', 'collections', 'patches', 'cm'):
            m = mock.MagicMock(name=f'matplotlib.{sub}')
            setattr(mpl, sub, m)
            sys.modules[f'matplotlib.{sub}'] = m
    stubs = {'rasterio': ('features', 'transform'), 'shapely': ('ops', 'geometry'), 'bs4': ()}
    for top, subs in stubs.items():
        if absent(top):
            m = mock.MagicMock(name=top)
            m.__path__ = []
            sys.modules[top] = m
            for sub in subs:
                sys.modules[f'{top}.{sub}'] = getattr(m, sub)


def reference_package():
    """Namespace with the reference's own `core`, `modules`, `scene`, `ragged`, `spaces`, `cuda` (its extension) and
    `envs` (explorer / deathmatch modules), or None when oracle/_ref (extension) or baseline/_ref (package) is not built."""
    global _REFPKG
    if _REFPKG is not None:
        return _REFPKG or None
    import sys
    import types
    site = os.path.join(ROOT, 'baseline', '_ref')
    ext = reference_module()
    if ext is None or not os.path.exists(os.path.join(site, 'megastep', 'modules.py')):
        _REFPKG = False
        return None
    _stub_absent_modules()
    if site not in sys.path:
        sys.path.insert(0, site)                               # for `rebar` (arrdict / dotdict)
    pkg = types.ModuleType('megastep')
    pkg.__path__ = [os.path.join(site, 'megastep')]            # submodules come from the installed reference ...
    pkg.cuda = ext                                             # ... but `megastep.cuda` is the pre-built extension
    sys.modules['megastep'] = pkg
    sys.modules['megastep.cuda'] = ext
    ns = types.SimpleNamespace(cuda=ext)
    for name in ('ragged', 'core', 'spaces', 'scene', 'modules', 'cubicasa'):
        setattr(ns, name, importlib.import_module(f'megastep.{name}'))
    demo = types.ModuleType('megastep.demo')                   # (its __init__ pulls in the RL learner)
    demo.__path__ = [os.path.join(site, 'megastep', 'demo')]
    sys.modules['megastep.demo'] = demo
    envs = types.ModuleType('megastep.demo.envs')
    envs.__path__ = [os.path.join(site, 'megastep', 'demo', 'envs')]
    sys.modules['megastep.demo.envs'] = envs
    ns.explorer = importlib.import_module('megastep.demo.envs.explorer')
    ns.deathmatch = importlib.import_module('megastep.demo.envs.deathmatch')
    import rebar.arrdict
    import rebar.dotdict
    ns.arrdict, ns.dotdict = rebar.arrdict, rebar.dotdict
    _REFPKG = ns
    return ns


def reference_core(pkg, arrays, st, res, fov, fps=10., device='cuda'):
    """A reference `core.Core` (its own Python, its own extension) over the same scene arrays and agent state."""
    s = reference_scenery(pkg.cuda, arrays, device)
    c = pkg.core.Core(s, res=res, fov=fov, fps=fps)
    load_state(c, st)
    return c
