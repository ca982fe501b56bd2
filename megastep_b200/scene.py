"""Scene construction: a list of geometries -> the ragged `lights / lines / textures` tensors of a `cuda.Scenery`
(reference: megastep/scene.py:25-100). One-off CPU work followed by an upload and a `cuda.bake`.

`scene_arrays` is the pure-numpy half (no device, no extension) so it can be tested — and fed to the CPU oracle —
without a GPU; `scenery` uploads its result and bakes the static lighting.
"""
import numpy as np

from . import constants as core      # (the CPU half must not load the native library: bench.py's reference arm builds scenes too)

# ten bland wall colours (iwanthue), reference scene.py:10-20
COLORS = ['#c185ae', '#73a171', '#5666a4', '#9f7c4a', '#809cd5', '#566e40', '#8e537b', '#4f9fa4', '#b56d66', '#5a728c']

_NAMED = {'g': (0., .5, 0.), 'r': (1., 0., 0.), 'b': (0., 0., 1.), 'k': (0., 0., 0.), 'w': (1., 1., 1.)}


def to_rgb(c):
    """The subset of matplotlib.colors.to_rgb the scenes use: '#rrggbb', single-letter names, and grey levels
    given as a float string ('.25')."""
    if isinstance(c, str):
        if c.startswith('#') and len(c) == 7:
            return tuple(int(c[i:i + 2], 16) / 255 for i in (1, 3, 5))
        if c in _NAMED:
            return _NAMED[c]
        g = float(c)
        assert 0 <= g <= 1
        return (g, g, g)
    return tuple(float(x) for x in c[:3])


def lengths(lines):
    return ((lines[..., 0, :] - lines[..., 1, :]) ** 2).sum(-1) ** .5


def agent_model():
    """The agent's outline: an octagon of AGENT_WIDTH across, as (8, 2, 2) line endpoints in agent-local metres."""
    corners = np.array([[-.5, -1.], [+.5, -1.], [+1., -.5], [+1., +.5], [+.5, +1.], [-.5, +1.], [-1., +.5], [-1., -.5]])
    walls = np.stack([corners, np.roll(corners, -1, 0)], 1)
    return core.AGENT_WIDTH / 2 * walls


def agent_colors():
    """Dark grey body with green flanks and red nose/tail faces."""
    k, g, r = '.25', 'g', 'r'
    return np.stack([to_rgb(s) for s in (k, g, k, r, k, r, k, g)])


def resolutions(lines):
    """Texels per line at TEXTURE_RES metres per texel."""
    return np.ceil(lengths(lines) / core.TEXTURE_RES).astype(int)


def wall_pattern(n, l=.5, random=np.random):
    """A random-walk brightness pattern with jumps every ~l metres; makes depth perceivable."""
    p = core.TEXTURE_RES / l
    jumps = random.choice(np.array([0., 1.]), p=np.array([1 - p, p]), size=n)
    jumps = jumps * random.normal(size=n)
    return .5 + .5 * (jumps.cumsum() % 1)


def init_textures(agentlines, agentcolors, walls, random=np.random):
    colormap = np.array([to_rgb(c) for c in COLORS])
    wallcolors = colormap[np.arange(len(walls)) % len(colormap)]
    colors = np.concatenate([agentcolors, wallcolors])

    texwidths = resolutions(np.concatenate([agentlines, walls]))
    owner = np.repeat(np.arange(len(texwidths)), texwidths)
    textures = core.gamma_decode(colors[owner])

    pattern = wall_pattern(textures.shape[0], random=random)
    pattern[:texwidths[:len(agentlines)].sum()] = 1.
    return textures * pattern[:, None], texwidths


def random_lights(lights, random=np.random):
    """(I, 2) positions -> (I, 3) with a random intensity in [.5, 2)."""
    return np.concatenate([lights, random.uniform(.5, 2., (len(lights), 1))], -1)


def scene_arrays(geometries, n_agents=1, random=np.random):
    """Flat numpy arrays of the scene, in the layout the kernels (and the CPU oracle) consume:
    dict(n_agents, model, lines, line_widths, lights, light_widths, textures, tex_widths)."""
    model = agent_model()
    agentlines = np.tile(model, (n_agents, 1, 1))
    agentcolors = np.tile(agent_colors(), (n_agents, 1))
    lines, lwidths, lights, iwidths, textures, twidths = [], [], [], [], [], []
    for g in geometries:
        lt = random_lights(np.asarray(g['lights']).reshape(-1, 2), random)
        walls = np.asarray(g['walls']).reshape(-1, 2, 2)
        tex, tw = init_textures(agentlines, agentcolors, walls, random)
        lines.append(np.concatenate([agentlines, walls]))
        lwidths.append(len(agentlines) + len(walls))
        lights.append(lt)
        iwidths.append(len(lt))
        textures.append(tex)
        twidths.append(tw)
    f32 = lambda xs, tail: (np.concatenate(xs) if xs else np.zeros((0, *tail))).astype(np.float32)
    return dict(
        n_agents=n_agents, model=model.astype(np.float32),
        lines=f32(lines, (2, 2)), line_widths=np.array(lwidths, dtype=np.int32),
        lights=f32(lights, (3,)), light_widths=np.array(iwidths, dtype=np.int32),
        textures=f32(textures, (3,)), tex_widths=(np.concatenate(twidths) if twidths else np.zeros(0)).astype(np.int32))


def upload(arrays, device='cuda'):
    """scene_arrays(...) -> cuda.Scenery on `device` (not yet baked)."""
    import torch
    from . import cuda
    t = lambda k, dtype: torch.as_tensor(arrays[k], dtype=dtype).contiguous().to(device)
    return cuda.Scenery(
        n_agents=arrays['n_agents'],
        lights=cuda.Ragged2D(t('lights', torch.float32), t('light_widths', torch.int32)),
        lines=cuda.Ragged3D(t('lines', torch.float32), t('line_widths', torch.int32)),
        textures=cuda.Ragged2D(t('textures', torch.float32), t('tex_widths', torch.int32)),
        model=t('model', torch.float32))


def tiled_scenery(arrays, n_envs, device='cuda', params=None):
    """A baked `cuda.Scenery` of `n_envs` environments cycling through the envs of `arrays` (a `scene_arrays` result) —
    the reference likewise cycles through its ~5k floorplans (cubicasa.py:218-224). The distinct envs are uploaded and
    baked ONCE; the repetition happens on the device (torch `repeat`), so a 262,144-env scenery is built in seconds:
    the host never holds more than the distinct envs, and `bake` (a pure function of an env's static lines and lights)
    is not redone for the copies."""
    import torch
    from . import cuda
    base = upload(arrays, device)
    cuda.bake(base, params=params or cuda.make_params(core.AGENT_RADIUS, 64, 130., 10.))
    u = len(base.lines)
    if n_envs == u:
        return base
    reps, rest = divmod(n_envs, u)
    lw, iw, tw = base.lines.widths, base.lights.widths, base.textures.widths
    l_rest, i_rest = int(lw[:rest].sum()), int(iw[:rest].sum())
    t_rest = int(tw[:l_rest].long().sum())

    def cyc(t, n_rest):
        whole = t.repeat(reps, *([1] * (t.dim() - 1)))
        return whole if n_rest == 0 else torch.cat([whole, t[:n_rest]]).contiguous()

    s = cuda.Scenery(
        n_agents=base.n_agents,
        lights=cuda.Ragged2D(cyc(base.lights.vals, i_rest), cyc(iw, rest)),
        lines=cuda.Ragged3D(cyc(base.lines.vals, l_rest), cyc(lw, rest)),
        textures=cuda.Ragged2D(cyc(base.textures.vals, t_rest), cyc(tw, l_rest)),
        model=base.model)
    s.baked.vals.copy_(cyc(base.baked.vals, t_rest))
    return s


def scenery(geometries, n_agents=1, device='cuda', random=np.random):
    """Geometries -> baked `cuda.Scenery` (reference scene.py:75-100)."""
    from . import cuda
    s = upload(scene_arrays(geometries, n_agents, random), device)
    cuda.bake(s, params=cuda.make_params(core.AGENT_RADIUS, 64, 130., 10.))   # bake uses no per-Core parameter
    return s
