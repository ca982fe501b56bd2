"""Synthetic, seeded, cubicasa-shaped floorplans.

The reference draws its geometries from the Cubicasa5k dataset (megastep/cubicasa.py:177-224), which needs a
network download and a licence prompt; none of it is in the repository. Benchmarks and parity tests therefore run
on generated floorplans with the same statistics as the one real sample the reference documents
(megastep/core.py:104-106: 307 lines at 4 agents = 275 wall segments, 21 lights): rectilinear rooms from a binary
space partition of a ~15 x 12 m footprint, walls as thin closed rectangles (4 segments each, like the SVG wall
polygons of geometry.py:43-57) broken by door gaps, one light at each room's centre.

Geometries use the reference's dict format (see geometry.py): walls (W,2,2), lights (I,2), masks, res — plus
`rooms` (the room rectangles), which makes spawn sampling analytic.
"""
import numpy as np

from . import geometry
from .arrdict import arrdict

THICKNESS = (.08, .2)     # wall thickness range, metres
DOOR = (.8, 1.1)          # door width range, metres


def _rect_segments(x0, y0, x1, y1):
    c = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]])
    return np.stack([c, np.roll(c, -1, 0)], 1)


def _wall_pieces(lo, hi, random, door=True):
    """Break the 1-D extent [lo, hi] of a wall around one door gap."""
    pieces = [(lo, hi)]
    if door and hi - lo > 1.6:
        w = random.uniform(*DOOR)
        d = random.uniform(lo + .3, hi - .3 - w)
        pieces = [(lo, d), (d + w, hi)]
    return [(a, b) for a, b in pieces if b - a > .15]


def floorplan(seed, with_masks=False, target_walls=None):
    """One synthetic floorplan geometry. Deterministic in `seed`."""
    random = np.random.RandomState(seed)
    width, height = random.uniform(11, 17), random.uniform(9, 14)
    n_rooms = int(np.clip(np.round(random.normal(21, 5)), 8, 32))
    target = float(np.clip(random.normal(275, 60), 60, 600)) if target_walls is None else float(target_walls)
    m = geometry.MARGIN

    rooms = [(m, m, m + width, m + height)]
    partitions = []   # (axis, position, lo, hi): axis 0 = vertical wall at x=position spanning y in [lo, hi]
    while len(rooms) < n_rooms:
        areas = np.array([(r[2] - r[0]) * (r[3] - r[1]) for r in rooms])
        x0, y0, x1, y1 = rooms.pop(int(random.choice(len(rooms), p=areas / areas.sum())))
        if max(x1 - x0, y1 - y0) < 2.4:
            rooms.append((x0, y0, x1, y1))
            if all(max(r[2] - r[0], r[3] - r[1]) < 2.4 for r in rooms):
                break
            continue
        if (x1 - x0) > (y1 - y0):
            p = random.uniform(x0 + .35 * (x1 - x0), x0 + .65 * (x1 - x0))
            rooms += [(x0, y0, p, y1), (p, y0, x1, y1)]
            partitions.append((0, p, y0, y1))
        else:
            p = random.uniform(y0 + .35 * (y1 - y0), y0 + .65 * (y1 - y0))
            rooms += [(x0, y0, x1, p), (x0, p, x1, y1)]
            partitions.append((1, p, x0, x1))

    # exterior shell, no doors
    shell = [(0, m, m, m + height), (0, m + width, m, m + height), (1, m, m, m + width), (1, m + height, m, m + width)]
    pieces = []   # (axis, position, half thickness, lo, hi)
    for k, (axis, pos, lo, hi) in enumerate(shell + partitions):
        t = random.uniform(*THICKNESS) / 2
        pieces += [(axis, pos, t, a, b) for a, b in _wall_pieces(lo, hi, random, door=k >= len(shell))]
    # real floorplans break walls at every junction; split the longest pieces until the segment count is on target
    while 4 * len(pieces) < target:
        k = int(np.argmax([b - a for *_, a, b in pieces]))
        axis, pos, t, a, b = pieces[k]
        if b - a < .6:
            break
        mid = random.uniform(a + .3 * (b - a), a + .7 * (b - a))
        pieces[k:k + 1] = [(axis, pos, t, a, mid), (axis, pos, t, mid, b)]
    segs = [_rect_segments(*((pos - t, a, pos + t, b) if axis == 0 else (a, pos - t, b, pos + t)))
            for axis, pos, t, a, b in pieces]
    walls = np.concatenate(segs)

    rooms = np.array(rooms)
    lights = np.stack([(rooms[:, 0] + rooms[:, 2]) / 2, (rooms[:, 1] + rooms[:, 3]) / 2], -1)
    g = arrdict(walls=walls, lights=lights, rooms=rooms, res=geometry.RES, id=f'synthetic-{seed}')
    if with_masks:
        polys = [np.array([[r[0], r[1]], [r[2], r[1]], [r[2], r[3]], [r[0], r[3]]]) for r in rooms]
        g['masks'] = geometry.masks(walls, polys)
    return g


def sample(n_geometries, seed=1, n_unique=1024, with_masks=False):
    """A list of `n_geometries` floorplans; at most `n_unique` distinct ones, repeated cyclically (the reference
    likewise cycles through its ~5k designs, cubicasa.py:218-224)."""
    unique = [floorplan(seed * 1_000_003 + i, with_masks) for i in range(min(n_geometries, n_unique))]
    return [unique[i % len(unique)] for i in range(n_geometries)]


def spawns(geometries, n_agents, random, clearance=.35):
    """Random collision-free-ish poses: positions (N, A, 2) uniform inside rooms (kept `clearance` from the room's
    walls), angles (N, A) uniform in [-180, 180). Plays the role of RandomSpawns without needing masks. Envs that share
    a geometry object (`sample` cycles through its distinct floorplans) are drawn together, so that very large batches
    take seconds."""
    N = len(geometries)
    pos = np.zeros((N, n_agents, 2), np.float32)
    groups = {}
    for n, g in enumerate(geometries):
        groups.setdefault(id(g), (g, []))[1].append(n)
    for g, envs in groups.values():
        rooms = g['rooms']
        inner = np.stack([rooms[:, 0] + clearance, rooms[:, 1] + clearance, rooms[:, 2] - clearance, rooms[:, 3] - clearance], -1)
        ok = (inner[:, 2] > inner[:, 0]) & (inner[:, 3] > inner[:, 1])
        inner = inner[ok] if ok.any() else inner
        areas = np.maximum(inner[:, 2] - inner[:, 0], 1e-3) * np.maximum(inner[:, 3] - inner[:, 1], 1e-3)
        which = random.choice(len(inner), size=(len(envs), n_agents), p=areas / areas.sum())
        u = random.uniform(size=(len(envs), n_agents, 2))
        pos[envs, :, 0] = inner[which, 0] + u[..., 0] * (inner[which, 2] - inner[which, 0])
        pos[envs, :, 1] = inner[which, 1] + u[..., 1] * (inner[which, 3] - inner[which, 1])
    ang = random.uniform(-180, 180, (N, n_agents)).astype(np.float32)
    return pos, ang


def tile_arrays(arrays, n_envs):
    """Repeat the envs of a `scene.scene_arrays` result cyclically up to `n_envs` environments."""
    u = len(arrays['line_widths'])
    if n_envs == u:
        return arrays
    reps, rest = divmod(n_envs, u)
    lw, iw, tw = arrays['line_widths'], arrays['light_widths'], arrays['tex_widths']
    l_rest, i_rest = int(lw[:rest].sum()), int(iw[:rest].sum())
    t_rest = int(tw[:l_rest].astype(np.int64).sum())

    def cyc(vals, n_rest):                                  # the whole array `reps` times, then its first `n_rest` rows
        return np.concatenate([np.tile(vals, (reps,) + (1,) * (vals.ndim - 1)), vals[:n_rest]])

    return dict(
        n_agents=arrays['n_agents'], model=arrays['model'],
        lines=cyc(arrays['lines'], l_rest), line_widths=cyc(lw, rest),
        lights=cyc(arrays['lights'], i_rest), light_widths=cyc(iw, rest),
        textures=cyc(arrays['textures'], t_rest), tex_widths=cyc(tw, l_rest),
        **({'baked': cyc(arrays['baked'], t_rest)} if 'baked' in arrays else {}))
