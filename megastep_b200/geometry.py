"""Geometry helpers: the `geometry` dict format, mask rasterisation and grid<->metre transforms.

A geometry is a dotdict with (reference: docs/concepts.rst:236-259, megastep/geometry.py:99-108)
    walls   (W, 2, 2) endpoints in metres, everything inside the +quadrant with a MARGIN border
    lights  (I, 2)    light positions
    masks   (H, W) int16 grid at `res` metres/cell: -1 wall, 0 free space, k >= 1 room ids; row 0 is the TOP
    res     cell size in metres

`masks` restates megastep/geometry.py:74-93 without shapely/rasterio (absent here and unpinned in the reference's
setup.py:23-27, so parity for masks is pinned only by its documented semantics): rooms are painted first, walls —
buffered by 1 cm — last, and a cell is painted when the shape touches it at all (`all_touched=True`).
SVG floorplan parsing (geometry.py:14-57) is out of scope: it needs the Cubicasa5k download.
"""
import numpy as np

MARGIN = 1.   # metres of free border around the floorplan
RES = .2      # metres per mask cell
SCALE = 100.  # SVG units (cm) per metre
WALL_BUFFER = .01


def cyclic_pairs(xs):
    """[(x0, x1), (x1, x2), ..., (xn, x0)]"""
    return [(xs[i], xs[(i + 1) % len(xs)]) for i in range(len(xs))]


def mask_shape(*pointsets):
    """(H, W) of the mask covering all points plus the margin (geometry.py:74-79)."""
    pts = np.concatenate([np.asarray(p).reshape(-1, 2) for ps in pointsets for p in ps])
    assert pts.min() > 0, 'Masker currently requires the points to be in the top-right quadrant'
    r, t = pts.max(0) + MARGIN
    return int(t / RES) + 1, int(r / RES) + 1


def _segment_hits_boxes(p, q, lo, hi):
    """Slab test: does segment p->q touch each axis-aligned box [lo, hi]? lo/hi: (..., 2)."""
    d = q - p
    t0 = np.zeros(lo.shape[:-1])
    t1 = np.ones(lo.shape[:-1])
    ok = np.ones(lo.shape[:-1], dtype=bool)
    for ax in range(2):
        if abs(d[ax]) < 1e-12:
            ok &= (p[ax] >= lo[..., ax]) & (p[ax] <= hi[..., ax])
        else:
            ta, tb = (lo[..., ax] - p[ax]) / d[ax], (hi[..., ax] - p[ax]) / d[ax]
            t0 = np.maximum(t0, np.minimum(ta, tb))
            t1 = np.minimum(t1, np.maximum(ta, tb))
    return ok & (t0 <= t1)


def _inside(poly, pts):
    """Even-odd point-in-polygon for pts (..., 2)."""
    x, y = pts[..., 0], pts[..., 1]
    inside = np.zeros(x.shape, dtype=bool)
    for (x0, y0), (x1, y1) in cyclic_pairs([tuple(p) for p in poly]):
        if y0 == y1:
            continue
        crosses = ((y0 > y) != (y1 > y)) & (x < (x1 - x0) * (y - y0) / (y1 - y0) + x0)
        inside ^= crosses
    return inside


def _cells(shape, lo_xy, hi_xy, res):
    """Index window and [lo, hi] corner arrays of the cells overlapping the bbox lo_xy..hi_xy."""
    h, w = shape
    j0, j1 = max(int(np.floor(lo_xy[0] / res)), 0), min(int(np.floor(hi_xy[0] / res)), w - 1)
    # row i covers y in [(h-1-i) res, (h-i) res]
    i0, i1 = max(h - 1 - int(np.floor(hi_xy[1] / res)), 0), min(h - 1 - int(np.floor(lo_xy[1] / res)), h - 1)
    if j1 < j0 or i1 < i0:
        return None
    ii, jj = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), indexing='ij')
    lo = np.stack([jj * res, (h - 1 - ii) * res], -1)
    return ii, jj, lo, lo + res


def masks(walls, spaces, res=RES):
    """(H, W) int16: -1 on walls, k on the k-th space (1-based), 0 elsewhere."""
    walls = np.asarray(walls, dtype=float)
    shape = mask_shape(walls, *[[s] for s in spaces]) if len(spaces) else mask_shape(walls)
    out = np.zeros(shape, dtype=np.int16)
    for k, poly in enumerate(spaces):
        poly = np.asarray(poly, dtype=float)
        win = _cells(shape, poly.min(0), poly.max(0), res)
        if win is None:
            continue
        ii, jj, lo, hi = win
        touched = _inside(poly, (lo + hi) / 2)
        for p, q in cyclic_pairs(list(poly)):
            touched |= _segment_hits_boxes(p, q, lo, hi)
        out[ii[touched], jj[touched]] = k + 1
    for a, b in walls:
        win = _cells(shape, np.minimum(a, b) - WALL_BUFFER, np.maximum(a, b) + WALL_BUFFER, res)
        if win is None:
            continue
        ii, jj, lo, hi = win
        touched = _segment_hits_boxes(a, b, lo - WALL_BUFFER, hi + WALL_BUFFER)
        out[ii[touched], jj[touched]] = -1
    return out


def centers(indices, shape, res):
    """Mask (i, j) indices -> (x, y) of the cell centres (geometry.py:110-122)."""
    i, j = indices[..., 0] + .5, indices[..., 1] + .5
    return res * np.stack([j, shape[0] - i], -1)


def indices(coords, shape, res):
    """(x, y) coordinates -> (i, j) of the containing cell (geometry.py:124-137)."""
    x, y = coords[..., 0], coords[..., 1]
    i = (shape[0] - y / res).clip(0, shape[0] - 1)
    j = (x / res).clip(0, shape[1] - 1)
    return np.stack([i, j], -1).astype(int)


def centroid(poly):
    """Area centroid of a simple polygon (the reference uses shapely's, geometry.py:95-97)."""
    p = np.asarray(poly, dtype=float)
    x0, y0 = p[:, 0], p[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    cr = x0 * y1 - x1 * y0
    area = cr.sum() / 2
    if abs(area) < 1e-12:
        return p.mean(0)
    return np.array([((x0 + x1) * cr).sum(), ((y0 + y1) * cr).sum()]) / (6 * area)
