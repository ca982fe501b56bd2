"""BASELINE.json configs[0] (C1): ONE floorplan's scene build + geometry.masks on the CPU — plumbing, no CUDA step.
Times, single-threaded on the host: the synthetic floorplan standing in for a Cubicasa geometry (the dataset needs a
download), `geometry.masks` (this package's shapely/rasterio-free restatement of megastep/geometry.py:81-93),
`scene.scene_arrays` (megastep/scene.py:75-100's CPU half) and the C oracle's `bake` of that one env; then checks the
shapes / semantics the reference documents. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megastep_b200 import geometry, scene, synthetic   # noqa: E402


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return out, min(ts) * 1e3


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_004
    g, t_plan = best(lambda: synthetic.floorplan(seed))
    polys = [np.array([[r[0], r[1]], [r[2], r[1]], [r[2], r[3]], [r[0], r[3]]]) for r in g.rooms]
    masks, t_masks = best(lambda: geometry.masks(g.walls, polys))
    arrays, t_arrays = best(lambda: scene.scene_arrays([g], 4, np.random.RandomState(1)))
    out = {'config': 'C1: one floorplan, scene build + geometry.masks on CPU', 'host_cores': os.cpu_count(), 'threads_used': 1,
           'walls': int(len(g.walls)), 'lights': int(len(g.lights)), 'texels': int(len(arrays['textures'])),
           'mask_shape': list(masks.shape), 'floorplan_ms': t_plan, 'masks_ms': t_masks, 'scene_arrays_ms': t_arrays}
    try:
        from oracle import oracle
        _, out['oracle_bake_ms'] = best(lambda: oracle.bake(arrays), reps=3)
        out['oracle_threads'] = oracle.num_threads()
    except Exception as e:  # noqa: BLE001
        out['oracle_bake_ms'] = None
        out['oracle_error'] = str(e)[:100]
    # documented semantics (docs/concepts.rst:251-259): -1 walls, 0 free, k >= 1 room ids; 0.2 m cells; 1 m margin
    assert masks.dtype == np.int16 and masks.min() == -1 and masks.max() == len(g.rooms)
    assert (masks > 0).mean() > .4 and (masks == -1).mean() > .02
    h, w = masks.shape
    assert abs(h * geometry.RES - (g.walls[..., 1].max() + geometry.MARGIN)) <= 2 * geometry.RES
    assert arrays['line_widths'][0] == 32 + len(g.walls) and arrays['tex_widths'].sum() == len(arrays['textures'])
    out['checks'] = 'ok'
    print(json.dumps(out))


if __name__ == '__main__':
    main()
