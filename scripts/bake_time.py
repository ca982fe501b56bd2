"""One-off costs at the benchmark's size (not a benchmark): scene build on the host, upload, spatial table + visibility
grid, bake (over the table / brute force) next to the reference's own bake. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench     # noqa: E402
import common    # noqa: E402
from megastep_b200 import cuda, scene    # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) * 1e3


def main():
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cfg = dict(bench.WORKLOADS['deathmatch'])
    t = time.perf_counter()
    gs, arrays, pos, ang = bench.build_scene(cfg, n_envs, 256, 0)
    out = {'envs': n_envs, 'host_scene_build_ms': (time.perf_counter() - t) * 1e3, 'texels': int(len(arrays['textures'])),
           'lines': int(len(arrays['lines']))}
    params = cuda.make_params(bench.AGENT_RADIUS, cfg['res'], cfg['fov'], bench.FPS)
    s, out['upload_ms'] = timed(lambda: scene.upload(arrays))
    cuda.set_option('timing', 1)
    _, out['table_and_visibility_ms'] = timed(lambda: s._struct())
    out['vis_kernel_ms'] = cuda.get_option('time_ns_bake') / 1e6
    cuda.set_option('timing', 0)
    _, _ = timed(lambda: cuda.bake(s, params=params))                       # warm-up (attribute set-up)
    _, out['bake_table_ms'] = timed(lambda: cuda.bake(s, params=params))
    ours = s.baked.vals.clone()
    cuda.set_option('bake_brute', 1)
    _, out['bake_brute_ms'] = timed(lambda: cuda.bake(s, params=params))
    cuda.set_option('bake_brute', 0)
    out['table_equals_brute'] = bool(torch.equal(ours, s.baked.vals))
    ref = common.reference_module()
    if ref is not None:
        ref.initialize(bench.AGENT_RADIUS, cfg['res'], cfg['fov'], bench.FPS)
        rs = common.reference_scenery(ref, arrays)
        _, _ = timed(lambda: ref.bake(rs))
        _, out['reference_bake_ms'] = timed(lambda: ref.bake(rs))
        out['equals_reference'] = bool(torch.equal(ours, rs.baked.vals))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
