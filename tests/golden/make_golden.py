"""Generates tests/golden/reference_*.npz: outputs of the reference's OWN CUDA build (oracle/_ref, compiled from its
unmodified sources by oracle/build_ref.sh) on small seeded scenes. Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/*.npz gpurun_out/'

The inputs are rebuilt from seeds by tests/common.py, so the files only hold the reference's outputs:
render (indices, locations, dots, distances, screen, lines-after-draw), bake, and three physics ticks.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402

CASES = {
    # name: (kind, n_envs, n_agents, res, fov, seed)
    'box': ('box', 2, 1, 64, 130., 101),
    'column': ('column', 2, 2, 32, 90., 102),
    'explorer': ('synthetic', 6, 1, 64, 130., 103),
    'deathmatch': ('synthetic', 4, 4, 128, 70., 104),
}


def inputs(kind, N, A, seed):
    if kind == 'synthetic':
        gs, arrays = common.synthetic_scene(N, A, seed=seed, bake=False)
    else:
        from megastep_b200 import scene, toys
        gs = [getattr(toys, kind)()] * N
        arrays = scene.scene_arrays(gs, A, np.random.RandomState(seed))
    return gs, arrays, common.random_state(gs, A, seed=seed + 1)


def kicks(shape, seed):
    rng = np.random.RandomState(seed)
    return [(2 * rng.normal(size=shape)).astype(np.float32) for _ in range(3)]


if __name__ == '__main__':
    ref = common.reference_module()
    assert ref is not None, 'oracle/_ref is not built'
    for name, (kind, N, A, res, fov, seed) in CASES.items():
        gs, arrays, st = inputs(kind, N, A, seed)
        ref.initialize(common.AGENT_RADIUS, res, fov, 10.)
        rs = common.reference_scenery(ref, arrays)
        ref.bake(rs)
        out = {'baked': rs.baked.vals.cpu().numpy()}
        ra = common.reference_agents(ref, st)
        r = ref.render(rs, ra)
        for k in ('indices', 'locations', 'dots', 'distances', 'screen'):
            out[f'render_{k}'] = getattr(r, k).cpu().numpy()
        out['render_lines'] = rs.lines.vals.cpu().numpy()
        for t, kick in enumerate(kicks(st['velocity'].shape, seed + 2)):
            ra.velocity.add_(torch.as_tensor(kick).cuda())
            p = ref.physics(rs, ra)
            out[f'physics{t}_progress'] = p.progress.cpu().numpy()
            for k in ('angles', 'positions', 'angvelocity', 'velocity'):
                out[f'physics{t}_{k}'] = getattr(ra, k).cpu().numpy()
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(HERE, f'reference_{name}.npz'), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith('render_i') or k == 'baked'})
