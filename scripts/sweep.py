"""BASELINE.json configs[4] (reduced): ray-width x env-count sweep on ONE GPU, ours vs the reference's own CUDA build,
kernel-level (physics + render through each API, outputs materialised, no L2 flush, 20 iterations after 3 warm-ups).
Writes gpurun_out/sweep.json. Cells the reference cannot run (32-bit accessors) are recorded as null."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common  # noqa: E402
from megastep_b200 import core as core_, cuda, scene, synthetic  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ref = common.reference_module()
    A, fov = 4, 70.
    Ns = [int(x) for x in os.environ.get('SWEEP_N', '256,1024,4096,16384,65536').split(',')]
    Rs = [int(x) for x in os.environ.get('SWEEP_R', '16,64,128,512').split(',')]
    base = scene.scene_arrays(synthetic.sample(256, seed=1), A, np.random.RandomState(1))
    out = []
    for N in Ns:
        gs = synthetic.sample(N, seed=1, n_unique=256)
        arrays = synthetic.tile_arrays(base, N)
        pos, ang = synthetic.spawns(gs, A, np.random.RandomState(2))
        vel = torch.as_tensor((1.5 * np.random.RandomState(3).normal(size=(N, A, 2))).astype(np.float32)).cuda()
        s = scene.upload(arrays)
        cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 64, fov, 10.))
        rs = None
        if ref is not None:
            try:
                rs = common.reference_scenery(ref, arrays)
                rs.baked.vals.copy_(s.baked.vals)
            except Exception as e:  # noqa: BLE001
                print('reference scenery failed at N =', N, type(e).__name__, flush=True)
        for R in Rs:
            c = core_.Core(s, res=R, fov=fov, fps=10.)
            c.agents.positions.copy_(torch.as_tensor(pos))
            c.agents.angles.copy_(torch.as_tensor(ang))

            def ours():
                c.agents.velocity.copy_(vel)
                c.physics()
                c.render()
            rec = {'n_envs': N, 'n_agents': A, 'res': R, 'ours_s': timeit(ours)}
            rec['ours_afps'] = N * A / rec['ours_s']
            rec['ref_s'] = rec['ref_afps'] = None
            if rs is not None and R <= 1024:
                try:
                    ref.initialize(common.AGENT_RADIUS, R, fov, 10.)
                    ra = ref.Agents(angles=torch.as_tensor(ang).cuda(), positions=torch.as_tensor(pos).cuda(),
                                    angvelocity=torch.zeros(N, A, device='cuda'), velocity=torch.zeros(N, A, 2, device='cuda'))

                    def theirs():
                        ra.velocity.copy_(vel)
                        ref.physics(rs, ra)
                        ref.render(rs, ra)
                    rec['ref_s'] = timeit(theirs, iters=5, warm=1)
                    rec['ref_afps'] = N * A / rec['ref_s']
                except Exception as e:  # noqa: BLE001
                    rec['ref_error'] = f'{type(e).__name__}: {str(e)[:100]}'
            rec['speedup'] = rec['ref_s'] / rec['ours_s'] if rec['ref_s'] else None
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del c
        del s, rs
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
