"""BASELINE.json configs[4]: ray-width x env-count sweep at 1 / 2 / 4 / 8 GPUs, ours next to the reference's own CUDA build
(one GPU, rank 0 of a single-process run), kernel-level: physics + render through each API, all five Render outputs
materialised, 20 iterations after 3 warm-ups, CUDA events, max over ranks. Envs are sharded evenly over the ranks (N is
the TOTAL env count); no communication. The scenes are 256 distinct synthetic floorplans cycled and repeated on the
device (scene.tiled_scenery), so that 262,144 envs build in seconds. Cells the reference cannot run (its 32-bit accessors
overflow beyond 2^31 texel floats; an out-of-memory failure) are recorded as null.

    python scripts/sweep.py                       # 1 GPU, with the reference
    torchrun --nproc-per-node 8 scripts/sweep.py  # 8 GPUs (ours only)
    SWEEP_N=..., SWEEP_R=..., SWEEP_OUT=path      # subsets / output file (JSON list)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common  # noqa: E402
from megastep_b200 import core as core_, cuda, scene, sharding, synthetic  # noqa: E402


def timeit(fn, iters, warm, world):
    import torch.distributed as dist
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e-3], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        sharding.initialize()
    ref = common.reference_module() if world == 1 and not os.environ.get('SWEEP_NO_REF') else None
    A, fov = 4, 70.
    Ns = [int(x) for x in os.environ.get('SWEEP_N', '256,1024,4096,16384,65536,262144').split(',')]
    Rs = [int(x) for x in os.environ.get('SWEEP_R', '16,32,64,128,256,512').split(',')]
    base_gs = synthetic.sample(256, seed=1)
    base = scene.scene_arrays(base_gs, A, np.random.RandomState(1))
    params = cuda.make_params(common.AGENT_RADIUS, 64, fov, 10.)
    out = []
    for N in Ns:
        lo, hi = sharding.shard_range(N, rank, world)
        n = hi - lo
        if n == 0 or N % world:
            continue
        # this rank's envs: global env i uses floorplan i % 256; start the cycle at lo % 256 by rolling the base
        order = (np.arange(256) + lo) % 256
        local_base = sharding_take(base, order) if lo % 256 else base
        gs = [base_gs[(lo + i) % 256] for i in range(n)]
        pos, ang = synthetic.spawns(gs, A, np.random.RandomState(2 + rank))
        vel = torch.as_tensor((1.5 * np.random.RandomState(3 + rank).normal(size=(n, A, 2))).astype(np.float32)).cuda()
        s = scene.tiled_scenery(local_base, n, params=params)
        rs = None
        if ref is not None:
            try:
                distinct = scene.upload(local_base)
                cuda.bake(distinct, params=params)
                rs = common.reference_scenery_tiled(ref, local_base, n, baked=distinct.baked.vals)
                del distinct
            except Exception as e:  # noqa: BLE001
                print(f'reference scenery failed at N = {N}: {type(e).__name__}: {str(e)[:120]}', flush=True)
        for R in Rs:
            c = core_.Core(s, res=R, fov=fov, fps=10.)
            c.agents.positions.copy_(torch.as_tensor(pos))
            c.agents.angles.copy_(torch.as_tensor(ang))

            def ours():
                c.agents.velocity.copy_(vel)
                c.physics()
                c.render()
            rec = {'n_gpus': world, 'n_envs': N, 'n_agents': A, 'res': R, 'ours_s': timeit(ours, 20, 3, world)}
            rec['ours_afps'] = N * A / rec['ours_s']
            rec['ref_s'] = rec['ref_afps'] = None
            if rs is not None and R <= 1024:
                try:
                    ref.initialize(common.AGENT_RADIUS, R, fov, 10.)
                    ra = ref.Agents(angles=torch.as_tensor(ang).cuda(), positions=torch.as_tensor(pos).cuda(),
                                    angvelocity=torch.zeros(n, A, device='cuda'), velocity=torch.zeros(n, A, 2, device='cuda'))

                    def theirs():
                        ra.velocity.copy_(vel)
                        ref.physics(rs, ra)
                        ref.render(rs, ra)
                    rec['ref_s'] = timeit(theirs, 5 if N * R < (1 << 24) else 2, 1, 1)
                    rec['ref_afps'] = N * A / rec['ref_s']
                    del ra
                except Exception as e:  # noqa: BLE001
                    rec['ref_error'] = f'{type(e).__name__}: {str(e)[:100]}'
            rec['speedup'] = rec['ref_s'] / rec['ours_s'] if rec['ref_s'] else None
            if rank == 0:
                print(json.dumps(rec), flush=True)
            out.append(rec)
            del c
            torch.cuda.empty_cache()
        del s, rs
        torch.cuda.empty_cache()
    if rank == 0:
        path = os.environ.get('SWEEP_OUT', os.path.join(ROOT, 'gpurun_out', f'sweep_{world}gpu.json'))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(out, open(path, 'w'), indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


def sharding_take(arrays, order):
    """The envs of a scene_arrays dict in another order."""
    parts = [sharding.shard_arrays(arrays, int(i), int(i) + 1) for i in order]
    return dict(n_agents=arrays['n_agents'], model=arrays['model'],
                **{k: np.concatenate([p[k] for p in parts]) for k in ('lines', 'line_widths', 'lights', 'light_widths', 'textures', 'tex_widths')})


if __name__ == '__main__':
    main()
