"""A/B timing of kernel variants in ONE process (not a benchmark): the fused step and its kernels under option sets.

    python scripts/ab.py [--workload deathmatch] [--steps 300] 'persist=2' 'persist=1,merge_dyn=2' ''

For each option set (comma-separated name=value pairs of msb_set_option; '' = defaults) prints one JSON line:
step_us (mean over steps, L2 flushed between steps, CUDA events per step), p50/p95, and the kernels timed apart.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench                                                    # noqa: E402
from megastep_b200 import cuda, modules, scene, core as core_   # noqa: E402

ALL = ('nch', 'threads', 'stage_rec', 'idx64', 'persist', 'merge_dyn', 'dyn_groups', 'stages', 'no_sched', 'dyn_warps', 'pdl',
       'no_prefetch', 'no_env_order', 'dyn_window', 'debug_skip_dyn', 'no_vis', 'fused_step', 'view_ctas_per_sm', 'bake_brute')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='deathmatch')
    ap.add_argument('--envs', type=int, default=None)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--no-raw', action='store_true')
    ap.add_argument('--stats', action='store_true', help='one extra step with the device counters on: where tick_kernel\'s warp-cycles go')
    ap.add_argument('sets', nargs='*', default=[''])
    args = ap.parse_args()
    cfg = dict(bench.WORKLOADS[args.workload])
    n_envs = args.envs or cfg['n_envs']
    gs, arrays, pos, ang = bench.build_scene(cfg, n_envs, 256, 0)
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(bench.AGENT_RADIUS, cfg['res'], cfg['fov'], bench.FPS))
    N, A = pos.shape[:2]
    K = args.steps
    acts = torch.as_tensor(np.random.RandomState(3).randint(0, 7, (K + 20, N, A)).astype(np.int32)).cuda()
    flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device='cuda')
    for spec in args.sets:
        for name in ALL:
            cuda.set_option(name, 1 if name == 'pdl' else 0)
        for kv in filter(None, spec.split(',')):
            name, v = kv.split('=')
            cuda.set_option(name, int(v))
        c = core_.Core(s, res=cfg['res'], fov=cfg['fov'], fps=bench.FPS)
        c.agents.positions.copy_(torch.as_tensor(pos))
        c.agents.angles.copy_(torch.as_tensor(ang))
        step = modules.FusedStep(c, subsample=cfg['subsample'], raw=not args.no_raw)
        for i in range(20):
            step(acts[i])
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for i in range(K):
            step.actions.copy_(acts[20 + i])
            flush.fill_(0.)
            starts[i].record()
            step._plan()
            stops[i].record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in zip(starts, stops)])
        cuda.set_option('timing', 1)
        for i in range(min(K, 100)):
            step.actions.copy_(acts[20 + i])
            flush.fill_(0.)
            step._plan()
        torch.cuda.synchronize()
        km = {kind: round(cuda.get_option(f'time_ns_{kind}') / 1e3 / max(cuda.get_option(f'time_count_{kind}'), 1), 1)
              for kind in ('physics', 'render', 'dyn') if cuda.get_option(f'time_count_{kind}') > 0}
        cuda.set_option('timing', 0)
        st = None
        if args.stats:
            cuda.set_option('stats', 1)
            cuda.set_option('stats_reset', 0)
            step._plan()
            torch.cuda.synchronize()
            names = ('items', 'wait', 'dyn', 'drain', 'prep', 'total', 'n_dyn', 'n_items', 'max_warp', 'n_spun', 'fetch')
            raw = {n: cuda.get_option(f'stat_{16 + i}') for i, n in enumerate(names)}
            tot = max(raw['total'], 1)
            st = {n: round(raw[n] / tot, 4) for n in ('items', 'wait', 'dyn', 'drain', 'prep', 'fetch')}
            st.update(n_dyn=raw['n_dyn'], n_items=raw['n_items'], n_spun=raw['n_spun'], max_warp_cycles=raw['max_warp'],
                      tests=cuda.get_option('stat_tests'), groups=cuda.get_option('stat_groups'), total_cycles=tot)
            cuda.set_option('stats', 0)
        print(json.dumps({'options': spec, 'stats': st, 'workload': args.workload, 'envs': N, 'step_us': round(float(ms.mean()) * 1e3, 1),
                          'p50': round(float(np.percentile(ms, 50)) * 1e3, 1), 'p95': round(float(np.percentile(ms, 95)) * 1e3, 1),
                          'kernels_apart_us': km, 'checksum': float(step._plan.rgb.double().sum())}), flush=True)
        del step, c


if __name__ == '__main__':
    main()
