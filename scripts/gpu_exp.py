"""Timing experiments (not a benchmark): render/step on Deathmatch 4096x4x128 under debug switches."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common
from megastep_b200 import cuda, modules, scene, synthetic, core as core_

def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def setup(N=4096, A=4, res=128, fov=70.):
    gs = synthetic.sample(N, seed=1, n_unique=256)
    arrays = synthetic.tile_arrays(scene.scene_arrays(gs[:256], A, np.random.RandomState(1)), N)
    pos, ang = synthetic.spawns(gs, A, np.random.RandomState(2))
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, res, fov, 10.))
    c = core_.Core(s, res=res, fov=fov, fps=10.)
    c.agents.positions.copy_(torch.as_tensor(pos)); c.agents.angles.copy_(torch.as_tensor(ang))
    return c

if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'exp'
    if len(sys.argv) > 2: cuda.OCCLUDER_RUN = int(sys.argv[2])
    if mode == 'order':
        out = {}
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        for order in ('morton', 'str', 'morton', 'str'):
            cuda.TABLE_ORDER = order
            c = setup()
            step = modules.FusedStep(c, subsample=1, raw=True)
            for _ in range(20): step(acts)
            tag = order + ('2' if f'step_us/{order}' in out else '')
            out[f'step_us/{tag}'] = round(timeit(lambda: step(acts), iters=200), 1)
            out[f'render_us/{tag}'] = round(timeit(lambda: c.render(), iters=200), 1)
            out[f'physics_us/{tag}'] = round(timeit(lambda: c.physics(), iters=200), 1)
            cuda.set_option('debug_skip_dyn', 1); out[f'main_us/{tag}'] = round(timeit(lambda: c.render(), iters=200), 1); cuda.set_option('debug_skip_dyn', 0)
            cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
            c.render(); torch.cuda.synchronize()
            for k in ('stat_tests', 'stat_groups', 'stat_dyn_iters', 'stat_dyn_scans', 'stat_replays'): out[f'{k}/{tag}'] = cuda.get_option(k)
            cuda.set_option('stats', 0)
            del step, c
        print(json.dumps(out)); sys.exit(0)
    c = setup()
    if mode == 'e2e':
        out = {}
        N, A = 4096, 4
        acts_host = torch.as_tensor(np.random.RandomState(3).randint(0, 7, (400, N, A)).astype(np.int32)).pin_memory()
        step = modules.FusedStep(c, subsample=1, raw=True)
        step._capture(host_io=True)
        stream = torch.cuda.current_stream()
        for i in range(50): step.step_host(acts_host[i])
        t = {'stage': 0., 'replay': 0., 'sync': 0.}
        t0 = time.perf_counter()
        for i in range(300):
            a = time.perf_counter(); step.actions_host.copy_(acts_host[50 + i]); b = time.perf_counter()
            step._graph.replay(); c_ = time.perf_counter()
            stream.synchronize(); d = time.perf_counter()
            t['stage'] += b - a; t['replay'] += c_ - b; t['sync'] += d - c_
        out['host_io_total_us'] = round((time.perf_counter() - t0) / 300 * 1e6, 1)
        for k_, v in t.items(): out[f'host_io_{k_}_us'] = round(v / 300 * 1e6, 1)
        # device time of the same graph alone
        out['graph_device_us'] = round(timeit(lambda: step._graph.replay(), iters=300), 1)
        # plain: separate copies
        step2 = modules.FusedStep(c, subsample=1, raw=True, graph=True)
        res = torch.empty((N, A), dtype=torch.float32).pin_memory()
        t0 = time.perf_counter()
        for i in range(300):
            step2.actions.copy_(acts_host[50 + i], non_blocking=True); step2._graph.replay(); res.copy_(step2._plan.progress, non_blocking=True); stream.synchronize()
        out['separate_total_us'] = round((time.perf_counter() - t0) / 300 * 1e6, 1)
        print(json.dumps(out)); sys.exit(0)
    if mode == 'quick':
        out = {}
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        step = modules.FusedStep(c, subsample=1, raw=True)
        for _ in range(20): step(acts)
        for rep in range(2):
            out[f'step_us/{rep}'] = round(timeit(lambda: step(acts), iters=300), 1)
            out[f'render_us/{rep}'] = round(timeit(lambda: c.render(), iters=300), 1)
            cuda.set_option('debug_skip_dyn', 1); out[f'main_us/{rep}'] = round(timeit(lambda: c.render(), iters=300), 1); cuda.set_option('debug_skip_dyn', 0)
        out['physics_us'] = round(timeit(lambda: c.physics(), iters=300), 1)
        out['bake_ms'] = round(timeit(lambda: cuda.bake(c.scenery, params=c.params), iters=3, warm=1) / 1e3, 1)
        cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
        c.render(); torch.cuda.synchronize()
        for k in ('stat_tests', 'stat_groups', 'stat_dyn_iters', 'stat_dyn_scans', 'stat_dyn_entries', 'stat_replays'): out[k] = cuda.get_option(k)
        cuda.set_option('stats', 0)
        print(json.dumps(out)); sys.exit(0)
    if mode == 'ncu':
        for _ in range(3): c.render()
        torch.cuda.synchronize()
        sys.exit(0)
    if mode == 'ncu_step':
        step = modules.FusedStep(c, subsample=1, raw=True)
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        for _ in range(24): step(acts)
        torch.cuda.synchronize()
        sys.exit(0)
    if mode == 'launches':
        step = modules.FusedStep(c, subsample=1, raw=True)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        for _ in range(3):
            c.physics(); c.render(); step(acts)
        torch.cuda.synchronize()
        sys.exit(0)
    if mode == 'run':
        step = modules.FusedStep(c, subsample=1, raw=True)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        res = {'run': cuda.OCCLUDER_RUN, 'render_us': timeit(lambda: c.render()), 'step_us': timeit(lambda: step(acts)),
               'phys_then_render_us': timeit(lambda: (c.physics(), c.render())), 'physics_us': timeit(lambda: c.physics())}
        cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
        c.render(); torch.cuda.synchronize()
        for k in ('stat_tests', 'stat_groups', 'stat_dyn_rays', 'stat_dyn_iters'): res[k] = cuda.get_option(k)
        cuda.set_option('stats', 0)
        print(json.dumps(res))
        sys.exit(0)
    if mode == 'dynstat':
        out = {}
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        step = modules.FusedStep(c, subsample=1, raw=True)
        for _ in range(20): step(acts)
        for dw in (1, 2):
            cuda.set_option('dyn_warps', dw)
            cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
            step(acts); torch.cuda.synchronize()
            for k in ('stat_dyn_rays', 'stat_dyn_iters', 'stat_dyn_entries', 'stat_dyn_cycles', 'stat_dyn_maxcyc', 'stat_dyn_warpmax', 'stat_dyn_slow', 'stat_dyn_kernel', 'stat_dyn_scans', 'stat_dyn_scans_lit', 'stat_dyn_iters_lit'):
                out[f'{k}/dw{dw}'] = cuda.get_option(k)
            cuda.set_option('stats', 0)
            cuda.set_option('timing', 1)
            for _ in range(50): step(acts)
            torch.cuda.synchronize()
            out[f'dyn_us/dw{dw}'] = round(cuda.get_option('time_ns_dyn') / cuda.get_option('time_count_dyn') / 1e3, 1)
            cuda.set_option('timing', 0)
        cuda.set_option('dyn_warps', 0)
        for skip in (0, 1):
            cuda.set_option('debug_skip_dyn', skip)
            cuda.set_option('timing', 1)
            for _ in range(50): step(acts)
            torch.cuda.synchronize()
            for kind in ('physics', 'render', 'dyn'):
                cnt = cuda.get_option(f'time_count_{kind}')
                if cnt: out[f'kernel_us/skip{skip}/{kind}'] = round(cuda.get_option(f'time_ns_{kind}') / cnt / 1e3, 1)
            cuda.set_option('timing', 0)
        cuda.set_option('debug_skip_dyn', 0)
        for threads in (32, 64, 128):
            cuda.set_option('threads', threads)
            out[f'physics_us/t{threads}'] = round(timeit(lambda: c.physics()), 1)
        cuda.set_option('threads', 0)
        print(json.dumps(out)); sys.exit(0)
    if mode == 'pdl':
        out = {}
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        step = modules.FusedStep(c, subsample=1, raw=True)
        for _ in range(20): step(acts)
        for rep in range(2):
            for pdl in (0, 1):
                for dw in (2, 4):
                    cuda.set_option('pdl', pdl); cuda.set_option('dyn_warps', dw)
                    out[f'step_us/pdl{pdl}/dw{dw}/rep{rep}'] = round(timeit(lambda: step(acts), iters=200), 1)
                    out[f'render_us/pdl{pdl}/dw{dw}/rep{rep}'] = round(timeit(lambda: c.render(), iters=200), 1)
        cuda.set_option('pdl', 1); cuda.set_option('dyn_warps', 0)
        print(json.dumps(out)); sys.exit(0)
    if mode == 'vis':
        out = {}
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        step = modules.FusedStep(c, subsample=1, raw=True)
        for _ in range(20): step(acts)
        vis = c.scenery._vis[0]
        out['vis_cells'] = vis.numel(); out['vis_bits_per_cell'] = round(float(sum(((vis >> i) & 1).sum().item() for i in range(32))) / vis.numel(), 2)
        for no_vis in (1, 0):
            cuda.set_option('no_vis', no_vis)
            out[f'step_us/novis{no_vis}'] = round(timeit(lambda: step(acts), iters=200), 1)
            cuda.set_option('timing', 1)
            for _ in range(50): step(acts)
            torch.cuda.synchronize()
            for kind in ('physics', 'render', 'dyn'):
                out[f'kernel_us/novis{no_vis}/{kind}'] = round(cuda.get_option(f'time_ns_{kind}') / cuda.get_option(f'time_count_{kind}') / 1e3, 1)
            cuda.set_option('timing', 0)
            cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
            step(acts); torch.cuda.synchronize()
            for k in ('stat_dyn_rays', 'stat_dyn_iters', 'stat_dyn_entries', 'stat_dyn_scans', 'stat_dyn_scans_lit'): out[f'{k}/novis{no_vis}'] = cuda.get_option(k)
            cuda.set_option('stats', 0)
        import time
        from megastep_b200 import cuda as cu
        torch.cuda.synchronize(); t0 = time.time()
        cu._check(cu._lib.msb_build_visibility(__import__('ctypes').byref(c.scenery._c), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize(); out['build_vis_ms'] = round((time.time() - t0) * 1e3, 1)
        print(json.dumps(out)); sys.exit(0)
    if mode == 'view':
        out = {}
        for nch, threads in ((0, 0), (4, 128), (2, 256), (2, 128), (2, 64), (1, 128), (1, 256)):
            cuda.set_option('nch', nch); cuda.set_option('threads', threads)
            out[f'render_us/nch{nch}/t{threads}'] = round(timeit(lambda: c.render()), 1)
            cuda.set_option('debug_skip_dyn', 1); out[f'main_us/nch{nch}/t{threads}'] = round(timeit(lambda: c.render()), 1); cuda.set_option('debug_skip_dyn', 0)
        cuda.set_option('nch', 0); cuda.set_option('threads', 0)
        for w in (1, 2, 4, 8):
            cuda.set_option('dyn_window', w)
            out[f'render_us/window{w}'] = round(timeit(lambda: c.render()), 1)
            cuda.set_option('timing', 1)
            for _ in range(30): c.render()
            torch.cuda.synchronize()
            out[f'dyn_kernel_us/window{w}'] = round(cuda.get_option('time_ns_dyn') / max(1, cuda.get_option('time_count_dyn')) / 1e3, 1)
            cuda.set_option('timing', 0)
            cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
            c.render(); torch.cuda.synchronize()
            out[f'dyn_entries/window{w}'] = cuda.get_option('stat_dyn_entries'); out[f'dyn_iters/window{w}'] = cuda.get_option('stat_dyn_iters')
            cuda.set_option('stats', 0)
        cuda.set_option('dyn_window', 0)
        cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
        c.render(); torch.cuda.synchronize()
        for k in ('stat_tests', 'stat_groups', 'stat_dyn_rays', 'stat_dyn_iters', 'stat_dyn_entries', 'stat_replays'): out[k] = cuda.get_option(k)
        cuda.set_option('stats', 0)
        out['physics_us'] = round(timeit(lambda: c.physics()), 1)
        torch.manual_seed(0)
        acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
        for fused in (0, 1):
            cuda.set_option('fused_step', fused)
            c2 = setup()
            step = modules.FusedStep(c2, subsample=1, raw=True)
            out[f'step_us/fused{fused}'] = round(timeit(lambda: step(acts)), 1)
            cuda.set_option('timing', 1)
            for _ in range(50): step(acts)
            torch.cuda.synchronize()
            for kind in ('physics', 'render', 'dyn', 'step'):
                cnt = cuda.get_option(f'time_count_{kind}')
                if cnt: out[f'kernel_us/fused{fused}/{kind}'] = round(cuda.get_option(f'time_ns_{kind}') / cnt / 1e3, 1)
            cuda.set_option('timing', 0)
            cuda.set_option('stats', 1); cuda.set_option('stats_reset', 0)
            step(acts); torch.cuda.synchronize()
            for k in ('stat_tests', 'stat_groups', 'stat_dyn_rays', 'stat_dyn_iters', 'stat_dyn_entries'): out[f'after_steps/fused{fused}/{k}'] = cuda.get_option(k)
            cuda.set_option('stats', 0)
        cuda.set_option('fused_step', 0)
        print(json.dumps(out)); sys.exit(0)
    out = {}
    for skip in (0, 1):
        cuda.set_option('debug_skip_dyn', skip)
        for nch in (1, 2, 4):
            cuda.set_option('nch', nch)
            out[f'render_us/skip_dyn{skip}/nch{nch}'] = timeit(lambda: c.render())
    cuda.set_option('nch', 0); cuda.set_option('debug_skip_dyn', 0)
    print(json.dumps(out))
