"""Diagnostics on a real GPU: parity statistics against the reference build and the oracle (printed, not asserted),
candidate-test counters, and kernel-level timings of ours vs the reference. Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common  # noqa: E402
from megastep_b200 import cuda, modules, scene, synthetic, core as core_  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = {}


def exact(a, b):
    return float((((a == b) | (a.isnan() & b.isnan())).float().mean()))


def parity(name, kind, N, A, res, fov, ref):
    if kind == 'synthetic':
        gs, arrays = common.synthetic_scene(N, A, seed=11)
    else:
        gs, arrays = common.toy_scene(kind, N, A, seed=11)
    st = common.random_state(gs, A, seed=12)
    c = common.to_device(arrays, st, res, fov)
    rec = {}
    r = c.render()
    torch.cuda.synchronize()
    want = oracle.render(arrays, st, res=res, fov=fov)
    d, really = common.index_agreement(r.indices.cpu().numpy(), want['indices'], r.distances.cpu().numpy(), want['distances'])
    rec['oracle_idx_differ'], rec['oracle_idx_really'] = d, really
    if ref is not None:
        ref.initialize(common.AGENT_RADIUS, res, fov, 10.)
        rs, ra = common.reference_scenery(ref, arrays), common.reference_agents(ref, st)
        rr = ref.render(rs, ra)
        torch.cuda.synchronize()
        rec['ref_idx_exact'] = float((r.indices == rr.indices).float().mean())
        for k in ('locations', 'dots', 'distances', 'screen'):
            a, b = getattr(r, k), getattr(rr, k)
            rec[f'ref_{k}_exact'] = exact(a, b)
            fin = torch.isfinite(a) & torch.isfinite(b)
            rec[f'ref_{k}_maxabs'] = float((a[fin] - b[fin]).abs().max()) if fin.any() else 0.
        rec['ref_lines_exact'] = exact(c.scenery.lines.vals, rs.lines.vals)
        # physics, 3 ticks
        for tick in range(3):
            p, rp = c.physics(), ref.physics(rs, ra)
            torch.cuda.synchronize()
            rec[f'phys{tick}_flags_exact'] = float(((p.progress < 1) == (rp.progress < 1)).float().mean())
            rec[f'phys{tick}_progress_exact'] = exact(p.progress, rp.progress)
            for k in ('positions', 'angles', 'velocity', 'angvelocity'):
                a, b = getattr(c.agents, k), getattr(ra, k)
                rec[f'phys{tick}_{k}_exact'] = exact(a, b)
                rec[f'phys{tick}_{k}_maxabs'] = float((a - b).abs().max())
            kick = torch.randn_like(c.agents.velocity) * 2
            c.agents.velocity.add_(kick)
            ra.velocity.add_(kick)
    OUT[f'parity/{name}'] = rec
    print(name, json.dumps(rec), flush=True)


def bake_parity(ref):
    gs, arrays = common.synthetic_scene(6, 2, seed=21, bake=True)
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 64, 130., 10.))
    got = s.baked.vals.cpu().numpy()
    rec = {'oracle_frac_bad': float((np.abs(got - arrays['baked']) > 1e-4).mean())}
    if ref is not None:
        ref.initialize(common.AGENT_RADIUS, 64, 130., 10.)
        rs = common.reference_scenery(ref, {k: v for k, v in arrays.items() if k != 'baked'})
        ref.bake(rs)
        torch.cuda.synchronize()
        rec['ref_exact'] = exact(s.baked.vals, rs.baked.vals)
        rec['ref_maxabs'] = float((s.baked.vals - rs.baked.vals).abs().max())
    OUT['parity/bake'] = rec
    print('bake', json.dumps(rec), flush=True)


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def timings(ref, N=4096, A=4, res=128, fov=70., tag='deathmatch'):
    gs = synthetic.sample(N, seed=1, n_unique=256)
    arrays = synthetic.tile_arrays(scene.scene_arrays(gs[:256], A, np.random.RandomState(1)), N)
    pos, ang = synthetic.spawns(gs, A, np.random.RandomState(2))
    s = scene.upload(arrays)
    t = time.time()
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, res, fov, 10.))
    torch.cuda.synchronize()
    rec = {'bake_s': time.time() - t}
    c = core_.Core(s, res=res, fov=fov, fps=10.)
    c.agents.positions.copy_(torch.as_tensor(pos))
    c.agents.angles.copy_(torch.as_tensor(ang))
    vel = torch.randn_like(c.agents.velocity) * 1.5

    def phys():
        c.agents.velocity.copy_(vel)
        c.physics()
    rec['ours_physics_us'] = timeit(phys)
    c.agents.positions.copy_(torch.as_tensor(pos))
    for nch in (0, 1, 2, 4):
        for threads in (0, 64, 128, 256):
            cuda.set_option('nch', nch)
            cuda.set_option('threads', threads)
            rec[f'ours_render_us/nch{nch}/t{threads}'] = timeit(lambda: c.render())
    cuda.set_option('nch', 0)
    cuda.set_option('threads', 0)
    # candidate-test counters
    cuda.set_option('stats', 1)
    cuda.set_option('stats_reset', 0)
    c.render()
    torch.cuda.synchronize()
    for k in ('stat_tests', 'stat_groups', 'stat_dyn_rays', 'stat_dyn_iters'):
        rec[k] = cuda.get_option(k)
    rec['brute_force_tests'] = int(arrays['line_widths'].astype(np.int64).sum()) * A * ((res + 31) // 32)
    cuda.set_option('stats', 0)
    step = modules.FusedStep(c, subsample=1, raw=True)
    acts = torch.randint(0, 7, (N, A), dtype=torch.int32, device='cuda')
    rec['ours_fused_step_us'] = timeit(lambda: step(acts))
    step2 = modules.FusedStep(c, subsample=1, raw=False)
    rec['ours_fused_step_noraw_us'] = timeit(lambda: step2(acts))
    if ref is not None:
        ref.initialize(common.AGENT_RADIUS, res, fov, 10.)
        rs = common.reference_scenery(ref, arrays)
        rs.baked.vals.copy_(s.baked.vals)
        ra = ref.Agents(angles=torch.as_tensor(ang).cuda(), positions=torch.as_tensor(pos).cuda(),
                        angvelocity=torch.zeros(N, A, device='cuda'), velocity=torch.zeros(N, A, 2, device='cuda'))

        def rphys():
            ra.velocity.copy_(vel)
            ref.physics(rs, ra)
        rec['ref_physics_us'] = timeit(rphys, iters=20)
        ra.positions.copy_(torch.as_tensor(pos))
        rec['ref_render_us'] = timeit(lambda: ref.render(rs, ra), iters=20)
    OUT[f'timing/{tag}'] = rec
    print(tag, json.dumps(rec), flush=True)


if __name__ == '__main__':
    ref = common.reference_module()
    print('reference build loaded:', ref is not None, flush=True)
    which = sys.argv[1:] or ['parity', 'timing']
    if 'parity' in which:
        parity('box', 'box', 3, 1, 64, 130., ref)
        parity('column', 'column', 2, 2, 32, 90., ref)
        parity('syn-explorer', 'synthetic', 24, 1, 64, 130., ref)
        parity('syn-deathmatch', 'synthetic', 16, 4, 128, 70., ref)
        parity('syn-48', 'synthetic', 5, 3, 48, 100., ref)
        parity('syn-512', 'synthetic', 4, 2, 512, 70., ref)
        bake_parity(ref)
    if 'timing' in which:
        timings(ref)
        timings(ref, N=4096, A=1, res=64, fov=130., tag='explorer')
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(OUT, open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w'), indent=1)
