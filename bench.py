#!/usr/bin/env python
"""bench.py — agent-frames/s of the physics()+render() hot path on synthetic cubicasa-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload deathmatch|explorer|...]
                    [--envs E] [--gather [--obs-dtype float32|float16|uint8]]

A "step" is one whole environment tick over the batch: MomentumMovement (random actions) -> physics -> render (all
five Render tensors materialised, as the reference's render() does) -> RGB/Depth/IMU observation heads.
  * ours:      msb_step through the C ABI: movement+physics kernel, render+heads kernel, agent-hit lighting kernel —
               three back-to-back launches (replayed as one CUDA graph), no host round trip.
  * reference: the reference's OWN kernels.cu/wrappers.cpp built unmodified for sm_100a (oracle/_ref) driven by the
               reference's OWN unmodified Python (baseline/_ref: core.Core, modules.MomentumMovement / render / RGB /
               Depth / IMU) through tests/common.py::reference_package — the same shim the parity tests use. megastep
               has no CPU step path (docs/faq.rst:23-27), so this — not a CPU run — is the reference arm; if oracle/_ref
               is not loadable the arm falls back to timing the CPU oracle port on the host cores.
Under torchrun (N > 1) every rank steps a replica of the same shard (identical work per GPU: weak scaling; --distinct-shards
gives each rank its own floorplans); no per-step collective unless --gather (ShardedCore's packed all-gather).

Prints ONE JSON line (rank 0), the last line of stdout. See DESIGN.md §5 for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

WORKLOADS = {
    # BASELINE.json configs[2] (the configuration the metric is quoted on) and configs[1]
    'deathmatch': dict(n_envs=4096, n_agents=4, res=128, fov=70., subsample=1),
    'explorer': dict(n_envs=4096, n_agents=1, res=64, fov=130., subsample=1),
    # the demo-faithful variants (render at 4x, subsample by 4: demo/envs/deathmatch.py:26-28, explorer.py:13-15)
    'deathmatch-demo': dict(n_envs=4096, n_agents=4, res=512, fov=70., subsample=4),
    'explorer-demo': dict(n_envs=4096, n_agents=1, res=256, fov=130., subsample=4),
}
AGENT_RADIUS = .15 / 2 ** .5
FPS = 10.
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='deathmatch', choices=sorted(WORKLOADS))
    ap.add_argument('--envs', type=int, default=None, help='envs per GPU (default: the workload\'s)')
    ap.add_argument('--res', type=int, default=None)
    ap.add_argument('--unique', type=int, default=256, help='distinct floorplans, tiled cyclically')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-raw', action='store_true', help='skip materialising the five Render tensors (obs heads only)')
    ap.add_argument('--gather', action='store_true', help='all-gather the observations to every rank each step (NCCL, side stream, overlapping the next step)')
    ap.add_argument('--obs-dtype', default='float32', choices=['float32', 'float16', 'uint8'], help='precision of the gathered observations (uint8: rgb and depth quantised to 8 bits, imu as fp16)')
    ap.add_argument('--gather-transport', default='nccl', choices=['nccl', 'p2p'], help='nccl: all_gather_into_tensor; p2p: symmetric memory + copy-engine peer copies (no SMs)')
    ap.add_argument('--distinct-shards', action='store_true', help='every rank builds its own floorplans / poses / actions (seeds + rank) instead of a replica of rank 0\'s shard')
    ap.add_argument('--no-graph', action='store_true', help='e2e leg: plain launches instead of a CUDA-graph replay')
    ap.add_argument('--e2e', default='torch', choices=['native', 'torch'], help='e2e leg: the library\'s own host-driven graph (one call per tick) or a torch CUDA graph between PyTorch copies')
    ap.add_argument('--dry-run', action='store_true', help='build the scene and the CPU baseline only (no GPU)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# scene
# ---------------------------------------------------------------------------------------------------------------
def build_scene(cfg, n_envs, unique, rank):
    from megastep_b200 import scene, synthetic
    gs = synthetic.sample(n_envs, seed=1 + rank, n_unique=unique)
    base = scene.scene_arrays(gs[:min(unique, n_envs)], cfg['n_agents'], np.random.RandomState(1 + rank))
    arrays = synthetic.tile_arrays(base, n_envs)
    pos, ang = synthetic.spawns(gs, cfg['n_agents'], np.random.RandomState(2 + rank))
    return gs, arrays, pos, ang


def algorithmic_bytes(cfg, arrays, fused, raw=True):
    """Bytes one step must move, per SURVEY.md §8(d): B_env = 64A + 32AF + 32W + 68AR + 16 for the separate
    physics + render calls (every Render output materialised, texel gathers counted per ray). The fused kernel
    stages each env's segments once, so the second read of the static lines and of the agent state (16W + 12A + 8)
    drops out; the observation heads add 16A*R/sub + 12A written."""
    A, R, F = cfg['n_agents'], cfg['res'], 8
    N = len(arrays['line_widths'])
    W = float(arrays['line_widths'].mean()) - A * F
    b_env = 64 * A + 32 * A * F + 32 * W + 68 * A * R + 16
    if fused:
        b_env -= 16 * W + 12 * A + 8
    if not raw:
        b_env -= 28 * A * R
    b_env += 16 * A * R / cfg['subsample'] + 12 * A
    return b_env * N, W


def render_kernel_bytes(cfg, arrays, raw=True):
    """Algorithmic bytes of ONE launch of view_kernel (draw + raycast + shade + heads), the dominant kernel:
    the render terms of SURVEY.md §8(d) — 16AF (draw) + [12A + 16L + 8] + 16AR (raycast) + (40 + 12)AR (shade) — plus
    the fused heads, 16AR/sub + 12A. Texel gathers are counted per ray with no reuse, as §8(d) specifies."""
    A, R, F = cfg['n_agents'], cfg['res'], 8
    N = len(arrays['line_widths'])
    L = float(arrays['line_widths'].mean())
    b_env = 16 * A * F + (12 * A + 16 * L + 8) + 16 * A * R + 52 * A * R + 16 * A * R / cfg['subsample'] + 12 * A
    if not raw:
        b_env -= 28 * A * R
    return b_env * N


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of view_kernel, from the committed ncu --set full
    capture of the same workload (profiles/r02_kernels_full_summary.json, entry of full_view_kernel_raw.csv)."""
    path = os.path.join(ROOT, 'profiles', 'r02_kernels_full_summary.json')
    try:
        recs = [r for r in json.load(open(path)) if 'view_kernel' in r['Kernel Name'] and r.get('source') == 'full_view_kernel_raw.csv']
        mb = [float(r['dram__bytes_read.sum']) + float(r['dram__bytes_write.sum']) for r in recs]
        return sum(mb) / len(mb) * 1e6, os.path.relpath(path, ROOT)
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return None, None


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(', '))

    def mark(self, wait_s=5.):
        """Call right before the timed region: waits for nvidia-smi to deliver its first sample, then discards
        everything sampled so far."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < wait_s:
            time.sleep(.02)
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[getattr(self, 'first', 0):]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores, bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, arrays, pos, ang, budget_s=12.):
    from megastep_b200 import sharding
    from oracle import oracle
    A, R = cfg['n_agents'], cfg['res']
    rng = np.random.RandomState(7)

    def run(n):
        sub = sharding.shard_arrays(arrays, 0, n)
        sub['baked'] = np.ones(len(sub['textures']), np.float32) if 'baked' not in sub else sub['baked']
        st = dict(angles=ang[:n].copy(), positions=pos[:n].copy(), angvelocity=np.zeros((n, A), np.float32),
                  velocity=(2 * rng.normal(size=(n, A, 2))).astype(np.float32))
        t = time.perf_counter()
        oracle.physics(sub, st, fps=FPS)
        oracle.render(sub, st, res=R, fov=cfg['fov'])
        return time.perf_counter() - t

    n = min(64, len(pos))
    t = run(n)
    n = int(min(len(pos), max(n, n * budget_s / max(t, 1e-3))))
    reps, total = 0, 0.
    while total < 3. and reps < 64:
        total += run(n)
        reps += 1
    return {'value': n * A * reps / total, 'unit': 'agent-frames/s', 'cores': oracle.num_threads(), 'kind': 'port',
            'sample': f'oracle/megastep_oracle.c physics+render, {reps} step(s) over the first {n} of {len(pos)} envs x {A} agents x {R} rays, '
                      f'{oracle.num_threads()} OpenMP threads, {total:.2f} s wall'}


# ---------------------------------------------------------------------------------------------------------------
# the two arms
# ---------------------------------------------------------------------------------------------------------------
class Ours:
    name = 'ours'

    def __init__(self, cfg, arrays, pos, ang, device, raw):
        import torch
        from megastep_b200 import core as core_, cuda, modules, scene
        self.cuda = cuda
        for kv in filter(None, os.environ.get('MSB_OPTIONS', '').split(',')):   # A/B switches, e.g. MSB_OPTIONS=pdl=0,no_vis=1
            name, value = kv.split('=')
            cuda.set_option(name, int(value))
        s = scene.upload(arrays, device)
        cuda.bake(s, params=cuda.make_params(AGENT_RADIUS, cfg['res'], cfg['fov'], FPS))
        self.core = core_.Core(s, res=cfg['res'], fov=cfg['fov'], fps=FPS)
        self.core.agents.positions.copy_(torch.as_tensor(pos))
        self.core.agents.angles.copy_(torch.as_tensor(ang))
        self.stepper = modules.FusedStep(self.core, subsample=cfg['subsample'], raw=raw)
        self.actions = self.stepper.actions
        self.graphed = False
        self.fused = False                   # physics and render are separate launches: each stages the segments

    def use_graph(self, native=False):
        """The public API's CUDA-graph modes. native: FusedStep.enable_host_graph() — the tick captured inside the
        library and driven by ONE foreign call per tick (actions up, graph launch, progress down, stream sync);
        otherwise (the default: measured faster, 156 vs 165 us per tick) FusedStep(graph=True): a torch.cuda.CUDAGraph
        replay between PyTorch copies."""
        self.host_graph = False
        if native:
            try:
                self.stepper.enable_host_graph()
                self.host_graph = True
            except RuntimeError as e:
                print(f'native host graph unavailable ({e}); falling back to the torch graph', file=sys.stderr)
        if not self.host_graph:
            self.stepper._capture()
        self.graphed = True

    def step_host(self, actions_host):
        return self.stepper.step_host(actions_host).progress

    def step(self):
        self.stepper()                       # movement+physics | render+heads | agent-hit lighting: 3 launches, no host sync
                                             # (replayed as one CUDA graph once use_graph() has captured them)

    def plain_step(self):
        self.stepper._plan()                 # the same three launches, never through the graph

    def result(self):
        return self.stepper._plan.progress

    def obs(self):
        p = self.stepper._plan
        return {'rgb': p.rgb, 'd': p.depth, 'imu': p.imu}

    def launches(self):
        return self.cuda.launch_count()


class Reference:
    """The reference's own CUDA build (oracle/_ref/megastepcuda*.so) driven by the reference's own, unmodified Python
    (baseline/_ref/megastep: core.Core, modules.MomentumMovement :106-118, modules.render :126-136, Depth :170-184,
    RGB :211-224, IMU :263-270) — through tests/common.py::reference_package, the same shim the parity tests use.
    Nothing of this repository's library is on that path (it is not even loaded into the process)."""
    name = 'reference'

    def __init__(self, cfg, arrays, pos, ang, device, raw):
        import torch
        import common
        self.torch = torch
        pkg = common.reference_package()
        if pkg is None:
            raise RuntimeError('oracle/_ref is not built')
        self.pkg = pkg
        scenery = common.reference_scenery(pkg.cuda, arrays, device)
        self.core = pkg.core.Core(scenery, res=cfg['res'], fov=cfg['fov'], fps=FPS)     # -> cuda.initialize (core.py:86)
        pkg.cuda.bake(scenery)
        self.core.agents.positions.copy_(torch.as_tensor(pos))
        self.core.agents.angles.copy_(torch.as_tensor(ang))
        m = pkg.modules
        self.mover, self.rgb, self.depth, self.imu = (m.MomentumMovement(self.core), m.RGB(self.core, subsample=cfg['subsample']),
                                                      m.Depth(self.core, subsample=cfg['subsample']), m.IMU(self.core))
        N, A = pos.shape[:2]
        self.actions = torch.zeros((N, A), dtype=torch.int32, device=device)
        self.decision = pkg.arrdict.arrdict(actions=self.actions)
        self.fused = False
        self._p = self._obs = None

    def step(self):
        self._p = self.mover(self.decision)
        r = self.pkg.modules.render(self.core)
        self._obs = {'rgb': self.rgb(r), 'd': self.depth(r), 'imu': self.imu()}

    def result(self):
        return self._p.progress

    def obs(self):
        return self._obs

    def launches(self):
        return 0


def run_gpu(args, cfg, arm_cls, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    n_envs = args.envs or cfg['n_envs']
    # Weak scaling = the SAME work on every GPU: each rank steps a replica of the same shard (same floorplans, poses and
    # actions), so that the max-over-ranks time measures the system, not which rank drew the costliest floorplans.
    seed_rank = rank if args.distinct_shards else 0
    gs, arrays, pos, ang = build_scene(cfg, n_envs, args.unique, seed_rank)
    raw = not args.no_raw
    arm = arm_cls(cfg, arrays, pos, ang, device, raw)
    N, A = pos.shape[:2]
    K, W = args.steps, args.warmup
    rng = np.random.RandomState(3 + seed_rank)
    acts_host = torch.as_tensor(rng.randint(0, 7, (K + W, N, A)).astype(np.int32)).pin_memory()
    acts_dev = acts_host.to(device)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=device)
    gather = None
    if args.gather and world > 1:
        from megastep_b200 import sharding
        arm.actions.copy_(acts_dev[0])
        arm.step()
        # ShardedCore's collective: rows packed per env, ONE all_gather_into_tensor on a side stream, double-buffered
        packing = sharding.PackedObs(A, cfg['res'] // cfg['subsample'], getattr(torch, args.obs_dtype))
        gather = sharding.RowGather(packing, N, device, transport=args.gather_transport)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, L2 flushed between steps, device-timed per step -------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    per_step_launches = None
    if hasattr(arm, 'plain_step'):
        arm.actions.copy_(acts_dev[0])
        l0 = arm.launches()
        arm.plain_step()
        per_step_launches = arm.launches() - l0          # kernels of ours in one step (a graph replay launches the same ones)
        if not args.no_graph:
            arm.use_graph(native=False)                  # the step's launches as one CUDA-graph replay: no launch gaps, less host work
    for i in range(W):
        arm.actions.copy_(acts_dev[i])
        arm.step()
        if gather is not None:
            gather.start(arm.obs())           # (the collective's first calls set up its channels / peer mappings)
    if gather is not None:
        gather.wait()
    barrier()
    if sampler:
        sampler.mark()
    barrier()
    launches0 = arm.launches()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    t0 = time.perf_counter()
    for i in range(K):
        arm.actions.copy_(acts_dev[W + i])
        flush.fill_(0.)                       # evict the previous step's lines/outputs from the 126 MB L2
        starts[i].record()
        arm.step()
        if gather is not None:
            # pack + start the all-gather of this step's observations; it runs on the side stream under the next step
            # (its buffers are reused two steps later: if the gathers cannot keep up, that wait lands in the step's time)
            gather.start(arm.obs())
        stops[i].record()
    if gather is not None:
        gather.wait()
    barrier()
    wall = time.perf_counter() - t0
    launches = arm.launches() - launches0
    if per_step_launches is not None and getattr(arm, 'graphed', False):
        launches = K * per_step_launches                 # replayed from the graph: the library's own counter does not see them
    # ---- per-kernel durations: a short extra loop with CUDA events around every kernel of the library (same L2
    # flush between steps); kept out of the loop above so that the events' own launch gaps do not touch `value`
    kernel_ms = None
    if hasattr(arm, 'cuda'):
        arm.cuda.set_option('timing', 1)
        for i in range(min(K, 100)):
            arm.actions.copy_(acts_dev[W + i])
            flush.fill_(0.)
            arm.plain_step()
        torch.cuda.synchronize()
        kernel_ms = {kind: arm.cuda.get_option(f'time_ns_{kind}') / 1e6 / max(arm.cuda.get_option(f'time_count_{kind}'), 1)
                     for kind in ('physics', 'render', 'dyn') if arm.cuda.get_option(f'time_count_{kind}') > 0}
        arm.cuda.set_option('timing', 0)
    per_step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, stops)])
    step_ms = float(per_step_ms.sum())

    # ---- e2e: host actions in pinned memory -> H2D -> step through the public API -> D2H of the step's result -----
    if hasattr(arm, 'use_graph') and not args.no_graph and args.e2e == 'native':
        arm.use_graph(native=True)
    result_host = torch.empty((N, A), dtype=torch.float32).pin_memory()
    one_launch = getattr(arm, 'host_graph', False)

    def tick(i):
        if one_launch:                                     # H2D + kernels + D2H in one graph launch, then a stream sync
            arm.step_host(acts_host[i])
        else:
            arm.actions.copy_(acts_host[i], non_blocking=True)
            arm.step()
            result_host.copy_(arm.result(), non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the host consumes each step's result before the next

    for i in range(W):
        tick(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        tick(W + i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None

    # max over ranks
    t = torch.tensor([step_ms, e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = t.tolist()
    total_bytes, mean_w = algorithmic_bytes(cfg, arrays, arm.fused, raw)
    return dict(arm=arm, gather_bytes=gather.bytes_received() if gather is not None else None, kernel_ms=kernel_ms, N=N, A=A, step_ms=step_ms, e2e_ms=e2e_ms, wall_s=wall, launches=launches, clocks=clocks,
                bytes_per_step=total_bytes, mean_walls=mean_w, arrays=arrays, pos=pos, ang=ang,
                per_step_ms=per_step_ms)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650., 'fallback (B200_PROFILING.md)'


def main():
    # NCCL's own log lines (rank / channel banner under NCCL_DEBUG=INFO) stay as the launcher set them; the JSON line is
    # the LAST line rank 0 prints, after the process group is gone
    args = parse()
    cfg = dict(WORKLOADS[args.workload])
    if args.res:
        cfg['res'] = args.res
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    n_envs = args.envs or cfg['n_envs']

    if args.dry_run:
        gs, arrays, pos, ang = build_scene(cfg, n_envs, args.unique, 0)
        b, w = algorithmic_bytes(cfg, arrays, True)
        print(json.dumps({'envs': n_envs, 'mean_walls': w, 'bytes_per_step': b, 'cpu_baseline': cpu_baseline(cfg, arrays, pos, ang, 3.)}))
        return

    reference = args.impl == 'reference'
    if reference and rank != 0:
        return      # the reference is single-GPU: rank 0 alone runs it
    if world > 1 and not reference:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=__import__('torch').device('cuda', local_rank))
    eff_world = 1 if reference else world

    arm_cls = Reference if reference else Ours
    kind = 'reference'
    try:
        out = run_gpu(args, cfg, arm_cls, rank if not reference else 0, eff_world, local_rank)
    except RuntimeError as e:
        if not reference or 'oracle/_ref' not in str(e):
            raise
        out, kind = None, 'port'

    if rank != 0:
        if world > 1 and not reference:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    base_cfg = {'workload': f"{args.workload}: synthetic cubicasa-shaped floorplans, {n_envs} envs/GPU x {cfg['n_agents']} agents x "
                            f"{cfg['res']} rays, fov {cfg['fov']:g}, MomentumMovement + physics + render + RGB/Depth/IMU (subsample {cfg['subsample']})",
                'envs_per_gpu': n_envs, 'n_agents': cfg['n_agents'], 'res': cfg['res'], 'fov': cfg['fov'], 'subsample': cfg['subsample'],
                'raw_render_outputs': not args.no_raw, 'parallelism': f'env-sharded x{eff_world} ({"each rank its own floorplans" if args.distinct_shards else "each rank a replica of the same shard: identical work per GPU"}), no per-step collective' + (f' + obs all-gather ({args.obs_dtype})' if args.gather else ''),
                'l2': f'flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write); per-step CUDA events'}

    if out is None:
        # reference arm without the reference build: the CPU oracle port on the host cores
        gs, arrays, pos, ang = build_scene(cfg, n_envs, args.unique, 0)
        cb = cpu_baseline(cfg, arrays, pos, ang)
        line = {'metric': 'agent-frames/sec', 'value': cb['value'], 'unit': 'agent-frames/s', 'n_gpus': 0, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg, 'impl': 'reference', 'cpu_baseline': cb,
                'e2e': {'value': cb['value'], 'unit': 'agent-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    K = args.steps
    frames = out['N'] * out['A'] * eff_world
    value = frames * K / (out['step_ms'] * 1e-3)
    e2e = frames * K / (out['e2e_ms'] * 1e-3)
    step_achieved = out['bytes_per_step'] / (out['step_ms'] / K * 1e-3) / 1e9
    km = out.get('kernel_ms')
    if km and 'render' in km:
        # the dominant kernel, timed live with CUDA events on its stream (inside the library)
        rb = render_kernel_bytes(cfg, out['arrays'], not args.no_raw)
        achieved = rb / (km['render'] * 1e-3) / 1e9
        traffic, traffic_src = profiled_traffic() if (args.workload == 'deathmatch' and not args.envs and not args.res and not args.no_raw) else (None, None)
        roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                    'kernel': 'view_kernel (draw + raycast + shade + Depth/RGB/IMU heads)', 'kernel_ms': km['render'],
                    'kernel_share_of_step': km['render'] / sum(km.values()), 'algorithmic_bytes_per_launch': rb,
                    'traffic_source': traffic_src, 'all_kernels_ms': km, 'peak_source': peak_src,
                    'whole_step': {'achieved': step_achieved, 'frac': step_achieved / peak, 'algorithmic_bytes_per_step': out['bytes_per_step']},
                    'note': 'not HBM-bound: issue/latency-bound (ncu, profiles/r02_kernels_full_summary.json: issue-active 61%, DRAM 14% of peak, 61.6 M warp-instructions). Kernels are timed apart here (events between them); in the step dyn_kernel overlaps the tail of view_kernel, so all_kernels_ms sums to more than ms_per_step. See DESIGN.md'}
    else:
        roofline = {'bound': 'hbm', 'achieved': step_achieved, 'peak': peak, 'unit': 'GB/s', 'frac': step_achieved / peak, 'traffic': None,
                    'kernel': 'whole step (physics + render + ~40 ATen launches)', 'algorithmic_bytes_per_step': out['bytes_per_step'],
                    'mean_walls_per_env': out['mean_walls'], 'peak_source': peak_src}
    line = {
        'metric': 'agent-frames/sec', 'value': value, 'unit': 'agent-frames/s', 'n_gpus': eff_world, 'steps': K, 'warmup': args.warmup,
        'ms_per_step': out['step_ms'] / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': base_cfg, 'impl': args.impl,
        'e2e': {'value': e2e, 'unit': 'agent-frames/s', 'h2d_bytes_per_step': out['N'] * out['A'] * 4,
                'd2h_bytes_per_step': out['N'] * out['A'] * 4, 'ms_per_step': out['e2e_ms'] / K,
                'what': ('pinned-host actions -> H2D -> the reference\'s own modules.MomentumMovement / render / RGB / Depth / IMU on its own CUDA build' if reference else
                         'pinned-host actions -> H2D -> step via the public API (modules.FusedStep' + ('.step_host: one foreign call per tick around a CUDA graph captured inside the library' if getattr(out.get('arm'), 'host_graph', False) else ', CUDA-graph replay' if getattr(out.get('arm'), 'graphed', False) else '') + ')')
                        + ' -> D2H of progress (the physics result) + stream sync, every step; observations (33 MB/step at the default workload) stay on the device in both arms, as the reference\'s policy networks consume them there'},
        'gpu_launches': out['launches'],
        'roofline': roofline,
        'clocks': out['clocks'],
        'step_ms_percentiles': {p: float(np.percentile(out['per_step_ms'], p)) for p in (5, 50, 95)},
    }
    if out.get('gather_bytes') is not None:
        line['gather'] = {'what': 'observations packed per env (one launch), ' + ('one NCCL all_gather_into_tensor' if args.gather_transport == 'nccl' else 'copy-engine peer copies out of symmetric memory between two barriers') + ' per step on a side stream (overlaps the next step), every rank receives the whole batch', 'transport': args.gather_transport, 'dtype': args.obs_dtype, 'bytes_received_per_rank_per_step': out['gather_bytes'],
                          'achieved_GBps_per_rank': out['gather_bytes'] / (out['step_ms'] / K * 1e-3) / 1e9}
    if reference:
        line['cpu_baseline'] = {'value': value, 'unit': 'agent-frames/s', 'cores': 0, 'kind': 'reference',
                                'sample': 'the reference\'s own CUDA build (oracle/_ref) on one B200 of this box: megastep has no CPU step path'}
    elif not args.no_cpu_baseline and world == 1:                # rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1)
        line['cpu_baseline'] = cpu_baseline(cfg, out['arrays'], out['pos'], out['ang'])
    if world > 1 and not reference:
        import torch.distributed as dist
        dist.destroy_process_group()
        time.sleep(1.)                      # let the other ranks' NCCL teardown lines (NCCL_DEBUG=INFO) drain first
    sys.stdout.flush()
    sys.stdout.write('\n' + json.dumps(line) + '\n')      # the LAST line of stdout, on a line of its own
    sys.stdout.flush()


if __name__ == '__main__':
    main()
