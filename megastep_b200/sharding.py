"""Multi-GPU: environments shard embarrassingly.

Every kernel indexes by environment only (reference: kernels.cu:185,298,330,412), so a batch of N environments is
split into contiguous ranges, one per GPU, one process per GPU (torch.distributed). Each rank owns its slice of
every ragged array, its own agents and its own parameters; a step involves NO communication. The only collective
is optional: all-gathering the observation tensors over NCCL (NVLink 5 / NVSwitch) when the caller wants the whole
batch on every device, issued on a side stream so it overlaps the next tick.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_envs, rank, world_size):
    """Contiguous [lo, hi) of the envs owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_arrays(arrays, lo, hi):
    """The envs [lo, hi) of a `scene.scene_arrays` dict (numpy, host side)."""
    lw, iw, tw = arrays['line_widths'], arrays['light_widths'], arrays['tex_widths']
    ls = np.concatenate([[0], np.cumsum(lw)])
    is_ = np.concatenate([[0], np.cumsum(iw)])
    ts = np.concatenate([[0], np.cumsum(tw.astype(np.int64))])
    l0, l1 = ls[lo], ls[hi]
    out = dict(n_agents=arrays['n_agents'], model=arrays['model'],
               lines=arrays['lines'][l0:l1], line_widths=lw[lo:hi],
               lights=arrays['lights'][is_[lo]:is_[hi]], light_widths=iw[lo:hi],
               textures=arrays['textures'][ts[l0]:ts[l1]], tex_widths=tw[l0:l1])
    if 'baked' in arrays:
        out['baked'] = arrays['baked'][ts[l0]:ts[l1]]
    return out


def shard_scenery(scenery, lo, hi):
    """The envs [lo, hi) of a device-resident `cuda.Scenery`, as views (no copy) — reference Ragged slicing
    (common.h:136-144) applied at env granularity, textures/baked at the matching line range."""
    from . import cuda
    l0, l1 = int(scenery.lines.starts[lo].item()), int(scenery.lines.ends[hi - 1].item())
    s = cuda.Scenery(n_agents=scenery.n_agents, lights=scenery.lights[lo:hi], lines=scenery.lines[lo:hi],
                     textures=scenery.textures[l0:l1], model=scenery.model)
    s.baked.vals.copy_(scenery.baked[l0:l1].vals)
    return s


class ObsGather:
    """All-gathers equally-shaped per-rank observation tensors into one (world_size * n_local, ...) batch.

        g = ObsGather(example_local_obs)        # once; allocates the gathered buffers
        g.start(obs)                            # enqueue on a side stream (CUDA) — overlaps the next step
        full = g.wait()                         # arrdict of gathered tensors, valid on the current stream
    """

    def __init__(self, example, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.keys = list(example.keys())
        self.out = {k: v.new_empty((self.world * v.shape[0], *v.shape[1:])) for k, v in example.items()}
        self.cuda = next(iter(example.values())).is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self._work = []

    def start(self, obs):
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for k in self.keys:
                    obs[k].record_stream(self.stream)
                    dist.all_gather_into_tensor(self.out[k], obs[k].contiguous(), group=self.group)
        else:
            for k in self.keys:
                chunks = list(self.out[k].chunk(self.world, 0))
                self._work.append(dist.all_gather(chunks, obs[k].contiguous(), group=self.group, async_op=True))

    def wait(self):
        from .arrdict import arrdict
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        for w in self._work:
            w.wait()
        self._work = []
        return arrdict(self.out)


def gathered_bytes(example, world_size):
    """Bytes each rank receives per gather (cost model: / ~725 GB/s measured all-gather bus bandwidth)."""
    return sum(v.numel() * v.element_size() for v in example.values()) * (world_size - 1)


# ----------------------------------------------------------------------------------------------------------------------
# ShardedCore: one process per GPU, each stepping its own contiguous range of the batch's environments
# ----------------------------------------------------------------------------------------------------------------------
def initialize(device=None, devices=None, backend=None):
    """Single-host process group, as the reference sets it up for its learner (rebar/processes.py:18-29): rendezvous on
    127.0.0.1, port 29500 + the first device, rank = this process's position in `devices`. Under torchrun (RANK /
    WORLD_SIZE / MASTER_* in the environment) those are used instead. No-op if a group already exists."""
    import os
    if dist.is_initialized():
        return
    backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        dist.init_process_group(backend)
        return
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    if device is None:
        os.environ['MASTER_PORT'] = str(29500)
        dist.init_process_group(backend, rank=0, world_size=1)
    else:
        os.environ['MASTER_PORT'] = str(29500 + devices[0])
        dist.init_process_group(backend, rank=devices.index(device), world_size=len(devices))


class PackedObs:
    """The observation heads of one rank packed per environment — one row [rgb | d | imu] per env — so that ONE
    all-gather of the rows gives every rank the whole batch, and each observation of the whole batch is a plain strided
    view of the gathered rows (no unpacking pass). Optionally carried at reduced precision: `torch.float16`, or
    `torch.uint8` — rgb and depth (both in [0, 1]) quantised to 8 bits, what image observations are usually stored as
    (the reference's own replay-buffer estimate assumes 64 px x 3 channels, docs/faq.rst:77-79), the imu kept as fp16.

    views(rows) -> arrdict(rgb (N,A,3,1,ro), d (N,A,1,1,ro), imu (N,A,3))."""

    def __init__(self, n_agents, ro, dtype=torch.float32):
        self.A, self.ro, self.dtype = n_agents, ro, dtype
        A = n_agents
        if dtype == torch.uint8:
            n_img = A * 4 * ro
            pad = n_img % 2                                     # the fp16 imu starts on an even byte
            self.cols = {'rgb': (0, A * 3 * ro), 'd': (A * 3 * ro, n_img), 'imu': (n_img + pad, n_img + pad + 2 * A * 3)}
        else:
            self.cols = {'rgb': (0, A * 3 * ro), 'd': (A * 3 * ro, A * 4 * ro), 'imu': (A * 4 * ro, A * 4 * ro + A * 3)}
        self.width = self.cols['imu'][1]
        if dtype == torch.uint8:
            self.width += self.width % 2                        # rows stay 2-byte aligned

    def empty(self, n_envs, device):
        return torch.empty((n_envs, self.width), dtype=self.dtype, device=device)

    def pack(self, obs, out):
        """obs: arrdict(rgb, d, imu) of one rank (float32) -> out (n_local, width); three strided copies (with the cast)."""
        n = out.shape[0]
        if out.is_cuda and all(obs[k].dtype == torch.float32 and obs[k].is_contiguous() for k in ('rgb', 'd', 'imu')):
            from . import cuda                                    # one launch instead of three to seven
            mode = {torch.float32: 0, torch.float16: 1, torch.uint8: 2}[self.dtype]
            cuda.pack_obs(obs['rgb'], obs['d'], obs['imu'], out, mode, self.cols['imu'][0] if mode == 2 else 0)
            return out
        for k, (a, b) in self.cols.items():
            src = obs[k].reshape(n, -1)
            if self.dtype == torch.uint8:
                if k == 'imu':
                    out[:, a:b].view(torch.float16).copy_(src)
                else:
                    out[:, a:b].copy_((src * 255.).round_().clamp_(0., 255.))
            else:
                out[:, a:b].copy_(src)
        return out

    def views(self, rows):
        from .arrdict import arrdict
        A, ro = self.A, self.ro
        c = self.cols
        imu = rows[:, c['imu'][0]:c['imu'][1]]
        if self.dtype == torch.uint8:
            imu = imu.view(torch.float16)
        return arrdict(rgb=rows[:, c['rgb'][0]:c['rgb'][1]].unflatten(1, (A, 3, 1, ro)),
                       d=rows[:, c['d'][0]:c['d'][1]].unflatten(1, (A, 1, 1, ro)),
                       imu=imu.unflatten(1, (A, 3)))


class ShardedCore:
    """This rank's share of a batch of environments, behind one object.

        sharding.initialize(device, devices)                 # or torchrun
        sc = ShardedCore(arrays, n_envs_total, res=128, fov=70, subsample=1)
        out = sc.step(actions_local)                         # FusedStep on this rank's envs [sc.lo, sc.hi): no communication
        full = sc.gather()                                   # optional: every rank gets the whole batch's observations

    `arrays` is either the whole batch's `scene.scene_arrays` dict (each rank keeps its slice) or a callable
    `(lo, hi) -> arrays` that builds only this rank's envs. Envs are split into contiguous ranges (`shard_range`); each
    rank owns its scenery, agents and parameters. The only collective is `gather()`: the observation heads are packed
    one row per env (`PackedObs`) and all-gathered with ONE `all_gather_into_tensor` on a side stream, so that it
    overlaps the next `step()`; `gather_start()` / `gather_wait()` split it. `obs_dtype=torch.float16` / `torch.uint8`
    halve / quarter the bytes on the wire; `transport='p2p'` moves them with the copy engines (see `RowGather`).
    """

    def __init__(self, arrays, n_envs, n_agents=None, res=64, fov=130., fps=10., subsample=1, raw=False, obs_dtype=torch.float32,
                 group=None, device=None, graph=False, positions=None, angles=None, transport='nccl'):
        from . import core as core_, cuda, modules, scene
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_envs = n_envs
        self.lo, self.hi = shard_range(n_envs, self.rank, self.world)
        assert n_envs % self.world == 0, 'the all-gather needs equally sized shards: n_envs must be a multiple of the world size'
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        self.device = torch.device(device)
        local = arrays(self.lo, self.hi) if callable(arrays) else shard_arrays(arrays, self.lo, self.hi)
        s = scene.upload(local, self.device)
        if 'baked' in local:
            s.baked.vals.copy_(torch.as_tensor(local['baked']))
        else:
            cuda.bake(s, params=cuda.make_params(core_.AGENT_RADIUS, res, fov, fps))
        self.core = core_.Core(s, res=res, fov=fov, fps=fps)
        if positions is not None:
            self.core.agents.positions.copy_(torch.as_tensor(positions[self.lo:self.hi]))
        if angles is not None:
            self.core.agents.angles.copy_(torch.as_tensor(angles[self.lo:self.hi]))
        self.stepper = modules.FusedStep(self.core, subsample=subsample, raw=raw, graph=graph)
        self.packing = PackedObs(self.core.n_agents, res // subsample, obs_dtype)
        self.transport = transport
        self._gatherer = None
        self._out = None

    @property
    def n_local(self):
        return self.hi - self.lo

    def step(self, actions=None):
        """One tick of this rank's envs; `actions` (n_local, A) int. Returns FusedStep's arrdict (obs, progress, render)."""
        self._out = self.stepper(actions)
        return self._out

    def gather_start(self):
        """Pack the latest observations and start all-gathering them on the side stream."""
        if self._gatherer is None:
            self._gatherer = RowGather(self.packing, self.n_local, self.device, self.group, self.transport)
        self._gatherer.start(self._out.obs)

    def gather_wait(self):
        """The whole batch's observations (env-major, rank 0's envs first): arrdict(rgb, d, imu) of strided views."""
        return self._gatherer.wait()

    def gather(self):
        self.gather_start()
        return self.gather_wait()


class RowGather:
    """The all-gather of per-env packed rows, double-buffered so that the gather of tick t may still be in flight (side
    stream) while tick t+1 runs and packs into the other buffer. Two transports:

      'nccl'  one `all_gather_into_tensor` per step (NCCL over NVLink / NVSwitch): its kernel shares the SMs with the step;
      'p2p'   the rows live in symmetric memory (torch.distributed._symmetric_memory: every rank maps every peer's
              buffer) and each rank PULLS its peers' rows with plain device-to-device copies between two barriers — the
              copy engines move the bytes over NVLink, no SM is taken from the step that runs meanwhile (the pattern of
              torch's own `_low_contention_all_gather`).
    """

    def __init__(self, packing, n_local, device, group=None, transport='nccl'):
        self.packing, self.group, self.transport = packing, group, transport
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cuda = torch.device(device).type == 'cuda'
        if transport == 'p2p' and not (self.cuda and self.world > 1):
            self.transport = transport = 'nccl'
        self.handles = None
        if transport == 'p2p':
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            self.rows = [symm.empty((n_local, packing.width), dtype=packing.dtype, device=torch.device(device)) for _ in range(2)]
            self.handles = [symm.rendezvous(r, pg) for r in self.rows]
        else:
            self.rows = [packing.empty(n_local, device) for _ in range(2)]
        self.full = [packing.empty(n_local * self.world, device) for _ in range(2)]
        self.stream = torch.cuda.Stream(device=device, priority=-1) if self.cuda else None
        self.done = [None, None]
        self.tick = 0
        self._work = None

    def start(self, obs):
        i = self.tick & 1
        if self.cuda:
            main = torch.cuda.current_stream()
            if self.done[i] is not None:
                main.wait_event(self.done[i])                  # the gather that last read rows[i] / wrote full[i]
            self.packing.pack(obs, self.rows[i])
            self.stream.wait_stream(main)
            with torch.cuda.stream(self.stream):
                if self.world == 1:
                    self.full[i].copy_(self.rows[i])
                elif self.transport == 'p2p':
                    h, rows = self.handles[i], self.rows[i]
                    chunks = self.full[i].chunk(self.world)
                    h.barrier()                                # every rank has packed its rows
                    for step in range(self.world):             # own rows first, then the peers', each rank starting elsewhere
                        r = (self.rank - step) % self.world
                        chunks[r].copy_(h.get_buffer(r, rows.shape, rows.dtype))
                    h.barrier()                                # every rank has read them: they may be packed over
                else:
                    dist.all_gather_into_tensor(self.full[i], self.rows[i], group=self.group)
                self.done[i] = torch.cuda.Event()
                self.done[i].record(self.stream)
        else:
            self.packing.pack(obs, self.rows[i])
            if self.world > 1:
                self._work = dist.all_gather_into_tensor(self.full[i], self.rows[i], group=self.group, async_op=True) \
                    if dist.get_backend(self.group) != 'gloo' else \
                    dist.all_gather(list(self.full[i].chunk(self.world, 0)), self.rows[i], group=self.group, async_op=True)
            else:
                self.full[i].copy_(self.rows[i])
        self._pending = i
        self.tick += 1

    def wait(self):
        i = self._pending
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.done[i])
        elif self._work is not None:
            self._work.wait()
            self._work = None
        return self.packing.views(self.full[i])

    def bytes_received(self):
        return self.rows[0].numel() * self.rows[0].element_size() * (self.world - 1)
