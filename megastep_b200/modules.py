"""Chunks of functionality that environments are assembled from (reference: megastep/modules.py:10-381).

Movement modules set the agents' velocities and call `cuda.physics`; `render()` calls `cuda.render` and reshapes
its outputs for conv nets; `Depth` / `RGB` / `IMU` turn state into observations. `RGBD` and `FusedStep` are
additions: the same results with the surrounding elementwise work folded into the kernels.
"""
import numpy as np
import torch

from . import cuda, geometry, spaces
from .arrdict import arrdict, stack, torchify


def to_local_frame(angles, p):
    """Rotate global-frame vectors `p` into the frames of agents facing `angles` (degrees)."""
    a = np.pi / 180 * angles
    c, s = torch.cos(a), torch.sin(a)
    x, y = p[..., 0], p[..., 1]
    return torch.stack([c * x + s * y, -s * x + c * y], -1)


def to_global_frame(angles, p):
    """Rotate agent-local vectors `p` into the global frame."""
    a = np.pi / 180 * angles
    c, s = torch.cos(a), torch.sin(a)
    x, y = p[..., 0], p[..., 1]
    return torch.stack([c * x - s * y, s * x + c * y], -1)


# noop, forward/backward, strafe left/right, turn left/right (modules.py:45-47, 95-96)
_VELOCITY = [[0., 0.], [0., 1.], [0., -1.], [1., 0.], [-1., 0.], [0., 0.], [0., 0.]]
_ANGVELOCITY = [0., 0., 0., 0., 0., +1., -1.]


def _actionset(lin, ang, device):
    return arrdict(velocity=lin * torch.tensor(_VELOCITY), angvelocity=ang * torch.tensor(_ANGVELOCITY)).to(device)


def _physics(core):
    return cuda.physics(core.scenery, core.agents, params=getattr(core, 'params', None))


class SimpleMovement:

    def __init__(self, core, speed=10, ang_speed=180, n_agents=None):
        """Momentum-free movement: each action is a fixed displacement / turn per step."""
        self.core = core
        self._actionset = _actionset(speed / core.fps, ang_speed / core.fps, core.device)
        self.space = spaces.MultiDiscrete(n_agents or core.n_agents, 7)

    def __call__(self, decision):
        core = self.core
        delta = self._actionset[decision.actions.long()]
        core.agents.angvelocity[:] = delta.angvelocity
        core.agents.velocity[:] = to_global_frame(core.agents.angles, delta.velocity)
        return _physics(core)


class MomentumMovement:

    def __init__(self, core, accel=5, ang_accel=180, decay=.125, n_agents=None):
        """Movement with momentum: actions accelerate the agent; velocity decays by `decay` per step."""
        self.core = core
        self.accel, self.ang_accel, self.decay = accel, ang_accel, decay
        self._actionset = _actionset(accel / core.fps, ang_accel / core.fps, core.device)
        self.space = spaces.MultiDiscrete(n_agents or core.n_agents, 7)

    def __call__(self, decision):
        core = self.core
        delta = self._actionset[decision.actions.long()]
        core.agents.angvelocity[:] = (1 - self.decay) * core.agents.angvelocity + delta.angvelocity
        core.agents.velocity[:] = (1 - self.decay) * core.agents.velocity + to_global_frame(core.agents.angles, delta.velocity)
        return _physics(core)


class FusedMovement:
    """`MomentumMovement` as one launch (msb_move): the decay, the action impulses and physics() in the same kernel,
    bit-identical to `MomentumMovement(core)(decision)` (and to the reference's, tests/test_gpu_reference_python.py)."""

    def __init__(self, core, accel=5, ang_accel=180, decay=.125, n_agents=None):
        self.core = core
        self.actions = torch.zeros((core.n_envs, core.n_agents), dtype=torch.int32, device=core.device)
        self._plan = cuda.StepPlan(core.scenery, core.agents, core.params, actions=self.actions, accel=accel, ang_accel=ang_accel,
                                   decay=decay, raw=False, subsample=None)
        self.space = spaces.MultiDiscrete(n_agents or core.n_agents, 7)

    def __call__(self, decision):
        self.actions.copy_(decision.actions.reshape(self.actions.shape), non_blocking=True)
        self._plan.move_only()
        return arrdict(progress=self._plan.progress)


def unpack(d):
    """`cuda` result objects -> arrdicts with the same attributes."""
    if isinstance(d, torch.Tensor):
        return d
    return arrdict({k: unpack(getattr(d, k)) for k in dir(d) if not k.startswith('_')})


def render(core):
    """`cuda.render`, as an arrdict, with a height dim added and `screen` permuted to (N, A, C, H=1, W)."""
    r = unpack(cuda.render(core.scenery, core.agents, params=getattr(core, 'params', None)))
    r = arrdict({k: v.unsqueeze(2) for k, v in r.items()})
    r['screen'] = r.screen.permute(0, 1, 4, 2, 3)
    return r


def downsample(screen, subsample):
    """(..., W) -> (..., W/subsample, subsample); follow with an aggregation over the last dim."""
    return screen.view(*screen.shape[:-1], screen.shape[-1] // subsample, subsample)


class Depth:

    def __init__(self, core, n_agents=None, subsample=1, max_depth=10):
        """Depth observations in [0, 1]: 1 at the near plane (the agent radius), 0 at `max_depth` metres or beyond."""
        self.core = core
        self.space = spaces.MultiImage(n_agents or core.n_agents, 1, 1, core.res // subsample)
        self.max_depth = max_depth
        self.subsample = subsample

    def __call__(self, r=None):
        r = render(self.core) if r is None else r
        depth = 1 - ((r.distances - self.core.agent_radius) / self.max_depth).clamp(0, 1)
        self._last_obs = downsample(depth, self.subsample).mean(-1).unsqueeze(3)
        return self._last_obs

    def state(self, e=0):
        return self._last_obs[e].clone()


class RGB:

    def __init__(self, core, n_agents=None, subsample=1):
        """Linear-RGB observations, (N, A, 3, 1, res/subsample); gamma-encode before displaying."""
        self.core = core
        self.space = spaces.MultiImage(n_agents or core.n_agents, 3, 1, core.res // subsample)
        self.subsample = subsample

    def __call__(self, r=None):
        r = render(self.core) if r is None else r
        self._last_obs = downsample(r.screen, self.subsample).mean(-1)
        return self._last_obs

    def state(self, e=0):
        return self._last_obs[e].clone()


class IMU:

    def __init__(self, core, speed_scale=10., ang_scale=360., n_agents=None):
        """(angular velocity, forward velocity, lateral velocity), each scaled, as an (N, A, 3) tensor."""
        self.core = core
        self.space = spaces.MultiVector(n_agents or core.n_agents, 3)
        self.speed_scale = speed_scale
        self.ang_scale = ang_scale

    def __call__(self):
        agents = self.core.agents
        return torch.cat([agents.angvelocity[..., None] / self.ang_scale,
                          to_local_frame(agents.angles, agents.velocity) / self.speed_scale], -1)


class RGBD:
    """RGB + Depth (+ IMU) observations produced by the render kernel itself.

    The reference composes these from `render()` with ~10 elementwise PyTorch launches over the full-resolution
    outputs (modules.py:170-184, 211-224, as used at demo/envs/deathmatch.py:81-85); here the subsampled heads are
    written by the same launch that casts the rays, so the (N, A, res) intermediates need not be materialised.

    obs = RGBD(core, subsample=4)();  obs.rgb (N,A,3,1,res/sub), obs.d (N,A,1,1,res/sub), obs.imu (N,A,3)
    """

    def __init__(self, core, n_agents=None, subsample=1, max_depth=10, speed_scale=10., ang_scale=360., raw=False):
        n_agents = n_agents or core.n_agents
        self.core = core
        self.subsample, self.max_depth = subsample, max_depth
        self.space = arrdict(rgb=spaces.MultiImage(n_agents, 3, 1, core.res // subsample),
                             d=spaces.MultiImage(n_agents, 1, 1, core.res // subsample),
                             imu=spaces.MultiVector(n_agents, 3))
        self._plan = cuda.StepPlan(core.scenery, core.agents, core.params, actions=None, raw=raw, subsample=subsample,
                                   max_depth=max_depth, speed_scale=speed_scale, ang_scale=ang_scale)

    def __call__(self):
        p = self._plan
        p.render_only()
        self._last_obs = arrdict(rgb=p.rgb, d=p.depth, imu=p.imu)
        return self._last_obs

    @property
    def render(self):
        return self._plan.render

    def state(self, e=0):
        return self._last_obs[e].clone()


class FusedStep:
    """One whole tick — MomentumMovement, physics, render, RGB/Depth/IMU — as a single kernel launch.

        step = FusedStep(core, subsample=4)
        out = step(actions)      # actions: (N, A) int tensor
        out.obs.rgb, out.obs.d, out.obs.imu, out.progress, out.render (None unless raw=True)

    Results equal `MomentumMovement(core)(decision)` followed by `RGB/Depth/IMU` on `render(core)`; outputs live in
    persistent buffers owned by this object (overwritten by the next call).
    """

    def __init__(self, core, accel=5, ang_accel=180, decay=.125, subsample=1, max_depth=10, speed_scale=10.,
                 ang_scale=360., raw=False, graph=False):
        self.core = core
        self.actions = torch.zeros((core.n_envs, core.n_agents), dtype=torch.int32, device=core.device)
        self._plan = cuda.StepPlan(core.scenery, core.agents, core.params, actions=self.actions, accel=accel,
                                   ang_accel=ang_accel, decay=decay, raw=raw, subsample=subsample, max_depth=max_depth,
                                   speed_scale=speed_scale, ang_scale=ang_scale)
        self.space = spaces.MultiDiscrete(core.n_agents, 7)
        self._graph = None
        self.actions_host = self.progress_host = None
        if graph:
            self._capture()

    def _capture(self, host_io=False):
        """Capture the launch in a CUDA graph so a step costs one graph replay on the host. With `host_io` the graph
        also holds the copy of the actions from a pinned host buffer (`actions_host`) and of `progress` back into one
        (`progress_host`): a whole host-to-host tick is then ONE launch (see `step_host`)."""
        if host_io:
            self.actions_host = torch.zeros(tuple(self.actions.shape), dtype=torch.int32).pin_memory()
            self.progress_host = torch.zeros(tuple(self._plan.progress.shape), dtype=torch.float32).pin_memory()
        agents = (self.core.agents.angles, self.core.agents.positions, self.core.agents.angvelocity, self.core.agents.velocity)
        snapshot = [t.clone() for t in agents]                 # the warm-up below advances the state once: undone at the end
        side = torch.cuda.Stream(device=self.core.device)
        side.wait_stream(torch.cuda.current_stream(self.core.device))
        with torch.cuda.stream(side):
            self._plan()           # warm-up outside capture (first-launch attribute set-up)
        torch.cuda.current_stream(self.core.device).wait_stream(side)
        torch.cuda.synchronize(self.core.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            if host_io:
                self.actions.copy_(self.actions_host, non_blocking=True)
            self._plan()
            if host_io:
                self.progress_host.copy_(self._plan.progress, non_blocking=True)
        # capture does not execute, but the warm-up advanced the state once: restore it
        for dst, src in zip(agents, snapshot):
            dst.copy_(src)

    def __call__(self, actions=None):
        if actions is not None:
            self.actions.copy_(actions if not hasattr(actions, 'actions') else actions.actions, non_blocking=True)
        if self._graph is not None:
            self._graph.replay()
        else:
            self._plan()
        p = self._plan
        return arrdict(obs=arrdict(rgb=p.rgb, d=p.depth, imu=p.imu), progress=p.progress, render=p.render)

    def enable_host_graph(self):
        """The tick as a CUDA graph captured inside the library, driven by `step_host` with ONE foreign call per tick
        (upload of the actions, graph launch, download of `progress`, stream sync: msb_step_graph_run) instead of four
        PyTorch calls."""
        agents = (self.core.agents.angles, self.core.agents.positions, self.core.agents.angvelocity, self.core.agents.velocity)
        snapshot = [t.clone() for t in agents]                 # the warm-up tick is undone below
        self._plan()                                           # first-launch set-up outside the capture
        torch.cuda.synchronize(self.core.device)
        self._plan.graph_create()
        for dst, src in zip(agents, snapshot):
            dst.copy_(src)
        self.actions_host = torch.zeros(tuple(self.actions.shape), dtype=torch.int32).pin_memory()
        self.progress_host = torch.zeros(tuple(self._plan.progress.shape), dtype=torch.float32).pin_memory()
        self._native_graph = True
        self._pinned_storages = {}
        self._progress_host_ptr = self.progress_host.data_ptr()
        p = self._plan
        self._host_result = arrdict(obs=arrdict(rgb=p.rgb, d=p.depth, imu=p.imu), progress=self.progress_host, render=p.render)

    def _is_pinned(self, t):
        """`t.is_pinned()` asks the driver every time (microseconds); the answer per storage does not change."""
        key = t.untyped_storage().data_ptr()
        known = self._pinned_storages.get(key)
        if known is None:
            known = self._pinned_storages[key] = t.is_pinned()
        return known

    def step_host(self, actions=None):
        """One tick driven from the host (needs `enable_host_graph()` or `_capture(host_io=True)`): `actions` — a host
        tensor / array, or None if the caller filled `self.actions_host` itself — go up, the tick runs, `progress`
        comes back into `self.progress_host`, then the stream is synchronised. Observations stay on the device, as
        in `__call__`."""
        p = self._plan
        if getattr(self, '_native_graph', False):
            src = self.actions_host
            if actions is not None and actions is not self.actions_host:
                a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(actions)
                if (a.dtype == torch.int32 and not a.is_cuda and a.shape == self.actions.shape and a.is_contiguous()
                        and self._is_pinned(a)):
                    src = a                                     # straight from the caller's pinned buffer
                else:
                    self.actions_host.copy_(a)
            p.graph_run(src.data_ptr(), self._progress_host_ptr, sync=True)
            return self._host_result
        if actions is not None and actions is not self.actions_host:
            self.actions_host.copy_(torch.as_tensor(actions))
        self._graph.replay()
        torch.cuda.current_stream(self.core.device).synchronize()
        return arrdict(obs=arrdict(rgb=p.rgb, d=p.depth, imu=p.imu), progress=self.progress_host, render=p.render)


def random_empty_positions(geometries, n_agents, n_points, random=np.random):
    """(n_geometries, n_agents, n_points, 2) random free-space points, in metres (modules.py:272-293)."""
    points = []
    for g in geometries:
        if 'masks' not in g:
            # a geometry without the rasterised free-space grid (synthetic.sample(n) builds it only on request): draw
            # the points inside its room rectangles instead
            from . import synthetic
            pts = [synthetic.spawns([g], n_agents, random)[0][0] for _ in range(n_points)]
            points.append(np.stack(pts, 1).astype(float))
            continue
        free = np.stack((g.masks > 0).nonzero(), -1)
        n_possible = min(len(free) // n_agents, n_points)
        sample = free[random.choice(np.arange(len(free)), (n_possible, n_agents), replace=True)]
        sample = np.concatenate([sample] * int(n_points / len(sample) + 1))[-n_points:]
        sample = random.permutation(sample)
        points.append(geometry.centers(sample, g.masks.shape, g.res).transpose(1, 0, 2))
    return stack(points)


class RandomSpawns:

    def __init__(self, geometries, core, n_spawns=100, fused=False, seed=0):
        """Respawns agents at random pre-computed free positions (modules.py:295-326). `fused`: the draw and the
        update as one kernel (cuda.env_respawn; a counter-based hash of (seed, call number, agent) picks the spawn)."""
        self.core = core
        self.fused, self.seed, self._calls = fused, seed, 0
        positions = random_empty_positions(geometries, core.n_agents, n_spawns)
        angles = core.random.uniform(-180, +180, (len(geometries), core.n_agents, n_spawns))
        self._spawns = torchify(arrdict(positions=positions, angles=angles)).to(core.device)

    def __call__(self, reset):
        """`reset`: (n_envs, n_agents) bool mask of agents to respawn; their velocities are zeroed.

        Same effect as the reference (modules.py:312-326) without its `nonzero()` — a device-to-host sync in the middle
        of every step: a spawn is drawn for every agent and blended in under the mask, all on the device."""
        core = self.core
        if self.fused:
            if not hasattr(self, '_flat'):
                self._flat = (self._spawns.positions.float().contiguous(), self._spawns.angles.float().contiguous())
            self._calls += 1
            cuda.env_respawn(core.scenery, core.agents, reset.reshape(core.n_envs, core.n_agents), self._flat[0], self._flat[1], self.seed, self._calls)
            return
        choices = torch.randint(0, self._spawns.angles.shape[-1], reset.shape, device=reset.device)
        angles = self._spawns.angles.gather(-1, choices[..., None]).squeeze(-1)
        positions = self._spawns.positions.gather(2, choices[..., None, None].expand(-1, -1, 1, 2)).squeeze(2)
        core.agents.angles.copy_(torch.where(reset, angles, core.agents.angles))
        core.agents.positions.copy_(torch.where(reset[..., None], positions, core.agents.positions))
        core.agents.velocity.masked_fill_(reset[..., None], 0.)
        core.agents.angvelocity.masked_fill_(reset, 0.)


class RandomLifespans:

    def __init__(self, core, max_lifespan, min_lifespan=None):
        """Flags agents that outlive a random lifespan in [min_lifespan, max_lifespan) (modules.py:328-381)."""
        self.min_lifespan = max_lifespan // 2 if min_lifespan is None else min_lifespan
        self.max_lifespan = max_lifespan
        self._max_lifespans = torch.zeros((core.n_envs, core.n_agents), dtype=torch.int, device=core.device)
        self._lifespans = torch.zeros_like(self._max_lifespans)
        self._reset(core.agent_full(True))

    def _reset(self, reset):
        self._lifespans[reset] = 0
        self._max_lifespans[reset] = torch.randint_like(self._max_lifespans, self.min_lifespan, self.max_lifespan)[reset]

    def __call__(self, reset=None):
        self._lifespans += 1
        reset = torch.zeros_like(self._lifespans, dtype=torch.bool) if reset is None else reset
        reset = (self._lifespans >= self._max_lifespans) | reset
        self._reset(reset)
        return reset

    def state(self, e):
        return arrdict(lifespan=self._lifespans[e], max_lifespans=self._max_lifespans[e]).clone()
