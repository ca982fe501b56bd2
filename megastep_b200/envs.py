"""The reference's two demo environments on this core: Explorer (megastep/demo/envs/explorer.py) and Deathmatch
(megastep/demo/envs/deathmatch.py) — SURVEY.md §8(f)3. Same constructor arguments, `reset()` / `step(decision)`
protocol and result trees (`obs`, `reward`, `reset`); plotting is left out.

The game rules are plain tensor functions at module level (`texel_indices`, `ExplorationLedger`, `shots`) so that they
are testable without a GPU; the classes only wire them to `core.Core` and the observation modules. Two departures
from the reference, both about host round trips inside a step: `RandomSpawns` is the sync-free one of this package,
and Explorer's texel bookkeeping sends the rays that hit nothing to a spare slot instead of indexing with their -1
(which in the reference marks the batch's last texel as seen).
"""
import numpy as np
import torch

from . import core as core_, modules, scene, spaces
from .arrdict import arrdict, torchify
from .dotdict import dotdict

CLEARANCE = 1.          # metres outside the floorplan before an agent starts losing health (deathmatch.py:8)


# ----------------------------------------------------------------------------------------------------------------------
# Explorer: reward = newly seen texels (explorer.py:34-58)
# ----------------------------------------------------------------------------------------------------------------------
def texel_indices(line_starts, tex_starts, tex_widths, indices, locations):
    """Global texel index under every ray, -1 where the ray hit nothing (explorer.py:34-43).

    indices (N, A, R) int: line hit within the env, -1 for none; locations (N, A, R): position along it in [0, 1].
    The texel is floor(width * location), clamped to the line's last texel.
    """
    # evaluated for every ray and masked afterwards: selecting the hits first (as the reference does) makes a
    # data-dependent shape, i.e. a device-to-host sync in every step
    hit = indices >= 0
    line = line_starts.long()[:, None, None] + indices.clamp(min=0).long()
    width = tex_widths[line].float()
    offset = torch.minimum(torch.floor(width * locations.nan_to_num(0.)), width - 1).long()
    return torch.where(hit, tex_starts.long()[line] + offset, torch.full_like(line, -1))


class ExplorationLedger:
    """Which texels each env has seen so far, and the reward that follows (explorer.py:29-31, 45-58, 73-77)."""

    def __init__(self, texel_env, n_envs, rays_per_obs):
        self.texel_env = texel_env.long()                       # (T,) env of every texel
        self._marks = torch.zeros(len(self.texel_env) + 1, dtype=torch.bool, device=texel_env.device)
        self.seen = self._marks[:-1]                            # (the extra slot takes the rays that hit nothing)
        self.potential = torch.zeros(n_envs, dtype=torch.float32, device=texel_env.device)
        self.rays_per_obs = rays_per_obs

    def reward(self, texels, reset):
        """texels: any-shape long tensor of texel indices (-1 ignored); reset (N,) bool envs whose reward is void."""
        # (no `texels[texels >= 0]`: a data-dependent shape is a device-to-host sync on the GPU. The reference indexes
        # with the -1s as they are, which marks the batch's LAST texel as seen whenever any ray misses.)
        self._marks[torch.where(texels >= 0, texels, len(self.texel_env))] = True
        potential = torch.zeros_like(self.potential).scatter_add_(0, self.texel_env, self.seen.float())
        gained = (potential - self.potential) / self.rays_per_obs
        self.potential = potential
        gained[reset] = 0.
        return gained

    def forget(self, reset):
        """reset (N,) bool: those envs start exploring afresh."""
        self.seen[reset[self.texel_env]] = False
        self.potential[reset] = 0.


class BitLedger:
    """`ExplorationLedger` on the device: one bit per texel of the whole scene and a running count per env, kept by two
    small kernels (`cuda.env_ledger_mark` / `env_ledger_clear`) — a step touches the texels its rays landed on instead
    of summing a flag for every texel of every env (the reference's `scatter_add_`, explorer.py:49-50: 28 M elements at
    4096 envs). Same rewards, potentials and `seen` set as `ExplorationLedger`."""

    def __init__(self, scenery, n_envs, rays_per_obs):
        from . import cuda
        dev = scenery.model.device
        self._cuda, self.scenery = cuda, scenery
        self._bits = torch.zeros(cuda.ledger_words(scenery), dtype=torch.int32, device=dev)
        self._count = torch.zeros(n_envs, dtype=torch.int32, device=dev)
        self._gained = torch.zeros(n_envs, dtype=torch.int32, device=dev)
        self.rays_per_obs = rays_per_obs

    @property
    def potential(self):
        return self._count.float()

    @property
    def seen(self):
        """(T,) bool, as ExplorationLedger.seen (unpacked on demand: plotting / tests)."""
        shifts = torch.arange(32, device=self._bits.device, dtype=torch.int32)
        bits = ((self._bits[:, None] >> shifts[None, :]) & 1).bool().reshape(-1)
        return bits[:self.scenery.textures.vals.size(0)]

    def reward(self, indices, locations, reset):
        """indices / locations: (N, A, R) outputs of render(); reset (N,) bool envs whose reward is void."""
        self._gained.zero_()
        self._cuda.env_ledger_mark(self.scenery, indices, locations, self._bits, self._count, self._gained)
        gained = self._gained.float() / self.rays_per_obs
        return torch.where(reset, torch.zeros_like(gained), gained)

    def forget(self, reset):
        self._cuda.env_ledger_clear(self.scenery, reset, self._bits, self._count)


class Explorer:
    """One agent per env, rewarded for every texel of wall it sees for the first time (explorer.py:8-107).

    `fused` (default on a CUDA device): movement + physics in one launch (`modules.FusedMovement`), render + RGB / Depth /
    IMU heads in one pass (`modules.RGBD`), the reward ledger as bits kept by two small kernels (`BitLedger`), respawns
    drawn on the device — about a dozen launches per step and no host round trip, where the reference's env runs ~70
    PyTorch ops, a `nonzero()` and a scatter over every texel of the scene. `fused=False` is the op-for-op restatement."""

    def __init__(self, geometries, *args, fused=None, **kwargs):
        """`geometries`: a list of geometries (`cubicasa.sample(n)` in the reference; here e.g.
        `synthetic.sample(n, with_masks=True)` — without the masks, spawn points are drawn inside the room rectangles)."""
        s = scene.scenery(geometries, 1)
        self.core = core_.Core(s, *args, res=4 * 64, fov=130, **kwargs)
        self.fused = (self.core.device.type == 'cuda') if fused is None else fused
        self._rgb = modules.RGB(self.core, n_agents=1, subsample=4)
        self._depth = modules.Depth(self.core, n_agents=1, subsample=4)
        self._mover = modules.FusedMovement(self.core) if self.fused else modules.MomentumMovement(self.core)
        self._imu = modules.IMU(self.core)
        self._respawner = modules.RandomSpawns(geometries, self.core, fused=self.fused)
        self.action_space = self._mover.space
        self.obs_space = dotdict(rgb=self._rgb.space, d=self._depth.space, imu=self._imu.space)
        sc = self.core.scenery
        rays_per_obs = self.core.res // self._rgb.subsample
        if self.fused:
            self._rgbd = modules.RGBD(self.core, subsample=4, raw=True)
            self._ledger = BitLedger(sc, self.core.n_envs, rays_per_obs)
        else:
            texel_env = sc.lines.inverse[sc.textures.inverse.long()]
            self._ledger = ExplorationLedger(texel_env, self.core.n_envs, rays_per_obs)
        self._lengths = torch.zeros(self.core.n_envs, device=self.core.device, dtype=torch.int)
        self.device = self.core.device

    def _observe(self, reset):
        if self.fused:
            obs = self._rgbd()                                  # render + the three heads, one pass
            r = self._rgbd.render
            self._rgb._last_obs, self._depth._last_obs = obs.rgb, obs.d
            return arrdict(rgb=obs.rgb, d=obs.d, imu=obs.imu), self._ledger.reward(r.indices, r.locations, reset)
        r = modules.render(self.core)
        obs = arrdict(rgb=self._rgb(r), d=self._depth(r), imu=self._imu())
        sc = self.core.scenery
        texels = texel_indices(sc.lines.starts, sc.textures.starts, sc.textures.widths, r.indices.squeeze(2), r.locations.squeeze(2))
        return obs, self._ledger.reward(texels, reset)

    def _reset(self, reset):
        self._respawner(reset.unsqueeze(-1))
        self._ledger.forget(reset)
        self._lengths[reset] = 0

    @torch.no_grad()
    def reset(self):
        reset = self.core.env_full(True)
        self._reset(reset)
        obs, reward = self._observe(reset)
        return arrdict(obs=obs, reset=reset, reward=reward)

    @torch.no_grad()
    def step(self, decision):
        self._mover(decision)
        self._lengths += 1
        reset = self._lengths >= self._ledger.potential + 200           # an episode lasts 200 steps plus one per texel found
        self._reset(reset)
        obs, reward = self._observe(reset)
        return arrdict(obs=obs, reset=reset, reward=reward)

    def state(self, e=0):
        led = self._ledger
        sc = self.core.scenery
        texel_env = led.texel_env if hasattr(led, 'texel_env') else sc.lines.inverse[sc.textures.inverse.long()].long()
        return arrdict(core=self.core.state(e), rgb=self._rgb.state(e), d=self._depth.state(e),
                       potential=led.potential[e].clone(), seen=led.seen[texel_env == e].clone(),
                       length=self._lengths[e].clone(), max_length=led.potential[e].add(200).clone())


# ----------------------------------------------------------------------------------------------------------------------
# Deathmatch: whoever has an opponent in the middle of its view hits it (deathmatch.py:54-72)
# ----------------------------------------------------------------------------------------------------------------------
def shots(opponents, n_agents):
    """opponents (N, A, 1, R') long: which agent each (downsampled) pixel of each agent's view shows, -1 for walls /
    nothing. Agent i hits agent j when j shows in one of the two middle pixels of i's view. Returns matchings
    (N, A, A) bool [shooter, target], hits (N, A) and wounds (N, A) as floats (deathmatch.py:54-63)."""
    res = opponents.size(-1)
    middle = opponents[..., res // 2 - 1:res // 2 + 1]                                   # (N, A, 1, 2)
    agents = torch.arange(n_agents, device=opponents.device)
    matchings = (middle[:, :, None] == agents[None, None, :, None, None]).flatten(3).any(-1)
    return matchings, matchings.sum(2).float(), matchings.sum(1).float()


def seen_agents(indices, n_model, n_agents, subsample):
    """(N, A, 1, R / subsample) long: the agent whose model line the centre ray of every pooled pixel hit, else -1
    (deathmatch.py:75-80)."""
    lines = modules.downsample(indices, subsample)[..., subsample // 2]
    agent = lines // n_model
    return torch.where((lines >= 0) & (agent < n_agents), agent, torch.full_like(lines, -1)).long()


def _extent(g):
    """(height, width) of the geometry's mask grid in metres (deathmatch.py:44); from the walls when there is no grid."""
    if 'masks' in g:
        return np.asarray(g.masks.shape) * g.res
    from . import geometry
    return np.asarray(geometry.mask_shape(np.asarray(g.walls, dtype=float))) * g.res


class Deathmatch:
    """Several agents per env shooting at whoever is in their crosshairs (deathmatch.py:20-119). The env is flattened
    to `n_envs * n_agents` single-agent environments at the interface, as in the reference."""

    def __init__(self, geometries, n_agents, *args, fused=None, **kwargs):
        """`fused` (default on a CUDA device): as `Explorer` — movement + physics in one launch, render + heads in one pass,
        the crosshair rule with its health / damage updates as one kernel (`cuda.env_shoot`), respawns drawn on the
        device."""
        s = scene.scenery(geometries, n_agents)
        self.core = core_.Core(s, *args, res=4 * 128, fov=70, **kwargs)
        self.fused = (self.core.device.type == 'cuda') if fused is None else fused
        self._rgb = modules.RGB(self.core, n_agents=1, subsample=4)
        self._depth = modules.Depth(self.core, n_agents=1, subsample=4)
        self._imu = modules.IMU(self.core, n_agents=1)
        self._movement = modules.FusedMovement(self.core, n_agents=1) if self.fused else modules.MomentumMovement(self.core, n_agents=1)
        self._spawner = modules.RandomSpawns(geometries, self.core, fused=self.fused)
        if self.fused:
            self._rgbd = modules.RGBD(self.core, subsample=4, raw=True)
            N, A = self.core.n_envs, self.core.n_agents
            self._matchings = torch.zeros((N, A, A), dtype=torch.uint8, device=self.core.device)
            self._hits = torch.zeros((N, A), dtype=torch.float32, device=self.core.device)
        self.action_space = self._movement.space
        self.obs_space = dotdict(rgb=self._rgb.space, d=self._depth.space, imu=self._imu.space, health=spaces.MultiVector(1, 1))
        self._bounds = torchify(np.stack([_extent(g) for g in geometries])).to(self.core.device).contiguous()
        self._health = self.core.agent_full(np.nan)
        self._damage = self.core.agent_full(np.nan)
        self.n_envs = self.core.n_envs * self.core.n_agents
        self.device = self.core.device

    @staticmethod
    def _expand(tree):
        return tree.map(lambda x: x.reshape(x.shape[0] * x.shape[1], 1, *x.shape[2:]))

    def _collapse(self, tree):
        A = self.core.n_agents
        return tree.map(lambda x: x.reshape(x.shape[0] // A, A, *x.shape[2:]))

    def _reset(self, reset=None):
        reset = (self._health <= 0) if reset is None else reset
        self._spawner(reset)
        self._health[reset] = 1.
        self._damage[reset] = 0.
        return reset.reshape(-1)

    def _shoot(self, opponents):
        self.matchings, hits, wounds = shots(opponents, self.core.n_agents)
        self._damage += .05 * hits
        pos = self.core.agents.positions
        outside = (pos < -CLEARANCE).any(-1) | (pos > (self._bounds[:, None] + CLEARANCE)).any(-1)
        self._health += -.05 * (wounds + outside) - .001                  # 5 % per wound or excursion, 0.1 % per step
        return hits.reshape(-1)

    def _observe(self):
        if self.fused:
            from . import cuda
            obs = self._rgbd()
            cuda.env_shoot(self.core.scenery, self.core.agents, self._rgbd.render.indices, self._rgb.subsample, self._bounds, CLEARANCE,
                           self._matchings, self._hits, self._health, self._damage)
            self.matchings = self._matchings.bool()
            self._rgb._last_obs, self._depth._last_obs = obs.rgb, obs.d
            return arrdict(rgb=obs.rgb, d=obs.d, imu=obs.imu, health=self._health.unsqueeze(-1).clone()), self._hits.reshape(-1).clone()
        r = modules.render(self.core)
        opponents = seen_agents(r.indices, len(self.core.scenery.model), self.core.n_agents, self._rgb.subsample)
        hits = self._shoot(opponents)
        obs = arrdict(rgb=self._rgb(r), d=self._depth(r), imu=self._imu(), health=self._health.unsqueeze(-1).clone())
        return obs, hits

    @torch.no_grad()
    def reset(self):
        reset = self._reset(self.core.agent_full(True))
        obs, reward = self._observe()
        return arrdict(obs=self._expand(obs), reward=reward, reset=reset)

    @torch.no_grad()
    def step(self, decision):
        reset = self._reset()
        self._movement(self._collapse(decision))
        obs, reward = self._observe()
        return arrdict(obs=self._expand(obs), reward=reward, reset=reset)

    def state(self, e=0):
        return arrdict(core=self.core.state(e), rgb=self._rgb.state(e), d=self._depth.state(e),
                       health=self._health[e].clone(), damage=self._damage[e].clone(),
                       matchings=self.matchings[e].clone(), bounds=self._bounds[e].clone())
