"""`Core`: the state of a batch of environments (reference: megastep/core.py:10-150).

The tensors hanging off the Core (`core.agents.*`, `core.scenery.*`) are the world state; `cuda.physics` /
`cuda.render` (usually via `modules`) advance and observe it.
"""
import numpy as np
import torch

from . import cuda
from .arrdict import clone
from .dotdict import dotdict

from .constants import AGENT_RADIUS, AGENT_WIDTH, TEXTURE_RES, gamma_decode, gamma_encode  # noqa: F401  (core.py:10-22)


def _init_agents(n_envs, n_agents, device='cuda'):
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)
    return cuda.Agents(angles=z(n_envs, n_agents), positions=z(n_envs, n_agents, 2),
                       angvelocity=z(n_envs, n_agents), velocity=z(n_envs, n_agents, 2))


class Core:

    def __init__(self, scenery, res=64, fov=130, fps=10):
        """:param scenery: a `cuda.Scenery` (see `scene.scenery`)
        :param res: rays (horizontal pixels) per agent. No 1024 cap (the reference's one-thread-per-ray limit).
        :param fov: field of view in degrees, < 180
        :param fps: steps per simulated second
        """
        self.n_envs = len(scenery.lines.widths)
        self.n_agents = scenery.n_agents
        self.res = res
        self.fov = fov
        self.agent_radius = AGENT_RADIUS
        self.fps = fps
        self.random = np.random.RandomState(1)
        self.device = scenery.model.device
        assert fov < 180, 'FOV should be less than 180°'

        # process-wide default, as the reference (core.py:86) ...
        cuda.initialize(self.agent_radius, self.res, self.fov, self.fps)
        # ... plus this Core's own copy, which `modules` passes explicitly so several Cores can coexist
        self.params = cuda.make_params(self.agent_radius, self.res, self.fov, self.fps)

        self.scenery = scenery
        self.agents = _init_agents(self.n_envs, self.n_agents, self.device)
        self.progress = torch.ones((self.n_envs, self.n_agents), device=self.device)

    def physics(self):
        return cuda.physics(self.scenery, self.agents, params=self.params)

    def render(self):
        return cuda.render(self.scenery, self.agents, params=self.params)

    def state(self, e):
        """A dotdict tree describing environment `e` (for plotting / debugging)."""
        options = {k: getattr(self, k) for k in ('n_envs', 'n_agents', 'res', 'fov', 'agent_radius', 'fps')}
        return clone(dotdict(**options, scenery=self.scenery.state(e), agents=self.agents.state(e),
                             progress=self.progress[e]))

    def env_full(self, x):
        """(n_envs,) tensor full of `x` on the Core's device."""
        return torch.full((self.n_envs,), x, device=self.device, dtype=_DTYPES[type(x)])

    def agent_full(self, x):
        """(n_envs, n_agents) tensor full of `x` on the Core's device."""
        return torch.full((self.n_envs, self.n_agents), x, device=self.device, dtype=_DTYPES[type(x)])


_DTYPES = {bool: torch.bool, int: torch.int32, float: torch.float32}
