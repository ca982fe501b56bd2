"""megastep_b200 — a B200-native simulation core behind megastep's Python API.

Drop-in for the per-step hot path of andyljones/megastep: `core.Core`, `ragged.Ragged`,
`modules.{MomentumMovement, RGB, Depth, RGBD, IMU, render, ...}` and the `cuda` extension surface
(`initialize, bake, physics, render, Agents, Scenery, Render, Physics, Ragged{1,2,3}D`), with the native work done
by hand-written sm_100a kernels in libmegastep_b200.so (see include/megastep_b200.h).

Importing the package does not need a GPU; `megastep_b200.cuda` needs the built library
(`python -m megastep_b200.build`) and its kernels need a CUDA device.
"""
import importlib

__version__ = '0.1.0'

_SUBMODULES = ('cuda', 'core', 'ragged', 'modules', 'spaces', 'scene', 'toys', 'geometry', 'synthetic', 'sharding',
               'arrdict', 'dotdict', 'build', 'envs', 'cubicasa', 'constants')


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module(f'{__name__}.{name}')
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
