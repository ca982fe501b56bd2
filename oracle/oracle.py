"""ctypes front-end of the CPU oracle (oracle/megastep_oracle.c) on numpy arrays.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
legs; never by anything under megastep_b200/.

A `scene` here is a plain dict of numpy arrays in the reference's ragged layout (megastep/src/common.h:185-214):
    n_agents, model (F,2,2), lines (sumL,2,2), line_widths (N,), lights (sumI,3), light_widths (N,),
    textures (sumT,3), tex_widths (sumL,), baked (sumT,)
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Config(ctypes.Structure):
    _fields_ = [('n_envs', ctypes.c_int32), ('n_agents', ctypes.c_int32), ('n_model', ctypes.c_int32),
                ('res', ctypes.c_int32), ('agent_radius', ctypes.c_float), ('half_screen', ctypes.c_float),
                ('fps', ctypes.c_float)]


def build():
    subprocess.run(['make', '-C', _HERE, '--no-print-directory'], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libmegastep_oracle.so')
        src = os.path.join(_HERE, 'megastep_oracle.c')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.mso_num_threads.restype = ctypes.c_int
        _LIB.mso_half_screen.restype = ctypes.c_float
        _LIB.mso_half_screen.argtypes = [ctypes.c_float]
    return _LIB


def half_screen(fov):
    """tanf(pi/180*fov/2.) exactly as the reference's initialize() computes it on the host (kernels.cu:22)."""
    return float(lib().mso_half_screen(fov))


def config(n_envs, n_agents, n_model, res, fov, fps, agent_radius):
    return Config(n_envs, n_agents, n_model, res, agent_radius, half_screen(fov), fps)


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t))


def starts(widths, dtype=np.int32):
    w = np.asarray(widths).astype(np.int64)
    return (np.cumsum(w) - w).astype(dtype)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def physics(scene, state, fps=10., agent_radius=.15 / 2 ** .5):
    """In-place on `state` (dict of angles (N,A), positions (N,A,2), angvelocity (N,A), velocity (N,A,2), all
    float32 C-contiguous). Returns progress (N,A)."""
    N, A = state['angles'].shape
    cfg = config(N, A, scene['model'].shape[0], 0, 90., fps, agent_radius)
    lines = _c(scene['lines'], np.float32)
    lw = _c(scene['line_widths'], np.int32)
    ls = starts(lw)
    for k in ('angles', 'positions', 'angvelocity', 'velocity'):
        assert state[k].dtype == np.float32 and state[k].flags.c_contiguous
    progress = np.empty((N, A), np.float32)
    lib().mso_physics(ctypes.byref(cfg), _p(lines), _p(lw, ctypes.c_int32), _p(ls, ctypes.c_int32),
                      _p(state['angles']), _p(state['positions']), _p(state['angvelocity']), _p(state['velocity']),
                      _p(progress))
    return progress


def bake(scene, fps=10., agent_radius=.15 / 2 ** .5):
    """Returns the baked (sumT,) light map for `scene` (does not modify it)."""
    N = len(scene['line_widths'])
    cfg = config(N, scene['n_agents'], scene['model'].shape[0], 0, 90., fps, agent_radius)
    lines = _c(scene['lines'], np.float32)
    lw = _c(scene['line_widths'], np.int32)
    ls = starts(lw)
    lights = _c(scene['lights'], np.float32)
    iw = _c(scene['light_widths'], np.int32)
    is_ = starts(iw)
    tw = _c(scene['tex_widths'], np.int32)
    ts = starts(tw, np.int64)
    baked = np.empty(int(tw.astype(np.int64).sum()), np.float32)
    lib().mso_bake(ctypes.byref(cfg), _p(lines), _p(lw, ctypes.c_int32), _p(ls, ctypes.c_int32),
                   _p(lights), _p(iw, ctypes.c_int32), _p(is_, ctypes.c_int32),
                   _p(tw, ctypes.c_int32), _p(ts, ctypes.c_int64), _p(baked))
    return baked


def render(scene, state, res, fov, fps=10., agent_radius=.15 / 2 ** .5):
    """Returns dict(indices, locations, dots, distances, screen, lines) — `lines` is the scene's line array after the
    agents' models have been drawn into it (the input scene is left untouched)."""
    N, A = state['angles'].shape
    cfg = config(N, A, scene['model'].shape[0], res, fov, fps, agent_radius)
    lines = np.array(scene['lines'], dtype=np.float32, order='C', copy=True)
    lw = _c(scene['line_widths'], np.int32)
    ls = starts(lw)
    lights = _c(scene['lights'], np.float32)
    iw = _c(scene['light_widths'], np.int32)
    is_ = starts(iw)
    tex = _c(scene['textures'], np.float32)
    tw = _c(scene['tex_widths'], np.int32)
    ts = starts(tw, np.int64)
    baked = _c(scene['baked'], np.float32)
    model = _c(scene['model'], np.float32)
    ang = _c(state['angles'], np.float32)
    pos = _c(state['positions'], np.float32)
    out = dict(indices=np.empty((N, A, res), np.int32), locations=np.empty((N, A, res), np.float32),
               dots=np.empty((N, A, res), np.float32), distances=np.empty((N, A, res), np.float32),
               screen=np.empty((N, A, res, 3), np.float32))
    lib().mso_render(ctypes.byref(cfg), _p(lines), _p(lw, ctypes.c_int32), _p(ls, ctypes.c_int32),
                     _p(lights), _p(iw, ctypes.c_int32), _p(is_, ctypes.c_int32),
                     _p(tex), _p(tw, ctypes.c_int32), _p(ts, ctypes.c_int64), _p(baked), _p(model), _p(ang), _p(pos),
                     _p(out['indices'], ctypes.c_int32), _p(out['locations']), _p(out['dots']), _p(out['distances']),
                     _p(out['screen']))
    out['lines'] = lines
    return out


def num_threads():
    return lib().mso_num_threads()


def set_num_threads(n):
    lib().mso_set_num_threads(int(n))
