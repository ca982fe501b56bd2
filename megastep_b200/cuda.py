"""The native interface: same names, arguments and error behaviour as the reference's `megastep.cuda` extension
module (megastep/src/wrappers.cpp:30-172), implemented over the C ABI of libmegastep_b200.so
(include/megastep_b200.h) instead of pybind11/ATen.

    initialize(agent_radius, res, fov, fps)      wrappers.cpp:53
    bake(scenery)                                wrappers.cpp:61
    physics(scenery, agents) -> Physics          wrappers.cpp:69
    render(scenery, agents) -> Render            wrappers.cpp:82
    Ragged1D / Ragged2D / Ragged3D               wrappers.cpp:14-28,99-101 (common.h:102-155)
    Agents, Scenery, Render, Physics             wrappers.cpp:103-172 (common.h:162-226)

There is no CPU path and no fallback: physics/render/bake on anything but CUDA tensors raise RuntimeError, as the
reference's TensorProxy does (common.h:12-14,33-37), and a missing library makes the import itself fail.
ctypes releases the GIL for the duration of each foreign call, like the reference's `gil_scoped_release` guards.

Beyond the reference: `step(...)` runs movement + physics + render + observation heads as one launch, the optional
`params=` keyword lets several Cores with different res/fov coexist (the reference keeps those in process globals,
wrappers.cpp:57-59), and texel offsets are 64-bit.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get('MEGASTEP_B200_LIB') or os.path.join(_HERE, 'libmegastep_b200.so')   # override: A/B builds

c_f32p = ctypes.c_void_p


class Params(ctypes.Structure):
    _fields_ = [('res', ctypes.c_int32), ('agent_radius', ctypes.c_float), ('half_screen', ctypes.c_float),
                ('fps', ctypes.c_float), ('fov', ctypes.c_float), ('reserved', ctypes.c_int32 * 3)]


class _Scenery(ctypes.Structure):
    _fields_ = [('n_envs', ctypes.c_int32), ('n_agents', ctypes.c_int32), ('n_model', ctypes.c_int32),
                ('max_lines', ctypes.c_int32), ('max_lights', ctypes.c_int32), ('occ_run', ctypes.c_int32),
                ('lines', ctypes.c_void_p), ('line_widths', ctypes.c_void_p), ('line_starts', ctypes.c_void_p),
                ('lights', ctypes.c_void_p), ('light_widths', ctypes.c_void_p), ('light_starts', ctypes.c_void_p),
                ('textures', ctypes.c_void_p), ('tex_widths', ctypes.c_void_p), ('tex_starts', ctypes.c_void_p),
                ('baked', ctypes.c_void_p), ('model', ctypes.c_void_p),
                ('n_lines', ctypes.c_int64), ('n_texels', ctypes.c_int64),
                ('occ_lines', ctypes.c_void_p), ('occ_starts', ctypes.c_void_p), ('occ_boxes', ctypes.c_void_p),
                ('box_starts', ctypes.c_void_p), ('occ_meta', ctypes.c_void_p), ('occ_rec', ctypes.c_void_p),
                ('vis', ctypes.c_void_p), ('vis_starts', ctypes.c_void_p), ('vis_meta', ctypes.c_void_p),
                ('env_order', ctypes.c_void_p)]


class _Agents(ctypes.Structure):
    _fields_ = [('angles', ctypes.c_void_p), ('positions', ctypes.c_void_p), ('angvelocity', ctypes.c_void_p),
                ('velocity', ctypes.c_void_p)]


class _RenderOut(ctypes.Structure):
    _fields_ = [('indices', ctypes.c_void_p), ('locations', ctypes.c_void_p), ('dots', ctypes.c_void_p),
                ('distances', ctypes.c_void_p), ('screen', ctypes.c_void_p)]


class _ObsOut(ctypes.Structure):
    _fields_ = [('rgb', ctypes.c_void_p), ('depth', ctypes.c_void_p), ('imu', ctypes.c_void_p),
                ('subsample', ctypes.c_int32), ('max_depth', ctypes.c_float), ('speed_scale', ctypes.c_float),
                ('ang_scale', ctypes.c_float)]


class _Workspace(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('bytes', ctypes.c_int64)]


class _Movement(ctypes.Structure):
    _fields_ = [('actions', ctypes.c_void_p), ('accel', ctypes.c_float), ('ang_accel', ctypes.c_float),
                ('decay', ctypes.c_float)]


def _load():
    if not os.path.exists(_LIBPATH):
        raise ImportError(
            f'{_LIBPATH} is missing. Build it with `python -m megastep_b200.build` (needs nvcc); '
            'there is no CPU or PyTorch fallback for the simulation kernels.')
    lib = ctypes.CDLL(_LIBPATH)
    P = ctypes.POINTER
    lib.msb_abi_version.restype = ctypes.c_int
    lib.msb_last_error.restype = ctypes.c_char_p
    lib.msb_launch_count.restype = ctypes.c_int64
    lib.msb_params_init.argtypes = [P(Params), ctypes.c_float, ctypes.c_int32, ctypes.c_float, ctypes.c_float]
    lib.msb_bake.argtypes = [P(Params), P(_Scenery), ctypes.c_void_p]
    lib.msb_build_visibility.argtypes = [P(_Scenery), ctypes.c_void_p]
    lib.msb_build_table.argtypes = [P(_Scenery), ctypes.c_void_p]
    lib.msb_physics.argtypes = [P(Params), P(_Scenery), P(_Agents), ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_move.argtypes = [P(Params), P(_Scenery), P(_Agents), P(_Movement), ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_render.argtypes = [P(Params), P(_Scenery), P(_Agents), P(_RenderOut), P(_ObsOut), P(_Workspace), ctypes.c_void_p]
    lib.msb_step.argtypes = [P(Params), P(_Scenery), P(_Agents), P(_Movement), ctypes.c_void_p, P(_RenderOut),
                             P(_ObsOut), P(_Workspace), ctypes.c_void_p]
    lib.msb_step_graph_create.argtypes = [P(Params), P(_Scenery), P(_Agents), P(_Movement), ctypes.c_void_p, P(_RenderOut),
                                          P(_ObsOut), P(_Workspace), P(ctypes.c_void_p)]
    lib.msb_step_graph_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.msb_step_graph_destroy.argtypes = [ctypes.c_void_p]
    lib.msb_workspace_bytes.argtypes = [P(Params), P(_Scenery), ctypes.c_int32]
    lib.msb_workspace_bytes.restype = ctypes.c_int64
    lib.msb_env_ledger_mark.argtypes = [P(_Scenery), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_env_ledger_clear.argtypes = [P(_Scenery), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_env_shoot.argtypes = [P(_Scenery), P(_Agents), ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_float,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_env_respawn.argtypes = [P(_Scenery), P(_Agents), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                    ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
    lib.msb_pack_obs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.msb_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    lib.msb_get_option.argtypes = [ctypes.c_char_p]
    lib.msb_get_option.restype = ctypes.c_int64
    if lib.msb_abi_version() != 5:
        raise ImportError(f'{_LIBPATH} has ABI version {lib.msb_abi_version()}, expected 5; rebuild it')
    return lib


_lib = _load()


def _check(code):
    if code != 0:
        raise RuntimeError(_lib.msb_last_error().decode())


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


# --------------------------------------------------------------------------------------------------------------
# Ragged (common.h:88-155)
# --------------------------------------------------------------------------------------------------------------
def _inverses(widths, total):
    # common.h:88-99: scatter ones at the row starts, cumsum, minus one
    starts = widths.cumsum(0) - widths.to(torch.long)
    flags = torch.ones(starts.size(0), dtype=torch.int32, device=widths.device)
    indices = torch.zeros(total, dtype=torch.int32, device=widths.device)
    return (indices.scatter(0, starts, flags).cumsum(0).to(torch.int32) - 1)


class _Ragged:
    _ndim = None

    def __init__(self, vals, widths):
        _require(isinstance(vals, torch.Tensor) and isinstance(widths, torch.Tensor), 'vals and widths must be tensors')
        _require(vals.is_contiguous(), 'vals must be contiguous')
        _require(widths.is_contiguous(), 'widths must be contiguous')
        _require(widths.dtype == torch.int32, 'widths must be an int32 tensor')
        _require(widths.dim() == 1, 'widths must be 1-dimensional')
        _require(vals.dtype == torch.float32, 'vals must be a float32 tensor')
        _require(vals.dim() == self._ndim, f'vals must be {self._ndim}-dimensional')
        total = int(widths.sum(0).item()) if widths.numel() else 0
        _require(total == vals.size(0), 'the widths must sum to the length of vals')
        with torch.no_grad():
            ends = widths.cumsum(0).to(torch.int32)
            self._vals, self._widths = vals, widths
            self._starts, self._ends = ends - widths, ends
        self._total = total
        self._inverse = None                                    # computed on first use: one int32 per value (GBs for the texels of a big batch)
        self._starts64 = None

    vals = property(lambda self: self._vals)
    widths = property(lambda self: self._widths)
    starts = property(lambda self: self._starts)
    ends = property(lambda self: self._ends)

    @property
    def inverse(self):
        if self._inverse is None:
            with torch.no_grad():
                self._inverse = _inverses(self._widths, self._total)
        return self._inverse

    def _long_starts(self):
        """64-bit exclusive prefix sum of the widths (the kernels' texel offsets); int32 `starts` wraps at 2^31."""
        if self._starts64 is None:
            w = self._widths.to(torch.long)
            self._starts64 = (w.cumsum(0) - w).contiguous()
        return self._starts64

    def __getitem__(self, x):
        if isinstance(x, slice):
            start, stop, step = x.indices(self._widths.size(0))
            _require(step == 1, 'ragged slices must have step 1')
            lo, hi = int(self._starts[start].item()), int(self._ends[stop - 1].item())
            return type(self)(self._vals[lo:hi], self._widths[start:stop])
        n = int(x)
        return self._vals[int(self._starts[n].item()):int(self._ends[n].item())]

    def clone(self):
        return type(self)(self._vals.clone(), self._widths.clone())

    def numpyify(self):
        from .ragged import RaggedNumpy
        from .arrdict import numpyify
        return RaggedNumpy(numpyify(self._vals), numpyify(self._widths))

    def __len__(self):
        return self._widths.size(0)


class Ragged1D(_Ragged):
    _ndim = 1


class Ragged2D(_Ragged):
    _ndim = 2


class Ragged3D(_Ragged):
    _ndim = 3


# --------------------------------------------------------------------------------------------------------------
# state containers (common.h:162-226)
# --------------------------------------------------------------------------------------------------------------
def _proxy(t, ndim, name, dtype=torch.float32):
    """TensorProxy's checks (common.h:33-37)."""
    _require(isinstance(t, torch.Tensor), f'{name} must be a tensor')
    _require(t.is_cuda, f'{name} must be a CUDA tensor')
    _require(t.is_contiguous(), f'{name} must be contiguous')
    _require(t.dtype == dtype, f'{name} must have dtype {dtype}')
    _require(t.dim() == ndim, f'{name} must be {ndim}-dimensional')
    return t


class Agents:
    """Holds the state of the agents; the four tensors stay caller-visible and are updated in place."""

    def __init__(self, angles, positions, angvelocity, velocity):
        self._angles = _proxy(angles, 2, 'angles')
        self._positions = _proxy(positions, 3, 'positions')
        self._angvelocity = _proxy(angvelocity, 2, 'angvelocity')
        self._velocity = _proxy(velocity, 3, 'velocity')
        n, a = angles.shape
        _require(tuple(positions.shape) == (n, a, 2) and tuple(velocity.shape) == (n, a, 2)
                 and tuple(angvelocity.shape) == (n, a), 'agent tensors disagree on (n_envs, n_agents)')
        self._c = _Agents(angles.data_ptr(), positions.data_ptr(), angvelocity.data_ptr(), velocity.data_ptr())

    angles = property(lambda self: self._angles)
    positions = property(lambda self: self._positions)
    angvelocity = property(lambda self: self._angvelocity)
    velocity = property(lambda self: self._velocity)

    def state(self, e):
        from .arrdict import arrdict
        return arrdict(angles=self._angles[e], positions=self._positions[e],
                       angvelocity=self._angvelocity[e], velocity=self._velocity[e])


class Scenery:
    """Holds the static scene plus the agents' model lines (common.h:185-214)."""

    def __init__(self, n_agents, lights, lines, textures, model):
        _require(isinstance(lights, Ragged2D), 'lights must be a Ragged2D')
        _require(isinstance(lines, Ragged3D), 'lines must be a Ragged3D')
        _require(isinstance(textures, Ragged2D), 'textures must be a Ragged2D')
        self._n_agents = int(n_agents)
        self._lights, self._lines, self._textures = lights, lines, textures
        self._model = _proxy(model, 3, 'model')
        _require(tuple(lines.vals.shape[1:]) == (2, 2), 'lines.vals must be (n, 2, 2)')
        _require(lights.vals.shape[1] == 3 and textures.vals.shape[1] == 3, 'lights/textures vals must be (n, 3)')
        _require(tuple(model.shape[1:]) == (2, 2), 'model must be (n, 2, 2)')
        _require(len(lights) == len(lines), 'lights and lines disagree on the number of environments')
        _require(len(textures) == lines.vals.size(0), 'textures must have one row per line')
        self._baked = Ragged1D(torch.ones_like(textures.vals[:, 0]).contiguous(), textures.widths)
        self._params = None   # set by core.Core so that several Cores can coexist
        self._c = None
        self._ws = {}

    n_agents = property(lambda self: self._n_agents)
    lights = property(lambda self: self._lights)
    lines = property(lambda self: self._lines)
    textures = property(lambda self: self._textures)
    baked = property(lambda self: self._baked)
    model = property(lambda self: self._model)

    def state(self, e):
        from .dotdict import dotdict
        se, ee = int(self._lines.starts[e].item()), int(self._lines.ends[e].item())
        return dotdict(n_agents=self._n_agents, lights=self._lights[e], lines=self._lines[e],
                       textures=self._textures[se:ee], model=self._model, baked=self._baked[se:ee])

    def invalidate(self):
        """Call after editing the STATIC geometry in place (`lines.vals` beyond the agents' rows, `lights`, the texture
        widths): the side tables derived from it — the spatial table, the light-visibility grid, the launch order —
        are built once and cached, so the kernels would otherwise keep seeing the old walls (the reference re-reads
        `lines` on every call; here only the agents' rows, the texels and `baked` are re-read). Rebuilt on the next call;
        re-bake if the lighting should follow."""
        self._c = None
        self._ws = {}

    def _struct(self):
        """The msb_scenery view of this object; built once (device pointers are stable, we hold the tensors). The static
        arrays are treated as immutable from here on: see `invalidate`."""
        if self._c is None:
            for name, t in (('lines', self._lines.vals), ('lights', self._lights.vals), ('textures', self._textures.vals)):
                _require(t.is_cuda, f'{name} must be a CUDA tensor')
            lw, iw = self._lines.widths, self._lights.widths
            self._tex_starts = self._textures._long_starts()
            self._c = _Scenery(
                n_envs=len(self._lines), n_agents=self._n_agents, n_model=self._model.size(0),
                max_lines=int(lw.max().item()) if lw.numel() else 0,
                max_lights=int(iw.max().item()) if iw.numel() else 0, occ_run=OCCLUDER_RUN,
                lines=self._lines.vals.data_ptr(), line_widths=lw.data_ptr(), line_starts=self._lines.starts.data_ptr(),
                lights=self._lights.vals.data_ptr(), light_widths=iw.data_ptr(), light_starts=self._lights.starts.data_ptr(),
                textures=self._textures.vals.data_ptr(), tex_widths=self._textures.widths.data_ptr(),
                tex_starts=self._tex_starts.data_ptr(), baked=self._baked.vals.data_ptr(), model=self._model.data_ptr(),
                n_lines=self._lines.vals.size(0), n_texels=self._textures.vals.size(0))
            if self._lines.vals.size(0) > 0:
                n_dynamic = self._n_agents * self._model.size(0)
                if TABLE_ORDER == 'str' and OCCLUDER_RUN == 16:
                    # the library's own builder (msb_build_table); the caller's part is the allocation
                    self._occ = _empty_table(lw, n_dynamic)
                else:
                    self._occ = _occluder_table(self._lines, n_dynamic, OCCLUDER_RUN, self._textures.widths, self._tex_starts)
                (self._c.occ_lines, self._c.occ_starts, self._c.occ_boxes, self._c.box_starts,
                 self._c.occ_meta, self._c.occ_rec) = (t.data_ptr() for t in self._occ)
                if TABLE_ORDER == 'str' and OCCLUDER_RUN == 16:
                    with _on_device(self._model) as stream:
                        _check(_lib.msb_build_table(ctypes.byref(self._c), stream))
                # launch order: the envs with the most lines first (their CTAs take longest), so the grid tails off on cheap ones
                self._env_order = torch.argsort(lw, descending=True, stable=True).int().contiguous()
                self._c.env_order = self._env_order.data_ptr()
                if USE_VISIBILITY_GRID and self._n_agents > 1:      # only rays that hit ANOTHER agent ask for dynamic light
                    self._vis = _visibility_grid(self._occ[2], self._occ[3], self._lines.widths, self._n_agents * self._model.size(0))
                    self._c.vis, self._c.vis_starts, self._c.vis_meta = (t.data_ptr() for t in self._vis)
                    with _on_device(self._model) as stream:
                        _check(_lib.msb_build_visibility(ctypes.byref(self._c), stream))
                # the kernels read these tables ahead of their programmatic-dependency wait (they are static): make sure
                # the one-off builders are done before anything is launched behind them
                torch.cuda.current_stream(self._model.device).synchronize()
        return self._c


def make_workspace(scenery, params, subsample=1):
    """Zeroed scratch for the second (dynamic-light) pass; returns (tensor, _Workspace struct)."""
    s = scenery._struct()
    nbytes = int(_lib.msb_workspace_bytes(ctypes.byref(params), ctypes.byref(s), int(subsample)))
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=scenery.model.device)
    return buf, _Workspace(buf.data_ptr(), nbytes)


def _shared_workspace(scenery, params):
    """The per-(scenery, res) scratch used by plain render() calls on the default stream."""
    key = (params.res, torch.cuda.current_stream(scenery.model.device).cuda_stream)
    if key not in scenery._ws:
        while len(scenery._ws) >= 4:                    # a handful of (res, stream) pairs at most: drop the oldest
            scenery._ws.pop(next(iter(scenery._ws)))
        scenery._ws[key] = make_workspace(scenery, params, 1)
    return scenery._ws[key][1]


def _morton16(x, y):
    """Interleave the low 16 bits of two int64 tensors."""
    def spread(v):
        v = v & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


@torch.no_grad()
def _empty_table(line_widths, n_dynamic, run=16):
    """The spatial table's arrays, allocated and with box_starts filled, for msb_build_table to fill: (occ_lines,
    occ_starts, occ_boxes, box_starts, occ_meta, occ_rec) of include/megastep_b200.h."""
    dev = line_widths.device
    W = (line_widths.long() - n_dynamic).clamp(min=0)
    nb = (W + run - 1) // run
    box_starts = (nb.cumsum(0) - nb).int().contiguous()
    nbox = int(nb.sum().item())
    n = line_widths.size(0)
    return (torch.empty((nbox * run, 4), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
            torch.empty((nbox, 4), dtype=torch.float32, device=dev), box_starts,
            torch.empty((n, 2), dtype=torch.float32, device=dev), torch.empty((nbox * run, 4), dtype=torch.int32, device=dev))


@torch.no_grad()
def _occluder_table(lines, n_dynamic, run=16, tex_widths=None, tex_starts=None):
    """(occ_lines, occ_starts, occ_boxes, box_starts, occ_meta, occ_rec) of include/megastep_b200.h in plain PyTorch
    — the restatement msb_build_table is tested against (and the only builder for the 'morton' order / other run
    lengths): every env's static segments packed into runs of `run`, each env padded to a whole number of runs, plus
    the bounding box of each run and, per row, its line's texel offset / count and original line index. Shadow tests
    and collisions ask order-free questions (any occluder / nearest obstacle); render() uses the line indices to
    restore the reference's line-order rule."""
    vals, widths = lines.vals.reshape(-1, 4), lines.widths.long()
    dev = vals.device
    env = lines.inverse.long()
    local = torch.arange(vals.size(0), device=dev) - lines.starts.long()[env]
    static = local >= n_dynamic
    sv, senv, sid = vals[static], env[static], local[static]
    mid = (sv[:, :2] + sv[:, 2:]) * .5
    W = (widths - n_dynamic).clamp(min=0)
    wstarts = W.cumsum(0) - W
    nb = (W + run - 1) // run
    if TABLE_ORDER == 'morton':
        lo = mid.min(0).values if len(mid) else torch.zeros(2, device=dev)
        cell = ((mid - lo) / .25).clamp(0, 65535).long()
        key = (senv << 32) | _morton16(cell[:, 0], cell[:, 1])
        order = torch.argsort(key)
    else:
        # sort-tile-recursive packing (the R-tree bulk load): ~sqrt(nb) vertical strips of equal population, each
        # sorted by y and cut into runs — tighter, less overlapping run boxes than a space-filling curve gives
        by_x = torch.argsort(mid[:, 0], stable=True)
        by_x = by_x[torch.argsort(senv[by_x], stable=True)]                 # by (env, x)
        rank_x = torch.empty_like(by_x)
        rank_x[by_x] = torch.arange(by_x.size(0), device=dev) - wstarts[senv[by_x]]
        strips = nb.double().sqrt().ceil().long().clamp(min=1)
        per_strip = ((nb + strips - 1) // strips) * run                      # segments per strip: whole runs
        strip = rank_x // per_strip[senv].clamp(min=1)
        by_y = torch.argsort(mid[:, 1], stable=True)
        order = by_y[torch.argsort((senv[by_y] << 20) | strip[by_y], stable=True)]   # by (env, strip, y)
    box_starts = nb.cumsum(0) - nb
    occ_starts = box_starts * run
    nbox = int(nb.sum().item())
    oenv = senv[order]
    rank = torch.arange(order.size(0), device=dev) - wstarts[oenv]
    row = occ_starts[oenv] + rank                       # where each sorted segment lands in the padded table
    big = torch.finfo(torch.float32).max
    occ = torch.full((nbox * run, 4), 1e30, dtype=torch.float32, device=dev)    # padding: a far, zero-length segment
    occ[row] = sv[order]
    rec = torch.zeros((nbox * run, 4), dtype=torch.int32, device=dev)
    rec[:, 3] = -1
    rec[row, 3] = sid[order].int()
    if tex_widths is not None:
        gl = lines.starts.long()[oenv] + sid[order]            # global line of each sorted row
        rec[row, :2] = tex_starts[gl].contiguous().view(torch.int32).reshape(-1, 2)     # little-endian {lo, hi}
        rec[row, 2] = tex_widths[gl]
    box = box_starts[oenv] + rank // run
    so = sv[order]
    xmin = torch.full((nbox,), big, device=dev).scatter_reduce(0, box, torch.minimum(so[:, 0], so[:, 2]), 'amin')
    ymin = torch.full((nbox,), big, device=dev).scatter_reduce(0, box, torch.minimum(so[:, 1], so[:, 3]), 'amin')
    xmax = torch.full((nbox,), -big, device=dev).scatter_reduce(0, box, torch.maximum(so[:, 0], so[:, 2]), 'amax')
    ymax = torch.full((nbox,), -big, device=dev).scatter_reduce(0, box, torch.maximum(so[:, 1], so[:, 3]), 'amax')
    boxes = torch.stack([xmin, ymin, xmax, ymax], -1).contiguous()
    # per env: the longest segment extent and the extent of the env (they size the shadow cull's safety margin)
    n = widths.size(0)
    ext = torch.maximum((so[:, 2] - so[:, 0]).abs(), (so[:, 3] - so[:, 1]).abs())
    vmax = torch.zeros(n, device=dev).scatter_reduce(0, oenv, ext, 'amax')
    lo_x = torch.full((n,), big, device=dev).scatter_reduce(0, oenv, torch.minimum(so[:, 0], so[:, 2]), 'amin')
    lo_y = torch.full((n,), big, device=dev).scatter_reduce(0, oenv, torch.minimum(so[:, 1], so[:, 3]), 'amin')
    hi_x = torch.full((n,), -big, device=dev).scatter_reduce(0, oenv, torch.maximum(so[:, 0], so[:, 2]), 'amax')
    hi_y = torch.full((n,), -big, device=dev).scatter_reduce(0, oenv, torch.maximum(so[:, 1], so[:, 3]), 'amax')
    diam = torch.where(W > 0, torch.maximum(hi_x - lo_x, hi_y - lo_y), torch.zeros_like(vmax))
    meta = torch.stack([vmax, diam], -1).contiguous()
    return occ, occ_starts.int().contiguous(), boxes, box_starts.int().contiguous(), meta, rec


VIS_CELL = .25     # metres; must match VIS_CELL in csrc/megastep_b200.cu


@torch.no_grad()
def _visibility_grid(boxes, box_starts, line_widths, n_dynamic):
    """(vis, vis_starts, vis_meta) of include/megastep_b200.h: an (unfilled) grid of VIS_CELL cells over each env's
    static geometry — the union of its run boxes — for msb_build_visibility to fill."""
    dev = boxes.device
    n = line_widths.size(0)
    nb = ((line_widths.long() - n_dynamic).clamp(min=0) + OCCLUDER_RUN - 1) // OCCLUDER_RUN
    env = torch.repeat_interleave(torch.arange(n, device=dev), nb)
    big = torch.finfo(torch.float32).max
    lo_x = torch.full((n,), big, device=dev).scatter_reduce(0, env, boxes[:, 0], 'amin')
    lo_y = torch.full((n,), big, device=dev).scatter_reduce(0, env, boxes[:, 1], 'amin')
    hi_x = torch.full((n,), -big, device=dev).scatter_reduce(0, env, boxes[:, 2], 'amax')
    hi_y = torch.full((n,), -big, device=dev).scatter_reduce(0, env, boxes[:, 3], 'amax')
    empty = nb == 0
    gx = torch.where(empty, torch.zeros_like(nb), ((hi_x - lo_x).clamp(min=0) / VIS_CELL).ceil().long().clamp(min=1, max=4096))
    gy = torch.where(empty, torch.zeros_like(nb), ((hi_y - lo_y).clamp(min=0) / VIS_CELL).ceil().long().clamp(min=1, max=4096))
    cells = gx * gy
    starts = (cells.cumsum(0) - cells).contiguous()
    meta = torch.stack([torch.where(empty, torch.zeros_like(lo_x), lo_x), torch.where(empty, torch.zeros_like(lo_y), lo_y),
                        gx.float(), gy.float()], -1).contiguous()
    vis = torch.zeros(max(int(cells.sum().item()), 1), dtype=torch.int32, device=dev)
    return vis, starts, meta


class Render:
    """The result of a render() call. Exactly five public attributes (modules.unpack walks dir())."""
    __slots__ = ('screen', 'indices', 'locations', 'dots', 'distances')

    def __init__(self, indices, locations, dots, distances, screen):
        self.indices, self.locations, self.dots, self.distances, self.screen = indices, locations, dots, distances, screen


class Physics:
    __slots__ = ('progress',)

    def __init__(self, progress):
        self.progress = progress


# --------------------------------------------------------------------------------------------------------------
# functions
# --------------------------------------------------------------------------------------------------------------
_PARAMS = None
OCCLUDER_RUN = 16        # segments per run / bounding box of the spatial table
TABLE_ORDER = 'str'      # how the table's runs are formed: 'str' (sort-tile-recursive) or 'morton'
USE_WORKSPACE = True   # False: agent-hit rays are lit inline by the first pass (same results; used by tests)
USE_VISIBILITY_GRID = True   # False: sceneries are built without the light-visibility grid (same results, more shadow scans)


def make_params(agent_radius, res, fov, fps):
    p = Params()
    _check(_lib.msb_params_init(ctypes.byref(p), float(agent_radius), int(res), float(fov), float(fps)))
    return p


def initialize(agent_radius, res, fov, fps):
    """Sets the process-wide default parameters used by bake/physics/render (kernels.cu:18-27)."""
    global _PARAMS
    _PARAMS = make_params(agent_radius, res, fov, fps)


def _params(scenery, params):
    p = params if params is not None else _PARAMS
    _require(p is not None, 'initialize(agent_radius, res, fov, fps) must be called before bake/physics/render')
    return p


class _on_device:
    """Makes the tensors' device current for the launch (the reference launches on whatever is current)."""
    __slots__ = ('idx', 'prev')

    def __init__(self, tensor):
        self.idx = tensor.device.index

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return torch.cuda.current_stream(self.idx).cuda_stream

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)


def bake(scenery, params=None):
    """Pre-computes the lighting of the static geometry into scenery.baked."""
    _require(isinstance(scenery, Scenery), 'scenery must be a Scenery')
    p = _params(scenery, params)
    s = scenery._struct()
    with _on_device(scenery.model) as stream:
        _check(_lib.msb_bake(ctypes.byref(p), ctypes.byref(s), stream))


def physics(scenery, agents, params=None):
    """Advances the agents by one tick, resolving collisions; returns Physics(progress)."""
    _require(isinstance(scenery, Scenery) and isinstance(agents, Agents), 'expected (Scenery, Agents)')
    p = _params(scenery, params)
    s = scenery._struct()
    n, a = agents._angles.shape
    _require(n == s.n_envs and a == s.n_agents, 'agents and scenery disagree on (n_envs, n_agents)')
    progress = torch.empty((n, a), dtype=torch.float32, device=agents._angles.device)
    with _on_device(agents._angles) as stream:
        _check(_lib.msb_physics(ctypes.byref(p), ctypes.byref(s), ctypes.byref(agents._c), progress.data_ptr(), stream))
    return Physics(progress)


def _alloc_render(n, a, r, device):
    f = dict(dtype=torch.float32, device=device)
    return Render(indices=torch.empty((n, a, r), dtype=torch.int32, device=device),
                  locations=torch.empty((n, a, r), **f), dots=torch.empty((n, a, r), **f),
                  distances=torch.empty((n, a, r), **f), screen=torch.empty((n, a, r, 3), **f))


def _render_struct(r):
    return _RenderOut(r.indices.data_ptr(), r.locations.data_ptr(), r.dots.data_ptr(), r.distances.data_ptr(),
                      r.screen.data_ptr())


def render(scenery, agents, params=None):
    """Renders the scenery onto the agents' cameras; returns a Render. Also moves the agents' model lines inside
    scenery.lines to the agents' current poses (the reference's draw_kernel side effect)."""
    _require(isinstance(scenery, Scenery) and isinstance(agents, Agents), 'expected (Scenery, Agents)')
    p = _params(scenery, params)
    s = scenery._struct()
    n, a = agents._angles.shape
    _require(n == s.n_envs and a == s.n_agents, 'agents and scenery disagree on (n_envs, n_agents)')
    out = _alloc_render(n, a, p.res, agents._angles.device)
    c = _render_struct(out)
    with _on_device(agents._angles) as stream:
        ws = _shared_workspace(scenery, p) if USE_WORKSPACE else None
        _check(_lib.msb_render(ctypes.byref(p), ctypes.byref(s), ctypes.byref(agents._c), ctypes.byref(c), None,
                               ctypes.byref(ws) if ws is not None else None, stream))
    return out


class StepPlan:
    """A pre-bound fused step: MomentumMovement -> physics -> render -> RGB/Depth/IMU heads in ONE kernel launch,
    writing into persistent output buffers (so it can also be captured in a CUDA graph).

    Built by `modules.FusedStep`; kept here because it owns ctypes structs.
    """

    def __init__(self, scenery, agents, params, actions=None, accel=5., ang_accel=180., decay=.125,
                 raw=True, subsample=None, max_depth=10., speed_scale=10., ang_scale=360.):
        n, a = agents.angles.shape
        dev = agents.angles.device
        self.scenery, self.agents, self.params = scenery, agents, params
        self.progress = torch.empty((n, a), dtype=torch.float32, device=dev)
        self.render = _alloc_render(n, a, params.res, dev) if raw else None
        self._out = _render_struct(self.render) if raw else None
        self.actions = actions
        self._mv = None
        if actions is not None:
            _proxy(actions, 2, 'actions', torch.int32)
            self._mv = _Movement(actions.data_ptr(), accel, ang_accel, decay)
        self.rgb = self.depth = self.imu = None
        self._obs = None
        if subsample is not None:
            ro = params.res // subsample
            self.rgb = torch.empty((n, a, 3, 1, ro), dtype=torch.float32, device=dev)
            self.depth = torch.empty((n, a, 1, 1, ro), dtype=torch.float32, device=dev)
            self.imu = torch.empty((n, a, 3), dtype=torch.float32, device=dev)
            self._obs = _ObsOut(self.rgb.data_ptr(), self.depth.data_ptr(), self.imu.data_ptr(), subsample,
                                max_depth, speed_scale, ang_scale)
        self._s = scenery._struct()
        self._wsbuf, self._ws = make_workspace(scenery, params, subsample or 1) if USE_WORKSPACE else (None, None)

    def step(self):
        """movement (if actions were bound) + physics + render + heads, one launch"""
        with _on_device(self.progress) as stream:
            _check(_lib.msb_step(ctypes.byref(self.params), ctypes.byref(self._s), ctypes.byref(self.agents._c),
                                 ctypes.byref(self._mv) if self._mv is not None else None, self.progress.data_ptr(),
                                 ctypes.byref(self._out) if self._out is not None else None,
                                 ctypes.byref(self._obs) if self._obs is not None else None,
                                 ctypes.byref(self._ws) if self._ws is not None else None, stream))

    __call__ = step

    def graph_create(self):
        """Captures step() once as a CUDA graph inside the library (msb_step_graph_create); step() must have run before."""
        self.graph_destroy()
        h = ctypes.c_void_p()
        with _on_device(self.progress):
            _check(_lib.msb_step_graph_create(ctypes.byref(self.params), ctypes.byref(self._s), ctypes.byref(self.agents._c),
                                              ctypes.byref(self._mv) if self._mv is not None else None, self.progress.data_ptr(),
                                              ctypes.byref(self._out) if self._out is not None else None,
                                              ctypes.byref(self._obs) if self._obs is not None else None,
                                              ctypes.byref(self._ws) if self._ws is not None else None, ctypes.byref(h)))
        self._graph = h

    def graph_run(self, actions_host_ptr, progress_host_ptr, sync=True):
        """actions (pinned host, or 0) up, the captured step, progress (pinned host, or 0) down, stream sync: one call"""
        dev = self.progress.device.index
        if torch.cuda.current_device() != dev:
            with _on_device(self.progress) as stream:
                _check(_lib.msb_step_graph_run(self._graph, actions_host_ptr or None, progress_host_ptr or None, stream, int(sync)))
            return
        stream = torch.cuda.current_stream(dev).cuda_stream
        if _lib.msb_step_graph_run(self._graph, actions_host_ptr or None, progress_host_ptr or None, stream, int(sync)):
            _check(1)

    def graph_destroy(self):
        if getattr(self, '_graph', None):
            _lib.msb_step_graph_destroy(self._graph)
        self._graph = None

    def __del__(self):
        try:
            self.graph_destroy()
        except Exception:       # noqa: BLE001  (interpreter shutdown)
            pass

    def move_only(self):
        """movement + physics, one launch (msb_move); the agents are not rendered"""
        with _on_device(self.progress) as stream:
            _check(_lib.msb_move(ctypes.byref(self.params), ctypes.byref(self._s), ctypes.byref(self.agents._c), ctypes.byref(self._mv),
                                 self.progress.data_ptr(), stream))

    def render_only(self):
        """render + heads, one launch; the agents are not moved"""
        with _on_device(self.progress) as stream:
            _check(_lib.msb_render(ctypes.byref(self.params), ctypes.byref(self._s), ctypes.byref(self.agents._c),
                                   ctypes.byref(self._out) if self._out is not None else None,
                                   ctypes.byref(self._obs) if self._obs is not None else None,
                                   ctypes.byref(self._ws) if self._ws is not None else None, stream))


# --------------------------------------------------------------------------------------------------------------
# environment rules on the device (include/megastep_b200.h, "Environment rules")
# --------------------------------------------------------------------------------------------------------------
def ledger_words(scenery):
    """Length of the int32 tensor that holds one 'seen' bit per texel of the scenery."""
    return (scenery.textures.vals.size(0) + 31) // 32


def env_ledger_mark(scenery, indices, locations, seen, potential, gained):
    """Explorer's bookkeeping (explorer.py:34-58): marks the texel under every ray in `seen` (int32 words, one bit per
    texel) and adds the number of newly seen texels per env to `potential` and `gained` (int32 (N,))."""
    s = scenery._struct()
    n, a, r = indices.shape
    _require(indices.dtype == torch.int32 and indices.is_contiguous() and locations.is_contiguous(), 'indices / locations must be contiguous (N, A, R) tensors')
    with _on_device(indices) as stream:
        _check(_lib.msb_env_ledger_mark(ctypes.byref(s), indices.data_ptr(), locations.data_ptr(), a, r, seen.data_ptr(), potential.data_ptr(),
                                        gained.data_ptr(), stream))


def env_ledger_clear(scenery, reset, seen, potential):
    """explorer.py:73-77: the envs flagged in `reset` (uint8 / bool (N,)) forget what they have seen."""
    s = scenery._struct()
    reset = reset.to(torch.uint8) if reset.dtype != torch.uint8 else reset
    with _on_device(seen) as stream:
        _check(_lib.msb_env_ledger_clear(ctypes.byref(s), reset.data_ptr(), seen.data_ptr(), potential.data_ptr(), stream))


def env_shoot(scenery, agents, indices, subsample, bounds, clearance, matchings, hits, health, damage):
    """Deathmatch's crosshair rule + health / damage updates (deathmatch.py:54-72, 75-80), one launch."""
    s = scenery._struct()
    n, a, r = indices.shape
    _require(indices.dtype == torch.int32 and indices.is_contiguous(), 'indices must be a contiguous int32 (N, A, R) tensor')
    with _on_device(indices) as stream:
        _check(_lib.msb_env_shoot(ctypes.byref(s), ctypes.byref(agents._c), indices.data_ptr(), r, int(subsample), bounds.data_ptr(),
                                  float(clearance), matchings.data_ptr(), hits.data_ptr(), health.data_ptr(), damage.data_ptr(), stream))


def env_respawn(scenery, agents, reset, spawn_positions, spawn_angles, seed, tick, choices=None):
    """RandomSpawns (modules.py:312-326) on the device: no nonzero(), no host sync."""
    s = scenery._struct()
    reset = reset.to(torch.uint8) if reset.dtype != torch.uint8 else reset
    with _on_device(reset) as stream:
        _check(_lib.msb_env_respawn(ctypes.byref(s), ctypes.byref(agents._c), reset.contiguous().data_ptr(), spawn_positions.data_ptr(),
                                    spawn_angles.data_ptr(), spawn_angles.shape[-1], int(seed) & 0xffffffff, int(tick) & 0xffffffff,
                                    choices.data_ptr() if choices is not None else None, stream))


def pack_obs(rgb, depth, imu, rows, mode, imu_offset=0):
    """rgb (N, A, 3, 1, ro), depth (N, A, 1, 1, ro), imu (N, A, 3), all fp32 -> rows (N, width): one launch (msb_pack_obs)."""
    n, a, ro = rgb.shape[0], rgb.shape[1], rgb.shape[-1]
    _require(rgb.is_contiguous() and depth.is_contiguous() and imu.is_contiguous() and rows.is_contiguous(), 'pack_obs needs contiguous tensors')
    _require(rgb.dtype == torch.float32 and depth.dtype == torch.float32 and imu.dtype == torch.float32, 'pack_obs packs fp32 observations')
    with _on_device(rows) as stream:
        _check(_lib.msb_pack_obs(rgb.data_ptr(), depth.data_ptr(), imu.data_ptr(), n, a, ro, rows.data_ptr(), rows.stride(0) * rows.element_size(),
                                 int(mode), int(imu_offset), stream))


def set_option(name, value):
    _check(_lib.msb_set_option(name.encode(), int(value)))


def get_option(name):
    return int(_lib.msb_get_option(name.encode()))


def launch_count():
    return int(_lib.msb_launch_count())


def library_path():
    return _LIBPATH
