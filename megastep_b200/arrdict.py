"""dotdicts of arrays/tensors: indexing, slicing and arithmetic broadcast over the values.

Mirrors the behaviour of rebar/arrdict.py:11-162 that the kept API relies on: `d[idx]` indexes every value,
`d[idx] = other_arrdict` assigns into every value, binary operators apply value-wise (against a scalar or a
same-keyed dict), plus the `torchify / numpyify / stack / cat / clone` helpers.
"""
import numpy as np
import torch

from .dotdict import dotdict, mapping

__all__ = ['arrdict', 'torchify', 'numpyify', 'stack', 'cat', 'clone']


def _is_key(x):
    return isinstance(x, str) or (isinstance(x, tuple) and len(x) > 0 and all(isinstance(p, str) for p in x))


class arrdict(dotdict):

    def __getitem__(self, x):
        if isinstance(x, str):
            return super().__getitem__(x)
        return type(self)((k, v[x]) for k, v in self.items())

    def __setitem__(self, x, y):
        if _is_key(x):
            super().__setitem__(x, y)
        elif isinstance(y, arrdict):
            for k in self:
                super().__getitem__(k)[x] = y[k]
        else:
            raise ValueError('Set items with a string key, or index-assign another arrdict')

    def __setattr__(self, name, value):
        raise ValueError('Setting by attribute is not allowed; set by key instead')

    def _binary(self, name, rhs):
        if isinstance(rhs, dict):
            return self.starmap(name, rhs)
        return type(self)((k, getattr(v, name)(rhs)) for k, v in self.items())


def _install_operators():
    names = ['lt', 'le', 'eq', 'ne', 'ge', 'gt', 'add', 'sub', 'mul', 'matmul', 'truediv', 'floordiv', 'mod',
             'divmod', 'pow', 'lshift', 'rshift', 'and', 'or', 'xor']
    reflected = ['radd', 'rsub', 'rmul', 'rmatmul', 'rtruediv', 'rfloordiv', 'rmod', 'rdivmod', 'rpow', 'rand',
                 'ror', 'rxor']
    for n in names + reflected:
        dunder = f'__{n}__'
        setattr(arrdict, dunder, (lambda d: lambda self, rhs: self._binary(d, rhs))(dunder))
    arrdict.__hash__ = None


_install_operators()


@mapping
def torchify(a):
    """numpy (or nested dicts of numpy) -> CPU tensors; floats to float32, ints to int32, bools to bool."""
    if hasattr(a, 'torchify'):
        return a.torchify()
    a = np.asarray(a)
    for kind, dtype in ((np.floating, torch.float32), (np.integer, torch.int32), (np.bool_, torch.bool)):
        if np.issubdtype(a.dtype, kind):
            return torch.as_tensor(np.array(a), dtype=dtype)
    raise ValueError(f"Can't torchify an array of dtype {a.dtype}")


@mapping
def numpyify(t):
    """tensors (or nested dicts of tensors) -> numpy arrays (copied to host)."""
    if isinstance(t, tuple):
        return tuple(numpyify(x) for x in t)
    if isinstance(t, torch.Tensor):
        return t.detach().clone().cpu().numpy()
    if hasattr(t, 'numpyify'):
        return t.numpyify()
    return t


def _combine(xs, torch_fn, numpy_fn, scalar_fn, args, kwargs):
    head = xs[0]
    if isinstance(head, dict):
        return type(head)((k, _combine([x[k] for x in xs], torch_fn, numpy_fn, scalar_fn, args, kwargs)) for k in head)
    if isinstance(head, torch.Tensor):
        return torch_fn(list(xs), *args, **kwargs)
    if isinstance(head, np.ndarray):
        return numpy_fn(list(xs), *args, **kwargs)
    if np.isscalar(head):
        return scalar_fn(xs)
    raise ValueError(f"Can't combine values of type {type(head)}")


def stack(xs, *args, **kwargs):
    """Stack a sequence of arrays / tensors / dicts of them along a new axis."""
    return _combine(xs, torch.stack, np.stack, np.array, args, kwargs)


def cat(xs, *args, **kwargs):
    """Concatenate a sequence of arrays / tensors / dicts of them; python scalars become a 1-D array."""
    return _combine(xs, torch.cat, np.concatenate, np.array, args, kwargs)


@mapping
def clone(t):
    if hasattr(t, 'clone'):
        return t.clone()
    if hasattr(t, 'copy'):
        return t.copy()
    return t
