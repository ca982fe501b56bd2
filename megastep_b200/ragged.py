"""Ragged arrays: a flat `vals` array plus per-row `widths` (reference: megastep/ragged.py:7-75).

`Ragged(vals, widths)` dispatches on the argument type: numpy arrays give a `RaggedNumpy`, torch tensors give the
`cuda.Ragged{1,2,3}D` the kernels consume.
"""
import numbers

import numpy as np


class RaggedNumpy:
    """numpy-backed ragged array with the same attributes as the tensor-backed one:
    vals (V, ...), widths (W,), starts (W,), ends (W,), inverse (V,)."""

    def __init__(self, vals, widths):
        widths = np.asarray(widths)
        assert widths.sum() == vals.shape[0], 'the widths must sum to the length of vals'
        ends = widths.cumsum().astype(int)
        self.vals, self.widths = vals, widths
        self.starts, self.ends = ends - widths, ends
        marks = np.zeros(int(widths.sum()), dtype=self.starts.dtype)
        marks[self.starts[self.starts < len(marks)]] = 1
        self.inverse = marks.cumsum().astype(int) - 1

    def __getitem__(self, x):
        if isinstance(x, numbers.Integral):
            return self.vals[self.starts[x]:self.ends[x]]
        if isinstance(x, slice):
            assert x.step in (None, 1), 'ragged slices must have step 1'
            lo = x.start or 0
            hi = len(self.ends) if x.stop is None else x.stop
            return RaggedNumpy(self.vals[self.starts[lo]:self.ends[hi - 1]], self.widths[lo:hi])
        raise ValueError(f'Can\'t handle index "{x}"')

    def __len__(self):
        return len(self.widths)

    def torchify(self):
        from .arrdict import torchify
        return Ragged(torchify(self.vals), torchify(self.widths))

    def __repr__(self):
        return f'{type(self).__name__}({self.widths})'


def Ragged(vals, widths):
    """numpy in -> RaggedNumpy; tensors in -> cuda.Ragged{vals.ndim}D."""
    if isinstance(vals, np.ndarray):
        return RaggedNumpy(vals, widths)
    from . import cuda
    return getattr(cuda, f'Ragged{vals.ndim}D')(vals, widths)
