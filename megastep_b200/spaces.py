"""Shape-only descriptors of observation and action spaces (reference: megastep/spaces.py:3-28).

A space says what an env's observation / action tensors look like per environment: `MultiVector(n_agents, dim)` is
`dim` floats per agent, `MultiImage(n_agents, C, H, W)` a (C, H, W) image per agent, `MultiDiscrete(n_agents, n_actions)`
one of `n_actions` choices per agent, `MultiConstant(n_agents)` one number per agent, `MultiEmpty()` nothing. Policy
heads dispatch on the class and read `.shape`; the classes carry nothing else. They are generated from one table — the
constructor's argument names, which are also the names of the shape's dimensions.
"""

_DIMENSIONS = {
    'MultiEmpty': (),
    'MultiVector': ('n_agents', 'dim'),
    'MultiImage': ('n_agents', 'C', 'H', 'W'),
    'MultiConstant': ('n_agents',),
    'MultiDiscrete': ('n_agents', 'n_actions'),
}


def _space(name, dims):
    def __init__(self, *args, **kwargs):
        given = dict(zip(dims, args))
        if len(args) > len(dims) or set(kwargs) - set(dims) or set(kwargs) & set(given):
            raise TypeError(f'{name}({", ".join(dims)}) got {args} {kwargs}')
        given.update(kwargs)
        missing = [d for d in dims if d not in given]
        if missing:
            raise TypeError(f'{name}() missing {missing}')
        if dims:
            self.shape = tuple(given[d] for d in dims)

    def __repr__(self):
        return f'{name}{getattr(self, "shape", ())}'

    def __eq__(self, other):
        return type(other) is type(self) and getattr(other, 'shape', ()) == getattr(self, 'shape', ())

    return type(name, (), {'__init__': __init__, '__repr__': __repr__, '__eq__': __eq__, '__hash__': lambda self: hash((name, getattr(self, 'shape', ()))),
                           '__doc__': f'{name}({", ".join(dims)}): shape = ({", ".join(dims)}{"," if len(dims) == 1 else ""})', 'dimensions': dims})


globals().update({name: _space(name, dims) for name, dims in _DIMENSIONS.items()})
__all__ = list(_DIMENSIONS)
