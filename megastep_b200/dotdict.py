"""Attribute-access dictionaries used as the state/observation containers of the API.

The reference hands these across its extension boundary (`Agents.state` / `Scenery.state` return them,
megastep/src/common.h:168-175,201-212) and its Python modules build observations out of them, so the drop-in API
needs an equivalent. Behaviour mirrored from rebar/dotdict.py:7-177: keys readable as attributes; an unknown
attribute is looked up on every value and the results re-wrapped; calling the dict calls every value; `map` /
`starmap` apply a function over the leaves of a (possibly nested) tree.
"""
from collections import OrderedDict
import functools

__all__ = ['dotdict', 'mapping', 'starmapping', 'leaves', 'treestr']

_WIDTH = 119
_HEIGHT = 200


def _describe(v, room):
    if isinstance(v, dotdict):
        return str(v)
    if isinstance(v, (list, set, dict)):
        return f'{type(v).__name__}({len(v)},)'
    shape = getattr(v, 'shape', None)
    if shape is not None:
        dtype = getattr(v, 'dtype', None)
        inner = f'{tuple(shape)}' if dtype is None else f'{tuple(shape)}, {dtype}'
        return f'{type(v).__name__}({inner})'
    text = str(v).splitlines() or ['']
    if len(text) > 1 or len(text[0]) > room:
        return text[0][:room] + ' ...'
    return text[0]


def treestr(tree):
    """Render a tree of dotdicts as an indented key/summary table."""
    pad = 4 + max((len(str(k)) for k in tree), default=0)
    out = [f'{type(tree).__name__}:']
    for k, v in tree.items():
        first, *rest = _describe(v, _WIDTH - pad).splitlines() or ['']
        out.append(str(k).ljust(pad) + first)
        out.extend(' ' * pad + line for line in rest)
        if len(out) >= _HEIGHT - 1:
            out.append('...')
            break
    return '\n'.join(out)


def mapping(f):
    """Lift `f` (a callable, or a method name) to act on every leaf of a tree of dicts."""
    def lifted(x, *args, **kwargs):
        if isinstance(x, dict):
            return type(x)((k, lifted(v, *args, **kwargs)) for k, v in x.items())
        if isinstance(f, str):
            return getattr(x, f)(*args, **kwargs)
        return f(x, *args, **kwargs)
    if callable(f):
        functools.update_wrapper(lifted, f)
    return lifted


def starmapping(f):
    """Lift `f` to act key-by-key across several same-shaped trees of dicts."""
    def lifted(x, *others):
        if isinstance(x, dict):
            return type(x)((k, lifted(x[k], *(o[k] for o in others))) for k in x)
        if isinstance(f, str):
            return getattr(x, f)(*others)
        return f(x, *others)
    if callable(f):
        functools.update_wrapper(lifted, f)
    return lifted


def leaves(tree):
    """Flat list of the non-dict values of a tree."""
    if isinstance(tree, dict):
        return [leaf for v in tree.values() for leaf in leaves(v)]
    return [tree]


class dotdict(OrderedDict):
    """An ordered dict whose string keys can be read as attributes."""

    def __getattr__(self, name):
        if name in self:
            return self[name]
        if name.startswith('__'):
            raise AttributeError(name)
        try:
            return type(self)((k, getattr(v, name)) for k, v in self.items())
        except AttributeError:
            raise AttributeError(f"No key '{name}', and not every value has an attribute '{name}'") from None

    def __dir__(self):
        return sorted(set(super().__dir__()) | {k for k in self if isinstance(k, str)})

    def __call__(self, *args, **kwargs):
        return type(self)((k, v(*args, **kwargs)) for k, v in self.items())

    def __str__(self):
        return treestr(self)

    __repr__ = __str__

    def __getstate__(self):
        return self

    def __setstate__(self, state):
        self.update(state)

    def copy(self):
        return type(self)(self.items())

    def pipe(self, f, *args, **kwargs):
        return f(self, *args, **kwargs)

    def map(self, f, *args, **kwargs):
        return mapping(f)(self, *args, **kwargs)

    def starmap(self, f, *others):
        return starmapping(f)(self, *others)
