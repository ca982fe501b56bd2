"""The on-disk geometry cache of the reference (megastep/cubicasa.py:107-224), so that real Cubicasa5k geometries
drop in when a cache file is at hand.

Format (cubicasa.py:149-174): gzip( np.savez( flat ) ) where `flat` maps "<id>/walls", "<id>/lights", "<id>/masks",
"<id>/res" to arrays — one geometry per id, the dict format geometry.py documents:

    walls  (W, 2, 2) float  wall segments, metres
    lights (I, 2)    float  light positions, metres
    masks  (H, W)    int16  -1 walls, 0 outside, k >= 1 the k-th room
    res    ()        float  metres per mask cell

`sample(n, split, seed)` follows the reference's: ids sorted, permuted by RandomState(seed), 90/10 training/test, cycled
when more geometries are asked for than there are. There is no download here (the reference fetches a ~10 MB file
from the network and asks for a licence confirmation, cubicasa.py:33-60,166-168): the cache has to exist already —
`megastep_b200.synthetic.sample` is the stand-in when it does not.
"""
import ast
import gzip
from io import BytesIO
from pathlib import Path
from zipfile import ZipFile

import numpy as np

from .dotdict import dotdict

CACHE = Path('.cache/cubicasa-geometry.npz.gz')     # where the reference keeps it (cubicasa.py:153)
FIELDS = ('walls', 'lights', 'masks', 'res')


def flatten(tree, prefix=''):
    """{'a': {'b': x}} -> {'a/b': x}: the archive's member names (reference: cubicasa.py:107-115)."""
    flat = {}
    stack = [(prefix, tree)]
    while stack:
        path, node = stack.pop()
        for key, value in node.items():
            name = f'{path}/{key}' if path else str(key)
            if isinstance(value, dict):
                stack.append((name, value))
            else:
                flat[name] = value
    return flat


def unflatten(flat):
    """{'a/b': x} -> {'a': {'b': x}}, nested containers of the same type as `flat` (reference: cubicasa.py:117-125)."""
    kind = type(flat)
    tree = kind()
    for name in flat:
        *parents, leaf = name.split('/')
        node = tree
        for part in parents:
            if part not in node:
                node[part] = kind()
            node = node[part]
        node[leaf] = flat[name]
    return tree


def fastload(raw):
    """One .npy member of the archive -> array, parsing the (version 1.0) header by hand (cubicasa.py:135-147);
    falls back to numpy's own reader for anything else (newer header versions, Fortran order)."""
    raw = bytes(raw)
    if raw[:6] == b'\x93NUMPY' and raw[6] == 1:
        headerlen = int(np.frombuffer(raw[8:10], dtype='<u2')[0])
        header = ast.literal_eval(raw[10:10 + headerlen].decode('latin1'))
        if not header.get('fortran_order', False):
            return np.frombuffer(raw[10 + headerlen:], dtype=np.dtype(header['descr'])).reshape(header['shape'])
    return np.load(BytesIO(raw), allow_pickle=False)


def save_geometries(geometries, path=CACHE):
    """Writes {id: geometry} in the reference's cache format (cubicasa.py:158-165). Only the four geometry fields are
    kept (the reference's cache holds nothing else; `id` is re-attached by `sample`)."""
    flat = flatten({str(k): {f: np.asarray(g[f]) for f in FIELDS} for k, g in geometries.items()})
    bs = BytesIO()
    np.savez(bs, **flat)
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    path.write_bytes(gzip.compress(bs.getvalue()))
    return path


def load_geometries(path=CACHE):
    """{id: dotdict(walls, lights, masks, res)} from a cache file (cubicasa.py:170-174)."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(
            f'{path} not found. This build does not download the Cubicasa5k geometry cache (no network, and the '
            'dataset is licensed for non-commercial use only: see the megastep FAQ); put the reference\'s '
            'cubicasa-geometry.npz.gz there, or use megastep_b200.synthetic.sample for cubicasa-shaped stand-ins.')
    raw = gzip.decompress(path.read_bytes())
    with ZipFile(BytesIO(raw)) as zf:
        flat = dotdict({n[:-4]: fastload(zf.read(n)) for n in zf.namelist()})
    return unflatten(flat)


_cache = {}


def sample(n_geometries, split='training', seed=1, path=CACHE):
    """A seeded sample of cached geometries, as the reference draws them (cubicasa.py:177-224)."""
    key = str(Path(path))
    if key not in _cache:
        gs = load_geometries(path)
        _cache[key] = type(gs)({k: type(v)({'id': k, **v}) for k, v in gs.items()})
    cache = _cache[key]
    ids = np.random.RandomState(seed).permutation(sorted(cache))         # the reference's order: sorted ids, then shuffled
    n_train = int(.9 * len(cache))
    if split not in ('training', 'test', 'all'):
        raise ValueError('Split must be train/test/all')
    chosen = {'training': ids[:n_train], 'test': ids[n_train:], 'all': ids}[split]
    return [cache[chosen[i % len(chosen)]] for i in range(n_geometries)]
