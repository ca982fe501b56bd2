"""Ahead-of-time build of the native library (in-tree, so the .so travels with the source snapshot).

    python -m megastep_b200.build

produces megastep_b200/libmegastep_b200.so from csrc/*.cu for sm_100a only. There is no JIT and no fallback: if
the library is missing, `import megastep_b200.cuda` raises.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'megastep_b200.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'msb_math.cuh'), os.path.join(HERE, '..', 'include', 'megastep_b200.h')]
LIB = os.path.join(HERE, 'libmegastep_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',   # B200 only
    '-O3', '-lineinfo', '-std=c++17',
    '-ftz=true',                                    # the reference is built --use_fast_math; see csrc/msb_math.cuh
    '-Xcompiler', '-fPIC', '-shared',
]


def nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build libmegastep_b200.so')


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, '-o', LIB, SRC]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{res.stdout}\n{res.stderr}')
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
