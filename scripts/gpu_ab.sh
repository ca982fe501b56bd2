#!/bin/bash
# A/B timing of library builds x runtime options (Deathmatch 4096x4x128 render)
for lib in build_variants/lib_mb4.so build_variants/lib_mb5.so build_variants/lib_mb6.so; do
    echo -n "$lib: "
    MEGASTEP_B200_LIB=$PWD/$lib timeout 300 python - <<PY 2>&1 | tail -1
import sys; sys.argv=['x','none','16']
sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import json, torch
import gpu_exp as g
from megastep_b200 import cuda
c = g.setup()
out = {}
for split in (1, 0):
    cuda.set_option('split_render', split)
    for nch in (4, 2, 1):
        cuda.set_option('nch', nch)
        out[f'split{split}/nch{nch}'] = round(g.timeit(lambda: c.render()), 1)
print(json.dumps(out))
PY
done
