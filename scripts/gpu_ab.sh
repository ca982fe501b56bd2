#!/bin/bash
# A/B timing of library builds x runtime variants (Deathmatch 4096x4x128 render, skip_dyn isolates the main kernel)
for lib in build_variants/lib_mb3.so build_variants/lib_mb4.so; do
  for v in 0 1 2 3; do
    echo -n "$lib variant $v: "
    MEGASTEP_B200_LIB=$PWD/$lib timeout 200 python - <<PY 2>&1 | tail -1
import sys; sys.argv=['x','none','16','$v']
sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import json, torch
import gpu_exp as g
from megastep_b200 import cuda
c = g.setup()
cuda.set_option('variant', $v)
out = {}
cuda.set_option('debug_skip_dyn', 1); out['main_only_us'] = round(g.timeit(lambda: c.render()), 1)
cuda.set_option('debug_skip_dyn', 0); out['render_us'] = round(g.timeit(lambda: c.render()), 1)
print(json.dumps(out))
PY
  done
done
