#!/bin/bash
# A/B timing of library builds (Deathmatch 4096x4x128): main render kernel alone (agent-hit lighting skipped), render, step
for lib in build_variants/*.so; do
    echo -n "$lib: "
    MEGASTEP_B200_LIB=$PWD/$lib timeout 300 python - <<PY 2>&1 | tail -1
import sys; sys.argv=['x','none','16']
sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import json, torch
import gpu_exp as g
from megastep_b200 import cuda, modules
c = g.setup()
out = {}
cuda.set_option('debug_skip_dyn', 1); out['main_us'] = round(g.timeit(lambda: c.render()), 1); cuda.set_option('debug_skip_dyn', 0)
out['render_us'] = round(g.timeit(lambda: c.render()), 1)
step = modules.FusedStep(c, subsample=1, raw=True)
acts = torch.randint(0, 7, (4096, 4), dtype=torch.int32, device='cuda')
out['step_us'] = round(g.timeit(lambda: step(acts)), 1)
print(json.dumps(out))
PY
done
