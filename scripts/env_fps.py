"""Env-level throughput (the reference's published figures — 180k FPS Explorer, 1.2M FPS Deathmatch on a
2080 Ti, docs/index.rst:13-25 — are env-level): whole `env.step(random actions)` loops of this package's envs (fused device
rules) next to the reference's own unmodified demo envs on its own CUDA build, same box, same synthetic floorplans.

    python scripts/env_fps.py explorer 4096 [steps]      python scripts/env_fps.py deathmatch 1024 [steps]
Prints one JSON line; FPS = agent-frames/s = n_envs * n_agents * steps / wall seconds (host launches included)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common  # noqa: E402


def run(env, actions, steps, make_decision):
    env.reset()
    for i in range(5):
        env.step(make_decision(actions[i]))
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(steps):
        env.step(make_decision(actions[5 + i]))
    torch.cuda.synchronize()
    return time.perf_counter() - t


def main():
    kind, n = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    from megastep_b200 import envs, synthetic
    from megastep_b200.arrdict import arrdict
    gs = synthetic.sample(n, seed=9, n_unique=256, with_masks=True)
    for g in gs:
        g['res'] = np.array(g['res'])
    A = 1 if kind == 'explorer' else 4
    out = {'env': kind, 'n_envs': n, 'n_agents': A, 'steps': steps}
    acts = torch.as_tensor(np.random.RandomState(1).randint(0, 7, (steps + 5, n * A if kind == 'deathmatch' else n, 1))).cuda()
    for fused in (True, False):
        np.random.seed(3)
        env = envs.Explorer(gs, fused=fused) if kind == 'explorer' else envs.Deathmatch(gs, A, fused=fused)
        secs = run(env, acts, steps, lambda a: arrdict(actions=a))
        out['ours_fused_fps' if fused else 'ours_unfused_fps'] = n * A * steps / secs
        del env
        torch.cuda.empty_cache()
    pkg = common.reference_package()
    if pkg is not None:
        np.random.seed(3)
        mod = pkg.explorer if kind == 'explorer' else pkg.deathmatch
        mod.cubicasa.sample = lambda k: gs[:k]
        env = mod.Explorer(n) if kind == 'explorer' else mod.Deathmatch(4 * n, A)
        rsteps = max(steps // 4, 20)
        secs = run(env, acts, rsteps, lambda a: pkg.arrdict.arrdict(actions=a))
        out['reference_fps'] = n * A * rsteps / secs
        out['speedup'] = out['ours_fused_fps'] / out['reference_fps']
    print(json.dumps(out))


if __name__ == '__main__':
    main()
