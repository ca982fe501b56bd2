"""CPU model of view_kernel's culling (csrc/megastep_b200.cu: view_agent / cast_batch), for counting — not for results.

Replays, in float64 numpy, which run boxes a warp opens, which segments a chunk's ballot lets through and how the
chunks' farthest hits tighten, on the benchmark's synthetic floorplans, and reports per warp-item (an agent's block of
`32 * NCH` rays) the two numbers the kernel's `stats` option counts on the GPU: batches ("groups") and warp-level
candidate tests. GPU minutes are scarce; this answers "what would run length 8 / another packing / four chunks per
warp do to the work" before any CUDA is written.

    python scripts/cull_sim.py [--envs 64] [--run 16] [--runs-per-batch 2] [--nch 2] [--order str|morton]

Measured on a B200 (stats option, Deathmatch 4096 x 4 x 128, STR packing, run 16, NCH 2): 2.17 groups and 16.3 tests
per item (agents' batches included when another agent is in sight).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megastep_b200 import scene, synthetic  # noqa: E402

AGENT_RADIUS = .15 / 2 ** .5
CULL_EPS = 4e-4


def pack(walls, run, order):
    """Static segments (W, 4) -> rows in table order, (nb, 4) run boxes. 'str' as msb_build_table, 'morton' as before."""
    w32 = walls.astype(np.float32)
    mid = (w32[:, :2] + w32[:, 2:]) * np.float32(.5)          # in float32, as the table builder ranks them
    n = len(walls)
    nb = -(-n // run)
    if order == 'morton':
        cell = np.clip(((mid - mid.min(0)) / np.float32(.25)), 0, 65535).astype(np.int64)

        def spread(v):
            v = (v | (v << 8)) & 0x00FF00FF
            v = (v | (v << 4)) & 0x0F0F0F0F
            v = (v | (v << 2)) & 0x33333333
            return (v | (v << 1)) & 0x55555555
        idx = np.argsort(spread(cell[:, 0]) | (spread(cell[:, 1]) << 1), kind='stable')
    else:
        strips = int(np.ceil(np.sqrt(nb)))
        per = -(-nb // strips) * run
        ox = np.lexsort((np.arange(n), mid[:, 0]))
        rank = np.empty(n, int)
        rank[ox] = np.arange(n)
        idx = np.lexsort((np.arange(n), mid[:, 1], rank // per))
    rows = walls[idx]
    boxes = np.empty((nb, 4))
    for b in range(nb):
        s = rows[b * run:(b + 1) * run]
        boxes[b] = (min(s[:, 0].min(), s[:, 2].min()), min(s[:, 1].min(), s[:, 3].min()),
                    max(s[:, 0].max(), s[:, 2].max()), max(s[:, 1].max(), s[:, 3].max()))
    return rows, boxes


def item(rows, boxes, pos, ang, r0, nch, R, hs, run, per_batch, others, model_radius, hint=None):
    """One warp-item. Returns (batches of static rows, candidate tests on them, agent batches, the rays' hit parameters);
    the tests on the agents' own lines are not modelled."""
    cs, sn = np.cos(np.deg2rad(ang)), np.sin(np.deg2rad(ang))
    xclip = .5 * AGENT_RADIUS / np.sqrt(1 + hs * hs)
    B0, dB = (R - 2 * r0) * hs / R, 64. * hs / R
    bounds = B0 - np.arange(nch + 1) * dB                      # chunk c spans slopes (bounds[c+1], bounds[c])
    rays = r0 + np.arange(32 * nch)
    y = (R - 2 * rays - 1) * hs / R
    ru = np.stack([cs - sn * y, sn + cs * y], -1)               # (rays, 2), un-normalised
    nearp = AGENT_RADIUS / np.hypot(ru[:, 0], ru[:, 1])
    best = np.where(rays < R, np.inf, 0.)
    cmax = np.array([np.inf if r0 + 32 * c < R else 0. for c in range(nch)])
    if hint is not None:                                        # what a perfect per-ray guess from the previous frame would buy
        best = np.minimum(best, hint * (1 + 1e-3) + 1e-3)
        cmax = np.array([best[32 * c:32 * c + 32].max() for c in range(nch)])

    def camera(p):
        d = p - pos
        return d[..., 0] * cs + d[..., 1] * sn, d[..., 1] * cs - d[..., 0] * sn

    def masks(X, Y):
        """chunk mask of a convex set given its corners in camera space (last axis = corners)"""
        e = Y[..., None, :] - X[..., None, :] * bounds[:, None]            # (..., nch + 1, corners)
        left, right = (e > 0).all(-1), (e < 0).all(-1)
        return ~left[..., :-1] & ~right[..., 1:]                           # (..., nch)

    # box round
    corners = np.stack([boxes[:, [0, 1]], boxes[:, [2, 1]], boxes[:, [2, 3]], boxes[:, [0, 3]]], 1)     # (nb, 4, 2)
    X, Y = camera(corners)
    xmax, xmin = X.max(1), X.min(1)
    bcm = masks(X, Y) & ~(xmax < xclip)[:, None]
    bsmin = np.maximum(xmin - 1e-3 - 1e-4 * np.maximum(np.abs(xmin), np.abs(xmax)), 0.)
    todo = bcm.any(1)
    groups = tests = 0
    whatif = np.zeros(4)        # tests that moved some ray's hit; tests left if culled against the span's own farthest hit;
                                # iterations if candidates with disjoint ray spans shared one (= deepest overlap); chunks
    segX, segY = camera(rows.reshape(-1, 2, 2))                            # (W, 2 endpoints)
    seg_cm = masks(segX, segY) & ~((segX < xclip).all(1))[:, None]
    seg_smin = segX.min(1) - 1e-3 - 1e-4 * np.abs(segX).max(1)
    V = rows[:, 2:] - rows[:, :2]
    PQ = rows[:, :2] - pos
    snum = V[:, 1] * PQ[:, 0] - V[:, 0] * PQ[:, 1]
    while True:
        alive = todo & (bcm & ~(bsmin[:, None] > cmax + CULL_EPS)).any(1)
        if not alive.any():
            break
        pick = np.argsort(np.where(alive, bsmin, np.inf), kind='stable')[:per_batch]
        pick = pick[alive[pick]]
        todo[pick] = False
        groups += 1
        segs = np.concatenate([np.arange(b * run, min((b + 1) * run, len(rows))) for b in pick])
        for c in range(nch):
            cand = segs[seg_cm[segs, c] & ~(seg_smin[segs] > cmax[c] + CULL_EPS)]
            tests += len(cand)
            lanes = slice(32 * c, 32 * c + 32)
            depth = np.zeros(32, int)
            for j in cand:
                UxV = ru[lanes, 0] * V[j, 1] - ru[lanes, 1] * V[j, 0]
                with np.errstate(divide='ignore', invalid='ignore'):
                    s = snum[j] / UxV
                    t = (ru[lanes, 1] * PQ[j, 0] - ru[lanes, 0] * PQ[j, 1]) / UxV
                span = (np.abs(UxV) >= 1e-3) & (t >= 0) & (t <= 1)
                hit = span & (nearp[lanes] < s) & (s < best[lanes])
                whatif[0] += hit.any()
                if span.any() and not (seg_smin[j] > best[lanes][span].max() + CULL_EPS):
                    whatif[1] += 1
                    depth += span
                best[lanes] = np.where(hit, s, best[lanes])
            if len(cand):
                cmax[c] = best[lanes].max()
                whatif[2] += depth.max()
                whatif[3] += 1
    # the agents' own batch: skipped unless another agent's disc is in sight (view_agent's test)
    rho = model_radius * 1.001 + 1e-3
    Bn = bounds[-1]
    agent_groups = 0
    for o in others:
        Xo, Yo = camera(o)
        behind = Xo + rho < xclip
        left = (Yo - Xo * B0) - rho * np.sqrt(1 + B0 * B0) > 0
        right = (Yo - Xo * Bn) + rho * np.sqrt(1 + Bn * Bn) < 0
        hidden = Xo - rho - 1e-3 - 1e-4 * (abs(Xo) + rho) > cmax.max() + CULL_EPS
        if not (behind or left or right or hidden):
            agent_groups = 1
            break
    return groups, tests, agent_groups, best, whatif


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=64)
    ap.add_argument('--agents', type=int, default=4)
    ap.add_argument('--res', type=int, default=128)
    ap.add_argument('--fov', type=float, default=70.)
    ap.add_argument('--run', type=int, default=16)
    ap.add_argument('--runs-per-batch', type=int, default=2)
    ap.add_argument('--nch', type=int, default=2)
    ap.add_argument('--order', default='str', choices=['str', 'morton'])
    ap.add_argument('--hint', action='store_true', help='second pass with every ray starting from (just behind) its final hit')
    args = ap.parse_args()
    gs = synthetic.sample(args.envs, seed=1, n_unique=args.envs)
    pos, ang = synthetic.spawns(gs, args.agents, np.random.RandomState(2))
    hs = np.tan(np.deg2rad(args.fov) / 2)
    model_radius = np.abs(scene.agent_model()).reshape(-1, 2)
    model_radius = np.hypot(model_radius[:, 0], model_radius[:, 1]).max()
    tot = np.zeros(3)
    wi = np.zeros(4)
    items = 0
    for n, g in enumerate(gs):
        walls = np.asarray(g.walls, dtype=np.float64).reshape(-1, 4)
        rows, boxes = pack(walls, args.run, args.order)
        for a in range(args.agents):
            others = [pos[n, b].astype(np.float64) for b in range(args.agents) if b != a]
            for r0 in range(0, args.res, 32 * args.nch):
                common_args = (rows, boxes, pos[n, a].astype(np.float64), float(ang[n, a]), r0, args.nch, args.res, hs, args.run,
                               args.runs_per_batch, others, model_radius)
                res = item(*common_args)
                if args.hint:
                    res = item(*common_args, hint=res[3])
                tot += res[:3]
                wi += res[4]
                items += 1
    g, t, ag = tot / items
    print(f'{vars(args)}\nitems {items}: static batches {g:.2f} + agent batches {ag:.2f} = {g + ag:.2f} per item '
          f'({(g + ag) * items / (args.envs * args.agents):.2f} per agent); candidate tests {t:.1f} per item ({t * items / (args.envs * args.agents):.1f} per agent) '
          f'[static only]; segments binned {g * args.run * args.runs_per_batch:.0f} per item')
    print(f'what if, per item: tests that move some ray\'s hit {wi[0] / items:.1f}; tests left when a segment is culled against the farthest hit '
          f'of ITS OWN ray span {wi[1] / items:.1f}; iterations if span-disjoint candidates shared one {wi[2] / items:.1f} '
          f'(over {wi[3] / items:.2f} non-empty chunk rounds per item)')


if __name__ == '__main__':
    main()
