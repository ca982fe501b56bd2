"""Which order does ATen's mean() add the last `sub` values in? (A probe for pool_sum in csrc/megastep_b200.cu.)"""
import itertools, json, sys
import torch
torch.manual_seed(0)
out = {}
for sub in (2, 4, 8, 16):
    x = torch.rand(4096, 4, 3, 1, 128 // 1, device='cuda')           # like screen (N, A, 3, 1, R)
    v = x.view(*x.shape[:-1], x.shape[-1] // sub, sub)
    m = v.mean(-1)
    c = [v[..., i] for i in range(sub)]
    def tree(vals):      # neighbours first
        while len(vals) > 1: vals = [vals[i] + vals[i + 1] for i in range(0, len(vals), 2)]
        return vals[0]
    def seq(vals):
        a = vals[0]
        for b in vals[1:]: a = a + b
        return a
    def strided(vals, k):   # k interleaved partial sums, then combined
        parts = [seq(vals[i::k]) for i in range(k)]
        return parts
    cands = {'sequential': seq(c), 'tree': tree(c), 'reverse': seq(c[::-1])}
    for k in (2, 4):
        if sub > k:
            p = strided(c, k)
            cands[f'strided{k}_seq'] = seq(p); cands[f'strided{k}_tree'] = tree(p)
    if sub == 4:
        cands['(02)(13)'] = (c[0] + c[2]) + (c[1] + c[3]); cands['(03)(12)'] = (c[0] + c[3]) + (c[1] + c[2])
        cands['0+(1+(2+3))'] = c[0] + (c[1] + (c[2] + c[3])); cands['(0+(1+2))+3'] = (c[0] + (c[1] + c[2])) + c[3]
    res = {}
    for name, s in cands.items():
        res[name] = float(((s * (1.0 / sub)) == m).float().mean()); res[name + '/div'] = float(((s / sub) == m).float().mean())
    out[sub] = {k: round(v, 4) for k, v in sorted(res.items(), key=lambda kv: -kv[1])[:6]}
print(json.dumps(out, indent=1))
