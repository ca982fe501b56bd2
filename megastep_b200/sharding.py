"""Multi-GPU: environments shard embarrassingly.

Every kernel indexes by environment only (reference: kernels.cu:185,298,330,412), so a batch of N environments is
split into contiguous ranges, one per GPU, one process per GPU (torch.distributed). Each rank owns its slice of
every ragged array, its own agents and its own parameters; a step involves NO communication. The only collective
is optional: all-gathering the observation tensors over NCCL (NVLink 5 / NVSwitch) when the caller wants the whole
batch on every device, issued on a side stream so it overlaps the next tick.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_envs, rank, world_size):
    """Contiguous [lo, hi) of the envs owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_arrays(arrays, lo, hi):
    """The envs [lo, hi) of a `scene.scene_arrays` dict (numpy, host side)."""
    lw, iw, tw = arrays['line_widths'], arrays['light_widths'], arrays['tex_widths']
    ls = np.concatenate([[0], np.cumsum(lw)])
    is_ = np.concatenate([[0], np.cumsum(iw)])
    ts = np.concatenate([[0], np.cumsum(tw.astype(np.int64))])
    l0, l1 = ls[lo], ls[hi]
    out = dict(n_agents=arrays['n_agents'], model=arrays['model'],
               lines=arrays['lines'][l0:l1], line_widths=lw[lo:hi],
               lights=arrays['lights'][is_[lo]:is_[hi]], light_widths=iw[lo:hi],
               textures=arrays['textures'][ts[l0]:ts[l1]], tex_widths=tw[l0:l1])
    if 'baked' in arrays:
        out['baked'] = arrays['baked'][ts[l0]:ts[l1]]
    return out


def shard_scenery(scenery, lo, hi):
    """The envs [lo, hi) of a device-resident `cuda.Scenery`, as views (no copy) — reference Ragged slicing
    (common.h:136-144) applied at env granularity, textures/baked at the matching line range."""
    from . import cuda
    l0, l1 = int(scenery.lines.starts[lo].item()), int(scenery.lines.ends[hi - 1].item())
    s = cuda.Scenery(n_agents=scenery.n_agents, lights=scenery.lights[lo:hi], lines=scenery.lines[lo:hi],
                     textures=scenery.textures[l0:l1], model=scenery.model)
    s.baked.vals.copy_(scenery.baked[l0:l1].vals)
    return s


class ObsGather:
    """All-gathers equally-shaped per-rank observation tensors into one (world_size * n_local, ...) batch.

        g = ObsGather(example_local_obs)        # once; allocates the gathered buffers
        g.start(obs)                            # enqueue on a side stream (CUDA) — overlaps the next step
        full = g.wait()                         # arrdict of gathered tensors, valid on the current stream
    """

    def __init__(self, example, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.keys = list(example.keys())
        self.out = {k: v.new_empty((self.world * v.shape[0], *v.shape[1:])) for k, v in example.items()}
        self.cuda = next(iter(example.values())).is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self._work = []

    def start(self, obs):
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for k in self.keys:
                    obs[k].record_stream(self.stream)
                    dist.all_gather_into_tensor(self.out[k], obs[k].contiguous(), group=self.group)
        else:
            for k in self.keys:
                chunks = list(self.out[k].chunk(self.world, 0))
                self._work.append(dist.all_gather(chunks, obs[k].contiguous(), group=self.group, async_op=True))

    def wait(self):
        from .arrdict import arrdict
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        for w in self._work:
            w.wait()
        self._work = []
        return arrdict(self.out)


def gathered_bytes(example, world_size):
    """Bytes each rank receives per gather (cost model: / ~725 GB/s measured all-gather bus bandwidth)."""
    return sum(v.numel() * v.element_size() for v in example.values()) * (world_size - 1)
