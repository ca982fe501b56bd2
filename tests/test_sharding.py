"""The multi-GPU host logic on CPU: env partitioning, scene slicing, and the observation all-gather over a
world_size-2 gloo group (the N>1 path of bench.py uses the same code with NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from megastep_b200 import scene, sharding, synthetic
from megastep_b200.arrdict import arrdict


def test_shard_ranges_partition_the_envs():
    for n, w in [(4096, 8), (10, 3), (7, 8), (65536, 8)]:
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_arrays_slices_every_ragged_consistently():
    gs = synthetic.sample(7, seed=3)
    arrays = scene.scene_arrays(gs, 2, np.random.RandomState(0))
    arrays['baked'] = np.arange(len(arrays['textures']), dtype=np.float32)
    parts = [sharding.shard_arrays(arrays, *sharding.shard_range(7, r, 3)) for r in range(3)]
    for k in ('lines', 'line_widths', 'lights', 'light_widths', 'textures', 'tex_widths', 'baked'):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts]), arrays[k])
    for p in parts:
        assert p['line_widths'].sum() == len(p['lines']) == len(p['tex_widths'])
        assert p['tex_widths'].sum() == len(p['textures']) == len(p['baked'])
        assert p['light_widths'].sum() == len(p['lights'])


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(6, rank, world)
        # each rank "observes" its own envs: value = global env id
        ids = torch.arange(lo, hi, dtype=torch.float32)
        obs = arrdict(rgb=ids[:, None, None].expand(-1, 2, 3).contiguous(), imu=ids[:, None].repeat(1, 4))
        g = sharding.ObsGather(obs)
        g.start(obs)
        full = g.wait()
        ok = bool((full.rgb[:, 0, 0] == torch.arange(6.)).all() and (full.imu[:, 3] == torch.arange(6.)).all()
                  and full.rgb.shape == (6, 2, 3))
        # a second round reuses the buffers
        g.start(arrdict(rgb=obs.rgb + 10, imu=obs.imu + 10))
        full = g.wait()
        ok = ok and bool((full.imu[:, 0] == torch.arange(6.) + 10).all())
        q.put((rank, ok, sharding.gathered_bytes(obs, world)))
    finally:
        dist.destroy_process_group()


def test_obs_all_gather_over_gloo_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[:2] for r in results] == [(0, True), (1, True)]
    assert results[0][2] == (3 * 2 * 3 + 3 * 4) * 4
