"""The multi-GPU host logic on CPU: env partitioning, scene slicing, and the observation all-gather over a
world_size-2 gloo group (the N>1 path of bench.py uses the same code with NCCL)."""
import os
import socket

import numpy as np
import pytest
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


def _row_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        A, ro, n_local = 2, 4, 3
        packing = sharding.PackedObs(A, ro)
        g = sharding.RowGather(packing, n_local, 'cpu')
        ok = True
        for tick in range(3):                                  # three rounds: both buffers get reused
            base = 100. * tick + 10. * rank
            obs = arrdict(rgb=base + torch.arange(n_local * A * 3 * ro, dtype=torch.float32).reshape(n_local, A, 3, 1, ro),
                          d=base + 1000 + torch.arange(n_local * A * ro, dtype=torch.float32).reshape(n_local, A, 1, 1, ro),
                          imu=base + 2000 + torch.arange(n_local * A * 3, dtype=torch.float32).reshape(n_local, A, 3))
            g.start(obs)
            full = g.wait()
            ok = ok and full.rgb.shape == (world * n_local, A, 3, 1, ro) and full.imu.shape == (world * n_local, A, 3)
            for r in range(world):
                off = 100. * tick + 10. * r
                sl = slice(r * n_local, (r + 1) * n_local)
                ok = ok and bool((full.rgb[sl] == obs.rgb - base + off).all() and (full.d[sl] == obs.d - base + off).all()
                                 and (full.imu[sl] == obs.imu - base + off).all())
        q.put((rank, ok, g.bytes_received()))
    finally:
        dist.destroy_process_group()


def test_packed_row_gather_over_gloo_world_size_2():
    """ShardedCore's collective — per-env packed rows, one all-gather, strided views of the result — on CPU."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[:2] for r in results] == [(0, True), (1, True)]
    assert results[0][2] == 3 * (2 * 4 * 4 + 2 * 3) * 4


def test_packed_obs_round_trip_and_half_precision():
    p = sharding.PackedObs(4, 16)
    obs = arrdict(rgb=torch.rand(5, 4, 3, 1, 16), d=torch.rand(5, 4, 1, 1, 16), imu=torch.randn(5, 4, 3))
    v = p.views(p.pack(obs, p.empty(5, 'cpu')))
    assert all(torch.equal(v[k], obs[k]) for k in obs)
    h = sharding.PackedObs(4, 16, torch.float16)
    vh = h.views(h.pack(obs, h.empty(5, 'cpu')))
    assert vh.rgb.dtype == torch.float16 and all(torch.allclose(vh[k].float(), obs[k], atol=2e-3) for k in obs)
    for A, ro in ((4, 16), (1, 5), (3, 7)):                      # 8-bit images, fp16 imu; odd widths exercise the alignment padding
        obs = arrdict(rgb=torch.rand(5, A, 3, 1, ro), d=torch.rand(5, A, 1, 1, ro), imu=torch.randn(5, A, 3))
        q = sharding.PackedObs(A, ro, torch.uint8)
        vq = q.views(q.pack(obs, q.empty(5, 'cpu')))
        assert vq.rgb.dtype == torch.uint8 and vq.imu.dtype == torch.float16 and q.width % 2 == 0
        assert (vq.rgb.float() / 255 - obs.rgb).abs().max() <= .5 / 255 + 1e-6 and (vq.d.float() / 255 - obs.d).abs().max() <= .5 / 255 + 1e-6
        assert torch.allclose(vq.imu.float(), obs.imu, atol=4e-3)


def _scene_for_shards(n_envs, n_agents, seed=7):
    gs = synthetic.sample(n_envs, seed=seed)
    arrays = scene.scene_arrays(gs, n_agents, np.random.RandomState(seed))
    pos, ang = synthetic.spawns(gs, n_agents, np.random.RandomState(seed + 1))
    return arrays, pos, ang


@pytest.mark.gpu
def test_sharded_core_on_one_rank_equals_the_plain_core():
    from megastep_b200 import core as core_, cuda, modules
    N, A = 8, 4
    arrays, pos, ang = _scene_for_shards(N, A)
    sc = sharding.ShardedCore(arrays, N, res=128, fov=70., subsample=2, positions=pos, angles=ang)
    s = scene.upload(arrays)
    cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 128, 70., 10.))
    c = core_.Core(s, res=128, fov=70., fps=10.)
    c.agents.positions.copy_(torch.as_tensor(pos))
    c.agents.angles.copy_(torch.as_tensor(ang))
    plain = modules.FusedStep(c, subsample=2)
    rng = np.random.RandomState(0)
    for tick in range(4):
        actions = torch.as_tensor(rng.randint(0, 7, (N, A)).astype(np.int32)).cuda()
        a, b = sc.step(actions), plain(actions)
        full = sc.gather()
        torch.cuda.synchronize()
        for k in ('rgb', 'd', 'imu'):
            assert torch.equal(a.obs[k], b.obs[k]) and torch.equal(full[k], b.obs[k]), (tick, k)


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from megastep_b200 import core as core_, cuda, modules
        N, A = 12, 4
        arrays, pos, ang = _scene_for_shards(N, A)
        ok = True
        for dtype, tol, transport in ((torch.float32, 0., 'nccl'), (torch.float16, 2e-3, 'nccl'), (torch.uint8, 2.1e-3, 'nccl'),
                                      (torch.float32, 0., 'p2p'), (torch.uint8, 2.1e-3, 'p2p')):
            sc = sharding.ShardedCore(arrays, N, res=128, fov=70., subsample=1, obs_dtype=dtype, positions=pos, angles=ang, transport=transport)
            # the same batch, whole, on this rank's GPU
            s = scene.upload(arrays)
            cuda.bake(s, params=cuda.make_params(common.AGENT_RADIUS, 128, 70., 10.))
            c = core_.Core(s, res=128, fov=70., fps=10.)
            c.agents.positions.copy_(torch.as_tensor(pos))
            c.agents.angles.copy_(torch.as_tensor(ang))
            whole = modules.FusedStep(c, subsample=1)
            rng = np.random.RandomState(0)
            for tick in range(5):
                actions = torch.as_tensor(rng.randint(0, 7, (N, A)).astype(np.int32)).cuda()
                sc.step(actions[sc.lo:sc.hi])
                sc.gather_start()                                  # overlaps the reference computation below
                want = whole(actions)
                full = sc.gather_wait()
                torch.cuda.synchronize()
                for k in ('rgb', 'd', 'imu'):
                    got = full[k].float() / 255 if full[k].dtype == torch.uint8 else full[k].float()
                    good = torch.equal(full[k], want.obs[k]) if tol == 0. else torch.allclose(got, want.obs[k], atol=tol if k != 'imu' else 4e-3)
                    ok = ok and bool(good) and full[k].shape == want.obs[k].shape
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_core_two_ranks_over_nccl():
    """Two processes, two GPUs: each steps half of the batch, the all-gather gives both the whole batch's observations,
    bit-identical (fp32) to stepping the whole batch on one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]


@pytest.mark.gpu
def test_pack_obs_kernel_equals_the_torch_packing():
    """msb_pack_obs (one launch) against PackedObs.pack's PyTorch form, for the three row formats and an odd width."""
    torch.manual_seed(0)
    for A, ro in ((4, 128), (3, 7), (1, 5)):
        obs = arrdict(rgb=torch.rand(33, A, 3, 1, ro) * 1.2 - .1, d=torch.rand(33, A, 1, 1, ro), imu=torch.randn(33, A, 3))
        for dtype in (torch.float32, torch.float16, torch.uint8):
            p = sharding.PackedObs(A, ro, dtype)
            want = p.pack(obs, p.empty(33, 'cpu'))
            got = p.pack(arrdict({k: v.cuda() for k, v in obs.items()}), torch.zeros((33, p.width), dtype=dtype, device='cuda'))
            v, w = p.views(got.cpu()), p.views(want)
            for k in ('rgb', 'd', 'imu'):
                assert torch.equal(v[k], w[k]), (A, ro, dtype, k)
