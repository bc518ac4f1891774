"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups run the
resampling plan + all-gather of totals + all-to-all exchange with a NumPy stand-in for the device
stages, and must reproduce the single-process oracle's global systematic resampling exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcmh_localization_b200.sharded import count_thresholds_le, exchange, plan_resample, resample_threshold


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_problem(n_global, seed):
    rs = np.random.RandomState(seed)
    w = (rs.uniform(0, 1, n_global) ** 5).astype(np.float32)
    w[rs.randint(0, n_global, n_global // 7)] = 0.0
    parts = rs.uniform(-5, 5, (n_global, 3))
    r = rs.uniform(0, 1.0 / n_global)
    return w, parts, r


def _worker(rank, world, port, n_global, seed, zero_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import clib
        w, parts, r = _global_problem(n_global, seed)
        n = n_global // world
        if zero_rank is not None:
            w[zero_rank * n:(zero_rank + 1) * n] = 0.0
        wl, pl = w[rank * n:(rank + 1) * n], parts[rank * n:(rank + 1) * n]
        # stage 1: global max -> scale   (device: mcl_weights_max + all-reduce MAX)
        wmax = torch.tensor([float(wl.max())], dtype=torch.float32)
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        scale = clib.resample_scale(float(wmax.item()), n_global)
        # stage 2: local fixed-point cumulative sums (device: mcl_resample_scan)
        q_ = np.floor(wl.astype(np.float64) * scale).astype(np.uint64)
        cl = np.cumsum(q_, dtype=np.uint64)
        total = torch.tensor([int(cl[-1])], dtype=torch.int64)
        totals = torch.zeros(world, dtype=torch.int64)
        dist.all_gather_into_tensor(totals, total)
        offsets, grand, m_lo, m_hi, send = plan_resample(totals.tolist(), r, n_global, world)
        # stage 3: local search (device: mcl_resample_search) + gather
        cnt = m_hi[rank] - m_lo[rank]
        idx = np.empty(cnt, np.int64)
        for j in range(cnt):
            T = resample_threshold(m_lo[rank] + j, r, n_global, grand)
            idx[j] = min(int(np.searchsorted(cl + np.uint64(offsets[rank]), np.uint64(T), side="left")), n - 1)
        sendbuf = torch.from_numpy(np.ascontiguousarray(pl[idx]))
        out = exchange(sendbuf, send[rank], [send[j][rank] for j in range(world)]).numpy()
        # expected: the single-process oracle on the global arrays
        ref_idx = clib.systematic_resample_q(w, n_global, r, scale)
        exp = parts[ref_idx][rank * n:(rank + 1) * n]
        q.put((rank, bool(np.array_equal(out, exp)), out.shape, int(cnt)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global,zero_rank", [(2, 4000, None), (2, 3000, 0), (2, 3000, 1), (3, 3000, 1)])
def test_global_resampling_sharded_equals_single_process(world, n_global, zero_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, 11 + world, zero_rank, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape, cnt in res:
        assert ok, (rank, shape, cnt)
        assert shape == (n_global // world, 3)
    assert sum(c for _, _, _, c in res) == n_global


def test_plan_partitions_outputs_and_matches_thresholds():
    rs = np.random.RandomState(0)
    for world in (1, 2, 4, 8):
        totals = [int(t) for t in rs.randint(0, 2**40, world)]
        totals[rs.randint(world)] = 0
        n_out = 8 * 1000
        r = rs.uniform(0, 1.0 / n_out)
        offsets, grand, m_lo, m_hi, send = plan_resample(totals, r, n_out, world)
        assert m_lo[0] == 0 and m_hi[-1] == n_out
        assert all(m_hi[k] == m_lo[k + 1] for k in range(world - 1))
        assert all(sum(send[k]) == m_hi[k] - m_lo[k] for k in range(world))
        assert all(sum(send[k][d] for k in range(world)) == n_out // world for d in range(world))
        for k in range(world):
            for m in (m_lo[k], m_hi[k] - 1):
                if m_lo[k] < m_hi[k] and 0 < k < world - 1:
                    T = resample_threshold(m, r, n_out, grand)
                    assert offsets[k] < T <= offsets[k] + totals[k]
        assert count_thresholds_le(-1, r, n_out, grand) == 0
        assert count_thresholds_le(grand * 2, r, n_out, grand) == n_out
