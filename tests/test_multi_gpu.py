"""N > 1 host logic on CPU (gloo, world size 2): shard ranges, globally keyed initial conditions and the all-reduce
of the statistics vector reproduce the single-process job.  The per-shard filter work is done by the batch oracle
(test infrastructure) -- the GPU path itself is covered by tests/test_gpu_noise.py::test_noise_free_filter0_and_sharding."""
import os
import socket

import numpy as np
import pytest

from dvi_ekf_b200.sharding import NSTAT, allreduce_stats, mc_initial_states, shard_of, summarise_stats

GT = np.array([0, 0, 0, 0, 0, 20.0])


def _shard_stats(first, count, n_frames=4):
    """statistics vector of filters first .. first+count-1 (layout of include/eskf.h) from the batch oracle"""
    from oracle.batch_oracle import BatchOracle
    from tests.helpers import mandala_scenario

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sc = mandala_scenario(np.load(os.path.join(root, "tests", "golden", "reference_golden.npz")), n_frames=n_frames, ifv=2,
                          frozen_dofs=(0, 0, 0, 0, 0, 0))
    x0 = mc_initial_states(sc.x0, count, first, seed=7)
    bo = BatchOracle(sc.cfg, x0, sc.P0, sc.u0)
    bo.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas)
    s = np.zeros(NSTAT)
    d2 = (bo.x[:, 10:16] - GT) ** 2
    s[:6] = d2.sum(0)
    s[6] = (d2.sum(1) / 6.0).sum()
    s[9] = count * len(sc.n_prop)
    s[11] = count
    return s, bo.x


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_of(n_total, rank, world)
    s, x = _shard_stats(first, count)
    t = torch.from_numpy(s.copy())
    allreduce_stats(t)
    q.put((rank, first, count, t.numpy().copy(), x))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_of_covers_the_batch():
    for n, w in ((4096, 1), (4096, 8), (10, 4), (3, 8), (1048576, 8)):
        blocks = [shard_of(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_of(8, 2, 2)


def test_initial_states_are_keyed_by_the_global_id():
    x0 = np.arange(26.0)
    full = mc_initial_states(x0, 12, 0, seed=5)
    assert np.array_equal(full[0], x0)  # global filter 0 is the nominal (noise-free) run
    assert np.array_equal(mc_initial_states(x0, 5, 7, seed=5), full[7:12])
    assert not np.array_equal(mc_initial_states(x0, 5, 7, seed=6), full[7:12])


def test_world_size_2_gloo_reproduces_the_single_process_statistics():
    import torch.multiprocessing as mp

    n_total, world = 9, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, xref = _shard_stats(0, n_total)
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 4)]
    for r in res:
        assert np.allclose(r[3], ref, rtol=1e-13, atol=0)  # every rank holds the reduced vector
    assert np.array_equal(np.vstack([r[4] for r in res]), xref)  # sharding does not change any filter
    out = summarise_stats(res[0][3])
    assert out["filters"] == n_total and len(out["dof_rmse"]) == 6
    assert np.isclose(out["dof_metric_mean"], ref[6] / n_total)


def test_bench_secondary_plan_is_valid_on_every_rank():
    """bench.py's `configs` block (BASELINE configs 3 / 4 / 5 + calibration) at 1 / 2 / 4 / 8 GPUs: the launch geometry of EVERY
    rank passes the argument checks of eskf_run (mirrored by engine.check_run_geometry) and the shards add up to the whole job.
    (A plan valid on rank 0 only leaves the other ranks of a torchrun job waiting in the all-reduce: round 2 lost ten minutes
    of a 2-GPU box to exactly that.)"""
    import bench
    from dvi_ekf_b200.engine import check_run_geometry

    for world in (1, 2, 4, 8):
        plans = [bench.secondary_plan(r, world, 4096) for r in range(world)]
        for name in ("config3", "config4", "config5", "calibration"):
            assert sum(p[name]["n"] for p in plans) == plans[0][name]["total"]
            for p in plans:
                g = p[name]
                check_run_geometry(g["n"], g["n_traj"], g["filters_per_traj"], g["filter_id0"], 1390, 139)
        for name in ("config3", "config5", "calibration"):  # contiguous global ids
            ids = [(p[name]["filter_id0"], p[name]["n"]) for p in plans]
            assert ids[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(ids, ids[1:]))
        assert len({p["config4"]["seed_offset"] for p in plans}) == world  # independent noise per rank


def test_run_geometry_checks():
    from dvi_ekf_b200.engine import check_run_geometry

    check_run_geometry(40, 2, 20, 0, 110, 11, np.full(22, 10, dtype=np.int32))
    check_run_geometry(20, 2, 20, 20, 110, 11)  # the second trajectory's shard
    with pytest.raises(ValueError, match="exceeds n_traj"):
        check_run_geometry(40, 2, 10, 0, 110, 11)
    with pytest.raises(ValueError, match="exceeds n_traj"):
        check_run_geometry(4608, 9, 512, 4608, 6567, 199)  # the round-2 bug: a global id offset on local streams
    with pytest.raises(ValueError, match="multiple of filters_per_traj"):
        check_run_geometry(40, 2, 32, 8, 110, 11)
    with pytest.raises(ValueError, match="exceeds n_steps"):
        check_run_geometry(4, 1, 4, 0, 110, 11, np.array([10] * 10 + [15], dtype=np.int32))
