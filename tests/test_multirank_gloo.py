"""CPU test (no GPU) of the N > 1 path: world_size-2 gloo run of the sharding + statistics all-reduce logic.

Each rank integrates its interleaved shard (with the CPU oracle standing in for the kernel -- this is a test), forms the
per-t_eval sums, all-reduces them through the package's own wrapper, and the result must equal the single-process
statistics of the whole ensemble; the shards together must reproduce the unsharded per-trajectory results bit for bit."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, N_EVAL = 96, 8


def sums_numpy(y_eval, n_emitted):
    mask = np.arange(y_eval.shape[1])[None, :] < n_emitted[:, None]
    ye = np.where(mask[:, :, None], y_eval, 0.0)
    return np.stack([ye.sum(0), (ye * ye).sum(0)], axis=-1), mask.sum(0).astype(np.int64)


def solve_shard(deb, ob, idx):
    y0 = deb.perturbed_ensemble([1.0, 1.0, 1.0], idx, seed=2026)
    ivp = (deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 4.0, y0).t_eval(np.linspace(0.5, 4.0, N_EVAL))
           .method(deb.ExplicitRungeKutta.dopri5().rtol(1e-8).max_steps(211)))
    return ob.oracle_solve(ivp, n_threads=1)


def worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    deb = importlib.import_module("differential-equations_b200")
    import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = deb.shard_indices(N_TOTAL, rank, world)
    s = solve_shard(deb, ob, idx)
    sums, counts = sums_numpy(s.y_eval, s.n_emitted)
    t_sums, t_counts = torch.from_numpy(sums.copy()), torch.from_numpy(counts.copy())
    deb.allreduce_ensemble_stats(t_sums, t_counts, dist)
    totals = torch.tensor([int(s.accepted.sum()), int(s.rejected.sum())])
    dist.all_reduce(totals)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, y_final=s.y_final, accepted=s.accepted, status=s.status,
             sums=t_sums.numpy(), counts=t_counts.numpy(), totals=totals.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    deb = importlib.import_module("differential-equations_b200")
    import oracle_binding as ob
    ob.load_oracle()  # build once, before the workers race for it
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole = solve_shard(deb, ob, np.arange(N_TOTAL))
    sums, counts = sums_numpy(whole.y_eval, whole.n_emitted)
    assert (whole.status != 0).any() and (whole.status == 0).any()  # ragged rows are part of the test
    r = [np.load(os.path.join(tmp_path, f"rank{k}.npz")) for k in range(2)]
    # shards are disjoint, cover everything, and reproduce the unsharded trajectories exactly
    assert sorted(np.concatenate([r[0]["idx"], r[1]["idx"]]).tolist()) == list(range(N_TOTAL))
    for k in range(2):
        assert np.array_equal(r[k]["y_final"].view(np.uint64), whole.y_final[r[k]["idx"]].view(np.uint64))
        assert np.array_equal(r[k]["accepted"], whole.accepted[r[k]["idx"]])
        assert np.array_equal(r[k]["status"], whole.status[r[k]["idx"]])
    # every rank holds the same reduced statistics, equal to the single-process ones
    assert np.array_equal(r[0]["sums"], r[1]["sums"]) and np.array_equal(r[0]["counts"], r[1]["counts"])
    assert np.array_equal(r[0]["counts"], counts)
    np.testing.assert_allclose(r[0]["sums"], sums, rtol=1e-12, atol=1e-12)
    assert r[0]["totals"].tolist() == [int(whole.accepted.sum()), int(whole.rejected.sum())]
    mean, var = deb.stats_to_mean_var(r[0]["sums"], r[0]["counts"])
    assert mean.shape == (N_EVAL, 3) and (var >= -1e-12).all()


def test_shard_indices_properties():
    deb = importlib.import_module("differential-equations_b200")
    for n, w in ((10_000_000, 8), (10, 4), (7, 8), (0, 2)):
        parts = [deb.shard_indices(n, r, w) for r in range(w)]
        assert sum(p.size for p in parts) == n and max(p.size for p in parts) - min(p.size for p in parts) <= 1
        if n:
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
    # generating a shard directly equals slicing the full ensemble
    full = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(1000))
    assert np.array_equal(deb.perturbed_ensemble([1.0, 1.0, 1.0], deb.shard_indices(1000, 3, 8)), full[3::8])
