"""CPU test of the N>1 path: two gloo ranks shard a batch of fits by contiguous index ranges, each fits its own
shard, one all_gather assembles the result, and it equals the single-process result bit for bit.  On the CPU
the per-rank compute stand-in is the oracle (the product path needs a GPU); what is under test is the sharding
and gather logic of curvepointslam_b200/shard.py, which bench.py uses unchanged over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from curvepointslam_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = [1e-10, 1e-10, 1e-10, 1e-20, 1e-5]
TOTAL = 21  # not divisible by 2: ragged shards


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import DER, STEREO19, oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(TOTAL, rank, world)
    b = synth.make_stereo19_batch(hi - lo, K=32, seed=77, first=lo)  # every rank generates only its shard
    r = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    full = shard.gather_fits(torch.from_numpy(r["p"]), TOTAL, rank, world)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    root = shard.gather_rows_to_root(torch.from_numpy(r["p"]), TOTAL, rank, world, dst=0)
    assert (root is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(root, full)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_rank_gloo_shard_and_gather(tmp_path):
    from oracle_lib import DER, STEREO19, oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b = synth.make_stereo19_batch(TOTAL, K=32, seed=77, first=0)
    ref = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)["p"]
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npy"))
        assert got.shape == ref.shape and np.array_equal(got.view(np.int64), ref.view(np.int64))


def _gate_problem():
    from curvepointslam_b200 import ekf
    x, G, ci, pi = ekf.make_synthetic_ekf(60, d=3, seed=5)
    N = x.size
    P = G @ G.T / 64.0 * 0.05 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    rng = np.random.default_rng(9)
    lm_xyz = x[pi[0]: pi[0] + 3 * len(pi)].reshape(-1, 3)
    z = lm_xyz[rng.integers(0, len(pi), 37)] - x[:3] + rng.normal(0, 3.0, (37, 3))
    return x, P, pi, z


def _gate_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, P, pi, z = _gate_problem()
    lo, hi = shard.shard_range(len(z), rank, world)
    idx, _ = oracle().gate(x, P, pi, z[lo:hi], 11.345)
    full = shard.gather_gate(torch.from_numpy(idx), len(z), rank, world)
    np.save(os.path.join(out_dir, f"gate{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gate_sharded_by_measurement(tmp_path):
    """The gate shards by measurement with the landmarks replicated: two ranks gate ragged halves and all-gather the
    indices; every rank ends up with the single-process association."""
    from oracle_lib import oracle
    port = _free_port()
    mp.spawn(_gate_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    x, P, pi, z = _gate_problem()
    ref, _ = oracle().gate(x, P, pi, z, 11.345)
    assert (ref >= 0).any()
    for rank in range(2):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), f"gate{rank}.npy")), ref)
