"""CPU-only, world_size 2 over gloo: the host-side logic of the row-partitioned solve (row offsets, global value range,
the two MAX all-reduces per round, SUM termination test) with a numpy model engine in place of the CUDA shard."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, seed, n, m, k, maximize, out_dir, exchange):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from helpers import random_sparse_instance
    from partition_model import ModelShardEngine
    from sparse_linear_assignment_b200.distributed import PartitionedKhoslaSolver, shard_rows
    rng = np.random.default_rng(seed)
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=200)
    if maximize:
        pass
    begin, count = shard_rows(n, world, rank)
    a, b = int(rp[begin]), int(rp[begin + count])
    eng = ModelShardEngine(O, m, rp[begin:begin + count + 1].astype(np.int64) - a, c[a:b], v[a:b])
    res = PartitionedKhoslaSolver(eng, exchange=exchange).solve(maximize=maximize, eps=1.0 / (m + 1))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), p2o=res["p2o"], o2p=res["o2p"], prices=res["prices"],
             row_begin=res["stats"]["row_begin"], rounds=res["stats"]["rounds"], bids=res["stats"]["global_bids"],
             arcs=res["stats"]["global_bid_arcs"], unassigned=res["stats"]["global_num_unassigned"])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed,maximize,exchange", [(2, 0, False, "dense"), (2, 1, True, "dense"), (3, 2, False, "dense"),
                                                          (2, 3, False, "sparse"), (3, 4, True, "sparse")])
def test_partitioned_driver_equals_single_instance_model(oracle, tmp_path, world, seed, maximize, exchange):
    from helpers import random_sparse_instance
    n, m, k = 61, 90, 5
    port = 29600 + seed
    mp.spawn(_worker, args=(world, port, seed, n, m, k, maximize, str(tmp_path), exchange), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=200)
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, maximize=maximize, eps=1.0 / (m + 1))
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    p2o = np.concatenate([p["p2o"] for p in parts])
    assert np.array_equal(p2o, ref["p2o"])
    for p in parts:                                  # replicas are identical on every rank
        assert np.array_equal(p["o2p"], ref["o2p"])
        assert np.array_equal(p["prices"], ref["prices"])
        assert int(p["rounds"]) == ref["stats"]["rounds"]
        assert int(p["bids"]) == ref["stats"]["bids"] and int(p["arcs"]) == ref["stats"]["bid_arcs"]
        assert int(p["unassigned"]) == ref["stats"]["num_unassigned"]
    assert [int(p["row_begin"]) for p in parts] == [sum(len(q["p2o"]) for q in parts[:r]) for r in range(world)]


def test_shard_rows_partition():
    from sparse_linear_assignment_b200.distributed import shard_rows
    for n, w in ((10, 3), (16_000_000, 8), (5, 8), (8192, 8)):
        spans = [shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(w - 1))
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


# ---- mesh engine: MeshKhoslaSolver's host side and the protocol, with a numpy model of the shard ------------------------
def _mesh_worker(rank, world, port, seed, n, m, k, maximize, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from functools import partial
    from oracle import oracle as O
    from helpers import random_sparse_instance
    from mesh_model import ModelMeshShard, _HostCsr
    from sparse_linear_assignment_b200.distributed import MeshKhoslaSolver, shard_rows
    rng = np.random.default_rng(seed)
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=200)
    begin, count = shard_rows(n, world, rank)
    a, b = int(rp[begin]), int(rp[begin + count])
    csr = _HostCsr(m, rp[begin:begin + count + 1].astype(np.int64) - a, c[a:b], v[a:b])
    mesh = MeshKhoslaSolver(csr, shard_factory=partial(ModelMeshShard, oracle=O)).setup()
    res = None
    for _ in range(2):                                   # a second solve on the same shards
        res = mesh.solve(maximize=maximize, eps=1.0 / (m + 1), gather=True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), p2o=res["p2o"], gp2o=res["global_p2o"], go2p=res["global_o2p"],
             gprices=res["global_prices"], row_begin=res["row_begin"], rounds=res["stats"]["rounds"],
             bids=res["stats"]["global_bids"], arcs=res["stats"]["global_bid_arcs"], unassigned=res["stats"]["global_num_unassigned"],
             first_object=res["owned"]["first_object"], num_owned=res["owned"]["num_owned"])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed,maximize", [(2, 10, False), (3, 11, True), (4, 12, False)])
def test_mesh_driver_equals_single_instance_model(oracle, tmp_path, world, seed, maximize):
    from helpers import random_sparse_instance
    from sparse_linear_assignment_b200.distributed import object_shard
    n, m, k = 61, 90, 5
    mp.spawn(_mesh_worker, args=(world, 29700 + seed, seed, n, m, k, maximize, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=200)
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, maximize=maximize, eps=1.0 / (m + 1), khosla_scaling=False)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert np.array_equal(np.concatenate([p["p2o"] for p in parts]), ref["p2o"])
    shard = object_shard(m, world)
    for r, p in enumerate(parts):
        assert np.array_equal(p["gp2o"], ref["p2o"]) and np.array_equal(p["go2p"], ref["o2p"])
        assert np.array_equal(p["gprices"], ref["prices"])
        assert int(p["rounds"]) == ref["stats"]["rounds"]
        assert int(p["bids"]) == ref["stats"]["bids"] and int(p["arcs"]) == ref["stats"]["bid_arcs"]
        assert int(p["unassigned"]) == ref["stats"]["num_unassigned"]
        assert int(p["first_object"]) == min(r * shard, m) and int(p["num_owned"]) == max(0, min(shard, m - r * shard))
    assert [int(p["row_begin"]) for p in parts] == [sum(len(q["p2o"]) for q in parts[:r]) for r in range(world)]


def test_object_shard_is_a_power_of_two_cover():
    from sparse_linear_assignment_b200.distributed import object_shard
    for m, w in ((64_000_000, 8), (64_000_000, 2), (7000, 3), (5, 8), (1, 1)):
        s = object_shard(m, w)
        assert s & (s - 1) == 0 and s * w >= m and (s == 1 or (s // 2) * w < m)
