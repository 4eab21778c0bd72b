"""torchrun target: one KhoslaSolver instance row-partitioned over the ranks (NCCL), generated shard by shard in HBM.
    torchrun --nproc-per-node N scripts/run_partitioned.py ROWS COLS K [dense|sparse] [check]
Rank 0 prints one JSON line; with `check` the same instance is also solved on rank 0 alone and compared."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import _lib
from sparse_linear_assignment_b200.distributed import CudaShardEngine, PartitionedKhoslaSolver, shard_rows

rows, cols, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
exchange = sys.argv[4] if len(sys.argv) > 4 else "dense"
check = len(sys.argv) > 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
begin, count = shard_rows(rows, world, rank)
solver, _ = S.KhoslaSolver.new(count, cols, count * k, device=local)
ctx = solver._context()
_lib.check(ctx, _lib.load().sla_generate_device_shard(ctx, rows, cols, k, 1, 300, 1000, 0, begin, count))
solver._num_rows, solver._num_cols, solver._dirty, solver._device_only = count, cols, False, True
eng = CudaShardEngine(solver)
drv = PartitionedKhoslaSolver(eng, exchange=exchange)
times = []
for rep in range(3):
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    t = time.perf_counter()
    res = drv.solve(False, None, download=(check and rep == 2))     # resident results, like solve_resident on one GPU
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t)
st = res["stats"]
out = {"exchange": exchange, "world": world, "rows": rows, "cols": cols, "k": k, "rounds": st["rounds"], "bid_arcs": st["global_bid_arcs"],
       "unassigned": st["global_num_unassigned"], "solve_ms": [round(x * 1e3, 3) for x in times],
       "note": "solve_ms[2] includes the result download when `check` is given",
       "bid_arcs_per_s": st["global_bid_arcs"] / min(times)}
if check and rank == 0:
    single, z = S.KhoslaSolver.new(rows, cols, rows * k, device=local)
    S.generators.kregular_device(single, rows, cols, k, seed=1)
    s1 = single.solve_resident(False, None)
    single.download_solution(z)
    out["single_gpu_ms"] = s1["ms_solve"]
    out["matches_single_gpu"] = bool(np.array_equal(z.person_to_object[begin:begin + count], res["p2o"]) and
                                     np.array_equal(z.object_to_person, res["o2p"]) and
                                     np.array_equal(single.prices(), res["prices"]) and s1["bid_arcs"] == st["global_bid_arcs"])
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
