"""torchrun target: one KhoslaSolver instance over the ranks' GPUs with the mesh engine (peer memory over NVLink),
generated shard by shard in HBM.
    torchrun --nproc-per-node N scripts/run_mesh.py ROWS COLS K [check] [reps]
Rank 0 prints one JSON line; with `check` the same instance is also solved on rank 0's GPU alone and compared."""
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
from sparse_linear_assignment_b200.distributed import MeshKhoslaSolver, shard_rows

rows, cols, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
check = len(sys.argv) > 4 and sys.argv[4] == "check"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
begin, count = shard_rows(rows, world, rank)
solver, _ = S.KhoslaSolver.new(count, cols, count * k, device=local)
ctx = solver._context()
_lib.check(ctx, _lib.load().sla_generate_device_shard(ctx, rows, cols, k, 1, 300, 1000, 0, begin, count))
solver._num_rows, solver._num_cols, solver._dirty, solver._device_only = count, cols, False, True
mesh = MeshKhoslaSolver(solver).setup()
times, dev_ms = [], []
for rep in range(reps):
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    t = time.perf_counter()
    res = mesh.solve(False, None, download=(check and rep == reps - 1), gather=(check and rep == reps - 1))
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t)
    dev_ms.append(res["stats"]["ms_solve"])
st = res["stats"]
dm = torch.tensor(dev_ms, dtype=torch.float64, device=torch.device("cuda", local))
if world > 1:
    dist.all_reduce(dm, op=dist.ReduceOp.MAX)
out = {"engine": "mesh", "world": world, "rows": rows, "cols": cols, "k": k, "rounds": st["rounds"],
       "bid_arcs": st["global_bid_arcs"], "unassigned": st["global_num_unassigned"],
       "host_ms": [round(x * 1e3, 3) for x in times], "device_ms_max_over_ranks": [round(x, 4) for x in dm.tolist()],
       "graph_launches": st["graph_launches"], "kernel_launches": st["kernel_launches"],
       "scan_value_bytes": solver.scan_value_bytes(),
       "note": "host_ms: begin + solve + finish incl. the three small all-reduces; the last repetition downloads and gathers when `check` is given"}
if check and rank == 0:
    single, z = S.KhoslaSolver.new(rows, cols, rows * k, device=local)
    S.generators.kregular_device(single, rows, cols, k, seed=1)
    for _ in range(3):
        s1 = single.solve_resident(False, None)
    single.download_solution(z)
    out["single_gpu_ms"] = s1["ms_solve"]
    out["matches_single_gpu"] = bool(np.array_equal(z.person_to_object, res["global_p2o"]) and
                                     np.array_equal(z.object_to_person, res["global_o2p"]) and
                                     np.array_equal(single.prices(), res["global_prices"]) and
                                     s1["bid_arcs"] == st["global_bid_arcs"] and s1["rounds"] == st["rounds"])
if os.environ.get("SLA_MESH_TIMELINE"):
    import ctypes as C
    buf = (C.c_uint64 * (64 * 12))()
    _lib.check(ctx, _lib.load().sla_mesh_timeline(ctx, buf, 64 * 12))
    tl = np.array(buf[:], dtype=np.uint64).reshape(64, 12).astype(np.int64)
    rows_ = []
    for r in range(1, min(int(st["rounds"]) + 1, 64)):
        t = tl[r]
        if t[0] == 0:
            break
        # us: bid work, B1 wait, max+resolve work (from max start), B2 wait, finish work, whole round (to the next bid start)
        nxt = tl[r + 1][0] if r + 1 < 64 and tl[r + 1][0] else t[7]
        rows_.append([r, round((t[1] - t[0]) / 1e3, 1), round((t[2] - t[1]) / 1e3, 1), round((t[3] - t[2]) / 1e3, 1),
                      round((t[8] - t[3]) / 1e3, 1), round((t[4] - t[8]) / 1e3, 1), round((t[5] - t[4]) / 1e3, 1),
                      round((t[6] - t[5]) / 1e3, 1), round((t[7] - t[6]) / 1e3, 1), round((nxt - t[0]) / 1e3, 1)])
    out_tl = {"rank": rank, "columns": ["round", "bid", "B1", "gap->max", "max(+gap)", "resolve", "B2", "finish", "control", "round_total"],
              "us": rows_}
    with open(f"gpurun_out/{os.environ.get('SLA_TAG', 'mesh')}_timeline_rank{rank}.json", "w") as f:
        json.dump(out_tl, f)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
