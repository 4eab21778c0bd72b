"""ncu target: the mesh engine with two ranks stepped in lockstep on ONE GPU (every kernel a separate, serial launch, all
'remote' pointers local), shards generated in HBM.  argv: ROWS COLS K [solves]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import _lib
from sparse_linear_assignment_b200.distributed import MeshShard, mesh_lockstep_solve, shard_rows

rows, cols, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
solves = int(sys.argv[4]) if len(sys.argv) > 4 else 2
world = 2
begins = [shard_rows(rows, world, r)[0] for r in range(world)] + [rows]
shards = []
for r in range(world):
    b, cnt = begins[r], begins[r + 1] - begins[r]
    s, _ = S.KhoslaSolver.new(cnt, cols, cnt * k)
    ctx = s._context()
    _lib.check(ctx, _lib.load().sla_generate_device_shard(ctx, rows, cols, k, 1, 300, 1000, 0, b, cnt))
    s._num_rows, s._num_cols, s._dirty, s._device_only = cnt, cols, False, True
    shards.append(MeshShard(s, r, world, begins))
for sh in shards:
    sh.connect_pointers([t.block for t in shards])
for _ in range(solves):
    t = time.perf_counter()
    res = mesh_lockstep_solve(shards)
    print("ok", res["stats"]["rounds"], res["stats"]["bid_arcs"], round((time.perf_counter() - t) * 1e3, 2), "ms (lockstep, host-stepped)")
