#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the auction hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1]

A "step" is one complete solve of one synthetic instance (BASELINE.json configs; default cfg3 = KhoslaSolver
1M x 4M, k=16, the configuration the north star quotes its roofline target on).  Metric: bid-arcs/s, where a
bid-arc is one (column, value) pair examined for one bidding person in one round.

  value     device-resident throughput: CSR already in HBM, results stay in HBM; CUDA events on the solve stream
            around exactly K steps, max over ranks.
  e2e       the same metric through the public API (KhoslaSolver.solve): every step uploads the CSR from pinned
            host memory, solves, and copies person_to_object / object_to_person / prices back to the host.
  roofline  the dominant kernel (round-1 bid scan) timed with CUDA events inside the library ("profile" option):
            algorithmic bytes 12*A + 8*B over the launch duration, against the measured HBM copy bandwidth.
  cpu_baseline   the CPU oracle (C restatement of the reference, goldens verified) on one host core.

With N > 1 (torchrun, one rank per GPU) the default workload is the one the north star partitions: cfg5, ONE KhoslaSolver
instance of 16M x 64M, k=16, row-partitioned over the N GPUs with the mesh engine (objects owner-partitioned, bids pushed
into the owner's HBM over NVLink inside the bid kernel; csrc/sla_mesh.cuh) -- "scaling": "strong".  The same line carries
the one-GPU solve of the same instance measured on rank 0 in the same run (`one_gpu`), a slice-by-slice comparison of the two
solutions, and the cfg4 batch (8,192 independent instances) split over the ranks.  `--workload cfg3 --gpus N` still runs N
independent replicas ("replicas").
`--impl reference` times the reference's CPU algorithm (the oracle port; the Rust crate cannot be built here) on the same
workload: cfg3 in full at N = 1; for cfg5 a bounded sample (the first 2,000,000 persons against all 64M objects).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bid_arcs_per_sec"
UNIT = "bid-arcs/s"

WORKLOADS = {
    # name: (solver, rows, cols, k, planted, eps)
    "cfg1": ("khosla", 1_000, 10_000, 32, False, None),
    "cfg2": ("forward", 20_000, 20_000, 64, True, None),
    "cfg3": ("khosla", 1_000_000, 4_000_000, 16, False, None),
    "cfg5": ("khosla", 16_000_000, 64_000_000, 16, False, None),
}
CFG4 = dict(instances=8192, rows=512, cols=512, k=32)    # batch of independent instances, Forward solver, eps-scaled


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS) + ["cfg4"],
                    help="auto: cfg3 on one GPU, cfg5 partitioned over the GPUs (mesh engine) on several")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def describe(workload):
    kind, n, m, k, planted, eps = WORKLOADS[workload]
    return {
        "workload": f"{workload}: {'KhoslaSolver' if kind == 'khosla' else 'ForwardAuctionSolver'} {n}x{m} k={k} "
                    f"integer costs {'floor(700*Beta(3,3)+300)' if workload == 'cfg1' else 'uniform in [300,1000)'}"
                    f"{' planted perfect matching' if planted else ''}, minimize, eps=None",
        "rows": n, "cols": m, "k": k, "arcs": n * k,
        "csr_bytes": n * k * 12 + (n + 1) * 4,
        "l2": "inputs larger than the 126 MB L2" if n * k * 12 > 126e6 else "L2 flushed between timed steps",
    }


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons every 200 ms while the timed region runs: NVML (what nvidia-smi reads;
    no process spawn, so the sampler does not disturb microsecond-scale steps), nvidia-smi as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.device_index = device_index
        self.samples = []          # (sm_mhz, sm_max_mhz, set(reasons))
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_once(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            smax = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((sm, smax, {k for k, bit in self.REASONS.items() if mask & bit}))
            return
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                              str(self.device_index)], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            s = [x.strip() for x in out.split(",")]
            reasons = {name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9])
                       if val.lower().startswith("active")}
            self.samples.append((float(s[1]), float(s[2]), reasons))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.sample_once()
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        try:
            self.sample_once()      # at least one sample right at the end of the timed region
        except Exception:
            pass
        sm = sorted(s[0] for s in self.samples)
        reasons = set().union(*[s[2] for s in self.samples]) if self.samples else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[1] for s in self.samples), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def flush_l2(torch, device):
    if not hasattr(flush_l2, "buf"):
        flush_l2.buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flush_l2.buf.add_(1)


# ---------------------------------------------------------------------------------------------------------------------
def host_instance(workload, seed, pinned=True):
    """The synthetic instance in (pinned) host memory, generated by the library's counter-based generator."""
    import numpy as np
    import torch
    from sparse_linear_assignment_b200 import generators as G
    kind, n, m, k, planted, eps = WORKLOADS[workload]
    pin = pinned and torch.cuda.is_available()
    t_rp = torch.empty(n + 1, dtype=torch.int32, pin_memory=pin)
    t_c = torch.empty(n * k, dtype=torch.int32, pin_memory=pin)
    t_v = torch.empty(n * k, dtype=torch.float64, pin_memory=pin)
    rp, c, v = t_rp.numpy().view(np.uint32), t_c.numpy().view(np.uint32), t_v.numpy()
    # cfg1 is the reference's own bench shape: values floor(700 * Beta(3,3) + 300) (benches/benchmark.rs:60,73)
    G.kregular_host(n, m, k, seed=seed, planted=planted, out=(rp, c, v), value_dist="beta33" if workload == "cfg1" else "uniform")
    return (rp, c, v), (t_rp, t_c, t_v)


def run_oracle(workload, csr, reps):
    """The reference's CPU algorithm (oracle port) on the host instance: best-of-reps solve time, bid-arcs."""
    from oracle import oracle as O
    kind, n, m, k, planted, eps = WORKLOADS[workload]
    rp, c, v = csr
    times, arcs, obj = [], 0, None
    for _ in range(reps):
        s = O.OracleSolver(kind, n, m, n * k)
        s.load_csr(n, m, rp, c, v)          # the oracle negates its own copy of the values in place
        t = time.perf_counter()
        s.solve(maximize=False, eps=eps)
        times.append(time.perf_counter() - t)
        arcs, obj = s.bid_arcs, s.get_objective()
        del s
    return times, arcs, obj


REF_SAMPLE_ROWS = 2_000_000      # cfg5 on the CPU: persons [0, 2M) of the instance against all 64M objects per step


def main_reference_sample(args):
    """The reference's CPU solver on a bounded sample of cfg5 (the full instance takes ~8 s per solve and 9 GB of host
    arrays): the first REF_SAMPLE_ROWS persons of the same generated instance, all 64M objects -- so prices and owners
    are as cache-hostile as in the full problem and bid-arcs/s is comparable."""
    import numpy as np
    from oracle import oracle as O
    from sparse_linear_assignment_b200 import generators as G
    kind, n, m, k, planted, eps = WORKLOADS["cfg5"]
    rows = REF_SAMPLE_ROWS
    rp, c, v = G.kregular_host(n, m, k, seed=args.seed, planted=planted, row_begin=0, row_count=rows)
    total_arcs, total_t, obj = 0, 0.0, None
    t0 = time.perf_counter()
    for step in range(max(args.warmup, 0) + args.steps):
        s = O.OracleSolver(kind, rows, m, rows * k)
        s.load_csr(rows, m, rp, c, v)
        t = time.perf_counter()
        s.solve(maximize=False, eps=1.0 / m)          # the eps of the full instance (1 / num_cols, ksparse.rs:167)
        dt = time.perf_counter() - t
        if step >= max(args.warmup, 0):
            total_arcs += s.bid_arcs
            total_t += dt
        obj = s.get_objective()
        del s
    value = total_arcs / total_t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": describe("cfg5"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"persons [0, {rows}) of the cfg5 instance against all {m} objects per step, solve() only "
                                   f"(the reference is single-threaded; C restatement of src/ksparse.rs, goldens verified)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "objective_of_sample": obj, "wall_s": time.perf_counter() - t0}))
    return 0


def main_reference(args, partitioned=False):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if partitioned or args.workload == "cfg5":
        return main_reference_sample(args)
    csr, _keep = host_instance(args.workload, args.seed, pinned=False)
    for _ in range(max(args.warmup, 0)):
        run_oracle(args.workload, csr, 1)
    t0 = time.perf_counter()
    total_arcs, total_t = 0, 0.0
    for _ in range(args.steps):
        times, arcs, obj = run_oracle(args.workload, csr, 1)
        total_arcs += arcs
        total_t += times[0]
    wall = time.perf_counter() - t0
    value = total_arcs / total_t
    cfg = describe(args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"full {args.workload} instance, {args.steps} solves, solve() only (the reference is "
                                   f"single-threaded; C restatement of src/ksparse.rs / src/symmetric.rs, goldens verified)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "objective": obj, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


def cfg4_oracle_sample(sample, seed0, threads):
    """The reference's CPU solver on `sample` instances of cfg4, one solver per host thread (independent instances)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from sparse_linear_assignment_b200 import generators as G
    n, m, k = CFG4["rows"], CFG4["cols"], CFG4["k"]
    data = [G.kregular_host(n, m, k, seed=seed0 + i, planted=True) for i in range(sample)]
    O.lib()

    def one(i):
        s = O.OracleSolver("forward", n, m, n * k)
        s.load_csr(n, m, *data[i])
        s.solve(maximize=False, eps=None)
        return s.bid_arcs

    t = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        arcs = sum(ex.map(one, range(sample)))
    return arcs, time.perf_counter() - t


def main_cfg4(args):
    """cfg4: 8,192 independent 512 x 512 k=32 instances (Forward, eps-scaled), instance ranges split over the ranks."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cfg = {"workload": f"cfg4: batch of {CFG4['instances']} independent ForwardAuctionSolver instances "
                       f"{CFG4['rows']}x{CFG4['cols']} k={CFG4['k']}, integer costs [300,1000), planted perfect matching, eps=None",
           "instances": CFG4["instances"], "rows": CFG4["rows"], "cols": CFG4["cols"], "k": CFG4["k"],
           "l2": "inputs (1.6 GB of CSR) larger than the 126 MB L2"}
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = 4 * cores
        tot_arcs = tot_t = 0.0
        for step in range(args.warmup + args.steps):
            arcs, t = cfg4_oracle_sample(sample, 0, cores)
            if step >= args.warmup:
                tot_arcs += arcs
                tot_t += t
        value = tot_arcs / tot_t
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": f"{sample} of the {CFG4['instances']} instances per step, one oracle solver per host thread"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    import numpy as np
    import torch
    import sparse_linear_assignment_b200 as S
    from sparse_linear_assignment_b200 import generators as G
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S.build_library()
    device = torch.device("cuda", local_rank)
    n, m, k = CFG4["rows"], CFG4["cols"], CFG4["k"]
    count = CFG4["instances"] // world
    first = rank * count

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    bs = S.BatchSolver("forward", device=local_rank)
    bs.generate_device(count, first, n, m, k, seed=0, planted=True)
    for _ in range(max(args.warmup, 3)):
        bs.solve(download=False, per_instance=False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms = arcs = 0.0
    for _ in range(args.steps):
        tot = bs.solve(download=False, per_instance=False)["total"]
        ms += tot["ms_solve"]
        arcs += tot["bid_arcs"]
    barrier()
    clocks = sampler.summary()
    # ---- e2e: the rank's instances built on the host (page-locked), uploaded, solved, results copied back ----
    e2e_count = min(count, 1024)
    rp = S.solver.host_array(e2e_count * n + 1, np.uint32)
    c = S.solver.host_array(e2e_count * n * k, np.uint32)
    v = S.solver.host_array(e2e_count * n * k, np.float64)
    for i in range(e2e_count):
        a = i * n * k
        G.kregular_host(n, m, k, seed=first + i, planted=True, out=(rp[i * n:(i + 1) * n + 1], c[a:a + n * k], v[a:a + n * k]))
        rp[i * n:(i + 1) * n + 1] += np.uint32(a)
    row_off = (np.arange(e2e_count + 1, dtype=np.uint64) * n).astype(np.uint32)
    col_off = (np.arange(e2e_count + 1, dtype=np.uint64) * m).astype(np.uint32)
    eb = S.BatchSolver("forward", device=local_rank)
    e2e_arcs, e2e_s = 0.0, 0.0
    from sparse_linear_assignment_b200 import _lib
    for step in range(2 + args.steps):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        ctx = eb._context()
        _lib.check(ctx, _lib.load().sla_batch_upload(ctx, e2e_count, row_off.ctypes.data, col_off.ctypes.data, rp.ctypes.data,
                                                     c.ctypes.data, v.ctypes.data))
        eb.n_inst, eb.row_off, eb.col_off = e2e_count, row_off, col_off
        res = eb.solve(download=True, per_instance=False)
        torch.cuda.synchronize(device)
        if step >= 2:
            e2e_s += time.perf_counter() - t0
            e2e_arcs += res["total"]["bid_arcs"]
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a = torch.tensor([arcs, e2e_arcs], dtype=torch.float64, device=device)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        (ms, e2e_s), (arcs, e2e_arcs) = t.tolist(), a.tolist()
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bids = tot["bids"]
        alg = 12 * tot["bid_arcs"] + 8 * bids
        ach = alg / (tot["ms_solve"] * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            sample = 4 * cores
            o_arcs, o_t = cfg4_oracle_sample(sample, 0, cores)
            cpu = {"value": o_arcs / o_t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{sample} of the {CFG4['instances']} instances, one oracle solver per host thread "
                             f"({o_t * 1e3:.0f} ms wall)"}
        e2e_sample = f"{e2e_count} instances per rank per step"
        print(json.dumps({
            "metric": METRIC, "value": arcs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "parallelism": f"{world} GPU(s), {count} instances each, one CTA per instance, no collective",
            "e2e": {"value": e2e_arcs / e2e_s, "unit": UNIT, "sample": e2e_sample, "h2d_bytes_per_step": int(rp.nbytes + c.nbytes + v.nbytes),
                    "d2h_bytes_per_step": int(e2e_count * (4 * n + 12 * m)), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": args.steps * world,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "kernel": "batch_kernel", "peak_source": peak_src,
                         "note": "CSR bytes of all bid scans over the launch time; rows are re-read from L1/L2, the "
                                 "kernel is bound by per-round latency inside each CTA, not by HBM"},
            "cpu_baseline": cpu,
            "solve": {k_: tot[k_] for k_ in ("rounds", "bids", "bid_arcs", "num_unassigned", "ms_solve")},
        }))
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


def slice_checksums(np, p2o, o2p, prices, first_row, first_object):
    """Order-sensitive 64-bit checksums of a rank's slices of a solution (wrap-around arithmetic)."""
    def cs(a, off):
        a = np.ascontiguousarray(a)
        w = np.arange(off + 1, off + 1 + a.size, dtype=np.uint64)
        return int(np.bitwise_xor.reduce(a.view(np.uint64 if a.dtype.itemsize == 8 else a.dtype).astype(np.uint64) * w +
                                         (w << np.uint64(7)))) if a.size else 0
    return [cs(p2o, first_row), cs(o2p, first_object), cs(prices, first_object)]


def main_partitioned_nccl(args, solver, why):
    """Fallback of main_mesh: cfg5 row-partitioned with replicated object state and the sparse NCCL exchange."""
    import torch
    import torch.distributed as dist
    from sparse_linear_assignment_b200.distributed import CudaShardEngine, PartitionedKhoslaSolver
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    kind, n, m, k, planted, eps = WORKLOADS["cfg5"]
    drv = PartitionedKhoslaSolver(CudaShardEngine(solver), exchange="sparse")
    for _ in range(max(args.warmup, 3)):
        drv.solve(False, eps, download=False)
    dist.barrier(device_ids=[local_rank])
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = drv.solve(False, eps, download=False)
    torch.cuda.synchronize(device)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        st = res["stats"]
        print(json.dumps({
            "metric": METRIC, "value": st["global_bid_arcs"] * args.steps / float(dt.item()), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * float(dt.item()) / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe("cfg5"),
            "parallelism": f"FALLBACK (mesh engine unavailable: {why}): persons row-partitioned over {world} GPUs, object state "
                           f"replicated, winners exchanged with NCCL all-gathers every round (host-driven loop)",
            "e2e": None, "gpu_launches": int(st["kernel_launches"]) * args.steps * world, "roofline": None, "cpu_baseline": None,
            "solve": {"rounds": st["rounds"], "bid_arcs": st["global_bid_arcs"], "num_unassigned": st["global_num_unassigned"]}}))
    dist.barrier(device_ids=[local_rank])
    dist.destroy_process_group()
    return 0


def main_mesh(args):
    """cfg5 -- ONE KhoslaSolver instance, 16M persons x 64M objects, k=16 -- row-partitioned over the ranks' GPUs with the
    mesh engine; rank 0 also solves the whole instance alone (the honest comparison point: it fits one B200)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import sparse_linear_assignment_b200 as S
    from sparse_linear_assignment_b200 import _lib, generators as G
    from sparse_linear_assignment_b200.distributed import MeshKhoslaSolver, shard_rows

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S.build_library()
    device = torch.device("cuda", local_rank)
    kind, n, m, k, planted, eps = WORKLOADS["cfg5"]
    begin, count = shard_rows(n, world, rank)

    def barrier():
        dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    def allmax(x):
        t = torch.tensor(x, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def allsum(x):
        t = torch.tensor(x, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    # ---- value: shards generated in HBM, results left in HBM ---------------------------------------------------------------
    solver, _ = S.KhoslaSolver.new(count, m, count * k, device=local_rank)
    ctx = solver._context()
    _lib.check(ctx, _lib.load().sla_generate_device_shard(ctx, n, m, k, args.seed, G.VALUE_LO, G.VALUE_HI, 0, begin, count))
    solver._num_rows, solver._num_cols, solver._dirty, solver._device_only = count, m, False, True
    try:
        mesh = MeshKhoslaSolver(solver).setup()
    except S.SlaError as e:
        # no peer mappings on this box (the failure is raised on every rank together): the same instance through the
        # older row-partitioned engine, winners exchanged with NCCL all-gathers -- slower, but the workload is measured
        return main_partitioned_nccl(args, solver, str(e))
    stream = torch.cuda.ExternalStream(solver._context_stream(), device=device)
    for _ in range(max(args.warmup, 3)):
        mesh.solve(False, eps, download=False, totals=False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    ms_dev, launches, r1 = 0.0, 0, []
    for _ in range(args.steps):
        res = mesh.solve(False, eps, download=False, totals=False)
        ms_dev += res["stats"]["ms_solve"]
        launches += res["stats"]["kernel_launches"]
        r1.append(mesh.shard.round1_ms())
    ev1.record(stream)
    barrier()
    ms_value = ev0.elapsed_time(ev1)
    clocks = sampler.summary()
    st = mesh.solve(False, eps, download=False, totals=True)["stats"]
    objective = mesh.objective()
    vb = solver.scan_value_bytes()
    ms_value, ms_dev = allmax([ms_value, ms_dev])
    launches = allsum([float(launches)])[0]
    r1_max = allmax([sorted(r1)[len(r1) // 2]])[0]

    # ---- the same instance on ONE GPU (rank 0), and a slice-by-slice comparison of the two solutions --------------------
    res = mesh.solve(False, eps, download=True, totals=False)
    own = res["owned"]
    mine = slice_checksums(np, res["p2o"], res["o2p"], res["prices"], begin, own["first_object"])
    spans = torch.tensor([begin, count, own["first_object"], own["num_owned"]] + [x >> 1 for x in mine] + [x & 1 for x in mine],
                         dtype=torch.int64, device=device)
    allspans = [torch.zeros_like(spans) for _ in range(world)]
    dist.all_gather(allspans, spans)
    one = None
    if rank == 0:
        single, z = S.KhoslaSolver.new(n, m, n * k, device=local_rank)
        G.kregular_device(single, n, m, k, seed=args.seed)
        for _ in range(3):
            single.solve_resident(False, eps)
        t1 = sorted(single.solve_resident(False, eps)["ms_solve"] for _ in range(max(min(args.steps, 10), 3)))
        s1 = single.last_stats
        single.download_solution(z)
        pr = single.prices()
        same = s1["bid_arcs"] == st["global_bid_arcs"] and s1["rounds"] == st["rounds"] and s1["num_unassigned"] == st["global_num_unassigned"]
        for sp in allspans:
            b_, c_, fo, no = [int(x) for x in sp[:4].tolist()]
            theirs = [(int(h) << 1) | int(l) for h, l in zip(sp[4:7].tolist(), sp[7:10].tolist())]
            ref = slice_checksums(np, z.person_to_object[b_:b_ + c_], z.object_to_person[fo:fo + no], pr[fo:fo + no], b_, fo)
            same = same and theirs == ref
        one = {"ms_per_step": t1[len(t1) // 2], "bid_arcs_per_s": s1["bid_arcs"] / (t1[len(t1) // 2] * 1e-3),
               "objective": single.device_objective(), "rounds": s1["rounds"], "scan_value_bytes": single.scan_value_bytes(),
               "identical_to_partitioned": bool(same and single.device_objective() == objective),
               "compared": "bid_arcs, rounds, num_unassigned, objective, and checksums of every rank's person_to_object / "
                           "object_to_person / prices slices against the same slices of the one-GPU solution"}
        del single, z, pr
    barrier()

    # ---- e2e: every rank's shard in (page-locked) host memory -> upload + solve + download of its slices, per step ----------
    hs, hz = S.KhoslaSolver.new(count, m, count * k, device=local_rank)
    hs.init(count, m)
    hs._i_starts_stops.resize(count + 1, 0)
    hs._j_counts.resize(count, k)
    hs._column_indices.resize(count * k, 0)
    hs._values.resize(count * k, 0.0)
    G.kregular_host(n, m, k, seed=args.seed, row_begin=begin, row_count=count,
                    out=(hs._i_starts_stops.view, hs._column_indices.view, hs._values.view))
    hmesh = MeshKhoslaSolver(hs).setup()
    hv = hs.values()
    e2e_steps = max(min(args.steps, 5), 2)
    e2e_s, e2e_arcs = 0.0, 0
    for step in range(2 + e2e_steps):
        if hv[0] < 0:
            np.negative(hv, out=hv)                      # untimed: hand the solver the caller's original costs again
        hs._dirty = True                                 # a fresh problem every step
        barrier()
        t0 = time.perf_counter()
        r = hmesh.solve(False, eps, download=True, totals=False)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if step >= 2:
            e2e_s += dt
            e2e_arcs += r["stats"]["bid_arcs"]
    h2d, value_bytes = hs.last_upload()
    d2h = 4 * count + 12 * r["owned"]["num_owned"]
    e2e_s = allmax([e2e_s])[0]
    e2e_arcs, h2d_all, d2h_all = allsum([float(e2e_arcs), float(h2d), float(d2h)])

    # ---- the cfg4 batch split over the same ranks (secondary block) ----------------------------------------------------------
    per = CFG4["instances"] // world
    bs = S.BatchSolver("forward", device=local_rank)
    bs.generate_device(per, rank * per, CFG4["rows"], CFG4["cols"], CFG4["k"], seed=0, planted=True)
    for _ in range(3):
        bs.solve(download=False, per_instance=False)
    barrier()
    c4_ms = c4_arcs = 0.0
    c4_steps = max(min(args.steps, 5), 2)
    for _ in range(c4_steps):
        tot = bs.solve(download=False, per_instance=False)["total"]
        c4_ms += tot["ms_solve"]
        c4_arcs += tot["bid_arcs"]
    c4_ms = allmax([c4_ms])[0]
    c4_arcs, c4_un = allsum([c4_arcs, float(tot["num_unassigned"])])

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        arcs = st["global_bid_arcs"]
        a1, b1 = count * k, count                       # round 1 on one rank: every local person bids
        alg = (4 + vb) * a1 + 8 * b1
        ach = alg / (r1_max * 1e-3) / 1e9 if r1_max > 0 else None
        cfg = describe("cfg5")
        line = {
            "metric": METRIC, "value": arcs * args.steps / (ms_value * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "parallelism": f"one instance over {world} GPUs: persons row-partitioned ({count} per rank), objects owner-partitioned "
                           f"({own['shard_objects']} per rank); bids pushed into the owner's HBM over NVLink peer mappings inside "
                           f"the bid kernel, flag barriers between the kernels, no collective and no host sync per round",
            "clocks": clocks,
            "device_ms_per_step": ms_dev / args.steps,
            "e2e": {"value": e2e_arcs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "h2d_value_bytes": value_bytes,
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                    "what": "MeshKhoslaSolver.solve with every rank's shard in page-locked host memory: upload (u16 values on "
                            "the wire, in-place sign normalisation of the host copy), solve, download of the rank's "
                            "person_to_object rows and of the object_to_person / prices slices it owns"},
            "gpu_launches": int(launches),
            "roofline": None if ach is None else {
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "mesh_bid_kernel<PRICE_ZERO" + (",u16>" if vb == 2 else ">"),
                "launch": f"round 1 on one rank ({b1} local persons bid): CSR scan + push of the bids to the owners; event-record "
                          f"nodes around the kernel inside the solve graph, median over the timed solves, max over ranks",
                "bidders": b1, "arcs": a1, "algorithmic_bytes": alg, "launch_us": r1_max * 1e3, "peak_source": peak_src,
                "value_bytes_in_hbm": vb,
                "nvlink_bytes_pushed": int(16 * b1 * (world - 1) / world),
                "note": "the kernel also stores 16 B per bid into the owners' inboxes, (world-1)/world of them over NVLink"},
            "nvlink": {"bid_entries_bytes_per_solve": int(16 * st["global_bids"] * (world - 1) / world),
                       "reply_bytes_per_solve": int(st["global_bids"] / 8 * (world - 1) / world),
                       "price_gather_sector_bytes_upper_bound": int(32 * (arcs - n * k) * (world - 1) / world),
                       "rounds": st["rounds"],
                       "note": "bytes that cross NVLink for the whole job; gathers: 32 B sector per arc scanned after round 1 "
                               "(upper bound: the bound-pruned scan fetches fewer)"},
            "one_gpu": one,
            "vs_one_gpu": None if not one else one["ms_per_step"] / (ms_value / args.steps),
            "cfg4_batch": {"value": c4_arcs / (c4_ms * 1e-3), "unit": UNIT, "ms_per_step": c4_ms / c4_steps, "steps": c4_steps,
                           "instances": CFG4["instances"], "instances_per_rank": per, "num_unassigned": int(c4_un),
                           "scaling": "strong", "note": "8,192 independent 512x512 k=32 Forward instances split over the ranks, "
                                                        "one CTA per instance, no collective; device time, max over ranks"},
            "cpu_baseline": None,
            "solve": {k_: st[k_] for k_ in ("rounds", "graph_launches", "kernel_launches")} |
                     {"bids": st["global_bids"], "bid_arcs": st["global_bid_arcs"], "num_unassigned": st["global_num_unassigned"]},
            "objective": objective,
        }
        print(json.dumps(line))
    barrier()
    dist.destroy_process_group()
    return 0


def main_ours(args):
    import numpy as np
    import torch
    import sparse_linear_assignment_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 2
    S.build_library()
    kind, n, m, k, planted, eps = WORKLOADS[args.workload]
    cls = S.KhoslaSolver if kind == "khosla" else S.ForwardAuctionSolver
    device = torch.device("cuda", local_rank)
    seed = args.seed + rank

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    # ---- inputs --------------------------------------------------------------------------------------------------
    csr, keep = host_instance(args.workload, seed, pinned=True)
    rp, c, v = csr
    resident, _ = cls.new(n, m, n * k, device=local_rank)
    resident.load_csr(n, m, rp, c, v)
    resident._sync_device()                       # CSR now resident in HBM; the host copy is not touched again
    stream = torch.cuda.ExternalStream(resident._context_stream(), device=device)
    small_inputs = n * k * 12 <= 126e6

    # ---- value: device-resident solves ---------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        resident.solve_resident(False, eps)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arcs = launches = 0
    ms_solve_sum = 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        if small_inputs:
            flush_l2(torch, device)
            torch.cuda.synchronize(device)         # the flush runs on torch's stream, the solve on the context's own:
                                                   # without this they would overlap (no flush, and a solve timed under
                                                   # 512 MB of foreign HBM traffic)
        st = resident.solve_resident(False, eps)
        arcs += st["bid_arcs"]
        launches += st["kernel_launches"]
        ms_solve_sum += st["ms_solve"]
    ev1.record(stream)
    barrier()
    ms_value = ev0.elapsed_time(ev1)
    if small_inputs:
        ms_value = ms_solve_sum                    # the L2 flush runs on another stream: count the solves only
    clocks = sampler.summary()
    stats = dict(st)

    # ---- e2e: public API with host buffers ------------------------------------------------------------------------
    solver, solution = cls.new(n, m, n * k, device=local_rank)
    solver.load_csr(n, m, rp, c, v)                # the solver's own host-side CSR storage (page-locked by the library)
    hv = solver.values()
    e2e_warm = max(min(args.warmup, 3), 1)
    for _ in range(e2e_warm):
        solver._dirty = True
        solver.solve(solution, False, eps)
    barrier()
    e2e_arcs, e2e_s = 0, 0.0
    for _ in range(args.steps):
        if hv[0] < 0:
            np.negative(hv, out=hv)                # untimed: hand the solver the caller's original (positive) costs again
        solver._dirty = True                       # a fresh problem every step: upload + solve + download
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        solver.solve(solution, False, eps)         # includes the in-place sign normalisation of values (solver.rs:214-216)
        # solve() is synchronous: it returns after the library's own cudaStreamSynchronize, with person_to_object /
        # object_to_person in the caller's host arrays and the negation workers joined -- no second device sync here
        e2e_s += time.perf_counter() - t0
        e2e_arcs += solver.last_stats["bid_arcs"]
    h2d, value_bytes = solver.last_upload()        # bytes the library actually moved: integer costs cross PCIe as u16
                                                   # and are widened to f64 in HBM (sla_last_upload)
    d2h = 4 * n + 4 * m                            # person_to_object + object_to_person (prices stay resident until read)
    objective = solver.get_objective(solution)

    # ---- reduce over ranks ---------------------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms_value, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a = torch.tensor([float(arcs), float(e2e_arcs), float(launches)], dtype=torch.float64, device=device)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        ms_value, e2e_s = t.tolist()
        arcs, e2e_arcs, launches = a.tolist()

    if rank == 0:
        # ---- roofline of the dominant kernel (round-1 bid scan), CUDA events inside the library -------------------------
        peak, peak_src = measured_peak_gbs()
        roof = {}
        # the variant that runs in the timed region, the general gathering variant, and the same two with the values read
        # as f64 (SURVEY 8(d)'s 12 bytes per arc: what real-valued weights get) -- all event-bracketed, live, in this run
        for skip, key, narrow in ((1, "roofline", 1), (0, "roofline_general_gather", 1), (1, "roofline_f64_values", 0),
                                  (0, "roofline_f64_general_gather", 0)):
            resident.set_option("narrow_scan", narrow)
            if not narrow and resident.scan_value_bytes() == 8 and "roofline" in roof and roof["roofline"]["value_bytes_in_hbm"] == 8:
                roof[key] = roof["roofline" if skip else "roofline_general_gather"]      # values were f64 to begin with
                continue
            resident.set_option("profile", 1)
            resident.set_option("zero_price_skip", skip)
            ts, rec = [], None
            for _ in range(5):
                resident.solve_resident(False, eps)
                prof = [p for p in resident.round_profile() if p["engine"] == 0]
                if prof:
                    rec = prof[0]
                    ts.append(rec["bid_ms"])
            # the same bracket around 8 back-to-back launches of the (idempotent) scan: the bracket itself -- event, launch
            # latency of a host-launched kernel, event -- costs several microseconds that the graph-launched solve does
            # not pay; reported beside the single-launch figure, which stays the headline
            b2b = None
            if rec:
                resident.set_option("profile_repeat", 8)
                tb = []
                for _ in range(5):
                    resident.solve_resident(False, eps)
                    prof = [p for p in resident.round_profile() if p["engine"] == 0]
                    if prof:
                        tb.append(prof[0]["bid_ms"] / 8.0)
                resident.set_option("profile_repeat", 1)
                b2b = sorted(tb)[len(tb) // 2] if tb else None
            # the eager bracket of round 1 (event, host-launched kernel, event), for comparison with the graph-launched one
            eager = None
            if rec and key in ("roofline", "roofline_f64_values"):
                resident.set_option("profile_graph", 0)
                te = []
                for _ in range(5):
                    resident.solve_resident(False, eps)
                    prof = [p for p in resident.round_profile() if p["engine"] == 0]
                    if prof:
                        te.append(prof[0]["bid_ms"])
                resident.set_option("profile_graph", 1)
                eager = sorted(te)[len(te) // 2] if te else None
            resident.set_option("profile", 0)
            resident.set_option("zero_price_skip", 1)
            vb = resident.scan_value_bytes()
            resident.set_option("narrow_scan", 1)
            if rec:
                t_ms = sorted(ts)[len(ts) // 2]
                # bytes per arc the scan must read: 4 (column index) + the width the values are resident with -- 8 (f64,
                # 12*A + 8*B) or 2 when a u16 upload left its lossless copy in HBM and the scan reads that
                survey = 12 * rec["arcs"] + 8 * rec["bidders"]
                alg = (4 + vb) * rec["arcs"] + 8 * rec["bidders"]
                ach = alg / (t_ms * 1e-3) / 1e9
                kname = "bid_regular_kernel<PRICE_ZERO>" if skip else "bid_regular_kernel<PRICE_LDG>"
                if vb == 2:
                    kname = kname[:-1] + ",u16>"
                roof[key] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": ncu_traffic(kname) if args.workload == "cfg3" else None, "kernel": kname, "launch": "round 1 (all persons bid)",
                             "bidders": rec["bidders"], "arcs": rec["arcs"], "algorithmic_bytes": alg,
                             "launch_us": t_ms * 1e3, "peak_source": peak_src, "value_bytes_in_hbm": vb,
                             "bytes_at_f64_values": survey, "equivalent_f64_gbs": survey / (t_ms * 1e-3) / 1e9}
                roof[key]["bracket"] = ("one graph launch: fence kernel, event-record node, the scan, event-record node "
                                        "(CUDA events on the context's stream)")
                if eager:
                    roof[key]["launch_us_eager_bracket"] = eager * 1e3
                if b2b:
                    roof[key]["launch_us_back_to_back"] = b2b * 1e3
                    roof[key]["frac_back_to_back"] = alg / (b2b * 1e-3) / 1e9 / peak
        # whole solve: SURVEY 8(d)'s f64 bytes over the solve time, whatever width the wide scans read the values at
        whole = (12 * stats["bid_arcs"] + 8 * stats["bids"]) / (stats["ms_solve"] * 1e-3) / 1e9

        cpu = None
        if not args.no_cpu_baseline:
            times, o_arcs, o_obj = run_oracle(args.workload, csr, 3)
            cpu = {"value": o_arcs / min(times), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"full {args.workload} instance, solve() only, best of 3 "
                             f"({min(times) * 1e3:.1f} ms, {o_arcs} bid-arcs; the reference is single-threaded)",
                   "ms": min(times) * 1e3, "objective": o_obj, "objective_matches_gpu": bool(o_obj == objective)}
        cfg = describe(args.workload)
        line = {
            "metric": METRIC, "value": arcs / (ms_value * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_value / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "replicas", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "parallelism": "1 GPU" if world == 1 else f"{world} independent instances, one per GPU (no collective): replicas, "
                                                      f"not scaling of one instance",
            "clocks": clocks,
            "e2e": {"value": e2e_arcs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "h2d_value_bytes": value_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof.get("roofline"),
            "roofline_general_gather": roof.get("roofline_general_gather"),
            "roofline_f64_values": roof.get("roofline_f64_values"),
            "roofline_f64_general_gather": roof.get("roofline_f64_general_gather"),
            "roofline_whole_solve": {"achieved": whole, "unit": "GB/s", "frac": whole / peak,
                                     "note": "sum over rounds of 12*A + 8*B (values counted as f64) over the whole solve time, tail rounds included"},
            "cpu_baseline": cpu,
            "solve": {k_: stats[k_] for k_ in ("rounds", "wide_rounds", "tail_rounds", "cluster_rounds", "bids", "bid_arcs", "num_unassigned",
                                               "kernel_launches", "graph_launches", "ms_solve")},
            "objective": objective,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    mesh = (a.workload == "auto" and max(world_env, a.gpus) > 1) or (a.workload == "cfg5" and world_env > 1)
    if a.workload == "auto":
        a.workload = "cfg5" if mesh else "cfg3"
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ and a.impl == "ours":
        # started by hand without torchrun: re-launch as one rank per GPU (the driver launches torchrun itself)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:])
    if a.workload == "cfg4":
        sys.exit(main_cfg4(a))
    if a.impl == "reference":
        sys.exit(main_reference(a, partitioned=mesh))
    sys.exit(main_mesh(a) if mesh else main_ours(a))
