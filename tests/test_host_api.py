"""CPU-only: the host-side mirror of the reference API (CSR builder, validation, objective, eps-CS, clone) and the
C-ABI library's exported symbols.  No compute call is made here (there is no GPU in this container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import fixtures, goldens

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sla):
    from sparse_linear_assignment_b200 import _lib
    header = open(os.path.join(ROOT, "include", "sla.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sla_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sla.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signatures out of sync with include/sla.h"
    lib.sla_version.restype = C.c_char_p
    assert b"sm_100a" in lib.sla_version()


def test_host_only_library_and_rust_binding_cover_the_header(sla):
    """libsla_host.so (g++, no CUDA) exports the host-only entry points and maps no CUDA runtime; rust/src/ffi.rs
    declares every function of include/sla.h (the Rust side is not compiled here: no toolchain)."""
    from sparse_linear_assignment_b200 import _lib
    host = C.CDLL(_lib.build_host_library())
    for name in ("sla_generate_host", "sla_generate_host_ex"):
        assert hasattr(host, name)
    import subprocess
    needed = subprocess.run(["ldd", _lib.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in needed.lower()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sla.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(sla_[a-z0-9_]+)\s*\(", header))
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (sla_[a-z0-9_]+)\(", ffi))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    for rel in ("rust/src/device.rs", "rust/build.rs", "rust/apply.py", "rust/patches/ksparse.rs.ed",
                "rust/patches/symmetric.rs.ed", "rust/patches/lib.rs.ed", "rust/patches/Cargo.toml.ed"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel


def test_stats_struct_layout_agrees_across_the_three_bindings(sla, tmp_path):
    """sla_stats crosses the C ABI by layout: the ctypes mirror must have the compiler's size and field offsets for
    include/sla.h, and rust/src/ffi.rs must list the same fields in the same order with the same widths."""
    import subprocess
    from sparse_linear_assignment_b200 import _lib
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sla.h")).read(), flags=re.S)
    body = re.search(r"typedef struct sla_stats \{(.*?)\} sla_stats;", header, flags=re.S).group(1)
    c_fields = re.findall(r"\b(uint32_t|uint64_t|double|float)\s+([a-z_0-9]+)\s*;", body)
    assert [n for _, n in c_fields] == [n for n, _ in _lib.SlaStats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sla.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(sla_stats));\n'
                   + "".join(f'  printf(" %zu", offsetof(sla_stats, {n}));\n' for _, n in c_fields)
                   + '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    nums = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert nums[0] == C.sizeof(_lib.SlaStats) and nums[0] % 8 == 0
    assert nums[1:] == [getattr(_lib.SlaStats, n).offset for _, n in c_fields]
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    rs_body = re.search(r"pub struct sla_stats \{(.*?)\}", ffi, flags=re.S).group(1)
    rs_fields = re.findall(r"pub ([a-z_0-9]+): (u32|u64|f64|f32)", rs_body)
    width = {"uint32_t": "u32", "uint64_t": "u64", "double": "f64", "float": "f32"}
    assert rs_fields == [(n, width[ty]) for ty, n in c_fields]


def _strip_rust(src):
    """Rust source without comments, string and char literals (enough for brace counting and call parsing)."""
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    return re.sub(r"'(?:\\.|[^'\\])'", "' '", src)


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_rust_patches_apply_to_the_reference_and_are_consistent(sla, tmp_path):
    """rust/apply.py on a scratch copy of the reference crate (v0.1.5): the line-addressed patches land where they were
    cut for, the patched sources are balanced and no longer mention the host-side state they removed, the replaced
    solve bodies call the device mirror, and every `ffi::sla_*` call in src/device.rs matches the arity of its
    declaration in src/ffi.rs.  (No Rust toolchain here: this is the structural half of a compile check.  The reference
    is only present in the build container -- the test skips elsewhere.)"""
    import shutil
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isfile(os.path.join(ref, "src", "ksparse.rs")):
        pytest.skip("reference checkout not present")
    crate = tmp_path / "crate"
    shutil.copytree(ref, crate, ignore=shutil.ignore_patterns(".git", "target"))
    for dirpath, _, files in os.walk(crate):
        os.chmod(dirpath, 0o755)
        for f in files:
            os.chmod(os.path.join(dirpath, f), 0o644)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "rust", "apply.py"), str(crate)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    again = subprocess.run([sys.executable, os.path.join(ROOT, "rust", "apply.py"), str(crate)], capture_output=True, text=True)
    assert again.returncode != 0 and "already patched" in again.stderr      # line-addressed patches: never twice
    srcs = {rel: open(crate / rel).read() for rel in ("src/ksparse.rs", "src/symmetric.rs", "src/lib.rs", "src/device.rs",
                                                       "src/ffi.rs", "build.rs", "Cargo.toml")}
    for rel, text in srcs.items():
        if rel.endswith(".rs"):
            code = _strip_rust(text)
            for a, b in ("{}", "()", "[]"):
                assert code.count(a) == code.count(b), (rel, a, code.count(a), code.count(b))
    ks, sy = _strip_rust(srcs["src/ksparse.rs"]), _strip_rust(srcs["src/symmetric.rs"])
    assert "self.dev.khosla_solve(" in ks and "self.dev.forward_solve(" in sy
    for gone in ("ustack", "num_iter"):
        assert gone not in ks, gone
    for gone in ("best_bids", "best_bidders", "unassigned_people", "person_to_assignment_idx", "push_all_left", "num_iter",
                 "bid_and_assign"):
        assert gone not in sy, gone
    for code in (ks, sy):                                  # the trait surface is untouched
        for method in ("fn new(", "fn solve(", "fn prices(&self)", "fn prices_mut(&mut self)", "fn values_mut(&mut self)",
                       "fn num_rows(&self)", "fn num_cols_mut(&mut self)"):
            assert method in code, method
        assert "OnceCell<Vec<f64>>" in code and "dev: DeviceMirror" in code and "impl<I: UnsignedInt" in code
    assert "mod device;" in srcs["src/lib.rs"] and "mod ffi;" in srcs["src/lib.rs"]
    assert 'links = "sla_b200"' in srcs["Cargo.toml"] and 'build = "build.rs"' in srcs["Cargo.toml"]
    assert "rustc-link-lib=dylib=sla_b200" in srcs["build.rs"]
    # arity of every FFI call the device mirror makes
    ffi = _strip_rust(srcs["src/ffi.rs"])
    decl = {m.group(1): len(_split_args(m.group(2)))
            for m in re.finditer(r"pub fn (sla_[a-z0-9_]+)\((.*?)\)\s*(?:->[^;]*)?;", ffi, flags=re.S)}
    dev = _strip_rust(srcs["src/device.rs"])
    calls = 0
    for m in re.finditer(r"ffi::(sla_[a-z0-9_]+)\(", dev):
        name, i, depth = m.group(1), m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(dev[i], 0)
            i += 1
        assert name in decl, name
        assert len(_split_args(dev[m.end():i - 1])) == decl[name], (name, dev[m.end():i - 1])
        calls += 1
    assert calls >= 8
    # the struct fields the mirror reads exist in the binding
    for field in re.findall(r"\bst\.([a-z_]+)", dev + ks + sy):
        assert re.search(rf"pub {field}: ", ffi), field


def test_no_gpu_means_loud_failure(sla):
    """Without a device the product path must fail, never fall back to a CPU solve."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    solver, solution = sla.KhoslaSolver.new(4, 4, 16)
    solver.init(2, 2)
    solver.extend_from_values(0, [0, 1], [1.0, 2.0])
    solver.extend_from_values(1, [0, 1], [3.0, 1.0])
    with pytest.raises(sla.SlaError) as e:
        solver.solve(solution, False, None)
    assert e.value.code == 3   # SLA_ERR_NO_DEVICE


@pytest.mark.parametrize("cls_name", ["KhoslaSolver", "ForwardAuctionSolver"])
def test_cumulative_idx_diff_u16(sla, cls_name):
    g = goldens()["cumulative_idx_diff"]                      # src/symmetric.rs:525-534
    solver, _ = getattr(sla, cls_name).new(7, 7, 7, index_dtype=np.uint16)
    solver.init(7, 7)
    for r in g["rows"]:
        solver.add_value(r, 0, 0.0)
    assert list(solver.i_starts_stops()) == g["i_starts_stops"]
    assert list(solver.j_counts()) == g["j_counts"]
    assert solver.i_starts_stops().dtype == np.uint16
    assert solver.num_of_arcs() == 7


def test_builder_matches_oracle_builder(sla, oracle):
    fx = fixtures()
    rp, c, v = fx["large_row_ptr"], fx["large_cols"], fx["large_vals"]
    solver, _ = sla.ForwardAuctionSolver.new(90, 900, 90 * 32)
    solver.init(90, 900)
    o = oracle.OracleSolver("forward", 90, 900, 90 * 32)
    o.init(90, 900)
    for i in range(90):
        a, b = rp[i], rp[i + 1]
        if i % 2:
            solver.extend_from_values(i, c[a:b], v[a:b])
        else:
            for g in range(a, b):
                solver.add_value(i, c[g], v[g])
        o.extend_from_values(i, c[a:b], v[a:b])
    assert np.array_equal(solver.i_starts_stops(), o.i_starts_stops)
    assert np.array_equal(solver.j_counts(), o.j_counts)
    assert np.array_equal(solver.column_indices(), c)
    assert np.array_equal(solver.values(), v)
    bulk, _ = sla.KhoslaSolver.new(1, 1, 1)
    bulk.load_csr(90, 900, rp, c, v)
    assert np.array_equal(bulk.i_starts_stops(), solver.i_starts_stops())
    assert np.array_equal(bulk.j_counts(), solver.j_counts())


def test_builder_errors(sla):
    solver, _ = sla.KhoslaSolver.new(4, 4, 16)
    with pytest.raises(sla.SlaError):
        solver.init(5, 4)                                     # solver.rs:192
    solver.init(2, 4)
    with pytest.raises(sla.SlaError):
        solver.add_value(1, 0, 1.0)                           # solver.rs:55
    solver.add_value(0, 0, 1.0)
    with pytest.raises(sla.SlaError):
        solver.add_value(2, 0, 1.0)                           # solver.rs:44
    with pytest.raises(sla.SlaError):
        solver.extend_from_values(0, [1, 2], [1.0])           # solver.rs:75
    s16, _ = sla.KhoslaSolver.new(4, 4, 16, index_dtype=np.uint16)
    with pytest.raises(sla.SlaError):
        s16.init(0xFFFF, 0xFFFF)                              # solver.rs:193
    s16.init(2, 70000)
    with pytest.raises(sla.SlaError):
        s16.extend_from_values(0, np.zeros(70000), np.zeros(70000))   # I::from_usize, solver.rs:80-81
    empty, sol = sla.ForwardAuctionSolver.new(4, 4, 16)
    empty.init(2, 2)
    with pytest.raises(sla.SlaError):
        empty.solve(sol, False, None)                         # validate_input before any device work, solver.rs:234


def test_solution_new_and_clone(sla):
    _, z = sla.KhoslaSolver.new(3, 5, 9)
    assert z.person_to_object.size == 0 and z.object_to_person.size == 0    # solution.rs:46-53
    assert np.isnan(z.eps) and z.num_unassigned == 0xFFFFFFFF
    _, z16 = sla.KhoslaSolver.new(3, 5, 9, index_dtype=np.uint16)
    assert z16.num_unassigned == 0xFFFF
    solver, _ = sla.ForwardAuctionSolver.new(2, 2, 4)
    solver.init(2, 2)
    solver.extend_from_values(0, [0, 1], [1.0, 2.0])
    twin = solver.clone()
    twin.extend_from_values(1, [0], [5.0])
    assert solver.num_of_arcs() == 2 and twin.num_of_arcs() == 3


def test_get_objective_ecs_and_toleration_match_oracle(sla, oracle):
    """Host post-processing (solver.rs:110-189) against the oracle's, on a solved oracle state."""
    fx = fixtures()
    rp, c, v = fx["large_row_ptr"], fx["large_cols"], fx["large_vals"]
    o = oracle.OracleSolver("forward", 90, 900, 90 * 32)
    o.load_csr(90, 900, rp, c, v)
    o.solve(maximize=False)
    solver, z = sla.ForwardAuctionSolver.new(90, 900, 90 * 32)
    solver.load_csr(90, 900, rp, c, v)
    solver.init_solve(z, False)                               # host-side sign normalisation, solver.rs:207-230
    assert np.array_equal(solver.values(), o.values)
    assert z.num_unassigned == 90 and np.all(z.person_to_object == 0xFFFFFFFF)
    z.person_to_object = o.person_to_object
    solver.prices_mut()[:] = o.prices
    assert solver.get_objective(z) == o.get_objective() == goldens()["random_large"]["minimize"]
    tol = solver.get_toleration(10.0)
    assert tol == oracle.get_toleration(10.0)
    for eps in (1.0 / 90, 1e-9, 0.0):
        assert solver.ecs_satisfied(z.person_to_object, eps, tol) == o.ecs_satisfied(eps, tol)
    for cost in (0.0, 0.5, 1.0, 999.0, 1000.0, 1e6):
        assert solver.get_toleration(cost) == oracle.get_toleration(cost)


def test_host_narrowing_is_lossless_or_refuses(sla):
    """sla_host_narrow (the host half of the narrow upload): accepts exactly the arrays whose every value survives the
    round trip through u16 / f32 bit for bit, leaves the input untouched when it refuses, and negates in place only
    when it accepts.  Lengths around the SIMD width, values around the representable range, zeros of both signs, NaN."""
    import ctypes as C
    from sparse_linear_assignment_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)

    def run(v, tier, negate):
        v = np.ascontiguousarray(v, dtype=np.float64)
        work = v.copy()
        out = np.zeros(max(v.size, 1), dtype=np.uint16 if tier == 2 else np.float32)
        rc = lib.sla_host_narrow(work.ctypes.data, v.size, tier, out.ctypes.data, int(negate))
        with np.errstate(invalid="ignore", over="ignore"):
            if tier == 2:
                ok = bool(np.all((v >= 0) & (v <= 65535) & (np.floor(v) == v) & ~np.signbit(v)))
            else:
                ok = bool(np.array_equal(v.astype(np.float32).astype(np.float64).view(np.uint64), v.view(np.uint64)))
        assert rc == (1 if ok else 0), (tier, negate, v[:8], rc, ok)
        if ok:
            assert np.array_equal(out[: v.size].astype(np.float64).view(np.uint64), v.view(np.uint64))
            assert np.array_equal(work.view(np.uint64), (-v if negate else v).view(np.uint64))
        else:
            assert np.array_equal(work.view(np.uint64), v.view(np.uint64))      # untouched (NaN payloads included)

    for n in (1, 3, 7, 8, 9, 15, 16, 17, 31, 1000, 4099):
        ints = rng.integers(0, 65536, size=n).astype(np.float64)
        halves = ints + 0.5
        for negate in (False, True):
            run(ints, 2, negate)
            run(ints, 4, negate)
            run(halves, 2, negate)
            run(halves, 4, negate)
            run(-ints - 1.0, 2, negate)
            run(-ints - 1.0, 4, negate)
            for pos in {0, n // 2, n - 1}:
                for bad in (65536.0, -0.0, -1.0, 0.1, float("nan"), float("inf"), 2.0 ** 40, 1e300, 5e-324):
                    w = ints.copy()
                    w[pos] = bad
                    run(w, 2, negate)
                    run(w, 4, negate)
            run(rng.uniform(0, 1000, size=n), 4, negate)
            edge = ints.copy()
            edge[0] = 65535.0
            edge[-1] = 0.0
            run(edge, 2, negate)


def test_cached_addresses_follow_reallocation_and_outputs_are_recognised_by_identity(sla, monkeypatch):
    """The wrapper hands the C ABI cached base addresses (ndarray.ctypes costs more than the rest of a small solve's
    Python time): they must follow every reallocation of the growable vectors, the solution vectors of the previous
    solve are reused only while they are the very same arrays of the right size, and the host-thread count follows the
    environment it was computed from."""
    from sparse_linear_assignment_b200 import solver as SV
    v = SV._Vec(np.float64, 4)
    assert v.addr == v.a.ctypes.data
    v.extend(np.arange(4.0))
    first = v.addr
    v.extend(np.arange(1000.0))                               # grows: new block, new address
    assert v.addr == v.a.ctypes.data and v.len == 1004 and np.array_equal(v.view[:4], np.arange(4.0))
    assert v.addr != first or v.a.size >= 1004
    v.assign(np.arange(5000.0))
    assert v.addr == v.a.ctypes.data and v.view[-1] == 4999.0

    s, z = sla.KhoslaSolver.new(3, 5, 9)
    s.init(3, 5)
    s.extend_from_values(0, [0, 1], [1.0, 2.0])
    s.extend_from_values(1, [1, 2], [3.0, 4.0])
    s.extend_from_values(2, [3, 4], [5.0, 6.0])
    for name in ("_i_starts_stops", "_column_indices", "_values"):
        vec = getattr(s, name)
        assert vec.addr == vec.a.ctypes.data
    p2o, o2p, a1, a2 = s._outputs(z)
    assert p2o.size == 3 and o2p.size == 5 and a1 == p2o.ctypes.data and a2 == o2p.ctypes.data
    z.person_to_object, z.object_to_person = p2o, o2p         # what _finish leaves in the caller's solution
    again = s._outputs(z)
    assert again[0] is p2o and again[1] is o2p and again[2:] == (a1, a2)
    z.object_to_person = np.zeros(5, dtype=np.uint32)         # another array of the same shape: taken as it is
    other = s._outputs(z)
    assert other[1] is z.object_to_person and other[3] == z.object_to_person.ctypes.data
    s.init(2, 5)                                              # another shape: fresh vectors
    s.extend_from_values(0, [0], [1.0])
    s.extend_from_values(1, [1], [1.0])
    third = s._outputs(z)
    assert third[0].size == 2 and third[0] is not p2o

    s.last_stats = {"rounds": 3}
    assert s.last_stats == {"rounds": 3}
    st = SV.SlaStats()
    st.rounds = 7
    s._last_stats, s._last_struct = None, st                  # what _finish records: the dict is built on first access
    assert s.last_stats["rounds"] == 7 and s.last_stats is s.last_stats

    monkeypatch.setenv("SLA_HOST_THREADS", "3")
    assert SV.host_threads() == 3
    monkeypatch.setenv("SLA_HOST_THREADS", "64")
    assert SV.host_threads() == 16
    monkeypatch.delenv("SLA_HOST_THREADS")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "100000")
    assert SV.host_threads() == 1
    monkeypatch.delenv("LOCAL_WORLD_SIZE")
    assert 1 <= SV.host_threads() <= 16
