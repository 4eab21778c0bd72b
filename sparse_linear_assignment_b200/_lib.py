"""ctypes binding of libsla_b200.so (the C ABI in include/sla.h).

The library is built in-tree by `build_library()` (nvcc, sm_100a only).  There is no CPU fallback: if the shared
object is missing, or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libsla_b200.so")
HOST_LIB_PATH = os.path.join(_PKG, "libsla_host.so")   # host-only entry points (the instance generator), no CUDA
CSRC = os.path.join(_PKG, "csrc")
INCLUDE = os.path.join(_ROOT, "include")

SLA_OK, SLA_ERR_INVALID, SLA_ERR_CUDA, SLA_ERR_NO_DEVICE, SLA_ERR_STATE, SLA_ERR_ALLOC = range(6)
SLA_NONE = 0xFFFFFFFF
ALGO_KHOSLA, ALGO_FORWARD = 0, 1

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


class SlaError(RuntimeError):
    """A C-ABI call returned a non-zero status (the Rust shim maps this to anyhow::Error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"sla error {code}: {message}")
        self.code = code
        self.message = message


class SlaStats(C.Structure):
    _fields_ = [
        ("num_unassigned", C.c_uint32), ("nits", C.c_uint32), ("nreductions", C.c_uint32),
        ("optimal_soln_found", C.c_uint32), ("eps", C.c_double), ("rounds", C.c_uint64), ("bids", C.c_uint64),
        ("bid_arcs", C.c_uint64), ("dropped", C.c_uint32), ("values_negated", C.c_uint32),
        ("wide_rounds", C.c_uint64), ("tail_rounds", C.c_uint64), ("kernel_launches", C.c_uint32),
        ("graph_launches", C.c_uint32), ("ms_solve", C.c_float), ("ms_total", C.c_float), ("cluster_rounds", C.c_uint64),
        ("restarts", C.c_uint32), ("reserved_", C.c_uint32),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class SlaRoundProfile(C.Structure):
    _fields_ = [
        ("round", C.c_uint32), ("engine", C.c_uint32), ("bidders", C.c_uint32), ("rounds_covered", C.c_uint32),
        ("arcs", C.c_uint64), ("bid_ms", C.c_float), ("assign_ms", C.c_float),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "sla.h")]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/sla_api.cu into libsla_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    stale = force or not os.path.exists(LIB_PATH)
    if not stale:
        t = os.path.getmtime(LIB_PATH)
        stale = any(os.path.getmtime(s) > t for s in sources())
    if stale:
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-o", LIB_PATH, os.path.join(CSRC, "sla_api.cu")]
        subprocess.run(cmd, check=True)
    build_host_library(force)
    return LIB_PATH


def build_host_library(force: bool = False) -> str:
    """Compile csrc/sla_host.cpp into libsla_host.so with plain g++: the host-only entry points of include/sla.h."""
    srcs = [os.path.join(CSRC, f) for f in ("sla_host.cpp", "sla_host_impl.h", "synth.h")] + [os.path.join(INCLUDE, "sla.h")]
    stale = force or not os.path.exists(HOST_LIB_PATH)
    if not stale:
        t = os.path.getmtime(HOST_LIB_PATH)
        stale = any(os.path.getmtime(s) > t for s in srcs)
    if stale:
        gxx = shutil.which("g++") or "g++"
        subprocess.run([gxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", HOST_LIB_PATH, srcs[0]], check=True)
    return HOST_LIB_PATH


_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/sla.h
SIGNATURES = {
    "sla_ctx_create": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(_vp)]),
    "sla_ctx_destroy": (None, [_vp]),
    "sla_last_error": (C.c_char_p, [_vp]),
    "sla_ctx_stream": (_vp, [_vp]),
    "sla_ctx_device": (C.c_int, [_vp]),
    "sla_version": (C.c_char_p, []),
    "sla_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "sla_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "sla_host_free": (None, [_vp]),
    "sla_host_negate_f64": (None, [_vp, C.c_size_t, C.c_int]),
    "sla_upload_csr": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint64]),
    "sla_upload_csr_negating": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint64, C.c_int]),
    "sla_host_narrow": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp, C.c_int]),
    "sla_last_upload": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "sla_scan_value_bytes": (C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    "sla_upload_csr_device": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint64]),
    "sla_generate_device": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                      C.c_int]),
    "sla_generate_device_shard": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                            C.c_int, C.c_uint32, C.c_uint32]),
    "sla_generate_host": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                    _vp, _vp, _vp]),
    "sla_generate_host_ex": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_int, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "sla_khosla_solve": (C.c_int, [_vp, C.c_int, C.c_double, _vp, _vp, _vp, C.POINTER(SlaStats)]),
    "sla_forward_solve": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_uint32, _vp, _vp, _vp,
                                    C.POINTER(SlaStats)]),
    "sla_download_solution": (C.c_int, [_vp, _vp, _vp, _vp]),
    "sla_get_objective": (C.c_int, [_vp, _f64p]),
    "sla_ecs_satisfied": (C.c_int, [_vp, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "sla_validate_matching": (C.c_int, [_vp, _u32p, C.POINTER(C.c_int)]),
    "sla_get_round_profile": (C.c_int, [_vp, C.POINTER(SlaRoundProfile), C.c_size_t, C.POINTER(C.c_size_t)]),
    "sla_batch_upload": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "sla_batch_generate_device": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]),
    "sla_batch_solve": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32, _vp, _vp, _vp,
                                  C.POINTER(SlaStats), C.POINTER(SlaStats)]),
    "sla_part_begin": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                 C.c_double]),
    "sla_part_claim": (C.c_int, [_vp]),
    "sla_part_local_value_range": (C.c_int, [_vp, _f64p, _f64p, _f64p]),
    "sla_part_bid": (C.c_int, [_vp]),
    "sla_part_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "sla_part_assign": (C.c_int, [_vp, _u32p, _u32p]),
    "sla_part_finish": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(SlaStats)]),
    "sla_part_sparse_buffers": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "sla_part_collect": (C.c_int, [_vp, _u32p]),
    "sla_part_apply_sparse": (C.c_int, [_vp, C.c_int, C.c_uint32, _u32p, _u32p]),
    "sla_mesh_create": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_uint32, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "sla_ipc_export": (C.c_int, [_vp, _vp]),
    "sla_ipc_import": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "sla_ipc_release": (C.c_int, [C.c_int, _vp]),
    "sla_mesh_connect": (C.c_int, [_vp, _vp, _vp]),
    "sla_mesh_begin": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "sla_mesh_solve": (C.c_int, [_vp]),
    "sla_mesh_phase": (C.c_int, [_vp, C.c_int]),
    "sla_mesh_poll": (C.c_int, [_vp, C.POINTER(C.c_int), _u32p, _u32p]),
    "sla_mesh_finish": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(SlaStats)]),
    "sla_mesh_owned": (C.c_int, [_vp, _u32p, _u32p, _u32p, _u32p]),
    "sla_mesh_round1_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "sla_mesh_objective": (C.c_int, [_vp, _f64p]),
    "sla_mesh_timeline": (C.c_int, [_vp, _vp, C.c_size_t]),
}

_lib = None


def load() -> C.CDLL:
    """Load libsla_b200.so; raise if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlaError(SLA_ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                              f"g.build()'` (the CUDA extension is required, there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


_host_lib = None


def load_host() -> C.CDLL:
    """Load libsla_host.so (g++-built, no CUDA): the synthetic instance generator for CPU-only processes."""
    global _host_lib
    if _host_lib is None:
        lib = C.CDLL(build_host_library())
        for name in ("sla_generate_host", "sla_generate_host_ex"):
            res, args = SIGNATURES[name]
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _host_lib = lib
    return _host_lib


def check(ctx, rc: int) -> None:
    if rc != SLA_OK:
        msg = load().sla_last_error(ctx)
        raise SlaError(rc, msg.decode() if msg else "")
