"""B200-native auction hot path behind the public API of DXist/sparse_linear_assignment.

`KhoslaSolver` / `ForwardAuctionSolver` / `AuctionSolution` mirror the reference's Rust API (src/lib.rs:7-19);
the bidding loop runs in hand-written sm_100a CUDA through the C ABI of `libsla_b200.so` (include/sla.h).
"""
from ._lib import SlaError, build_library, load as load_library
from .solver import AuctionSolution, AuctionSolver, ForwardAuctionSolver, KhoslaSolver
from . import generators
from .batch import BatchSolver

__all__ = ["AuctionSolution", "AuctionSolver", "ForwardAuctionSolver", "KhoslaSolver", "SlaError", "build_library",
           "load_library", "generators", "BatchSolver"]
