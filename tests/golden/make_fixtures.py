"""Regenerates tests/golden/ksparse_fixtures.npz: the CSR inputs of the reference's three random tests
(populate_with_ksparse_input, /root/reference/src/solver.rs:261-292, ChaCha8 seeds 1 / 2) produced by the
restated RNG chain in oracle/fixture_rng.c.  The chain is pinned by the reference's golden objectives
(tests/test_oracle_goldens.py); the committed arrays in turn pin the chain against accidental edits.

    python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

out = {}
for name, (n, m, k) in {"small": (5, 5, 2), "no_perfect": (9, 9, 3), "large": (90, 900, 32)}.items():
    rp, c, v = O.fixture_ksparse(n, m, k, 10.0)
    out[name + "_row_ptr"], out[name + "_cols"], out[name + "_vals"] = rp, c, v
out["chacha8_seed1_u64"] = O.chacha8_u64_stream(1, 40)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ksparse_fixtures.npz"), **out)
print("wrote ksparse_fixtures.npz")
