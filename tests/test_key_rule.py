"""The integer-key form of the first-round choice rule (csrc/sla_common.cuh: scan8_keys / key_choice_group_reduce /
key_choice_to_f64), restated in numpy and compared with the reference's sequential f64 rule (ksparse.rs:199-214,
symmetric.rs:361-376) on rows full of ties and extreme values: same best position, same best value, same second-best
profit, for both signs of the on-the-fly negation and for every lane split the kernels use.  CPU only; the GPU tests
compare the kernel itself with the f64 kernels and the CPU model."""
import numpy as np
import pytest


def reference_rule(values):
    """Strict '>' in position order; second = second largest of the multiset."""
    best, second, pos = -np.inf, -np.inf, -1
    for t, v in enumerate(values):
        if v > best:
            second, best, pos = best, v, t
        elif v > second:
            second = v
    return pos, best, second


def key_rule(x, flip, lanes):
    """x: u16 values of one row; lanes: LPR8 (each lane scans 8-arc chunks lane, lane + lanes, ...)."""
    keyflip = 0xFFFF if flip else 0
    state = []
    for lane in range(lanes):
        best = second = -1
        for off in range(8 * lane, len(x), 8 * lanes):
            for t in range(8):
                pk = ((int(x[off + t]) ^ keyflip) << 15) + (32767 - off - t)
                lo = min(best, pk)
                best = max(best, pk)
                second = max(second, lo)
        state.append((best, second))
    m = lanes // 2
    while m >= 1:                                   # butterfly: every lane ends with the same pair
        nxt = []
        for lane in range(lanes):
            b, s_ = state[lane]
            ob, os_ = state[lane ^ m]
            nxt.append((max(b, ob), max(max(s_, os_), min(b, ob))))
        state = nxt
        m //= 2
    best, second = state[0]
    sign = -1.0 if flip else 1.0
    value = sign * float((best >> 15) ^ keyflip)
    sec = sign * float((second >> 15) ^ keyflip)
    return 32767 - (best & 32767), value, sec


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("k,lanes", [(8, 1), (16, 2), (24, 4), (32, 4), (64, 8), (256, 32)])
def test_key_rule_equals_reference_rule(flip, k, lanes):
    rng = np.random.default_rng(k + flip)
    for trial in range(200):
        hi = (2, 5, 1000, 65536)[trial % 4]         # few distinct values: ties for best and for second
        x = rng.integers(0, hi, size=k).astype(np.uint16)
        if trial % 7 == 0:
            x[rng.integers(0, k)] = 65535
        if trial % 5 == 0:
            x[:] = x[0]                             # all equal: position 0 wins, second == best
        v = -x.astype(np.float64) if flip else x.astype(np.float64)
        pos, best, second = reference_rule(v)
        kpos, kval, ksec = key_rule(x, flip, lanes)
        assert (kpos, kval, ksec) == (pos, best, second)
