"""CPU-only checks of the oracle's building blocks."""
import ctypes as C

import numpy as np


def test_bitvector_edlib_equals_plain_dp(built):
    """o_edlib_align (bit-vector, stored deltas) == o_edlib_align_dp (integer matrix) on random
    mutated pairs, including failures (distance above the cutoff) and ragged sizes."""
    import oracle_ffi
    L = oracle_ffi.lib()
    rng = np.random.default_rng(5)
    al = L.o_aligner_new(0.5)
    n_fail = 0
    for it in range(300):
        qn = int(rng.integers(1, 700))
        q = rng.integers(0, 4, qn, dtype=np.uint8)
        if it % 7 == 0:
            t = rng.integers(0, 4, int(rng.integers(1, 700)), dtype=np.uint8)       # unrelated -> may fail
        else:
            keep = rng.random(qn) > 0.12
            t = q[keep].copy()
            m = rng.random(t.size) < 0.1
            t[m] = (t[m] + 1) & 3
            t = np.concatenate([t, rng.integers(0, 4, int(rng.integers(0, 200)), dtype=np.uint8)])
        if t.size == 0:
            continue
        b1 = [C.create_string_buffer(4096) for _ in range(2)]
        b2 = [C.create_string_buffer(4096) for _ in range(2)]
        e1 = [C.c_int(), C.c_int()]
        e2 = [C.c_int(), C.c_int()]
        r1 = L.o_edlib_align(al, q.ctypes.data, qn, t.ctypes.data, t.size, b1[0], b1[1], C.byref(e1[0]), C.byref(e1[1]))
        r2 = L.o_edlib_align_dp(0.5, q.ctypes.data, qn, t.ctypes.data, t.size, b2[0], b2[1], C.byref(e2[0]), C.byref(e2[1]))
        assert r1 == r2
        n_fail += r1 == 0
        if r1:
            assert b1[0].value == b2[0].value and b1[1].value == b2[1].value
            assert (e1[0].value, e1[1].value) == (e2[0].value, e2[1].value)
    assert n_fail > 0
    L.o_aligner_free(al)
