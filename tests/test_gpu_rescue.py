"""GPU parity of the first half of the oc2cns -r 1 rescue: nb_rescue_local (k_rescue_local: Myers' O(nd) local alignment,
gapped_align/oc_daligner.c:34-79) against the oracle's o_rescue_local on every candidate of the long-indel golden set."""
import ctypes as C
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


class Task(C.Structure):
    _fields_ = [("qid", C.c_uint32), ("sid", C.c_uint32), ("qdir", C.c_int32), ("qstart", C.c_int32), ("tstart", C.c_int32),
                ("pad", C.c_int32)]


class Result(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ok", "qbeg", "qend", "tbeg", "tend", "diffs", "overflow", "pad")]


def test_rescue_local_matches_oracle(built, tmp_path):
    from necat_b200 import capi, formats
    import oracle_ffi
    wrk = util.golden_workdir_r1(tmp_path)
    os.chdir(str(tmp_path))
    vol = formats.read_volume(os.path.join(wrk, "vol0"))
    cand = formats.read_candidates(os.path.join(util.GOLDEN, "r1", "cand_k13.bin"))
    sizes = vol.info["size"].astype(np.int64)
    # normalise so that the subject is forward (gapped_candidate.c:72-93) and take the extension start
    tasks, codes = [], formats.unpack_bases(vol.pac, vol.db_size)
    starts = vol.info["offset"].astype(np.int64)
    for rec in cand:
        w = int(rec[0])
        sdir, qdir, flag = w >> 31, (w >> 30) & 1, (w >> 29) & 1
        sid, sbeg, send, qid, qbeg, qend = (int(x) for x in rec[1:7])
        if sdir:
            qdir ^= 1
            flag ^= 1
            qbeg, qend = int(sizes[qid]) - qend, int(sizes[qid]) - qbeg
            sbeg, send = int(sizes[sid]) - send, int(sizes[sid]) - sbeg
        qoff, soff = (qbeg, sbeg) if flag else (qend, send)
        tasks.append((qid, sid, qdir, qoff, soff))
    L = capi.lib()
    ctx = capi.Context(0)
    dvol = ctx.upload(vol)
    arr = (Task * len(tasks))(*[Task(q, s, d, a, b, 0) for q, s, d, a, b in tasks])
    res = (Result * len(tasks))()
    L.nb_rescue_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_void_p]
    rc = L.nb_rescue_local(ctx.h, dvol, arr, len(tasks), 0.5, 1000, res)
    assert rc == 0, L.nb_last_error(ctx.h)
    O = oracle_ffi.lib()
    O.o_rescue_new.restype = C.c_void_p
    O.o_rescue_new.argtypes = [C.c_double]
    O.o_rescue_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 5
    rs = O.o_rescue_new(0.5)
    comp = np.array([3, 2, 1, 0], np.uint8)
    n_ok = 0
    for i, (qid, sid, qdir, qoff, soff) in enumerate(tasks):
        q = codes[starts[qid]:starts[qid] + sizes[qid]]
        if qdir:
            q = comp[q[::-1]]
        q = np.ascontiguousarray(q)
        t = np.ascontiguousarray(codes[starts[sid]:starts[sid] + sizes[sid]])
        out = [C.c_int() for _ in range(5)]
        ok = O.o_rescue_local(rs, q.ctypes.data_as(C.c_void_p), qoff, int(sizes[qid]), t.ctypes.data_as(C.c_void_p), soff, int(sizes[sid]),
                              1000, *[C.byref(x) for x in out])
        qb, qe, tb, te, diffs = (x.value for x in out)
        r = res[i]
        assert not r.overflow
        assert (r.qbeg, r.qend, r.tbeg, r.tend, r.diffs, r.ok) == (qb, qe, tb, te, diffs, ok), f"task {i}"
        n_ok += ok
    assert len(tasks) > 500 and n_ok > 100
    ctx.free_volume(dvol)
