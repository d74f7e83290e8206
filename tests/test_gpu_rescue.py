"""GPU parity of the oc2cns -r 1 rescue: nb_rescue_local (k_rescue_local: Myers' O(nd) local alignment,
gapped_align/oc_daligner.c:34-79) and nb_rescue_global (edlib_go, edlib/edlib_wrapper.c:111-255) against the oracle on the
long-indel golden set, and oc2cns -r 1 end to end against the reference binary's FASTA."""
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


class Range(C.Structure):
    _fields_ = [("qid", C.c_uint32), ("sid", C.c_uint32)] + [(n, C.c_int32) for n in ("qdir", "qfrom", "qto", "tfrom", "tto", "tolerance")]


def test_rescue_global_matches_oracle(built, tmp_path):
    """edlib_go on the GPU (nb_rescue_global) against the oracle's o_rescue_global on the ranges the local alignment finds."""
    from necat_b200 import capi, formats
    import oracle_ffi
    wrk = util.golden_workdir_r1(tmp_path)
    os.chdir(str(tmp_path))
    vol = formats.read_volume(os.path.join(wrk, "vol0"))
    cand = formats.read_candidates(os.path.join(util.GOLDEN, "r1", "cand_k13.bin"))
    sizes = vol.info["size"].astype(np.int64)
    starts = vol.info["offset"].astype(np.int64)
    codes = formats.unpack_bases(vol.pac, vol.db_size)
    comp = np.array([3, 2, 1, 0], np.uint8)
    O = oracle_ffi.lib()
    O.o_rescue_new.restype = C.c_void_p
    O.o_rescue_new.argtypes = [C.c_double]
    O.o_rescue_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 5
    O.o_rescue_global.argtypes = ([C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] +
                                  [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_double), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int)])
    rs = O.o_rescue_new(0.5)
    ranges, expect = [], []
    for rec in cand[::6]:
        w = int(rec[0])
        sdir, qdir, flag = w >> 31, (w >> 30) & 1, (w >> 29) & 1
        sid, sbeg, send, qid, qbeg, qend = (int(x) for x in rec[1:7])
        if sdir:
            qdir ^= 1
            flag ^= 1
            qbeg, qend = int(sizes[qid]) - qend, int(sizes[qid]) - qbeg
            sbeg, send = int(sizes[sid]) - send, int(sizes[sid]) - sbeg
        qoff, soff = (qbeg, sbeg) if flag else (qend, send)
        q = codes[starts[qid]:starts[qid] + sizes[qid]]
        if qdir:
            q = comp[q[::-1]]
        q = np.ascontiguousarray(q)
        t = np.ascontiguousarray(codes[starts[sid]:starts[sid] + sizes[sid]])
        out = [C.c_int() for _ in range(5)]
        if not O.o_rescue_local(rs, q.ctypes.data_as(C.c_void_p), qoff, int(sizes[qid]), t.ctypes.data_as(C.c_void_p), soff, int(sizes[sid]),
                                1000, *[C.byref(x) for x in out]):
            continue
        qb, qe, tb, te, diffs = (x.value for x in out)
        o4 = [C.c_int() for _ in range(4)]
        ident, qa, ta, ln = C.c_double(), C.c_char_p(), C.c_char_p(), C.c_int()
        ok = O.o_rescue_global(rs, q.ctypes.data_as(C.c_void_p), qb, qe, t.ctypes.data_as(C.c_void_p), tb, te, diffs, 1000, 4,
                               *[C.byref(x) for x in o4], C.byref(ident), C.byref(qa), C.byref(ta), C.byref(ln))
        ranges.append(Range(qid, sid, qdir, qb, qe, tb, te, diffs))
        expect.append((ok, [x.value for x in o4], oracle_ffi.strings_to_ops(qa.value, ta.value) if ok else None))
    assert len(ranges) > 40 and sum(e[0] for e in expect) > 20
    L = capi.lib()
    ctx = capi.Context(0)
    dvol = ctx.upload(vol)
    arr = (Range * len(ranges))(*ranges)
    aln, ops, nb = C.c_void_p(), C.c_void_p(), C.c_uint64()
    L.nb_rescue_global.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.nb_rescue_global(ctx.h, dvol, arr, len(ranges), 0.5, 1000, 4, C.byref(aln), C.byref(ops), C.byref(nb))
    assert rc == 0, L.nb_last_error(ctx.h)
    a = np.frombuffer((C.c_char * (capi.ALIGNMENT.itemsize * len(ranges))).from_address(aln.value), capi.ALIGNMENT).copy()
    o = np.frombuffer((C.c_char * max(nb.value, 1)).from_address(ops.value), np.uint8).copy()
    for i, (ok, ends, eops) in enumerate(expect):
        assert int(a["ok"][i]) == ok, f"range {i}"
        if ok:
            assert [int(a[k][i]) for k in ("qoff", "qend", "toff", "tend")] == ends, f"range {i}"
            got = o[int(a["ops_offset"][i]):int(a["ops_offset"][i]) + int(a["aln_size"][i])]
            assert np.array_equal(got, eops), f"range {i}"
    ctx.free_volume(dvol)


def test_oc2cns_rescue_matches_golden(built, tmp_path):
    """oc2cns -r 1 on the long-indel golden set: FASTA identical to the reference binary's."""
    import json
    import shutil
    man = json.load(open(os.path.join(util.GOLDEN, "manifest.json")))["r1"]
    wrk = util.golden_workdir_r1(tmp_path)
    c = str(tmp_path / "c.bin")
    shutil.copy(os.path.join(util.GOLDEN, "r1", "cand_k13.bin"), c)
    util.run([os.path.join(util.BIN_DIR, "oc2pcan"), "-t", "1", wrk, c], cwd=str(tmp_path))
    opts = man["cns_r1"]["options"].split()
    util.run([os.path.join(util.BIN_DIR, "oc2cns")] + opts + [wrk, c, "cns.fa", "raw.fa"], cwd=str(tmp_path))
    assert util.fasta_canon(str(tmp_path / "cns.fa")) == util.golden_fasta("r1/cns_r1")
    assert util.fasta_canon(str(tmp_path / "raw.fa")) == util.golden_fasta("r1/raw_r1")
