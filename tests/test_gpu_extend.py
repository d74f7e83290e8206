"""GPU parity of the block-wise Myers extension (nb_extend_candidates) against the
oracle's o_onc_align on the same candidates: same ok flag, same end points, same
alignment column by column."""
import ctypes as C
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _extend_gpu(ctx, dvol, recs, error, min_align, tail):
    from necat_b200 import capi
    L = capi.lib()
    recs = np.ascontiguousarray(recs, np.uint32)
    aln = C.c_void_p(); ops = C.c_void_p(); nbytes = C.c_uint64()
    rc = L.nb_extend_candidates(ctx.h, dvol, recs.ctypes.data_as(C.c_void_p), C.c_uint64(recs.shape[0]),
                                C.c_double(error), C.c_int(min_align), C.c_int(tail),
                                C.byref(aln), C.byref(ops), C.byref(nbytes))
    assert rc == 0, L.nb_last_error(ctx.h).decode()
    n = recs.shape[0]
    a = np.frombuffer((C.c_char * (capi.ALIGNMENT.itemsize * n)).from_address(aln.value), capi.ALIGNMENT).copy()
    o = np.frombuffer((C.c_char * max(nbytes.value, 1)).from_address(ops.value), np.uint8)[:nbytes.value].copy()
    L.nb_free(aln); L.nb_free(ops)
    return a, o


@pytest.mark.parametrize("gen,tail,min_align,band_test", [
    (dict(genome=80_000, depth=16, seed=31), 4, 2000, 0),
    (dict(genome=60_000, depth=16, seed=32, err=0.08), 1, 500, 0),
    (dict(genome=60_000, depth=14, seed=33, err=0.12), 4, 400, 0),
    # the walk-back trusts only the two middle words of a column's band: blocks are redone with a moved band all the time
    (dict(genome=60_000, depth=14, seed=34, err=0.15), 4, 400, 1),
])
def test_extension_matches_oracle(built, tmp_path, monkeypatch, gen, tail, min_align, band_test):
    from necat_b200 import capi, formats
    import oracle_ffi
    monkeypatch.setenv("NB_EXT_BAND_TEST", str(band_test))
    wrk, _ = util.make_workdir(tmp_path, **gen)
    vol = formats.read_volume(os.path.join(wrk, "vol0"))
    recs = util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, 0, str(tmp_path / "c.bin"))
    sizes = vol.info["size"].astype(np.int64)
    part = formats.partition_records(recs, sizes)
    assert part.shape[0] > 50
    part = part[:1500]
    ctx = capi.Context(0)
    dvol = ctx.upload(vol)
    aln, ops = _extend_gpu(ctx, dvol, part, 0.5, min_align, tail)
    al = oracle_ffi.lib().o_aligner_new(0.5)
    n_ok = 0
    for i, r in enumerate(part):
        w = int(r[0])
        qdir = (w >> 30) & 1
        qid, sid = int(r[4]), int(r[1])
        if w & (1 << 29):
            qoff, soff = int(r[5]), int(r[2])
        else:
            qoff, soff = int(r[6]), int(r[3])
        exp = oracle_ffi.onc_align(al, vol.sequence(qid, qdir), qoff, vol.sequence(sid, 0), soff, min_align, tail)
        g = aln[i]
        got_ops = ops[int(g["ops_offset"]):int(g["ops_offset"]) + int(g["aln_size"])]
        assert int(g["ok"]) == exp["ok"], f"candidate {i}: ok {g['ok']} vs {exp['ok']}"
        assert (int(g["qoff"]), int(g["qend"]), int(g["toff"]), int(g["tend"])) == \
               (exp["qoff"], exp["qend"], exp["toff"], exp["tend"]), f"candidate {i}: ends differ"
        assert got_ops.size == exp["ops"].size and np.array_equal(got_ops, exp["ops"]), f"candidate {i}: columns differ"
        assert int(g["n_match"]) == int((exp["ops"] == 0).sum())
        n_ok += exp["ok"]
    assert n_ok > 10
    oracle_ffi.lib().o_aligner_free(al)
    ctx.free_volume(dvol)
    ctx.close()
