"""GPU parity of the k-mer index itself (nb_index_build / nb_index_download) against a numpy restatement of
build_lookup_table (lookup_table/lookup_table.c:15-164): every k-mer that occurs at most q times, its positions ascending;
CSR starts over all 4^k hash values.  The numpy restatement is pinned on the C oracle's index by the CPU test below.
Covers hash runs of every size class of the build: short runs, runs of more than 24 positions and runs of more than
2048 (a homopolymer stretch), and k-mers dropped by the count cut-off."""
import ctypes as C
import os

import numpy as np
import pytest

import util


def np_index(vol, k, q):
    """-> (starts[4^k + 1], offsets) of the reference's index, dense CSR form."""
    from necat_b200 import formats
    codes = formats.unpack_bases(vol.pac, vol.db_size).astype(np.uint64)
    n = codes.size
    valid = np.zeros(n, bool)
    for off, size in zip(vol.info["offset"], vol.info["size"]):
        if size >= k:
            valid[int(off):int(off) + int(size) - k + 1] = True
    h = np.zeros(n, np.uint64)
    for i in range(k):                                      # first base in the most significant position (lookup_table.c:8-12)
        h[:n - k + 1] = (h[:n - k + 1] << np.uint64(2)) | codes[i:n - k + 1 + i]
    pos = np.nonzero(valid)[0]
    hv = h[pos]
    cnt = np.bincount(hv.astype(np.int64), minlength=1 << (2 * k))
    keep = cnt[hv.astype(np.int64)] <= q
    pos, hv = pos[keep], hv[keep]
    order = np.lexsort((pos, hv))
    cnt[cnt > q] = 0
    starts = np.zeros((1 << (2 * k)) + 1, np.uint32)
    starts[1:] = np.cumsum(cnt).astype(np.uint32)
    return starts, pos[order].astype(np.uint32), cnt


def make_volume(tmp_path, seed):
    """Random reads with planted repeats and one 3000-base homopolymer: runs of 2..40, of a few hundred and of ~3000."""
    from necat_b200 import formats
    import gen_reads
    _, reads = gen_reads.generate(40_000, 10, seed, err=0.05, repeat_len=1500, repeat_copies=12, n_repeats=2)
    rng = np.random.default_rng(seed)
    reads = list(reads)
    reads[3] = np.concatenate([reads[3][:500], np.zeros(3000, np.uint8), reads[3][500:]])            # poly-A
    reads[5] = np.concatenate([reads[5][:100], np.tile(np.array([0, 1], np.uint8), 400), reads[5][100:]])   # (AC)n
    reads.append(rng.integers(0, 4, 7, dtype=np.uint8))                                                # shorter than k
    wrk = str(tmp_path / "vol")
    formats.write_workdir(wrk, reads)
    return formats.read_volume(os.path.join(wrk, "vol0"))


@pytest.mark.gpu
@pytest.mark.parametrize("k,q,seed", [(11, 100000, 61), (12, 500, 62), (13, 20, 63), (15, 500, 64)])
def test_index_matches_reference_definition(built, tmp_path, k, q, seed):
    from necat_b200 import capi
    vol = make_volume(tmp_path, seed)
    e_starts, e_offs, cnt = np_index(vol, k, q)
    assert cnt.max() > (2048 if q >= 100000 else 24 if q > 24 else 2)       # the long-run paths are exercised
    ctx = capi.Context(0)
    dvol = ctx.upload(vol)
    idx = ctx.build_index(dvol, k, q)
    starts, offs, n_dist = ctx.index_arrays(idx, k)
    ctx.free_index(idx); ctx.free_volume(dvol); ctx.close()
    assert n_dist == int((cnt > 0).sum())
    assert np.array_equal(starts, e_starts), "CSR starts differ"
    assert offs.shape == e_offs.shape and np.array_equal(offs, e_offs), "positions differ (order inside a run matters: k_probe searches them)"


def test_numpy_index_is_the_oracles_index(built, tmp_path):
    """CPU: the numpy restatement above against the C oracle (oracle_pm.c: o_index_build), which is itself pinned on the
    reference binaries through the candidates it yields."""
    import oracle_ffi
    L = oracle_ffi.lib()
    vol = make_volume(tmp_path, 65)

    class OIndex(C.Structure):
        _fields_ = [("k", C.c_int), ("n_keys", C.c_uint64), ("key", C.POINTER(C.c_uint32)), ("start", C.POINTER(C.c_uint64)),
                    ("n_off", C.c_uint64), ("off", C.POINTER(C.c_uint64))]
    info = (oracle_ffi.OSeqInfo * vol.nseq)()
    for i in range(vol.nseq):
        info[i].offset = int(vol.info["offset"][i]); info[i].size = int(vol.info["size"][i])
    pac = np.ascontiguousarray(vol.pac)
    ov = oracle_ffi.OVolume(pac.ctypes.data_as(C.POINTER(C.c_uint8)), vol.db_size, vol.nseq, info, None, 0)
    for k, q in ((11, 100000), (12, 40)):
        ix = OIndex()
        assert L.o_index_build(C.byref(ov), k, q, C.byref(ix)) == 0
        starts, offs, cnt = np_index(vol, k, q)
        keys = np.ctypeslib.as_array(ix.key, (ix.n_keys,))
        st = np.ctypeslib.as_array(ix.start, (ix.n_keys + 1,))
        of = np.ctypeslib.as_array(ix.off, (ix.n_off,))
        assert np.array_equal(keys, np.nonzero(cnt)[0].astype(np.uint32))
        assert np.array_equal(st, starts[np.append(keys, 1 << (2 * k))].astype(np.uint64) if False else
                              np.append(starts[keys.astype(np.int64)], starts[-1]).astype(np.uint64))
        assert np.array_equal(of, offs.astype(np.uint64))
        L.o_index_free(C.byref(ix))
