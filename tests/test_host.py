"""CPU: host-side logic -- file formats, candidate regrouping, the C-ABI surface -- no GPU compute."""
import ctypes as C
import os
import re
import shutil

import numpy as np
import pytest

import util


def test_packeddb_roundtrip(tmp_path):
    from necat_b200 import formats
    rng = np.random.default_rng(1)
    reads = [rng.integers(0, 4, n, dtype=np.uint8) for n in (5, 1, 4097, 33, 1000)]
    vols = formats.write_workdir(str(tmp_path / "w"), reads, vol_size=4000)
    assert len(vols) == 2                                  # a volume closes once it holds >= vol_size bases
    k = 0
    for p in vols:
        v = formats.read_volume(p)
        for i in range(v.nseq):
            assert np.array_equal(v.sequence(i), reads[k])
            assert np.array_equal(v.sequence(i, 1), 3 - reads[k][::-1])
            assert v.name(i) == f"r{k}"
            k += 1
    assert k == len(reads)


def test_golden_volume_readable_by_both_readers(built, tmp_path):
    """formats.read_volume and the C host reader see the same reads as the reference's oc2mkdb wrote them."""
    from necat_b200 import formats
    v = formats.read_volume(os.path.join(util.GOLDEN, "vol0"))
    assert v.nseq == 94 and v.db_size == int(v.info["size"].sum())
    assert v.info["offset"][1] == v.info["size"][0]


def test_partition_matches_oc2pcan(built, tmp_path):
    """formats.partition_records (numpy) == our oc2pcan host == the reference's oc2pcan + loader normalisation."""
    from necat_b200 import formats
    wrk = util.golden_workdir(tmp_path)
    cand = str(tmp_path / "c.bin")
    shutil.copy(os.path.join(util.GOLDEN, "cand_k15.bin"), cand)
    rec = formats.read_candidates(cand)
    both = np.concatenate([rec, formats.swap_roles(rec)])
    ours = os.path.join(util.BIN_DIR, "oc2pcan")
    if os.path.exists(ours):
        util.run([ours, wrk, cand], cwd=str(tmp_path))
        assert open(cand + ".partitions").read().split() == ["1"]
        got = formats.read_candidates(cand + ".p0")
        assert util.same_multiset(got, both)
    ref = util.ref_bin("oc2pcan")
    if ref:
        c2 = str(tmp_path / "c2.bin")
        shutil.copy(os.path.join(util.GOLDEN, "cand_k15.bin"), c2)
        util.run([ref, "-t", "1", wrk, c2], cwd=str(tmp_path))
        assert util.same_multiset(formats.read_candidates(c2 + ".p0"), both)
    # subject made forward, coordinates mirrored
    sizes = formats.read_volume(os.path.join(util.GOLDEN, "vol0")).info["size"].astype(np.int64)
    part = formats.partition_records(rec, sizes)
    assert part.shape[0] == 2 * rec.shape[0] and np.all((part[:, 0] >> 31) == 0)
    assert np.all(part[:, 3] <= sizes[part[:, 1]]) and np.all(part[:, 6] <= sizes[part[:, 4]])


def test_capi_exports_every_declared_symbol(built):
    """The shared library loads without a GPU and exports exactly what include/necat_b200.h declares."""
    from necat_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libnecat_b200.so not built (nvcc missing?)")
    hdr = open(os.path.join(util.ROOT, "include", "necat_b200.h")).read()
    declared = set(re.findall(r"\b(nb_[a-z_]+)\s*\(", hdr))
    lib = C.CDLL(capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared but not exported"
    assert declared == set(capi.EXPORTS), (declared ^ set(capi.EXPORTS))


def test_no_cpu_fallback(built):
    """Without a B200 the product path refuses loudly instead of computing on the host."""
    import torch
    from necat_b200 import capi
    if torch.cuda.is_available() or not os.path.exists(capi.LIB_PATH):
        pytest.skip("needs the library and a machine without GPU")
    with pytest.raises(capi.NbError, match="no CUDA device|no CPU fallback"):
        capi.Context(0)
    p = os.path.join(util.BIN_DIR, "oc2pmov")
    if os.path.exists(p):
        import subprocess
        r = subprocess.run([p, "-j", "0", util.GOLDEN, "0", "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode != 0


def test_product_does_not_touch_the_oracle():
    """Nothing under necat_b200/ or include/ may include, link or execute oracle/ (it is test infrastructure)."""
    bad = []
    for base in ("necat_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(util.ROOT, base)):
            for f in fs:
                if f.endswith((".c", ".h", ".cu", ".cuh", ".py")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle[/_]|liboracle|necat_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    assert "oracle" not in open(os.path.join(util.ROOT, "Makefile")).read()


def test_fasta_order_helper(built, tmp_path):
    """nbh_order_by_id (what oc2cns orders its FASTA records with): ascending ids, stable -- tests/host/test_order.c."""
    exe = str(tmp_path / "test_order")
    util.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-o", exe, os.path.join(util.ROOT, "tests", "host", "test_order.c"),
              os.path.join(util.ROOT, "necat_b200", "host", "nb_host.c")])
    assert util.run([exe]).stdout.strip() == b"ok"


@pytest.mark.parametrize("opts", ["-p 40", "-p 25 -f 2", "-p 1000"])
def test_pcan_many_partitions_match_reference(built, tmp_path, opts):
    """oc2pcan with small partitions (-p reads per partition, -f files per pass; pcan.c:39-146): same partition count and,
    partition by partition, the same records as the reference binary."""
    ref = util.ref_bin("oc2pcan")
    if not ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from necat_b200 import formats
    wrk = util.golden_workdir(tmp_path)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    shutil.copy(os.path.join(util.GOLDEN, "cand_k15.bin"), a)
    shutil.copy(os.path.join(util.GOLDEN, "cand_k15.bin"), b)
    util.run([ref] + opts.split() + ["-t", "2", wrk, a], cwd=str(tmp_path))
    util.run([os.path.join(util.BIN_DIR, "oc2pcan")] + opts.split() + ["-t", "2", wrk, b], cwd=str(tmp_path))
    n = int(open(a + ".partitions").read().split()[0])
    assert open(b + ".partitions").read() == open(a + ".partitions").read()
    total = 0
    for i in range(n):
        ra, rb = formats.read_candidates(a + ".p%d" % i), formats.read_candidates(b + ".p%d" % i)
        assert util.same_multiset(ra, rb), f"partition {i}"
        total += ra.shape[0]
    assert total == 2 * formats.read_candidates(a).shape[0]
    if opts == "-p 40":
        assert n >= 3


def test_swap_candidate_roles_matches_numpy(built):
    """nb_swap_candidate_roles (C-ABI, host code) == formats.swap_roles (numpy restatement of change_pcan_roles,
    gapped_candidate.c:54-69); capi.with_swapped_roles lays out [records, swapped records]."""
    from necat_b200 import capi, formats
    rng = np.random.default_rng(7)
    rec = rng.integers(0, 2**32, size=(1000, 7), dtype=np.uint64).astype(np.uint32)
    both = capi.with_swapped_roles([rec[:400], rec[400:]])
    assert both.shape == (2000, 7)
    assert np.array_equal(both[:1000], rec)
    assert np.array_equal(both[1000:], formats.swap_roles(rec))
    assert capi.with_swapped_roles([rec[:0]]).shape == (0, 7)


def test_pack_candidates_matches_numpy(built):
    """nb_pack_candidates (C-ABI, host code; what Context.map_volume_packed uses) == capi.pack_candidates (numpy restatement
    of pack_gapped_candidate, gapped_candidate.c:13-28), including the score cap and the three flag bits."""
    import ctypes as C
    from necat_b200 import capi
    rng = np.random.default_rng(11)
    n = 5000
    c = np.zeros(n, capi.CANDIDATE)
    for f in ("sid", "qid", "sbeg", "send", "qbeg", "qend", "ssize", "qsize"):
        if f in c.dtype.names:
            c[f] = rng.integers(0, 2**31 - 1, n)
    c["score"] = rng.integers(0, 3_000_000, n)
    c["qdir"] = rng.integers(0, 2, n); c["sdir"] = rng.integers(0, 2, n)
    c["qoff"] = np.where(rng.random(n) < 0.5, c["qbeg"], c["qend"])
    out = np.empty((n, 7), np.uint32)
    capi.lib().nb_pack_candidates(c.ctypes.data_as(C.c_void_p), C.c_uint64(n), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, capi.pack_candidates(c))


def test_volume_append_shifts_whole_words(built, tmp_path):
    """nbh_volume_append (host/nb_host.c; pdb_merge, packed_db.c:419-448): the merged read store of oc2cns, volumes of
    every length mod 4 appended at every bit offset, against a base-by-base merge; stray bits behind the last base of a
    volume file do not leak into the store."""
    src = tmp_path / "t.c"
    src.write_text(r"""
#include "nb_host.h"
#include <string.h>
static unsigned long long rs = 88172645463325252ull;
static unsigned rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }
static int base_at(const unsigned char* p, unsigned long long i) { return (p[i >> 2] >> (((~i) & 3) << 1)) & 3; }
int main(void) {
    for (int trial = 0; trial < 300; ++trial) {
        const int nv = 2 + rnd() % 4;
        unsigned char* all = malloc(1 << 16); unsigned long long total = 0;
        nbh_volume d; memset(&d, 0, sizeof d);
        for (int v = 0; v < nv; ++v) {
            nbh_volume s; memset(&s, 0, sizeof s);
            s.db_size = (trial < 8 && v == 0) ? (unsigned long long)(trial % 4) : 1 + rnd() % 300;
            if (trial % 7 == 0) s.db_size += 64 * (rnd() % 9);
            s.nseq = 1;
            s.offset = calloc(2, 8); s.size = calloc(2, 8); s.hdr_offset = calloc(2, 8);
            s.size[0] = s.db_size; s.hdr = strdup("r"); s.hdr_size = 2;
            const unsigned long long nb = (s.db_size + 3) >> 2;
            s.pac = calloc(nb + 16, 1);
            for (unsigned long long i = 0; i < nb; ++i) s.pac[i] = (unsigned char)rnd();      /* pad bits of the last byte: garbage */
            for (unsigned long long i = 0; i < s.db_size; ++i) all[total + i] = (unsigned char)base_at(s.pac, i);
            total += s.db_size;
            if (v == 0) d = s;
            else { if (nbh_volume_append(&d, &s)) return 2; nbh_volume_free(&s); }
        }
        if (d.db_size != total || d.nseq != (unsigned long long)nv) return 3;
        for (unsigned long long i = 0; i < total; ++i) if (base_at(d.pac, i) != all[i]) { fprintf(stderr, "trial %d base %llu\n", trial, i); return 4; }
        unsigned long long off = 0;
        for (int v = 0; v < nv; ++v) { if (d.offset[v] != off) return 5; off += d.size[v]; }
        if (nv > 1) {       /* behind the last base: zero (the first volume's own pad bits are masked when the next is appended) */
            if ((total & 3) && (d.pac[total >> 2] & (0xff >> (2 * (total & 3))))) return 6;
            for (int k = 0; k < 16; ++k) if (d.pac[((total + 3) >> 2) + k]) return 7;
        }
        nbh_volume_free(&d); free(all);
    }
    puts("ok");
    return 0;
}
""")
    host = os.path.join(util.ROOT, "necat_b200", "host")
    exe = str(tmp_path / "t")
    util.run(["gcc", "-O2", "-std=gnu99", "-I", host, "-o", exe, str(src), os.path.join(host, "nb_host.c")])
    assert util.run([exe]).stdout.strip() == b"ok"


def test_merge_volumes_matches_base_by_base():
    """pipeline.merge_volumes (the merged read store of the correction half, consensus/main.c:66): byte-shifted packed
    volumes == the bases of the volumes one after the other, offsets cumulative, nothing behind the last base."""
    from necat_b200 import formats, pipeline
    rng = np.random.default_rng(3)
    for trial in range(120):
        vols, allb = [], []
        for _ in range(int(rng.integers(2, 5))):
            n = int(rng.integers(0, 300)) if trial % 5 else int(rng.integers(0, 4))
            codes = rng.integers(0, 4, n, dtype=np.uint8)
            pac = formats.pack_bases(codes).copy()
            if n & 3 and pac.size:
                pac[-1] |= int(rng.integers(0, 256)) & (0xff >> (2 * (n & 3)))      # garbage behind the last base
            info = np.zeros(1, formats.SEQINFO)
            info["size"] = n
            vols.append(formats.Volume(pac, n, info, b"r\0"))
            allb.append(codes)
        m = pipeline.merge_volumes(vols)
        exp = np.concatenate(allb)
        assert m.db_size == exp.size and m.pac.size == (exp.size + 3) // 4
        assert np.array_equal(formats.unpack_bases(m.pac, m.db_size), exp)
        if exp.size & 3:
            assert (int(m.pac[-1]) & (0xff >> (2 * (exp.size & 3)))) == 0
        assert np.array_equal(m.info["offset"], np.cumsum([0] + [int(v.db_size) for v in vols])[:-1])
