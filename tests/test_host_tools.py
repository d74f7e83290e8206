"""CPU: the ingest / reorder hosts either side of the hot path -- necat_b200/bin/oc2mkdb and
necat_b200/bin/oc2ReorderCnsReads -- against golden files written by the reference binaries and, where oracle/_ref is
built, against those binaries on fresh inputs.  Volume files must be byte-identical (apart from the reference's
uninitialised padding bytes); reordered FASTA must have the same id sequence and, per id, the same records (the
reference's order among equal ids is whatever its unstable sort leaves)."""
import os
import re
import shutil
import struct
import subprocess
from collections import defaultdict

import numpy as np
import pytest

import util

HOST = os.path.join(util.GOLDEN, "host")
MKDB = os.path.join(util.BIN_DIR, "oc2mkdb")
REORDER = os.path.join(util.BIN_DIR, "oc2ReorderCnsReads")
needs_ref = pytest.mark.skipif(util.ref_bin("oc2mkdb") is None, reason="oracle/_ref not built (needs /root/reference)")


def mask_pads(buf):
    b = bytearray(buf)
    ns = struct.unpack_from("<Q", b, 31)[0]
    for i in range(ns):
        o = 31 + 16 + i * 32 + 28
        b[o:o + 4] = b"\0\0\0\0"
    return bytes(b)


def fasta(path):
    recs = []
    for line in open(path, "rb"):
        if line[:1] == b">":
            recs.append([line.rstrip(b"\n"), b""])
        else:
            recs[-1][1] += line.rstrip(b"\n")
    return recs


def rec_id(h):
    return int(re.search(rb"\d+", h).group())


def same_reorder(a_path, b_path):
    a, b = fasta(a_path), fasta(b_path)
    ia, ib = [rec_id(h) for h, _ in a], [rec_id(h) for h, _ in b]
    assert ia == ib and ia == sorted(ia)
    da, db = defaultdict(list), defaultdict(list)
    for h, s in a:
        da[rec_id(h)].append((h, s))
    for h, s in b:
        db[rec_id(h)].append((h, s))
    assert all(sorted(da[k]) == sorted(db[k]) for k in da)
    return len(set(ia)) == len(ia)          # unique ids: the files must then be identical byte for byte


def copy_inputs(tmp_path):
    for f in os.listdir(HOST):
        shutil.copy(os.path.join(HOST, f), str(tmp_path))


def test_mkdb_matches_golden(built, tmp_path):
    copy_inputs(tmp_path)
    p = util.run([MKDB, "w", "list1", "list2"], cwd=str(tmp_path))
    assert b"file_list: list1" in p.stdout and b"pack a.fa: 8 reads, 1180 bps" in p.stdout
    w = str(tmp_path / "w")
    assert open(os.path.join(w, "reads_info.txt")).read() == open(os.path.join(HOST, "reads_info.txt")).read()
    assert open(os.path.join(w, "volume_names.txt")).read() == open(os.path.join(HOST, "volume_names.txt")).read().replace("TMP/w", "w")
    assert mask_pads(open(os.path.join(w, "vol0"), "rb").read()) == open(os.path.join(HOST, "vol0.masked"), "rb").read()
    # and the volume is what the hot path's own readers expect
    from necat_b200 import formats
    v = formats.read_volume(os.path.join(w, "vol0"))
    assert v.nseq == 100 and v.name(0) == "r0" and v.name(3) == "empty" and int(v.info["size"][3]) == 0


def test_mkdb_errors(built, tmp_path):
    assert subprocess.run([MKDB, "w"], cwd=str(tmp_path), capture_output=True).returncode == 1          # usage
    open(str(tmp_path / "l"), "w").write("missing.fa\n")
    p = subprocess.run([MKDB, "w", "l"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode == 1 and b"ERROR" in p.stderr
    # a gzip stream cut short is an error, not a silently shorter volume
    import gzip
    with gzip.open(str(tmp_path / "x.fa.gz"), "wb") as f:
        f.write(b">x\n" + bytes(np.random.default_rng(1).choice(np.frombuffer(b"ACGT", np.uint8), 200_000)) + b"\n")
    data = open(str(tmp_path / "x.fa.gz"), "rb").read()
    open(str(tmp_path / "cut.fa.gz"), "wb").write(data[:len(data) // 2])
    open(str(tmp_path / "l"), "w").write("cut.fa.gz\n")
    p = subprocess.run([MKDB, "w", "l"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode == 1 and b"read error" in p.stderr


def test_mkdb_volume_split(built, tmp_path):
    """The volume cap (2 000 000 000 bases in the reference, makedb/main.c:8; checked once at full size against the
    reference binary) scaled down: a volume closes after the read that takes it to the cap or beyond; the
    volumes together hold what one volume would."""
    from necat_b200 import formats
    rng = np.random.default_rng(7)
    reads = [rng.integers(0, 4, int(rng.integers(50, 2000)), dtype=np.uint8) for _ in range(300)]
    with open(str(tmp_path / "in.fa"), "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">s%d\n" % i + np.frombuffer(b"ACGT", np.uint8)[r].tobytes() + b"\n")
    open(str(tmp_path / "l"), "w").write("in.fa\n")
    util.run([MKDB, "one", "l"], cwd=str(tmp_path))
    util.run([MKDB, "many", "l"], cwd=str(tmp_path), env=dict(os.environ, NB_MKDB_VOLUME_SIZE="50000"))
    formats.write_workdir(str(tmp_path / "py"), reads, names=["s%d" % i for i in range(len(reads))], vol_size=50000)
    nv, nr = (int(x) for x in open(str(tmp_path / "many" / "reads_info.txt")).read().split())
    assert nr == 300 and nv > 3
    assert open(str(tmp_path / "many" / "volume_names.txt")).read().replace("many/", "") == \
        open(str(tmp_path / "py" / "volume_names.txt")).read().replace(str(tmp_path / "py") + "/", "")
    one = formats.read_volume(str(tmp_path / "one" / "vol0"))
    codes = formats.unpack_bases(one.pac, one.db_size)
    got, first = [], 0
    for v in range(nv):
        x = formats.read_volume(str(tmp_path / "many" / ("vol%d" % v)))
        assert mask_pads(open(str(tmp_path / "many" / ("vol%d" % v)), "rb").read()) == mask_pads(open(str(tmp_path / "py" / ("vol%d" % v)), "rb").read())
        if v < nv - 1:
            assert x.db_size >= 50000 and x.db_size - int(x.info["size"][-1]) < 50000
        got.append(formats.unpack_bases(x.pac, x.db_size))
        first += x.nseq
    assert first == 300 and np.array_equal(np.concatenate(got), codes)


@needs_ref
def test_mkdb_matches_reference(built, tmp_path):
    rng = np.random.default_rng(11)
    A = np.frombuffer(b"ACGTacgtNn", np.uint8)
    import gzip
    with open(str(tmp_path / "p.fa"), "wb") as f:
        for i in range(500):
            s = A[rng.integers(0, 8 if i % 9 else 10, int(rng.integers(0, 30000)))].tobytes()
            w = int(rng.integers(40, 200))
            f.write(b">p%d extra words\n" % i + b"\n".join(s[k:k + w] for k in range(0, len(s), w)) + b"\n")
    with gzip.open(str(tmp_path / "q.fq.gz"), "wb", compresslevel=1) as f:
        for i in range(500):
            n = int(rng.integers(1, 20000))
            f.write(b"@q%d\n" % i + A[rng.integers(0, 4, n)].tobytes() + b"\n+\n" + bytes(rng.integers(33, 90, n, dtype=np.uint8)) + b"\n")
    open(str(tmp_path / "l1"), "w").write("p.fa\n")
    open(str(tmp_path / "l2"), "w").write("q.fq.gz\np.fa\n")
    util.run([util.ref_bin("oc2mkdb"), "ref_w", "l1", "l2"], cwd=str(tmp_path))
    util.run([MKDB, "our_w", "l1", "l2"], cwd=str(tmp_path))
    for f in ("reads_info.txt", "volume_names.txt"):
        assert open(str(tmp_path / "ref_w" / f)).read().replace("ref_w", "W") == open(str(tmp_path / "our_w" / f)).read().replace("our_w", "W")
    assert mask_pads(open(str(tmp_path / "ref_w" / "vol0"), "rb").read()) == mask_pads(open(str(tmp_path / "our_w" / "vol0"), "rb").read())


def test_reorder_matches_golden(built, tmp_path):
    copy_inputs(tmp_path)
    util.run([REORDER, "cns_in.fa", "raw_in.fa", "o_cns.fa", "o_raw.fa"], cwd=str(tmp_path))
    assert not same_reorder(os.path.join(HOST, "cns_out.fa"), str(tmp_path / "o_cns.fa"))       # many records per id
    assert same_reorder(os.path.join(HOST, "raw_out.fa"), str(tmp_path / "o_raw.fa"))           # unique ids ...
    assert open(os.path.join(HOST, "raw_out.fa"), "rb").read() == open(str(tmp_path / "o_raw.fa"), "rb").read()   # ... identical files
    # sorted input (what this repo's oc2cns writes) is streamed through: same bytes again, and the tool says so
    p = util.run([REORDER, "o_cns.fa", "o_raw.fa", "o2_cns.fa", "o2_raw.fa"], cwd=str(tmp_path), env=dict(os.environ, NB_TRACE="1"))
    assert p.stderr.count(b"already in id order: streamed") == 2
    for f in ("cns", "raw"):
        assert open(str(tmp_path / ("o_%s.fa" % f)), "rb").read() == open(str(tmp_path / ("o2_%s.fa" % f)), "rb").read()


def test_reorder_edges(built, tmp_path):
    assert subprocess.run([REORDER, "a", "b", "c"], cwd=str(tmp_path), capture_output=True).returncode == 1       # usage
    open(str(tmp_path / "e.fa"), "wb").close()
    util.run([REORDER, "e.fa", "e.fa", "o1.fa", "o2.fa"], cwd=str(tmp_path))
    assert os.path.getsize(str(tmp_path / "o1.fa")) == 0 and os.path.getsize(str(tmp_path / "o2.fa")) == 0
    open(str(tmp_path / "noid.fa"), "wb").write(b">abc\nACGT\n")
    p = subprocess.run([REORDER, "noid.fa", "e.fa", "o1.fa", "o2.fa"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode == 1 and b"no numeric id" in p.stderr
    # order broken only at the very end: the streamed prefix must be discarded, not duplicated
    with open(str(tmp_path / "late.fa"), "wb") as f:
        for i in list(range(1, 200)) + [0]:
            f.write(b">ontsa_id_%d_x\n" % i + b"ACGT" * (i % 7) + b"\n")
    util.run([REORDER, "late.fa", "e.fa", "o1.fa", "o2.fa"], cwd=str(tmp_path))
    assert [rec_id(h) for h, _ in fasta(str(tmp_path / "o1.fa"))] == list(range(200))


@needs_ref
def test_reorder_matches_reference(built, tmp_path):
    rng = np.random.default_rng(13)
    A = np.frombuffer(b"ACGT", np.uint8)
    ids = np.arange(3000)
    for s in range(0, 3000, 64):
        rng.shuffle(ids[s:s + 64])              # what cat of the reference's per-thread output looks like
    with open(str(tmp_path / "cns.fa"), "wb") as f, open(str(tmp_path / "raw.fa"), "wb") as g:
        for i in ids:
            for frag in range(int(rng.integers(1, 4))):
                f.write(b">ontsa_id_%d_(%d_1_2_3_4_5_0.5)\n" % (i, frag) + A[rng.integers(0, 4, int(rng.integers(500, 9000)))].tobytes() + b"\n")
            if i % 3 == 0:
                g.write(b">r%d_(0_0_0_0_0_0_0)\n" % i + A[rng.integers(0, 4, int(rng.integers(1000, 3000)))].tobytes() + b"\n")
    util.run([util.ref_bin("oc2ReorderCnsReads"), "cns.fa", "raw.fa", "r_cns.fa", "r_raw.fa"], cwd=str(tmp_path))
    util.run([REORDER, "cns.fa", "raw.fa", "o_cns.fa", "o_raw.fa"], cwd=str(tmp_path))
    same_reorder(str(tmp_path / "r_cns.fa"), str(tmp_path / "o_cns.fa"))
    assert same_reorder(str(tmp_path / "r_raw.fa"), str(tmp_path / "o_raw.fa"))
    assert open(str(tmp_path / "r_raw.fa"), "rb").read() == open(str(tmp_path / "o_raw.fa"), "rb").read()


@needs_ref
def test_mkdb_parser_fuzz(built, tmp_path):
    """Differential fuzz of the FASTA / FASTQ reader: random concatenations of header characters, line ends (\\n, \\r\\n,
    lone \\r), blanks, bases, quality-like runs -- whatever the reference's parser makes of them, oc2mkdb must make the
    same volume (1 900 cases of this generator were run once; a slice of them runs here)."""
    rng = np.random.default_rng(20261017)
    tokens = [b">", b"@", b"+", b"\n", b"\r\n", b"\r", b" ", b"\t", b"ACGT", b"acgt", b"N", b"-", b"A", b"C", b"G", b"T", b"x1", b"22",
              b"IIII", b"!!", b"\n\n", b">r", b"@q", b"+\n", b"ACGTACGTACGTAGCTAGCTAGCATCGATCGA\n"]
    w = np.array([3, 3, 2, 6, 1, 1, 2, 1, 6, 2, 1, 1, 2, 2, 2, 2, 2, 2, 2, 1, 1, 3, 3, 2, 4], float)
    w /= w.sum()
    d = str(tmp_path)
    open(os.path.join(d, "l"), "w").write("in.fx\n")
    for case in range(120):
        data = b"".join(tokens[i] for i in rng.choice(len(tokens), int(rng.integers(0, 60)), p=w))
        if rng.random() < 0.5:
            data = b">s0\n" + data
        open(os.path.join(d, "in.fx"), "wb").write(data)
        for x in ("rw", "ow"):
            shutil.rmtree(os.path.join(d, x), ignore_errors=True)
        pr = subprocess.run([util.ref_bin("oc2mkdb"), "rw", "l"], cwd=d, capture_output=True)
        po = subprocess.run([MKDB, "ow", "l"], cwd=d, capture_output=True)
        assert pr.returncode == po.returncode == 0, data
        assert open(os.path.join(d, "rw", "reads_info.txt")).read() == open(os.path.join(d, "ow", "reads_info.txt")).read(), data
        has = os.path.exists(os.path.join(d, "rw", "vol0"))
        assert has == os.path.exists(os.path.join(d, "ow", "vol0")), data
        if has:
            assert mask_pads(open(os.path.join(d, "rw", "vol0"), "rb").read()) == mask_pads(open(os.path.join(d, "ow", "vol0"), "rb").read()), data
