#!/usr/bin/env python3
"""Golden fixtures for the ingest / reorder hosts (oc2mkdb, oc2ReorderCnsReads), made by running the
UNMODIFIED reference binaries (oracle/_ref) on small hand-made inputs that hit the parser's corners.

    python tests/golden/make_golden_host.py      # needs oracle/_ref

Written to tests/golden/host/:
    a.fa b.fq.gz c.fq d.fa e.fa list1 list2      inputs of oc2mkdb (multi-line FASTA with comments / lower case / N / '-' /
                                                 CRLF / empty lines / empty records / no final newline; FASTQ with quality
                                                 lines that start with '@' or '>', multi-line records; a truncated FASTQ; an
                                                 empty file) and the two list files
    vol0.masked reads_info.txt volume_names.txt  what the reference's oc2mkdb wrote (the 4 uninitialised padding bytes of
                                                 every 32-byte sequence record zeroed)
    cns_in.fa raw_in.fa                          inputs of oc2ReorderCnsReads: ids out of order, many records per id
    cns_out.fa raw_out.fa                        what the reference wrote
"""
import gzip
import os
import shutil
import struct
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import util  # noqa: E402


def mask_pads(buf):
    b = bytearray(buf)
    ns = struct.unpack_from("<Q", b, 31)[0]
    for i in range(ns):
        o = 31 + 16 + i * 32 + 28
        b[o:o + 4] = b"\0\0\0\0"
    return bytes(b)


def write_inputs(d, seed=5):
    rng = np.random.default_rng(seed)

    def seq(n, alpha=b"ACGT"):
        return bytes(rng.choice(np.frombuffer(alpha, np.uint8), n))
    with open(os.path.join(d, "a.fa"), "wb") as f:
        f.write(b"garbage before the first header\n")
        f.write(b">r0 some comment\n" + seq(70) + b"\n" + seq(33) + b"\n")
        f.write(b">r1\tc\n" + seq(10, b"acgtnN-") + b"\r\n" + seq(7) + b"\r\n")
        f.write(b">r2\n\n\n" + seq(5) + b"\n\n")
        f.write(b">empty\n")
        f.write(b">\n" + seq(9) + b"\n")
        f.write(b">r5 x\n" + seq(1001, b"ACGTN") + b"\n")
        f.write(b">r6\n\r\n" + seq(3) + b"\n")
        f.write(b">last\n" + seq(41))
    with gzip.open(os.path.join(d, "b.fq.gz"), "wb") as f:
        for i in range(50):
            n = int(rng.integers(1, 300))
            s = seq(n)
            q = bytes(rng.integers(33, 74, n, dtype=np.uint8))
            if i % 7 == 0:
                q = b"@" + q[1:]
            if i % 11 == 0:
                q = b">" + q[1:]
            if i % 5 == 0 and n > 10:
                f.write(b"@q%d desc\n" % i + s[:n // 2] + b"\n" + s[n // 2:] + b"\n+\n" + q[:n // 3] + b"\n" + q[n // 3:] + b"\n")
            else:
                f.write(b"@q%d\n" % i + s + b"\n+q%d\n" % i + q + b"\n")
    with open(os.path.join(d, "c.fq"), "wb") as f:
        f.write(b"@t0\nACGTACGT\n+\nIIIIIIII\n@t1\nACGTAC\n+\nIII\n@t2\nACGT\n+\nIIII\n@t3\nACGTA\n+\nII\n")
    with open(os.path.join(d, "d.fa"), "wb") as f:
        for i in range(40):
            f.write(b">d%d\n" % i + seq(int(rng.integers(100, 3000))) + b"\n")
    open(os.path.join(d, "e.fa"), "wb").close()
    open(os.path.join(d, "list1"), "w").write("a.fa\nb.fq.gz\n")
    open(os.path.join(d, "list2"), "w").write("c.fq\nd.fa\ne.fa\n")

    def fasta(path, ids, alpha=b"ACGT"):
        with open(os.path.join(d, path), "wb") as f:
            for k, i in enumerate(ids):
                n = int(rng.integers(0, 400)) if k % 37 else 0
                f.write(b">ontsa_id_%d_(%d_%d_%d_%d_%d_%d_0.123)\n" % (i, k, k + n, n, n + 5, 7, 9) + seq(n, alpha) + b"\n")
    fasta("cns_in.fa", rng.integers(0, 60, 200))
    fasta("raw_in.fa", rng.permutation(90), alpha=b"ACGTacgtN")


def main():
    assert util.ref_bin("oc2mkdb") and util.ref_bin("oc2ReorderCnsReads"), "build oracle/_ref first (make -C oracle ref)"
    out = os.path.join(HERE, "host")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    write_inputs(out)
    td = tempfile.mkdtemp()
    try:
        subprocess.check_call([util.ref_bin("oc2mkdb"), os.path.join(td, "w"), "list1", "list2"], cwd=out, stdout=subprocess.DEVNULL)
        assert open(os.path.join(td, "w", "reads_info.txt")).read().split()[0] == "1"
        open(os.path.join(out, "vol0.masked"), "wb").write(mask_pads(open(os.path.join(td, "w", "vol0"), "rb").read()))
        shutil.copy(os.path.join(td, "w", "reads_info.txt"), out)
        open(os.path.join(out, "volume_names.txt"), "w").write(open(os.path.join(td, "w", "volume_names.txt")).read().replace(td, "TMP"))
        subprocess.check_call([util.ref_bin("oc2ReorderCnsReads"), "cns_in.fa", "raw_in.fa", "cns_out.fa", "raw_out.fa"], cwd=out)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    print("wrote", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
