"""CPU: the two host-only helpers necat.pl runs around a correction iteration -- this repo's drop-ins of
`fsa_rd_tools longest|n50` (fsa/read_tools.cpp:43-76,196-242) and of the vendored pigz -- against the reference's own
fsa_rd_tools (built into oracle/_ref/pipeline by `make -C oracle ref`) and gzip."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import util

FSA = os.path.join(util.BIN_DIR, "fsa_rd_tools")
PIGZ = os.path.join(util.BIN_DIR, "pigz")
REF_FSA = os.path.join(util.REF_DIR, "pipeline", "fsa_rd_tools")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_FSA), reason="oracle/_ref/pipeline is not built (make -C oracle ref, needs /root/reference)")


def _inputs(d):
    """FASTA with every liberty the reference's reader allows (lines of any width, blank lines, white space around lines,
    comments, a header with blanks before the name, lower case, bases outside ACGT, a record without sequence), gzipped
    FASTA, FASTQ (plain and gzipped, '+' lines with and without the name), and a list file naming all of them."""
    rng = np.random.default_rng(5)

    def seq(n, alpha="ACGT"):
        return "".join(rng.choice(list(alpha), n))

    with open(os.path.join(d, "a.fasta"), "w") as f:
        for i in range(300):
            n = int(rng.integers(50, 4000))
            s = (seq(n, "ACGT" if i % 7 else "ACGTN") if i % 11 else seq(n, "acgt"))
            hdr = ">r%d" % i
            if i % 3 == 0:
                hdr += " some comment  here "
            if i % 5 == 0:
                hdr = ">  spaced%d\tx y" % i
            f.write(hdr + "\n")
            w = int(rng.integers(30, 200))
            for k in range(0, n, w):
                line = s[k:k + w]
                if i % 13 == 0:
                    line = "  " + line + " \t"
                f.write(line + "\n")
                if i % 17 == 0:
                    f.write("\n   \n")
        f.write(">empty_seq\n>after_empty\nACGTACGT\n")
    with gzip.open(os.path.join(d, "b.fasta.gz"), "wt") as f:
        for i in range(200):
            f.write(">g%d\n%s\n" % (i, seq(int(rng.integers(100, 6000)))))
    with open(os.path.join(d, "c.fastq"), "w") as f:
        for i in range(150):
            n = int(rng.integers(100, 3000))
            f.write("@q%d extra\n%s\n+\n%s\n" % (i, seq(n, "ACGT" if i % 9 else "ACGTRY"), "I" * n))
    with gzip.open(os.path.join(d, "d.fastq.gz"), "wt") as f:
        for i in range(100):
            n = int(rng.integers(100, 3000))
            f.write("@z%d\n%s\n+z%d\n%s\n" % (i, seq(n), i, "@" * n))
    open(os.path.join(d, "list.txt"), "w").write("".join(os.path.join(d, x) + "\n" for x in ("a.fasta", "b.fasta.gz", "c.fastq", "d.fastq.gz")).replace("\n", "\n\n", 1))
    open(os.path.join(d, "e.fa"), "w").write(">x1\nACGT\n>x2\nAC\nGT\nNN\n")


@needs_ref
@pytest.mark.parametrize("args,out", [
    (["--ifname", "list.txt", "--min_length", "1000", "--base_size", "300000", "--discard_illegal_read"], "o.fasta"),    # necat.pl:122
    (["--ifname", "list.txt", "--min_length", "3000", "--base_size", "300000", "--discard_illegal_read"], "o.fasta"),
    (["--ifname=list.txt", "--base_size=100000"], "o.fasta"),                                                               # Necat.pm:22
    (["--ifname=list.txt", "--base_size=1"], "o.fasta"),
    (["--ifname=list.txt", "--base_size=100000000"], "o.fasta"),
    (["--ifname", "a.fasta", "--min_length", "500"], "o.fasta"),
    (["--ifname", "a.fasta", "--base_size", "50000", "--discard_illegal_read"], "o.fasta"),
    (["--ifname", "c.fastq", "--base_size", "50000", "--discard_illegal_read"], "o.fasta"),
    (["--ifname", "d.fastq.gz", "--base_size", "50000"], "o.fasta"),
    (["--ifname", "e.fa", "--base_size", "5", "--discard_illegal_read"], "o.fasta"),
    (["--ifname", "list.txt", "--min_length", "1000", "--base_size", "300000", "--discard_illegal_read"], "o.fasta.gz"),
])
def test_longest_matches_reference(built, tmp_path, args, out):
    d = str(tmp_path)
    _inputs(d)
    util.run([REF_FSA, "longest"] + args + ["--ofname", "ref_" + out], cwd=d)
    util.run([FSA, "longest"] + args + ["--ofname", "our_" + out], cwd=d)
    a, b = open(os.path.join(d, "ref_" + out), "rb").read(), open(os.path.join(d, "our_" + out), "rb").read()
    assert len(a) > 5 and a == b          # (byte for byte, the gzipped form too: same zlib, same bytes in)


@needs_ref
@pytest.mark.parametrize("args", [["--ifname", "list.txt"], ["--ifname", "b.fasta.gz", "--genome_size", "300000"], ["--ifname", "c.fastq"]])
def test_n50_matches_reference(built, tmp_path, args):
    d = str(tmp_path)
    _inputs(d)
    a = util.run([REF_FSA, "n50"] + args, cwd=d).stdout
    b = util.run([FSA, "n50"] + args, cwd=d).stdout
    assert b"N50: " in a and a == b


def test_longest_stops_on_an_empty_input(built, tmp_path):
    """The reference's tool dies on an assertion here (utility.cpp:47); the drop-in says why and fails as well."""
    open(str(tmp_path / "none.fasta"), "w").write("")
    p = subprocess.run([FSA, "longest", "--ifname", "none.fasta", "--ofname", "o.fasta", "--base_size", "100"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode != 0 and b"no read" in p.stderr


def test_pigz_round_trip(built, tmp_path):
    """pigz -f -p N file -> file.gz (members of 4 MiB, written in order), file removed; what zlib's gzread, gzip -d and the
    reference's fsa_rd_tools read back; an empty file; -k keeps the input; an existing output needs -f."""
    rng = np.random.default_rng(3)
    data = np.frombuffer(b"ACGT\n>", np.uint8)[rng.integers(0, 6, 9_500_000)].tobytes()      # three blocks
    p = str(tmp_path / "x.fasta")
    for threads in ("1", "3", "8"):
        open(p, "wb").write(data)
        util.run([PIGZ, "-f", "-p", threads, p])
        assert not os.path.exists(p)
        assert gzip.open(p + ".gz").read() == data
        assert subprocess.run(["gzip", "-dc", p + ".gz"], capture_output=True).stdout == data
    open(p, "wb").write(b"again")
    assert subprocess.run([PIGZ, "-p", "2", p], capture_output=True).returncode != 0 and os.path.exists(p)     # x.fasta.gz exists
    util.run([PIGZ, "-f", "-k", "-p", "2", p])
    assert os.path.exists(p) and gzip.open(p + ".gz").read() == b"again"
    e = str(tmp_path / "empty")
    open(e, "wb").close()
    util.run([PIGZ, "-f", "-p", "4", e])
    assert gzip.open(e + ".gz").read() == b""
