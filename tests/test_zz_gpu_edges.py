"""GPU parity on the edge sets the oracle is pinned on (tests/test_oracle_vs_ref.py::test_edge_sets_match_reference):
ragged reads (most below the length cut-offs, some shorter than a k-mer), starved coverage, an empty candidate file.
Runs last (file name) -- it was written after round 1's GPU budget was spent, its first run is the round-end run."""
import os
import shutil

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

PM13 = "-k 13 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 2 -j 0 -u 1 -i 1"
CNS = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


def _cns_pair(tmp_path, wrk, cand, tag):
    c = str(tmp_path / f"{tag}_part.bin")
    shutil.copy(cand, c)
    util.run([os.path.join(util.BIN_DIR, "oc2pcan"), "-t", "1", wrk, c])
    util.run([os.path.join(util.BIN_DIR, "oc2cns")] + CNS.split() + [wrk, c, str(tmp_path / f"{tag}_g_cns.fa"), str(tmp_path / f"{tag}_g_raw.fa")])
    util.run([util.ORACLE_CLI, "cns"] + CNS.split() + ["-M", "0", wrk, cand, str(tmp_path / f"{tag}_o_cns.fa"), str(tmp_path / f"{tag}_o_raw.fa")])
    for f in ("cns", "raw"):
        assert util.fasta_canon(str(tmp_path / f"{tag}_g_{f}.fa")) == util.fasta_canon(str(tmp_path / f"{tag}_o_{f}.fa")), f"{tag} {f}"
    return len(util.fasta_canon(str(tmp_path / f"{tag}_o_cns.fa")))


@pytest.mark.parametrize("gen", [
    dict(genome=30_000, depth=12, seed=54, err=0.10, min_len=60, max_len=4000, mean_log=6.6, sd_log=0.8),
    dict(genome=60_000, depth=2, seed=55, err=0.10),
])
def test_edge_sets_match_oracle(built, tmp_path, gen):
    wrk, reads = util.make_workdir(tmp_path, **gen)
    g = util.run_pmov(os.path.join(util.BIN_DIR, "oc2pmov"), PM13, wrk, 0, str(tmp_path / "g.bin"))
    o = util.run_pmov(util.ORACLE_CLI, PM13, wrk, 0, str(tmp_path / "o.bin"))
    assert o.shape[0] > 10 and util.same_multiset(g, o)
    assert _cns_pair(tmp_path, wrk, str(tmp_path / "o.bin"), "full") > 5
    empty = str(tmp_path / "empty.bin")
    open(empty, "wb").close()
    assert _cns_pair(tmp_path, wrk, empty, "empty") == 0


def test_reads_shorter_than_a_kmer(built, tmp_path):
    """Reads of 1..40 bases among normal ones: no probe, no index entry, no candidate for them -- and no crash."""
    from necat_b200 import formats
    import gen_reads
    _, reads = gen_reads.generate(30_000, 14, 56, err=0.10)
    rng = np.random.default_rng(3)
    tiny = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in (1, 2, 5, 12, 13, 14, 40)]
    mixed = []
    for i, r in enumerate(reads):
        mixed.append(r)
        if i < len(tiny):
            mixed.append(tiny[i])
    wrk = str(tmp_path / "vol")
    formats.write_workdir(wrk, mixed)
    g = util.run_pmov(os.path.join(util.BIN_DIR, "oc2pmov"), PM13, wrk, 0, str(tmp_path / "g.bin"))
    o = util.run_pmov(util.ORACLE_CLI, PM13, wrk, 0, str(tmp_path / "o.bin"))
    assert o.shape[0] > 50 and util.same_multiset(g, o)
    assert _cns_pair(tmp_path, wrk, str(tmp_path / "o.bin"), "tiny") > 5


def test_many_partitions_two_nodes(built, tmp_path):
    """oc2pcan -p 40 + oc2cns as two nodes (-mn 0 2, -mn 1 2; consensus/main.c:37-41,71-73), outputs concatenated as
    necat.pl does, against the oracle cut the same way (the CPU twin of this test pins the oracle on the reference)."""
    wrk = util.golden_workdir(tmp_path)
    cand = os.path.join(util.GOLDEN, "cand_k13.bin")
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([os.path.join(util.BIN_DIR, "oc2pcan"), "-p", "40", "-t", "1", wrk, c], cwd=str(tmp_path))
    assert int(open(c + ".partitions").read().split()[0]) >= 3
    got_c, got_r = [], []
    for node in (0, 1):
        util.run([os.path.join(util.BIN_DIR, "oc2cns")] + CNS.split() + [wrk, c, "n_cns.fa", "n_raw.fa", "-mn", str(node), "2"], cwd=str(tmp_path))
        got_c += util.fasta_canon(str(tmp_path / "n_cns.fa"))
        got_r += util.fasta_canon(str(tmp_path / "n_raw.fa"))
    util.run([util.ORACLE_CLI, "cns"] + CNS.split() + ["-M", "0", "-P", "40", wrk, cand, "o_cns.fa", "o_raw.fa"], cwd=str(tmp_path))
    assert sorted(got_c) == util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert sorted(got_r) == util.fasta_canon(str(tmp_path / "o_raw.fa"))


def test_candidate_text_output(built, tmp_path):
    """oc2pmov -j 0 -u 0 (text candidates with the seed position, gapped_candidate.h:26-41) against the oracle."""
    wrk = util.golden_workdir(tmp_path)
    opts = PM13.replace("-u 1", "-u 0")
    util.run([os.path.join(util.BIN_DIR, "oc2pmov")] + opts.split() + [wrk, "0", "g.txt"], cwd=str(tmp_path))
    util.run([util.ORACLE_CLI, "pmov"] + opts.split() + [wrk, "0", "o.txt"], cwd=str(tmp_path))
    g, o = sorted(open(str(tmp_path / "g.txt")).readlines()), sorted(open(str(tmp_path / "o.txt")).readlines())
    assert len(o) > 500 and g == o
