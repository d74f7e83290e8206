"""CPU: the oracle against the reference binaries run here (oracle/_ref; skipped where they are absent).
Complements the golden vectors with fresh seeded inputs, k = 13 keeps the reference's dense table small."""
import os

import pytest

import util

needs_ref = pytest.mark.skipif(util.ref_bin("oc2pmov") is None, reason="oracle/_ref not built (needs /root/reference)")

PM13 = "-k 13 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 2 -j 0 -u 1 -i 1"


@needs_ref
@pytest.mark.parametrize("gen", [dict(genome=50_000, depth=20, seed=51, err=0.12),
                                 dict(genome=40_000, depth=25, seed=52, err=0.10, repeat_len=2000, repeat_copies=5, n_repeats=2)])
def test_pipeline_matches_reference(built, tmp_path, gen):
    import shutil
    wrk, _ = util.make_workdir(tmp_path, **gen)
    r = util.run_pmov(util.ref_bin("oc2pmov"), PM13, wrk, 0, str(tmp_path / "r.bin"))
    o = util.run_pmov(util.ORACLE_CLI, PM13, wrk, 0, str(tmp_path / "o.bin"))
    assert r.shape[0] > 100 and util.same_multiset(r, o)
    c = str(tmp_path / "c.bin")
    shutil.copy(str(tmp_path / "r.bin"), c)
    util.run([util.ref_bin("oc2pcan"), "-t", "1", wrk, c])
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
    util.run([util.ref_bin("oc2cns")] + cns.split() + ["-t", "2", wrk, c, str(tmp_path / "r_cns.fa"), str(tmp_path / "r_raw.fa")])
    for mode in ("1", "0"):
        util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", mode, wrk, str(tmp_path / "r.bin"),
                  str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
        assert util.fasta_canon(str(tmp_path / "o_cns.fa")) == util.fasta_canon(str(tmp_path / "r_cns.fa")), f"mode {mode}"
        assert util.fasta_canon(str(tmp_path / "o_raw.fa")) == util.fasta_canon(str(tmp_path / "r_raw.fa"))


@needs_ref
def test_rescue_matches_reference(built, tmp_path):
    """oc2cns -r 1 on a fresh seeded set with long indels: oracle (oracle_rescue.c) == reference binary."""
    import shutil
    wrk, _ = util.make_workdir(tmp_path, genome=16_000, depth=16, seed=53, err=0.10, long_indel=0.5)
    r = util.run_pmov(util.ref_bin("oc2pmov"), PM13, wrk, 0, str(tmp_path / "r.bin"))
    assert r.shape[0] > 100
    c = str(tmp_path / "c.bin")
    shutil.copy(str(tmp_path / "r.bin"), c)
    util.run([util.ref_bin("oc2pcan"), "-t", "1", wrk, c])
    out = {}
    for rescue in ("0", "1"):
        cns = f"-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r {rescue} -f 0"
        util.run([util.ref_bin("oc2cns")] + cns.split() + ["-t", "2", wrk, c, str(tmp_path / "r_cns.fa"), str(tmp_path / "r_raw.fa")])
        util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "1", wrk, str(tmp_path / "r.bin"),
                  str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
        out[rescue] = util.fasta_canon(str(tmp_path / "r_cns.fa"))
        assert util.fasta_canon(str(tmp_path / "o_cns.fa")) == out[rescue], f"-r {rescue}"
        assert util.fasta_canon(str(tmp_path / "o_raw.fa")) == util.fasta_canon(str(tmp_path / "r_raw.fa"))
    assert len(out["1"]) > 20 and out["0"] != out["1"], "the set must exercise the rescue"


@needs_ref
@pytest.mark.parametrize("gen", [
    # ragged: most reads below the length cut-offs (-a 1000, -m 500, -l 500), many with no candidate at all
    dict(genome=30_000, depth=12, seed=54, err=0.10, min_len=60, max_len=4000, mean_log=6.6, sd_log=0.8),
    # starved: coverage below -x 3 almost everywhere, so nearly nothing is corrected and templates fall through to raw
    dict(genome=60_000, depth=2, seed=55, err=0.10),
])
def test_edge_sets_match_reference(built, tmp_path, gen):
    """Ragged and starved inputs, and an EMPTY candidate file: oracle == reference binaries, no crashes either side."""
    import shutil
    wrk, reads = util.make_workdir(tmp_path, **gen)
    r = util.run_pmov(util.ref_bin("oc2pmov"), PM13, wrk, 0, str(tmp_path / "r.bin"))
    o = util.run_pmov(util.ORACLE_CLI, PM13, wrk, 0, str(tmp_path / "o.bin"))
    assert util.same_multiset(r, o)
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
    for name, src in (("full", str(tmp_path / "r.bin")), ("empty", None)):
        plain = str(tmp_path / f"{name}.bin")
        if src:
            shutil.copy(src, plain)
        else:
            open(plain, "wb").close()
        c = str(tmp_path / f"{name}_part.bin")
        shutil.copy(plain, c)
        util.run([util.ref_bin("oc2pcan"), "-t", "1", wrk, c])
        util.run([util.ref_bin("oc2cns")] + cns.split() + ["-t", "2", wrk, c, str(tmp_path / "r_cns.fa"), str(tmp_path / "r_raw.fa")])
        util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "1", wrk, plain, str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
        assert util.fasta_canon(str(tmp_path / "o_cns.fa")) == util.fasta_canon(str(tmp_path / "r_cns.fa")), name
        assert util.fasta_canon(str(tmp_path / "o_raw.fa")) == util.fasta_canon(str(tmp_path / "r_raw.fa")), name


@needs_ref
def test_many_partitions_match_reference(built, tmp_path):
    """oc2pcan -p 40 cuts the golden set into several partitions; oc2cns then handles each on its own, and what it dumps
    as uncorrected depends on the cut (every id of [min_sid, max_sid) of a partition, consensus_one_partition.c:173-194).
    Run as two nodes (-mn 0 2, -mn 1 2; consensus/main.c:37-41,71-73) and concatenated, as necat.pl does."""
    import shutil
    wrk = util.golden_workdir(tmp_path)
    cand = os.path.join(util.GOLDEN, "cand_k13.bin")
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([util.ref_bin("oc2pcan"), "-p", "40", "-t", "1", wrk, c], cwd=str(tmp_path))
    assert int(open(c + ".partitions").read().split()[0]) >= 3
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
    got_c, got_r = [], []
    for node in (0, 1):
        util.run([util.ref_bin("oc2cns")] + cns.split() + ["-t", "2", wrk, c, "n_cns.fa", "n_raw.fa", "-mn", str(node), "2"], cwd=str(tmp_path))
        got_c += util.fasta_canon(str(tmp_path / "n_cns.fa"))
        got_r += util.fasta_canon(str(tmp_path / "n_raw.fa"))
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "1", "-P", "40", wrk, cand, "o_cns.fa", "o_raw.fa"], cwd=str(tmp_path))
    assert sorted(got_c) == util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert sorted(got_r) == util.fasta_canon(str(tmp_path / "o_raw.fa"))
    # and the cut matters: one partition dumps a different set of uncorrected reads
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "1", wrk, cand, "o1_cns.fa", "o1_raw.fa"], cwd=str(tmp_path))
    assert sorted(got_c) == util.fasta_canon(str(tmp_path / "o1_cns.fa"))


@needs_ref
def test_candidate_text_output_matches_reference(built, tmp_path):
    """oc2pmov -j 0 -u 0: the text form of the UNPACKED candidates (gapped_candidate.h:26-41, pm_worker.c:180-185), which
    still carries the seed position the 28-byte record drops."""
    wrk = util.golden_workdir(tmp_path)
    opts = PM13.replace("-u 1", "-u 0")
    util.run([util.ref_bin("oc2pmov")] + opts.split() + [wrk, "0", "r.txt"], cwd=str(tmp_path))
    util.run([util.ORACLE_CLI, "pmov"] + opts.split() + [wrk, "0", "o.txt"], cwd=str(tmp_path))
    r, o = sorted(open(str(tmp_path / "r.txt")).readlines()), sorted(open(str(tmp_path / "o.txt")).readlines())
    assert len(r) > 500 and r == o and all(len(x.split("\t")) == 13 for x in r)
