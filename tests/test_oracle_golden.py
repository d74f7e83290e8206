"""CPU: the oracle (CPU restatement) against the golden vectors produced by the unmodified reference
binaries (tests/golden/make_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

import util

MAN = json.load(open(os.path.join(util.GOLDEN, "manifest.json")))


@pytest.mark.parametrize("name", ["cand_k13.bin", "cand_k15.bin"])
def test_candidates_match_golden(built, tmp_path, name):
    from necat_b200 import formats
    wrk = util.golden_workdir(tmp_path)
    out = str(tmp_path / "o.bin")
    util.run([util.ORACLE_CLI, "pmov"] + MAN[name]["options"].split() + [wrk, "0", out], cwd=str(tmp_path))
    got = formats.canonical_records(formats.read_candidates(out))
    exp = formats.read_candidates(os.path.join(util.GOLDEN, name))
    assert exp.shape[0] == MAN[name]["records"] > 100
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("f,mode", [("0", "1"), ("0", "0"), ("1", "0")])
def test_consensus_matches_golden(built, tmp_path, f, mode):
    """mode 1 = klib introsort emulated (what the reference binary does), mode 0 = acceptance-order
    weight sums (what the GPU does): both reproduce the reference FASTA."""
    wrk = util.golden_workdir(tmp_path)
    cand = os.path.join(util.GOLDEN, "cand_k15.bin")
    opts = MAN[f"cns_f{f}"]["options"].split()
    util.run([util.ORACLE_CLI, "cns"] + opts + ["-M", mode, wrk, cand, "cns.fa", "raw.fa"], cwd=str(tmp_path))
    assert util.fasta_canon(str(tmp_path / "cns.fa")) == util.golden_fasta(f"cns_f{f}")
    assert util.fasta_canon(str(tmp_path / "raw.fa")) == util.golden_fasta(f"raw_f{f}")
    assert len(util.golden_fasta(f"cns_f{f}")) == MAN[f"cns_f{f}"]["records"] > 50


def test_m4_matches_golden(built, tmp_path):
    import gzip
    wrk = util.golden_workdir(tmp_path)
    util.run([util.ORACLE_CLI, "pmov"] + MAN["m4_k13"]["options"].split() + [wrk, "0", "o.m4"], cwd=str(tmp_path))
    got = sorted(open(str(tmp_path / "o.m4"), "rb").read().splitlines())
    exp = gzip.open(os.path.join(util.GOLDEN, "m4_k13.txt.gz")).read().splitlines()
    assert len(exp) == MAN["m4_k13"]["records"] > 100
    assert got == exp


@pytest.mark.parametrize("r", ["0", "1"])
def test_rescue_matches_golden(built, tmp_path, r):
    """oc2cns -r 1 (consensus_aux.c:169-201: O(nd) local alignment + global edlib where the block-wise extension
    fails or stops short) on reads that carry 300-700 base indels: the rescue changes most records, and the oracle's
    restatement of it (oracle/oracle_rescue.c) reproduces the reference's FASTA; -r 0 on the same set for contrast."""
    wrk = util.golden_workdir_r1(tmp_path)
    man = MAN["r1"]
    cand = os.path.join(util.GOLDEN, "r1", "cand_k13.bin")
    opts = man[f"cns_r{r}"]["options"].split()
    util.run([util.ORACLE_CLI, "cns"] + opts + ["-M", "1", wrk, cand, "cns.fa", "raw.fa"], cwd=str(tmp_path))
    exp = util.golden_fasta(f"r1/cns_r{r}")
    assert len(exp) == man[f"cns_r{r}"]["records"] > 40
    assert util.fasta_canon(str(tmp_path / "cns.fa")) == exp
    assert util.fasta_canon(str(tmp_path / "raw.fa")) == util.golden_fasta(f"r1/raw_r{r}")
    if r == "1":
        assert exp != util.golden_fasta("r1/cns_r0"), "the fixture must exercise the rescue"
