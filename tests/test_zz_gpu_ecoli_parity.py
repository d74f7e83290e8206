"""GPU parity at the BASELINE configuration (BASELINE.json configs[1]): E. coli-size 4.6 Mb x 40 synthetic ONT reads through
this repo's drop-in executables (oc2pmov SENSITIVE -> oc2pcan -> oc2cns -r 0 -f 0) against the UNMODIFIED reference
binaries (oracle/_ref, built from /root/reference by oracle/Makefile.ref; they travel to the GPU box) on the same work
directory: candidate records as a multiset, cns.fasta and raw.fasta in canonical order (the reference's
own output order changes from run to run).  About two minutes, most of them the reference's two 8 GiB table builds."""
import os
import shutil

import pytest

import util

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

PM = "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CNS = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


def test_ecoli_iteration_matches_reference_binaries(built, tmp_path):
    from necat_b200 import formats
    if not util.ref_bin("oc2cns"):
        pytest.skip("oracle/_ref is not built (make -C oracle ref, needs /root/reference)")
    threads = str(os.cpu_count() or 1)
    wrk, reads = util.make_workdir(tmp_path, genome=4_600_000, depth=40, seed=20260925)
    out = {}
    for tag, bins in (("gpu", util.BIN_DIR), ("ref", util.REF_DIR)):
        cand = str(tmp_path / f"{tag}_cand.bin")
        util.run([os.path.join(bins, "oc2pmov")] + PM.split() + ["-t", threads, wrk, "0", cand])
        part = cand + ".part"
        shutil.copy(cand, part)
        util.run([os.path.join(bins, "oc2pcan"), "-t", threads, wrk, part])
        cns, raw = str(tmp_path / f"{tag}_cns.fa"), str(tmp_path / f"{tag}_raw.fa")
        util.run([os.path.join(bins, "oc2cns")] + CNS.split() + ["-t", threads, wrk, part, cns, raw])
        out[tag] = (formats.read_candidates(cand), util.fasta_canon(cns), util.fasta_canon(raw))
    g, r = out["gpu"], out["ref"]
    assert r[0].shape[0] > 100_000 and len(r[1]) > 15_000
    assert util.same_multiset(g[0], r[0]), f"candidates differ: {g[0].shape[0]} vs {r[0].shape[0]}"
    assert len(g[1]) == len(r[1]) and len(g[2]) == len(r[2])
    bad = [i for i, (a, b) in enumerate(zip(g[1], r[1])) if a != b]
    assert not bad, f"{len(bad)} corrected records differ, first: {g[1][bad[0]][0]} vs {r[1][bad[0]][0]}"
    assert g[2] == r[2]
