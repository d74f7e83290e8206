"""GPU parity of the oc2pmov half: the drop-in executable (C host over the C-ABI,
CUDA kernels) against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

OC2PMOV = os.path.join(util.BIN_DIR, "oc2pmov")


@pytest.mark.parametrize("opts,gen", [
    (util.PM_SENSITIVE, dict(genome=150_000, depth=20, seed=11)),
    (util.PM_FAST, dict(genome=150_000, depth=20, seed=12)),
    # planted repeats + low count cutoff: dropped k-mers, saturated blocks, block zeroing
    ("-k 13 -z 5 -q 20 -b 1000 -s 3 -n 500 -a 500 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1",
     dict(genome=120_000, depth=25, seed=13, err=0.10, repeat_len=3000, repeat_copies=12, n_repeats=3)),
    ("-k 11 -z 10 -q 500 -b 2000 -s 2 -n 5 -a 500 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1",
     dict(genome=60_000, depth=30, seed=14, err=0.08)),
    # dense probing of a highly repetitive genome: > 4096 hits per (read, strand) take the global-memory sort path,
    # blocks saturate at 40 seeds, long index runs take the long-run sorter
    ("-k 12 -z 1 -q 2000 -b 500 -s 3 -n 500 -a 500 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1",
     dict(genome=60_000, depth=12, seed=15, err=0.06, repeat_len=2500, repeat_copies=14, n_repeats=1)),
])
def test_candidates_match_oracle(built, tmp_path, opts, gen):
    wrk, _ = util.make_workdir(tmp_path, **gen)
    got = util.run_pmov(OC2PMOV, opts, wrk, 0, str(tmp_path / "gpu.bin"))
    exp = util.run_pmov(util.ORACLE_CLI, opts, wrk, 0, str(tmp_path / "oracle.bin"), prefix=())
    assert exp.shape[0] > 0
    assert util.same_multiset(got, exp), f"gpu {got.shape} vs oracle {exp.shape}"


def test_ultra_long_reads(built, tmp_path):
    """Reads of 0.7-1.1 Mb: more than 65 535 probes each at -z 10 (24-bit probe numbers in the sort keys of
    k_build_blocks) and, beyond ~800 kb, probed from global memory instead of the shared-memory stage of k_probe; then the
    same reads through oc2pcan + oc2cns (extension chains of ~2000 blocks, templates of a million positions)."""
    import shutil
    wrk, reads = util.make_workdir(tmp_path, genome=1_300_000, depth=5, seed=31, min_len=700_000, max_len=1_100_000, mean_log=13.7, sd_log=0.15)
    assert max(r.size for r in reads) > 850_000 and min(r.size for r in reads) > 655_360 + 15
    cand = str(tmp_path / "oracle.bin")
    got = util.run_pmov(OC2PMOV, util.PM_SENSITIVE, wrk, 0, str(tmp_path / "gpu.bin"))
    exp = util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, 0, cand)
    assert exp.shape[0] >= 4
    assert util.same_multiset(got, exp), f"gpu {got.shape} vs oracle {exp.shape}"
    cns = "-a 2000 -x 2 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([os.path.join(util.BIN_DIR, "oc2pcan"), "-t", "1", wrk, c])
    util.run([os.path.join(util.BIN_DIR, "oc2cns")] + cns.split() + [wrk, c, str(tmp_path / "g_cns.fa"), str(tmp_path / "g_raw.fa")])
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "0", wrk, cand, str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
    exp_cns = util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert len(exp_cns) >= 1
    assert util.fasta_canon(str(tmp_path / "g_cns.fa")) == exp_cns
    assert util.fasta_canon(str(tmp_path / "g_raw.fa")) == util.fasta_canon(str(tmp_path / "o_raw.fa"))


def test_multi_volume(built, tmp_path):
    """Volume pairs (vid, i >= vid): cross-volume queries have no self-hit cut (word_finder.c:121-127)."""
    wrk, _ = util.make_workdir(tmp_path, genome=100_000, depth=24, seed=21, vol_size=900_000)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    assert nvol >= 2
    for vid in range(nvol):
        got = util.run_pmov(OC2PMOV, util.PM_SENSITIVE, wrk, vid, str(tmp_path / f"gpu{vid}.bin"))
        exp = util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, vid, str(tmp_path / f"or{vid}.bin"))
        assert util.same_multiset(got, exp), f"volume {vid}: gpu {got.shape} vs oracle {exp.shape}"


@pytest.mark.parametrize("opts", [
    "-k 13 -z 10 -q 500 -b 2000 -s 3 -n 100 -a 400 -d 0.25 -e 0.5 -m 500 -t 1 -j 1 -u 0 -i 1",
    "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 20 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 1 -u 0 -i 0",
])
def test_m4_matches_oracle(built, tmp_path, opts):
    """ALN mode (-j 1): M4 text lines identical to the oracle's (sorted), names with -i 1, ids with -i 0."""
    wrk, _ = util.make_workdir(tmp_path, genome=120_000, depth=18, seed=16, vol_size=1_200_000)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    assert nvol >= 2                      # includes a cross-volume pair: query and subject volumes differ
    for vid in range(nvol):
        g, o = str(tmp_path / f"g{vid}.m4"), str(tmp_path / f"o{vid}.m4")
        util.run([OC2PMOV] + opts.split() + [wrk, str(vid), g])
        util.run([util.ORACLE_CLI, "pmov"] + opts.split() + [wrk, str(vid), o])
        gl, ol = sorted(open(g).read().splitlines()), sorted(open(o).read().splitlines())
        assert len(ol) > 100 and gl == ol, f"volume {vid}: {len(gl)} vs {len(ol)} lines"


def test_m4_binary(built, tmp_path):
    opts = "-k 13 -z 10 -q 500 -b 2000 -s 3 -n 50 -a 400 -d 0.25 -e 0.5 -m 500 -t 1 -j 1 -u 1 -i 1"
    wrk, _ = util.make_workdir(tmp_path, genome=60_000, depth=15, seed=17)
    g, o = str(tmp_path / "g.bin"), str(tmp_path / "o.bin")
    util.run([OC2PMOV] + opts.split() + [wrk, "0", g])
    util.run([util.ORACLE_CLI, "pmov"] + opts.split() + [wrk, "0", o])
    dt = np.dtype([("qid", "<i4"), ("qdir", "<i4"), ("q", "<u8", 4), ("sid", "<i4"), ("sdir", "<i4"), ("s", "<u8", 4),
                   ("ident", "<f8"), ("vscore", "<i4"), ("pad", "<i4")])
    a, b = np.fromfile(g, dt), np.fromfile(o, dt)
    assert dt.itemsize == 96 and a.size == b.size > 50
    key = lambda r: sorted(map(lambda x: (x["qid"], x["sid"], x["qdir"], tuple(x["q"]), tuple(x["s"]), x["ident"], x["vscore"]), r))
    assert key(a) == key(b)
