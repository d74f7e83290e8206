"""GPU parity of the oc2cns half: oc2pcan + oc2cns drop-in executables against the
CPU oracle on the same candidates (canonical record order)."""
import os
import shutil

import pytest

import util

pytestmark = pytest.mark.gpu

OC2CNS = os.path.join(util.BIN_DIR, "oc2cns")
OC2PCAN = os.path.join(util.BIN_DIR, "oc2pcan")


def _run_cns(tmp_path, wrk, cand, opts, tag, env=None):
    c = str(tmp_path / f"cand_{tag}.bin")
    shutil.copy(cand, c)
    util.run([OC2PCAN, "-t", "1", wrk, c])
    cns, raw = str(tmp_path / f"{tag}_cns.fasta"), str(tmp_path / f"{tag}_raw.fasta")
    util.run([OC2CNS] + opts.split() + [wrk, c, cns, raw], env=dict(os.environ, **(env or {})))
    return cns, raw


def _run_oracle(tmp_path, wrk, cand, opts, tag):
    cns, raw = str(tmp_path / f"{tag}_cns.fasta"), str(tmp_path / f"{tag}_raw.fasta")
    util.run([util.ORACLE_CLI, "cns"] + opts.split() + ["-M", "0", wrk, cand, cns, raw])
    return cns, raw


@pytest.mark.parametrize("opts,gen", [
    (util.CNS_SENSITIVE, dict(genome=100_000, depth=25, seed=41)),
    ("-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 1", dict(genome=80_000, depth=25, seed=42)),
    ("-a 1000 -x 3 -y 8 -l 500 -e 0.5 -p 0.6 -u 1 -r 0 -f 0", dict(genome=60_000, depth=30, seed=43, err=0.10)),
    (util.CNS_SENSITIVE, dict(genome=80_000, depth=40, seed=44, err=0.10, repeat_len=4000, repeat_copies=6, n_repeats=2)),
])
def test_cns_matches_oracle(built, tmp_path, opts, gen):
    wrk, _ = util.make_workdir(tmp_path, **gen)
    cand = str(tmp_path / "cand.bin")
    util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, 0, cand)
    g_cns, g_raw = _run_cns(tmp_path, wrk, cand, opts, "gpu")
    o_cns, o_raw = _run_oracle(tmp_path, wrk, cand, opts, "oracle")
    ec, er = util.fasta_canon(o_cns), util.fasta_canon(o_raw)
    assert len(ec) > 20
    gc, gr = util.fasta_canon(g_cns), util.fasta_canon(g_raw)
    assert len(gc) == len(ec) and len(gr) == len(er), f"record counts: cns {len(gc)}/{len(ec)} raw {len(gr)}/{len(er)}"
    bad = [i for i, (a, b) in enumerate(zip(gc, ec)) if a != b]
    assert not bad, f"{len(bad)} corrected records differ, first: {gc[bad[0]][0]} vs {ec[bad[0]][0]}"
    assert gr == er


def test_cns_multi_volume(built, tmp_path):
    """Several volumes: oc2cns merges them (consensus/main.c:66) and ids stay global."""
    wrk, _ = util.make_workdir(tmp_path, genome=80_000, depth=24, seed=45, vol_size=700_000)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    assert nvol >= 2
    cand = str(tmp_path / "cand.bin")
    with open(cand, "wb") as out:
        for vid in range(nvol):
            p = str(tmp_path / f"pm{vid}.bin")
            util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, vid, p)
            out.write(open(p, "rb").read())
    g_cns, g_raw = _run_cns(tmp_path, wrk, cand, util.CNS_SENSITIVE, "gpu")
    o_cns, o_raw = _run_oracle(tmp_path, wrk, cand, util.CNS_SENSITIVE, "oracle")
    assert util.fasta_canon(g_cns) == util.fasta_canon(o_cns)
    assert util.fasta_canon(g_raw) == util.fasta_canon(o_raw)


@pytest.mark.parametrize("env", [
    {"NB_CNS_DP_TEST": "1"},                        # k_cns_dp with a score ring of 8: most templates are redone by k_cns_dp_serial
    {"NB_CNS_DP_TEST": "2"},                        # every template through k_cns_dp_serial
    {"NB_CNS_FLUSH_SLOW": "1"},                     # k_cns_links: the selection flush for every position
    {"NB_CNS_SPAN": "64"},                          # k_cns_links: one tile per unit, a checkpoint every 64 positions
    {"NB_CNS_SPAN": "192", "NB_CNS_DP_TEST": "1"},
    {"NB_CNS_LONG_PCT": "30"},                      # the longest templates' consensus (a warp per template) on its own stream beside the rest's
    {"NB_CNS_LONG_PCT": "10", "NB_CNS_SLICES": "5"},    # six slices: link kernels one after the other, best-path chains beside them
    {"NB_CNS_LONG_PCT": "30", "NB_CNS_SLICES": "1", "NB_CNS_DP_TEST": "1"},
    {"NB_CNS_LONG_PCT": "0"},                       # one slice
    {"NB_EXT_LONGQ": "2"},                          # k_extend: a quarter of every launch through the queue of the longest candidates (two-slot warps)
    {"NB_CNS_CHUNKS": "5", "NB_CNS_PIPES": "3"},    # the chunk pipeline: three workers on their own streams
    {"NB_CNS_CHUNKS": "2", "NB_CNS_PIPES": "1"},
    {"NB_CNS_GROUPS": "4", "NB_CNS_GROUPS_MIN": "8"},   # extension rounds of four template groups interleaved on the GPU
    {"NB_CNS_GROUPS": "3", "NB_CNS_GROUPS_MIN": "8", "NB_CNS_LONG": "1", "NB_CNS_LONG_MIN": "8"},   # + the longest templates' consensus beside the extension
])
def test_cns_kernel_variants(built, tmp_path, env):
    """The tunables of the consensus kernels (span length, score-ring size, flush path, pipelines) change
    the schedule, never the result."""
    wrk, _ = util.make_workdir(tmp_path, genome=70_000, depth=30, seed=46, err=0.12)
    cand = str(tmp_path / "cand.bin")
    util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, 0, cand)
    g_cns, g_raw = _run_cns(tmp_path, wrk, cand, util.CNS_SENSITIVE, "gpu", env=env)
    o_cns, o_raw = _run_oracle(tmp_path, wrk, cand, util.CNS_SENSITIVE, "oracle")
    assert len(util.fasta_canon(o_cns)) > 20
    assert util.fasta_canon(g_cns) == util.fasta_canon(o_cns)
    assert util.fasta_canon(g_raw) == util.fasta_canon(o_raw)
