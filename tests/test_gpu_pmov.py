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
])
def test_candidates_match_oracle(built, tmp_path, opts, gen):
    wrk, _ = util.make_workdir(tmp_path, **gen)
    got = util.run_pmov(OC2PMOV, opts, wrk, 0, str(tmp_path / "gpu.bin"))
    exp = util.run_pmov(util.ORACLE_CLI, opts, wrk, 0, str(tmp_path / "oracle.bin"), prefix=())
    assert exp.shape[0] > 0
    assert util.same_multiset(got, exp), f"gpu {got.shape} vs oracle {exp.shape}"


def test_multi_volume(built, tmp_path):
    """Volume pairs (vid, i >= vid): cross-volume queries have no self-hit cut (word_finder.c:121-127)."""
    wrk, _ = util.make_workdir(tmp_path, genome=100_000, depth=24, seed=21, vol_size=900_000)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    assert nvol >= 2
    for vid in range(nvol):
        got = util.run_pmov(OC2PMOV, util.PM_SENSITIVE, wrk, vid, str(tmp_path / f"gpu{vid}.bin"))
        exp = util.run_pmov(util.ORACLE_CLI, util.PM_SENSITIVE, wrk, vid, str(tmp_path / f"or{vid}.bin"))
        assert util.same_multiset(got, exp), f"volume {vid}: gpu {got.shape} vs oracle {exp.shape}"
