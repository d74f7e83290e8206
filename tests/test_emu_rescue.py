"""CPU: the warp-level core of the -r 1 local-alignment kernel (necat_b200/csrc/nb_rescue_core.cuh -- the code the
CUDA kernel k_rescue_local instantiates) run on the host, one std::thread per lane with barriers for the warp
collectives (tests/emu/emu_rescue.cpp), against the oracle (oracle/oracle_rescue.c) on the long-indel golden set: the O(nd) waves (k_rescue_local), and the two
building blocks of the global alignment that follows -- the strip-mined last-column scores of edlib's Hirschberg step and
the traceback leaf."""
import os
import subprocess

import util


def test_rescue_core_matches_oracle(built, tmp_path):
    subprocess.check_call(["make", "-s", "-C", util.ORACLE_DIR, "emu"])
    wrk = util.golden_workdir_r1(tmp_path)
    cand = os.path.join(util.GOLDEN, "r1", "cand_k13.bin")
    p = util.run([os.path.join(util.ORACLE_DIR, "_build", "emu_rescue"), wrk, cand, "3", "6", "2"], cwd=str(tmp_path))
    assert b"3 tasks, 0 mismatches; 6 columns, 0 mismatches" in p.stdout
    assert b"leaves 6, 0 mismatches" in p.stdout
    assert b"global alignments 2 (2 ok), 0 mismatches" in p.stdout      # host-driven Hirschberg (nb_rescue_driver.h) over the emulated cores
