#!/usr/bin/env python3
"""Replays GPU-marked tests of tests/ on the CPU, with the EMULATED build of the library (tests/emu/cudaemu: the
unmodified CUDA sources compiled for the host) standing in for necat_b200/bin.  Slow (minutes per test) and no
substitute for the B200 run -- it is how tests written without GPU time at hand were dry-run before being committed.

    python tools/run_gpu_tests_emulated.py [text partitions tiny edge0 edge1 flow]      # default: all of them
"""
import os
import pathlib
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import util  # noqa: E402

EMU = os.path.join(ROOT, "tests", "emu")
# NB_EMU_ASAN=1: the AddressSanitizer build with exact-size "device" allocations (tests/emu/README.md), ~3x slower still
ASAN = os.environ.get("NB_EMU_ASAN") == "1"
subprocess.check_call(["make", "-s", "-C", EMU] + (["asan"] if ASAN else []))
if ASAN:
    os.environ["NB_EMU_STRICT"] = "1"
    os.environ.setdefault("ASAN_OPTIONS", "detect_leaks=0")
util.ensure_oracle()
EMU_BIN = os.path.join(EMU, "_build", "cudaemu_asan" if ASAN else "cudaemu", "bin")
HOST_BIN = util.BIN_DIR
util.BIN_DIR = EMU_BIN                       # the three executables over the C-ABI come from the emulated build
import test_iteration_flow as F  # noqa: E402
import test_zz_gpu_edges as E  # noqa: E402

F.ours = lambda name: os.path.join(EMU_BIN if name in ("oc2pmov", "oc2pcan", "oc2cns") else HOST_BIN, name)


def tmp():
    return pathlib.Path(tempfile.mkdtemp())


def step(name, fn):
    t = time.time()
    try:
        fn()
        print(name, "PASS", "%.0fs" % (time.time() - t), flush=True)
        return 0
    except BaseException as e:  # noqa: BLE001
        print(name, "FAIL", type(e).__name__, str(e)[:400], "%.0fs" % (time.time() - t), flush=True)
        return 1


GENS = (dict(genome=30_000, depth=12, seed=54, err=0.10, min_len=60, max_len=4000, mean_log=6.6, sd_log=0.8),
        dict(genome=60_000, depth=2, seed=55, err=0.10))
TESTS = {
    "text": lambda: E.test_candidate_text_output(None, tmp()),
    "partitions": lambda: E.test_many_partitions_two_nodes(None, tmp()),
    "tiny": lambda: E.test_reads_shorter_than_a_kmer(None, tmp()),
    "edge0": lambda: E.test_edge_sets_match_oracle(None, tmp(), GENS[0]),
    "edge1": lambda: E.test_edge_sets_match_oracle(None, tmp(), GENS[1]),
    "flow": lambda: F.test_two_iterations_gpu(None, tmp()),
}

if __name__ == "__main__":
    bad = sum(step(k, TESTS[k]) for k in (sys.argv[1:] or list(TESTS)))
    sys.exit(1 if bad else 0)
