#!/usr/bin/env python3
"""Prepare a small work dir + candidates for profiling runs: tools/prep_small.py DIR GENOME DEPTH"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

d, genome, depth = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
os.makedirs(d, exist_ok=True)
wrk, _ = util.make_workdir(d, genome=genome, depth=depth, seed=20260925)
cand = os.path.join(d, "cand.bin")
subprocess.check_call([os.path.join(util.BIN_DIR, "oc2pmov")] + util.PM_SENSITIVE.split() + [wrk, "0", cand], stdout=subprocess.DEVNULL)
subprocess.check_call([os.path.join(util.BIN_DIR, "oc2pcan"), wrk, cand])
print(wrk, cand)
