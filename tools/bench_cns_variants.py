#!/usr/bin/env python3
"""A/B of library variants (tools/build_variant.sh) on the bench workload: candidates once, then the consensus job three
times per variant; prints the GPU times of the extension and consensus phases and the wall clock of the call.
usage: bench_cns_variants.py <lib.so> [<lib.so> ...]   (each in its own process: the library is loaded once per process)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))

if len(sys.argv) > 2 or (len(sys.argv) == 2 and not sys.argv[1].endswith(".so")):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "--one", lib])
    sys.exit(0)

import numpy as np
import bench
from necat_b200 import capi, formats
lib = sys.argv[2]
capi.LIB_PATH = lib
reads = bench.make_reads(4_600_000, 40)
vol = bench.reads_to_volume(reads)
ctx = capi.Context(0)
popt = capi.default_map_options(**bench.PM_OPTS)
copt = capi.default_cns_options(**bench.CNS_OPTS)
dvol = ctx.upload(vol)
idx = ctx.build_index(dvol, 15, 500)
packed = capi.pack_candidates(ctx.map_volume(idx, dvol, 0, dvol, 0, 0, vol.nseq, popt))
ctx.free_index(idx)
part = np.concatenate([packed, formats.swap_roles(packed)])
out = []
for it in range(4):
    t0 = time.perf_counter()
    recs, seq = ctx.consensus_job(dvol, part, copt)
    w = (time.perf_counter() - t0) * 1e3
    t = ctx.timings()
    bases = int(recs["seq_len"][recs["is_raw"] == 0].sum())
    seq.free()
    out.append((w, t.extend_ms, t.consensus_ms))
best = min(out[1:])
print("%-40s env %s | wall %.1f ext %.1f cns %.1f | all: %s | bases %d" % (lib[-40:], {k: v for k, v in os.environ.items() if k.startswith("NB_")},
      best[0], best[1], best[2], [tuple(round(x) for x in o) for o in out], bases), flush=True)
