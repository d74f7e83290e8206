#!/usr/bin/env python3
"""What one rank of N does in the correction half, on one GPU: the consensus job of the bench workload for the templates with
id % N == 0 (nb_consensus_job_sharded), N = 1, 2, 4, 8: GPU time of the extension and consensus phases and the wall clock
of the call.  Shows what does not shrink with N (the chains of the longest candidates / templates).
usage: bench_shard.py [N ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import bench
from necat_b200 import capi, formats

shards = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
reads = bench.make_reads(4_600_000, 40)
vol = bench.reads_to_volume(reads)
ctx = capi.Context(0)
popt = capi.default_map_options(**bench.PM_OPTS)
copt = capi.default_cns_options(**bench.CNS_OPTS)
dvol = ctx.upload(vol)
idx = ctx.build_index(dvol, 15, 500)
packed = capi.pack_candidates(ctx.map_volume(idx, dvol, 0, dvol, 0, 0, vol.nseq, popt))
ctx.free_index(idx)
part = capi.with_swapped_roles([packed])
env = {k: v for k, v in os.environ.items() if k.startswith("NB_")}
for n in shards:
    out = []
    for it in range(3):
        t0 = time.perf_counter()
        recs, seq = ctx.consensus_job(dvol, part, copt, n, 0)
        w = (time.perf_counter() - t0) * 1e3
        t = ctx.timings()
        seq.free()
        out.append((w, t.extend_ms, t.consensus_ms))
    best = min(out[1:])
    print("shard 0 of %d | env %s | wall %.1f ext %.1f cns %.1f" % (n, env, best[0], best[1], best[2]), flush=True)
