#!/usr/bin/env python3
"""Times the ingest / reorder hosts (necat_b200/bin/oc2mkdb, oc2ReorderCnsReads) beside the
reference binaries (oracle/_ref) on the E. coli-size read set of bench.py, checks that the outputs agree, and writes
one JSON document (default profiles/r01_host_tools.json).  CPU only -- these steps never touch the GPU.

    python tools/bench_host_tools.py [--genome 4600000 --depth 40 --repeat 3 --out profiles/r01_host_tools.json]
"""
import argparse
import json
import os
import re
import shutil
import struct
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import gen_reads  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OURS = os.path.join(ROOT, "necat_b200", "bin")


def timed(cmd, cwd, repeat):
    best = None
    for _ in range(repeat):
        t = time.perf_counter()
        subprocess.check_call(cmd, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best


def masked(path):
    b = bytearray(open(path, "rb").read())
    ns = struct.unpack_from("<Q", b, 31)[0]
    for i in range(ns):
        o = 31 + 16 + i * 32 + 28
        b[o:o + 4] = b"\0\0\0\0"
    return bytes(b)


def fasta_by_id(path):
    out, cur = {}, None
    ids = []
    for line in open(path, "rb"):
        if line[:1] == b">":
            i = int(re.search(rb"\d+", line).group())
            ids.append(i)
            cur = [line, b""]
            out.setdefault(i, []).append(cur)
        else:
            cur[1] += line
    return ids, {k: sorted(map(tuple, v)) for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--depth", type=float, default=40)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_host_tools.json"))
    a = ap.parse_args()
    assert os.path.exists(os.path.join(REF, "oc2mkdb")), "oracle/_ref is not built"
    d = tempfile.mkdtemp()
    try:
        _, reads = gen_reads.generate(a.genome, a.depth, 20260925)
        A = np.frombuffer(b"ACGT", np.uint8)
        with open(os.path.join(d, "reads.fa"), "wb") as f:
            for i, r in enumerate(reads):
                f.write(b">r%d\n" % i + A[r].tobytes() + b"\n")
        subprocess.check_call(["gzip", "-k", "-1", "reads.fa"], cwd=d)
        open(os.path.join(d, "plain"), "w").write("reads.fa\n")
        open(os.path.join(d, "gz"), "w").write("reads.fa.gz\n")
        # reorder input shaped like cat of the reference's oc2cns output: ids shuffled locally (threads interleave),
        # two fragments per read in cns, every fourth read in raw
        rng = np.random.default_rng(1)
        order = np.arange(len(reads))
        for s in range(0, len(order), 64):
            rng.shuffle(order[s:s + 64])
        with open(os.path.join(d, "cns_in.fa"), "wb") as f, open(os.path.join(d, "raw_in.fa"), "wb") as g:
            for k, i in enumerate(order):
                r = reads[i]
                h = r.size // 2
                f.write(b">ontsa_id_%d_(0_%d_%d_%d_20_12_0.8)\n" % (i, h, h, r.size) + A[r[:h]].tobytes() + b"\n")
                f.write(b">ontsa_id_%d_(%d_%d_%d_%d_20_12_0.8)\n" % (i, h, r.size, r.size - h, r.size) + A[r[h:]].tobytes() + b"\n")
                if k % 4 == 0:
                    g.write(b">ontsa_id_%d_(0_%d_%d_%d_0_0_0)\n" % (i, r.size, r.size, r.size) + A[r[:2000]].tobytes() + b"\n")
        bases = int(sum(r.size for r in reads))
        res = {"workload": f"E. coli-size {a.genome} b x {a.depth} synthetic reads (gen_reads seed 20260925): {len(reads)} reads, {bases} bases",
               "host": {"cores": os.cpu_count(), "note": "timed in the build container (no GPU involved), best of %d" % a.repeat},
               "steps": {}}
        for name, lst in (("oc2mkdb plain FASTA", "plain"), ("oc2mkdb gzip FASTA", "gz")):
            tr = timed([os.path.join(REF, "oc2mkdb"), "ref_w", lst], d, a.repeat)
            to = timed([os.path.join(OURS, "oc2mkdb"), "our_w", lst], d, a.repeat)
            same = masked(os.path.join(d, "ref_w", "vol0")) == masked(os.path.join(d, "our_w", "vol0"))
            res["steps"][name] = {"reference_s": round(tr, 3), "this_repo_s": round(to, 3), "speedup": round(tr / to, 2),
                                  "mbases_per_s": round(bases / to / 1e6, 1), "volume_identical": same}
        tr = timed([os.path.join(REF, "oc2ReorderCnsReads"), "cns_in.fa", "raw_in.fa", "r_cns.fa", "r_raw.fa"], d, a.repeat)
        to = timed([os.path.join(OURS, "oc2ReorderCnsReads"), "cns_in.fa", "raw_in.fa", "o_cns.fa", "o_raw.fa"], d, a.repeat)
        ia, da = fasta_by_id(os.path.join(d, "r_cns.fa"))
        ib, db = fasta_by_id(os.path.join(d, "o_cns.fa"))
        same = ia == ib and da == db and open(os.path.join(d, "r_raw.fa"), "rb").read() == open(os.path.join(d, "o_raw.fa"), "rb").read()
        res["steps"]["oc2ReorderCnsReads, shuffled input"] = {"reference_s": round(tr, 3), "this_repo_s": round(to, 3), "speedup": round(tr / to, 2),
                                                             "same_ids_and_records": same}
        ts = timed([os.path.join(OURS, "oc2ReorderCnsReads"), "o_cns.fa", "o_raw.fa", "o2_cns.fa", "o2_raw.fa"], d, a.repeat)
        res["steps"]["oc2ReorderCnsReads, input already in id order (what this repo's oc2cns writes): streamed"] = {
            "this_repo_s": round(ts, 3), "output_identical_to_input": open(os.path.join(d, "o_cns.fa"), "rb").read() == open(os.path.join(d, "o2_cns.fa"), "rb").read()}
        json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps(res, indent=1))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
