#!/usr/bin/env python3
"""Full-size parity run: our drop-in executables vs the unmodified reference binaries
(oracle/_ref) on the same synthetic read set, canonical-order comparison of the
candidate records and both FASTA outputs, with wall-clock per stage.

    python tools/full_parity.py --genome 4600000 --depth 40 --out gpurun_out/parity.json
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def timed(cmd, **kw):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if p.returncode:
        print(p.stdout[-4000:])
        raise SystemExit(f"{cmd[0]} failed with {p.returncode}")
    return time.perf_counter() - t0, p.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--depth", type=float, default=40.0)
    ap.add_argument("--seed", type=int, default=20260925)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--pm", default="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1")
    ap.add_argument("--cns", default="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0")
    a = ap.parse_args()
    import gen_reads
    import util
    from necat_b200 import formats
    threads = os.cpu_count() or 1
    res = {"genome": a.genome, "depth": a.depth, "seed": a.seed, "host_threads": threads}
    td = tempfile.mkdtemp(prefix="nbparity_")
    try:
        t0 = time.perf_counter()
        _, reads = gen_reads.generate(a.genome, a.depth, a.seed)
        wrk = os.path.join(td, "vol")
        formats.write_workdir(wrk, reads)
        res["reads"] = len(reads); res["bases"] = int(sum(r.size for r in reads)); res["gen_s"] = time.perf_counter() - t0
        B = util.BIN_DIR
        g = {}
        gc = os.path.join(td, "g_cand.bin")
        g["pmov_s"], log = timed([os.path.join(B, "oc2pmov")] + a.pm.split() + ["-t", str(threads), wrk, "0", gc])
        g["pmov_log"] = [ln for ln in log.splitlines() if "takes" in ln]
        g["pcan_s"], _ = timed([os.path.join(B, "oc2pcan"), "-t", str(threads), wrk, gc])
        g["cns_s"], log = timed([os.path.join(B, "oc2cns")] + a.cns.split() + ["-t", str(threads), wrk, gc,
                                os.path.join(td, "g_cns.fasta"), os.path.join(td, "g_raw.fasta")])
        g["cns_log"] = [ln for ln in log.splitlines() if "takes" in ln]
        g["total_s"] = g["pmov_s"] + g["pcan_s"] + g["cns_s"]
        gcand = formats.read_candidates(gc)
        g["n_candidates"] = int(gcand.shape[0])
        gcns = util.fasta_canon(os.path.join(td, "g_cns.fasta"))
        graw = util.fasta_canon(os.path.join(td, "g_raw.fasta"))
        g["cns_records"] = len(gcns); g["cns_bases"] = sum(len(s) for _, s in gcns); g["raw_records"] = len(graw)
        res["gpu"] = g
        print("GPU", json.dumps({k: v for k, v in g.items() if not k.endswith("log")}), flush=True)
        if not a.skip_ref and util.ref_bin("oc2cns"):
            R = util.REF_DIR
            r = {}
            rc = os.path.join(td, "r_cand.bin")
            r["pmov_s"], log = timed([os.path.join(R, "oc2pmov")] + a.pm.split() + ["-t", str(threads), wrk, "0", rc])
            r["pmov_log"] = [ln.split("INFO:")[-1].strip() for ln in log.splitlines() if "takes" in ln]
            r["pcan_s"], _ = timed([os.path.join(R, "oc2pcan"), "-t", str(threads), wrk, rc])
            r["cns_s"], log = timed([os.path.join(R, "oc2cns")] + a.cns.split() + ["-t", str(threads), wrk, rc,
                                    os.path.join(td, "r_cns.fasta"), os.path.join(td, "r_raw.fasta")])
            r["total_s"] = r["pmov_s"] + r["pcan_s"] + r["cns_s"]
            rcand = formats.read_candidates(rc)
            r["n_candidates"] = int(rcand.shape[0])
            rcns = util.fasta_canon(os.path.join(td, "r_cns.fasta"))
            rraw = util.fasta_canon(os.path.join(td, "r_raw.fasta"))
            r["cns_records"] = len(rcns); r["cns_bases"] = sum(len(s) for _, s in rcns); r["raw_records"] = len(rraw)
            res["ref"] = r
            res["candidates_identical"] = bool(util.same_multiset(gcand, rcand))
            res["cns_identical"] = gcns == rcns
            res["raw_identical"] = graw == rraw
            if not res["cns_identical"]:
                sg, sr = set(gcns), set(rcns)
                res["cns_only_gpu"] = len(sg - sr); res["cns_only_ref"] = len(sr - sg)
                res["cns_diff_headers"] = [h.decode() for h, _ in sorted(sr - sg)[:5]]
            res["speedup_total"] = r["total_s"] / g["total_s"]
            print("REF", json.dumps({k: v for k, v in r.items() if not k.endswith("log")}), flush=True)
        print(json.dumps({k: res[k] for k in res if k not in ("gpu", "ref")}))
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            json.dump(res, open(a.out, "w"), indent=1)
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()
