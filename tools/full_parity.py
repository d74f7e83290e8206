#!/usr/bin/env python3
"""Full-size parity run: our drop-in executables vs the unmodified reference binaries
(oracle/_ref) on the same synthetic read set, canonical-order comparison of the
candidate records and both FASTA outputs, with wall-clock per stage.

    python tools/full_parity.py --genome 4600000 --depth 40 --out gpurun_out/parity.json
    python tools/full_parity.py --iterations 2 --out gpurun_out/parity_iter2.json     # mkdb .. reorder, iteration 2 with -r 1
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


def run_arm(bins, wrk_parent, fasta_lists, pm, cns, threads, tag):
    """One correction iteration with the five executables of `bins`: mkdb -> pmov -> pcan -> cns -> reorder.
    Returns (stats, candidate file, reordered cns, reordered raw)."""
    g = {}
    wrk = os.path.join(wrk_parent, tag + "_wrk")
    g["mkdb_s"], _ = timed([os.path.join(bins, "oc2mkdb"), wrk] + fasta_lists)
    cand = os.path.join(wrk_parent, tag + "_cand.bin")
    g["pmov_s"], log = timed([os.path.join(bins, "oc2pmov")] + pm.split() + ["-t", str(threads), wrk, "0", cand])
    g["pmov_log"] = [ln.split("INFO:")[-1].strip() for ln in log.splitlines() if "takes" in ln]
    import shutil as _sh
    part = cand + ".part"
    _sh.copy(cand, part)
    g["pcan_s"], _ = timed([os.path.join(bins, "oc2pcan"), "-t", str(threads), wrk, part])
    c_tmp, r_tmp = os.path.join(wrk_parent, tag + "_cns.tmp.fa"), os.path.join(wrk_parent, tag + "_raw.tmp.fa")
    g["cns_s"], log = timed([os.path.join(bins, "oc2cns")] + cns.split() + ["-t", str(threads), wrk, part, c_tmp, r_tmp])
    g["cns_log"] = [ln.split("INFO:")[-1].strip() for ln in log.splitlines() if "takes" in ln]
    c_out, r_out = os.path.join(wrk_parent, tag + "_cns.fa"), os.path.join(wrk_parent, tag + "_raw.fa")
    g["reorder_s"], _ = timed([os.path.join(bins, "oc2ReorderCnsReads"), c_tmp, r_tmp, c_out, r_out])
    g["total_s"] = sum(g[k] for k in ("mkdb_s", "pmov_s", "pcan_s", "cns_s", "reorder_s"))
    return g, cand, c_out, r_out


def iterations_main(a):
    """--iterations N: N correction iterations the way necat.pl chains them (necat.pl:160-365): every iteration but the
    last with -f 1, iterations >= 2 with the FAST overlap options and -r 1; both arms use their OWN five executables
    (mkdb, pmov, pcan, cns, reorder) and are compared after every iteration."""
    import gen_reads
    import util
    from necat_b200 import formats
    threads = os.cpu_count() or 1
    res = {"genome": a.genome, "depth": a.depth, "seed": a.seed, "host_threads": threads, "iterations": []}
    td = tempfile.mkdtemp(prefix="nbparity_")
    try:
        _, reads = gen_reads.generate(a.genome, a.depth, a.seed)
        A = np.frombuffer(b"ACGT", np.uint8)
        with open(os.path.join(td, "reads.fa"), "wb") as f:
            for i, r in enumerate(reads):
                f.write(b">r%d\n" % i + A[r].tobytes() + b"\n")
        res["reads"] = len(reads); res["bases"] = int(sum(r.size for r in reads))
        arms = {"gpu": a.gpu_bins or util.BIN_DIR}
        if not a.skip_ref and util.ref_bin("oc2cns"):
            arms["ref"] = util.REF_DIR
        lists = {}
        for arm in arms:
            lists[arm] = os.path.join(td, arm + "_list0")
            open(lists[arm], "w").write(os.path.join(td, "reads.fa") + "\n")
        for it in range(1, a.iterations + 1):
            last = it == a.iterations
            pm = a.pm if it == 1 else a.pm2
            cns = (a.cns if it == 1 else a.cns2).replace("-f 0", "-f 0" if last else "-f 1")
            row = {"iteration": it, "pm": pm, "cns": cns}
            outs = {}
            for arm, bins in arms.items():
                st, cand, c_out, r_out = run_arm(bins, td, [lists[arm]], pm, cns, threads, "%s_it%d" % (arm, it))
                st["n_candidates"] = int(formats.read_candidates(cand).shape[0])
                cn = util.fasta_canon(c_out)
                st["cns_records"] = len(cn); st["cns_bases"] = sum(len(s) for _, s in cn)
                st["raw_records"] = len(util.fasta_canon(r_out))
                row[arm] = st
                outs[arm] = (cand, c_out, r_out)
                lists[arm] = os.path.join(td, "%s_list%d" % (arm, it))
                open(lists[arm], "w").write(c_out + "\n" + r_out + "\n")
                print(arm.upper(), "iteration", it, json.dumps({k: v for k, v in st.items() if not k.endswith("log")}), flush=True)
            if "ref" in outs:
                row["candidates_identical"] = bool(util.same_multiset(formats.read_candidates(outs["gpu"][0]), formats.read_candidates(outs["ref"][0])))
                row["cns_identical"] = util.fasta_canon(outs["gpu"][1]) == util.fasta_canon(outs["ref"][1])
                row["raw_identical"] = util.fasta_canon(outs["gpu"][2]) == util.fasta_canon(outs["ref"][2])
                row["speedup_total"] = row["ref"]["total_s"] / row["gpu"]["total_s"]
                print(json.dumps({k: row[k] for k in ("iteration", "candidates_identical", "cns_identical", "raw_identical", "speedup_total")}), flush=True)
            res["iterations"].append(row)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            json.dump(res, open(a.out, "w"), indent=1)
    finally:
        if a.keep:
            shutil.rmtree(a.keep, ignore_errors=True)
            shutil.move(td, a.keep)
        else:
            shutil.rmtree(td, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--depth", type=float, default=40.0)
    ap.add_argument("--seed", type=int, default=20260925)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--pm", default="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1")
    ap.add_argument("--cns", default="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0")
    ap.add_argument("--iterations", type=int, default=1, help=">= 2: chain correction iterations like necat.pl (see iterations_main)")
    ap.add_argument("--pm2", default="-k 15 -z 20 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1")
    ap.add_argument("--cns2", default="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 1 -f 0")
    ap.add_argument("--keep", default=None, help="--iterations: keep the work directory of the run under this path (for profiling)")
    ap.add_argument("--gpu-bins", default=None, help="directory of the executables of the first arm (default necat_b200/bin)")
    a = ap.parse_args()
    if a.iterations > 1:
        return iterations_main(a)
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
