#!/usr/bin/env python3
"""Scale check of the drop-in executables on ONE GPU at the volume sizes of BASELINE configs 4-5: a read set of more than
2 Gbp, so that oc2mkdb cuts it into V >= 2 volumes at the 2 Gbp limit and oc2pcan into several partitions of 100 000
templates (the layout necat.pl drives: oc2pmov once per index volume, `cat`, oc2pcan, oc2cns once per partition with
`-mn i count`, necat.pl:200-307), against the UNMODIFIED reference binaries (oracle/_ref) on the same FASTA files:

  * volumes / reads_info.txt of both oc2mkdb byte for byte (modulo the 4 uninitialised padding bytes per record),
  * the candidate records of ALL (index volume, query volume) pairs as a multiset,
  * cns.fasta / raw.fasta of the partitions --ref-parts names (default: partition 0) in canonical order.

The reads are made by P worker processes (gen_reads.mutate over one seeded genome), one FASTA file each.

    python tools/scale_volumes.py --genome 120000000 --depth 20 --out gpurun_out/scale_volumes.json
"""
import argparse
import json
import multiprocessing as mp
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

PM = "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CNS = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


def timed(cmd, **kw):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if p.returncode:
        print(p.stdout[-4000:])
        raise SystemExit("%s failed with %d" % (" ".join(cmd[:1]), p.returncode))
    return time.perf_counter() - t0, p.stdout


def _worker(args):
    """Reads of one worker: same genome for all (seed), own stream (seed + 1 + w), own FASTA file."""
    import gen_reads
    genome_size, seed, w, target, first_id_hint, path = args
    genome = np.random.default_rng(seed).integers(0, 4, genome_size, dtype=np.uint8)
    rng = np.random.default_rng(seed + 1 + w)
    A = np.frombuffer(b"ACGT", np.uint8)
    total = n = 0
    with open(path, "wb", buffering=1 << 24) as f:
        while total < target:
            L = int(np.clip(rng.lognormal(8.75, 0.55), 3000, 60000))
            s = int(rng.integers(0, genome_size - L + 1))
            frag = genome[s:s + L]
            if rng.random() < 0.5:
                frag = gen_reads.COMP[frag[::-1]]
            r = gen_reads.mutate(frag, rng, 0.15)
            if r.size < 3000:
                continue
            f.write(b">w%d_%d\n" % (w, n) + A[r].tobytes() + b"\n")
            total += r.size
            n += 1
    return n, total


def mem_available_gb():
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable:"):
            return int(ln.split()[1]) / 1e6
    return 0.0


def run_arm(bins, td, tag, fastas, threads, cns_parts, log):
    """mkdb -> pmov per index volume -> cat -> pcan -> cns per partition in cns_parts (None = all)."""
    st = {}
    wrk = os.path.join(td, tag + "_wrk")
    lst = os.path.join(td, tag + "_list")
    open(lst, "w").write("".join(p + "\n" for p in fastas))
    st["mkdb_s"], _ = timed([os.path.join(bins, "oc2mkdb"), wrk, lst])
    n_vol = len([ln for ln in open(os.path.join(wrk, "volume_names.txt")) if ln.strip()])
    st["volumes"] = n_vol
    cand = os.path.join(td, tag + "_cand.bin")
    st["pmov_s"] = []
    with open(cand, "wb") as out:
        for vid in range(n_vol):
            part = os.path.join(td, "%s_cand_v%d.bin" % (tag, vid))
            s, lg = timed([os.path.join(bins, "oc2pmov")] + PM.split() + ["-t", str(threads), wrk, str(vid), part])
            st["pmov_s"].append(s)
            log("%s oc2pmov v%d: %.1f s" % (tag, vid, s))
            with open(part, "rb") as f:
                shutil.copyfileobj(f, out, 1 << 24)
            os.remove(part)
    pc = cand + ".pcan"
    shutil.copy(cand, pc)
    st["pcan_s"], _ = timed([os.path.join(bins, "oc2pcan"), "-t", str(threads), wrk, pc])
    count = int(open(pc + ".partitions").read().split()[0])
    st["partitions"] = count
    parts = list(range(count)) if cns_parts is None else [p for p in cns_parts if p < count]
    st["cns_parts"] = parts
    st["cns_s"] = []
    outs = {}
    for i in parts:
        c, r = os.path.join(td, "%s_cns_%d.fa" % (tag, i)), os.path.join(td, "%s_raw_%d.fa" % (tag, i))
        s, _ = timed([os.path.join(bins, "oc2cns")] + CNS.split() + ["-t", str(threads), wrk, pc, c, r, "-mn", str(i), str(count)])
        st["cns_s"].append(s)
        log("%s oc2cns partition %d of %d: %.1f s" % (tag, i, count, s))
        outs[i] = (c, r)
    return st, wrk, cand, outs


def same_volumes(a, b):
    """Volume files byte for byte; reads_info.txt / volume_names.txt modulo the directory name."""
    res = {}
    names = sorted(f for f in os.listdir(a) if f.startswith("vol") and not f.endswith(".txt"))
    res["files"] = names == sorted(f for f in os.listdir(b) if f.startswith("vol") and not f.endswith(".txt"))
    ok = res["files"]
    for f in names if ok else []:
        if os.path.getsize(os.path.join(a, f)) != os.path.getsize(os.path.join(b, f)):
            ok = False
            break
        # the packed bases and the offsets must agree; the 32-byte sequence records carry 4 uninitialised padding
        # bytes in the reference -- compare with cmp and, on a difference, fall back to the parsed view
        if subprocess.run(["cmp", "-s", os.path.join(a, f), os.path.join(b, f)]).returncode:
            from necat_b200 import formats
            va, vb = formats.read_volume(os.path.join(a, f)), formats.read_volume(os.path.join(b, f))
            ok = ok and va.db_size == vb.db_size and np.array_equal(va.pac, vb.pac) and va.hdr == vb.hdr and \
                all(np.array_equal(va.info[k], vb.info[k]) for k in ("offset", "size", "hdr_offset"))
            res.setdefault("padding_only_differences", []).append(f)
    res["identical"] = bool(ok)
    res["reads_info_identical"] = open(os.path.join(a, "reads_info.txt")).read() == open(os.path.join(b, "reads_info.txt")).read()
    va, vb = ([[os.path.basename(ln.split("\t")[0])] + ln.split("\t")[1:] for ln in open(os.path.join(d, "volume_names.txt"))] for d in (a, b))
    res["volume_names_identical"] = va == vb
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=120_000_000)
    ap.add_argument("--depth", type=float, default=20.0)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--ref-parts", default="0", help="partitions the reference oc2cns runs (comma list, 'all', or 'none')")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--ref-cns-only", action="store_true",
                    help="the reference arm is oc2cns alone, on THIS repo's work directory and partition files (the candidates were "
                         "compared on the 2.4 Gbp run): what a full-size configuration-4 set affords; --ref-parts counts from the "
                         "end when negative (-1 = the last partition: reads beyond 2^32 bases of the merged store)")
    ap.add_argument("--ref-min-mem-gb", type=float, default=90.0, help="host memory the reference arm needs to be attempted")
    ap.add_argument("--gpu-bins", default=None, help="directory of the executables of the first arm (default necat_b200/bin)")
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import util
    from necat_b200 import formats
    threads = os.cpu_count() or 1
    W = a.workers or min(threads, 32)
    res = {"genome": a.genome, "depth": a.depth, "seed": a.seed, "host_threads": threads, "mem_available_gb": mem_available_gb(),
           "pm": PM, "cns": CNS}
    t_start = time.perf_counter()

    def log(msg):
        print("[%6.1f s] %s" % (time.perf_counter() - t_start, msg), flush=True)

    def dump():
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            json.dump(res, open(a.out, "w"), indent=1)

    td = tempfile.mkdtemp(prefix="nbscale_", dir=a.tmp)
    try:
        target = int(a.genome * a.depth)
        fastas = [os.path.join(td, "reads_%02d.fa" % w) for w in range(W)]
        with mp.Pool(W) as pool:
            made = pool.map(_worker, [(a.genome, a.seed, w, target // W, 0, fastas[w]) for w in range(W)])
        res["reads"] = int(sum(n for n, _ in made)); res["bases"] = int(sum(t for _, t in made))
        res["gen_s"] = time.perf_counter() - t_start
        log("%d reads, %.2f Gbp in %d FASTA files" % (res["reads"], res["bases"] / 1e9, W))
        g, gwrk, gcand, gout = run_arm(a.gpu_bins or util.BIN_DIR, td, "gpu", fastas, threads, None, log)
        gc = formats.read_candidates(gcand)
        g["n_candidates"] = int(gc.shape[0])
        g["cns_bases"] = 0; g["cns_records"] = 0
        for i, (c, r) in gout.items():
            for _, s in formats.read_fasta_records(c):
                g["cns_bases"] += len(s); g["cns_records"] += 1
        g["total_s"] = g["mkdb_s"] + sum(g["pmov_s"]) + g["pcan_s"] + sum(g["cns_s"])
        res["gpu"] = g
        log("GPU " + json.dumps(g))
        dump()
        ref_ok = not a.skip_ref and util.ref_bin("oc2cns") and a.ref_parts != "skip"
        if ref_ok and mem_available_gb() < a.ref_min_mem_gb:
            res["ref_skipped"] = "only %.0f GB of host memory available" % mem_available_gb()
            ref_ok = False
        if ref_ok and a.ref_cns_only:
            pc = gcand + ".pcan"
            count = g["partitions"]
            parts = [int(x) % count for x in a.ref_parts.split(",")]
            if not all(os.path.exists("%s.p%d" % (pc, i)) for i in parts):      # (oc2cns leaves them to necat.pl; be sure)
                shutil.copy(gcand, pc)
                timed([os.path.join(util.BIN_DIR, "oc2pcan"), "-t", str(threads), gwrk, pc])
            r = {"cns_parts": parts, "cns_s": [], "what": "reference oc2cns on this repo's volumes and partition files"}
            res["partitions_compared"] = []
            for i in parts:
                c, rr = os.path.join(td, "ref_cns_%d.fa" % i), os.path.join(td, "ref_raw_%d.fa" % i)
                sec, _ = timed([os.path.join(util.REF_DIR, "oc2cns")] + CNS.split() + ["-t", str(threads), gwrk, pc, c, rr, "-mn", str(i), str(count)])
                r["cns_s"].append(sec)
                log("ref oc2cns partition %d of %d: %.1f s" % (i, count, sec))
                row = {"partition": i, "gpu_s": g["cns_s"][i], "ref_s": sec}
                row["cns_identical"] = util.fasta_canon(gout[i][0]) == util.fasta_canon(c)
                row["raw_identical"] = util.fasta_canon(gout[i][1]) == util.fasta_canon(rr)
                res["partitions_compared"].append(row)
            res["ref"] = r
            ref_ok = False
        if ref_ok:
            parts = None if a.ref_parts == "all" else ([] if a.ref_parts == "none" else [int(x) for x in a.ref_parts.split(",")])
            r, rwrk, rcand, rout = run_arm(util.REF_DIR, td, "ref", fastas, threads, parts, log)
            rc = formats.read_candidates(rcand)
            r["n_candidates"] = int(rc.shape[0])
            res["ref"] = r
            log("REF " + json.dumps(r))
            res["volumes"] = same_volumes(gwrk, rwrk)
            res["candidates_identical"] = bool(util.same_multiset(gc, rc))
            del gc, rc
            res["partitions_compared"] = []
            for i, (c, rr) in rout.items():
                row = {"partition": i}
                row["cns_identical"] = util.fasta_canon(gout[i][0]) == util.fasta_canon(c)
                row["raw_identical"] = util.fasta_canon(gout[i][1]) == util.fasta_canon(rr)
                res["partitions_compared"].append(row)
            n = len(r["cns_parts"])
            if n:
                res["speedup_on_compared_stages"] = (sum(r["pmov_s"]) + r["pcan_s"] + sum(r["cns_s"])) / \
                    (sum(g["pmov_s"]) + g["pcan_s"] + sum(g["cns_s"][i] for i in r["cns_parts"]))
        log(json.dumps({k: v for k, v in res.items() if k not in ("gpu", "ref")}))
        dump()
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()
