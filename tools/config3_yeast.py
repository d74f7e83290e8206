#!/usr/bin/env python3
"""BASELINE.json config 3: S. cerevisiae-size genome (12 Mb x 50 synthetic ONT reads), NUM_ITER = 2 correction iterations on
2 B200s (necat.pl:160-365: iteration 1 SENSITIVE overlap options with -f 1, iteration 2 FAST options with -r 1), against
the unmodified reference binaries (oracle/_ref) run through the same chain on the host.

Launch:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/config3_yeast.py --out profiles/r02_config3_yeast_2gpu.json

Each iteration: the ranks read the PackedDB volumes of the work directory, necat_b200/pipeline.py does the overlap and
correction halves on N GPUs (index per volume, query reads split, exchange of the candidate records, templates of the one
partition id mod N), every rank writes its FASTA records, rank 0 puts them in id order with this repo's oc2ReorderCnsReads
and -- between iterations -- makes the next work directory with this repo's oc2mkdb, exactly the files necat.pl moves.
The reference arm (rank 0, afterwards): oc2mkdb, oc2pmov per volume, oc2pcan, oc2cns, oc2ReorderCnsReads of oracle/_ref.
Outputs are compared in canonical order after every iteration."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

PM1 = "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
PM2 = "-k 15 -z 20 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CNS1 = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 1"
CNS2 = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"


def pm_opts(z):
    return dict(kmer_size=15, scan_window=z, kmer_cnt_cutoff=500, block_size=2000, block_score_cutoff=3, num_candidates=500,
                align_size_cutoff=1000, ddfs_cutoff=0.25, error=0.5, num_output=500, num_threads=1, job=0, binary_output=1, use_hdr_as_id=1)


def cns_opts(r, f):
    return dict(min_align_size=2000, min_cov=4, max_cov=12, min_size=1000, full_consensus=f, error=0.5, mapping_ratio=0.8,
                rescue_long_indels=r, use_fixed_ident_cutoff=0)


def run(cmd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode:
        print(p.stdout[-3000:])
        raise SystemExit(f"{cmd[0]} failed with {p.returncode}")
    return time.perf_counter() - t0


def write_fasta(path, recs):
    with open(path, "wb") as f:
        for h, s in recs:
            f.write(h + b"\n" + s + b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=12_000_000)
    ap.add_argument("--depth", type=float, default=50.0)
    ap.add_argument("--seed", type=int, default=20260925)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import util
    from necat_b200 import capi, formats, pipeline
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    td = os.environ.get("NB_CONFIG3_DIR") or os.path.join(tempfile.gettempdir(), "nb_config3")
    B = util.BIN_DIR

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    res = {"config": "BASELINE.json configs[2]: %.1f Mb x %g synthetic ONT reads, 2 correction iterations" % (a.genome / 1e6, a.depth),
           "n_gpus": world, "iterations": []}
    if rank == 0:
        shutil.rmtree(td, ignore_errors=True)
        os.makedirs(td)
        import gen_reads
        t0 = time.perf_counter()
        _, reads = gen_reads.generate(a.genome, a.depth, a.seed)
        A = np.frombuffer(b"ACGT", np.uint8)
        with open(os.path.join(td, "reads.fa"), "wb") as f:
            for i, r in enumerate(reads):
                f.write(b">r%d\n" % i + A[r].tobytes() + b"\n")
        res["reads"] = len(reads); res["bases"] = int(sum(r.size for r in reads)); res["gen_s"] = time.perf_counter() - t0
        open(os.path.join(td, "list0"), "w").write(os.path.join(td, "reads.fa") + "\n")
        run([os.path.join(B, "oc2mkdb"), os.path.join(td, "gpu_wrk1"), os.path.join(td, "list0")])
        del reads
    barrier()
    ctx = capi.Context(local)
    outs = {}
    for it in (1, 2):
        wrk = os.path.join(td, "gpu_wrk%d" % it)
        nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
        vols = [formats.read_volume(os.path.join(wrk, "vol%d" % v)) for v in range(nvol)]
        names = [v.name(i).encode() for v in vols for i in range(v.nseq)]
        popt = capi.default_map_options(**pm_opts(10 if it == 1 else 20))
        copt = capi.default_cns_options(**cns_opts(0 if it == 1 else 1, 1 if it == 1 else 0))
        st = {}
        barrier()
        t0 = time.perf_counter()
        out = pipeline.correct_iteration(ctx, vols, popt, copt, rank, world, dev, stats=st)
        barrier()
        t_gpu = time.perf_counter() - t0
        cns, raw = [], []
        for recs, seq in out:
            c, r = formats.cns_fasta_records(recs, seq.array(), names)
            cns += c; raw += r
            seq.free()
        write_fasta(os.path.join(td, "gpu_it%d_cns.rank%d.fa" % (it, rank)), cns)
        write_fasta(os.path.join(td, "gpu_it%d_raw.rank%d.fa" % (it, rank)), raw)
        ncan = torch.tensor([st["n_candidates_local"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(ncan)
        barrier()
        row = {"iteration": it, "pm": PM1 if it == 1 else PM2, "cns": CNS1 if it == 1 else CNS2,
               "gpu": {"pipeline_s": t_gpu, "n_candidates": int(ncan.item()), "volumes": nvol,
                       "rank0_stage_ms": {k: st.get(k) for k in ("index_ms", "map_ms", "extend_ms", "consensus_ms", "wall_index_ms", "wall_map_ms", "wall_exchange_ms", "wall_consensus_ms")}}}
        if rank == 0:
            t1 = time.perf_counter()
            for kind in ("cns", "raw"):
                with open(os.path.join(td, "gpu_it%d_%s.tmp.fa" % (it, kind)), "wb") as o:
                    for r in range(world):
                        o.write(open(os.path.join(td, "gpu_it%d_%s.rank%d.fa" % (it, kind, r)), "rb").read())
            c_out, r_out = os.path.join(td, "gpu_it%d_cns.fa" % it), os.path.join(td, "gpu_it%d_raw.fa" % it)
            run([os.path.join(B, "oc2ReorderCnsReads"), os.path.join(td, "gpu_it%d_cns.tmp.fa" % it), os.path.join(td, "gpu_it%d_raw.tmp.fa" % it), c_out, r_out])
            if it == 1:
                open(os.path.join(td, "gpu_list1"), "w").write(c_out + "\n" + r_out + "\n")
                run([os.path.join(B, "oc2mkdb"), os.path.join(td, "gpu_wrk2"), os.path.join(td, "gpu_list1")])
            row["gpu"]["files_s"] = time.perf_counter() - t1
            row["gpu"]["total_s"] = t_gpu + row["gpu"]["files_s"]
            cn = util.fasta_canon(c_out)
            row["gpu"]["cns_records"] = len(cn); row["gpu"]["cns_bases"] = sum(len(s) for _, s in cn)
            row["gpu"]["raw_records"] = len(util.fasta_canon(r_out))
            outs[it] = (c_out, r_out)
            print("GPU iteration", it, json.dumps(row["gpu"]), flush=True)
        res["iterations"].append(row)
        barrier()
    ctx.close()
    if world > 1:
        dist.barrier()
    if rank == 0 and not a.skip_ref and util.ref_bin("oc2cns"):
        R = util.REF_DIR
        threads = os.cpu_count() or 1
        res["host_threads"] = threads
        lst = os.path.join(td, "list0")
        for it in (1, 2):
            wrk = os.path.join(td, "ref_wrk%d" % it)
            g = {}
            g["mkdb_s"] = run([os.path.join(R, "oc2mkdb"), wrk, lst])
            nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
            cand = os.path.join(td, "ref_it%d_cand.bin" % it)
            g["pmov_s"] = 0.0
            with open(cand, "wb") as o:
                for vid in range(nvol):
                    p = cand + ".v%d" % vid
                    g["pmov_s"] += run([os.path.join(R, "oc2pmov")] + (PM1 if it == 1 else PM2).split() + ["-t", str(threads), wrk, str(vid), p])
                    o.write(open(p, "rb").read())
            g["pcan_s"] = run([os.path.join(R, "oc2pcan"), "-t", str(threads), wrk, cand])
            ct, rt = os.path.join(td, "ref_it%d_cns.tmp.fa" % it), os.path.join(td, "ref_it%d_raw.tmp.fa" % it)
            g["cns_s"] = run([os.path.join(R, "oc2cns")] + (CNS1 if it == 1 else CNS2).split() + ["-t", str(threads), wrk, cand, ct, rt])
            c_out, r_out = os.path.join(td, "ref_it%d_cns.fa" % it), os.path.join(td, "ref_it%d_raw.fa" % it)
            g["reorder_s"] = run([os.path.join(R, "oc2ReorderCnsReads"), ct, rt, c_out, r_out])
            g["total_s"] = sum(g.values())
            g["n_candidates"] = os.path.getsize(cand) // 28
            row = res["iterations"][it - 1]
            row["ref"] = g
            row["candidates_same_count"] = g["n_candidates"] == row["gpu"]["n_candidates"]
            row["cns_identical"] = util.fasta_canon(outs[it][0]) == util.fasta_canon(c_out)
            row["raw_identical"] = util.fasta_canon(outs[it][1]) == util.fasta_canon(r_out)
            row["speedup_total"] = g["total_s"] / row["gpu"]["total_s"]
            print("REF iteration", it, json.dumps(g), json.dumps({k: row[k] for k in ("candidates_same_count", "cns_identical", "raw_identical", "speedup_total")}), flush=True)
            lst = os.path.join(td, "ref_list%d" % it)
            open(lst, "w").write(c_out + "\n" + r_out + "\n")
    if rank == 0:
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            json.dump(res, open(a.out, "w"), indent=1)
        shutil.rmtree(td, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
