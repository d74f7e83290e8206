#!/usr/bin/env python3
"""bench.py -- one correction iteration of NECAT's hot path (oc2pmov -> oc2pcan -> oc2cns)
on N B200s, BASELINE.json config[1]: E. coli-size 4.6 Mb genome, 40x synthetic ONT-like
reads (15 % error), OVLP_SENSITIVE + CNS_SENSITIVE options, -r 0 -f 0.

A "step" is one full pass over the read set: k-mer index build, candidate finding,
candidate regrouping, extension + consensus.  `value` = corrected bases per second
with the packed reads already resident in HBM; `e2e` = the same through the C-ABI
with host buffers (volume upload included; candidates and FASTA records always come
back to the host).  One process per GPU; N > 1 shards the query reads for the mapping
half, exchanges the 28-byte candidate records with an NCCL all-gather, and shards
the templates (id mod N) for the correction half.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

PM_OPTS = dict(kmer_size=15, scan_window=10, kmer_cnt_cutoff=500, block_size=2000, block_score_cutoff=3,
               num_candidates=500, align_size_cutoff=1000, ddfs_cutoff=0.25, error=0.5, num_output=500,
               num_threads=1, job=0, binary_output=1, use_hdr_as_id=1)            # necat.pl:20,32
CNS_OPTS = dict(min_align_size=2000, min_cov=4, max_cov=12, min_size=1000, full_consensus=0, error=0.5,
                mapping_ratio=0.8, rescue_long_indels=0, use_fixed_ident_cutoff=0)  # necat.pl:34,251-261
PM_STR = "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CNS_STR = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
SEED = 20260925


def make_reads(genome, depth):
    import gen_reads
    _, reads = gen_reads.generate(genome, depth, SEED)
    return reads


def reads_to_volume(reads):
    from necat_b200 import formats
    info = np.zeros(len(reads), formats.SEQINFO)
    hdr = bytearray()
    off = 0
    for i, r in enumerate(reads):
        info[i] = (off, r.size, len(hdr), 0, 0)
        hdr += b"r%d\0" % i
        off += r.size
    return formats.Volume(formats.pack_bases(np.concatenate(reads)), off, info, bytes(hdr))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(genome, depth, threads, td, reads=None):
    """The unmodified reference binaries (oracle/_ref) on a read set; returns timings + outputs."""
    from necat_b200 import formats
    ref = os.path.join(ROOT, "oracle", "_ref")
    if reads is None:
        reads = make_reads(genome, depth)
    wrk = os.path.join(td, "vol")
    formats.write_workdir(wrk, reads)
    cand = os.path.join(td, "cns_candidates.txt")
    t = {}
    t0 = time.perf_counter()
    log = subprocess.run([os.path.join(ref, "oc2pmov")] + PM_STR.split() + ["-t", str(threads), wrk, "0", cand],
                         check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    t["pmov_s"] = time.perf_counter() - t0
    for line in log.splitlines():
        if "'pairwise mapping v0 vs v0' takes" in line:
            t["pmov_mapping_s"] = float(line.split("takes")[1].split("secs")[0])
    t0 = time.perf_counter()
    subprocess.run([os.path.join(ref, "oc2pcan"), "-t", str(threads), wrk, cand], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t["pcan_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run([os.path.join(ref, "oc2cns")] + CNS_STR.split() + ["-t", str(threads), wrk, cand,
                   os.path.join(td, "cns.fasta"), os.path.join(td, "raw.fasta")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t["cns_s"] = time.perf_counter() - t0
    t["n_candidates"] = os.path.getsize(cand) // 28
    t["cns_bases"] = sum(len(s) for _, s in formats.read_fasta_records(os.path.join(td, "cns.fasta")))
    t["total_s"] = t["pmov_s"] + t["pcan_s"] + t["cns_s"]
    return t


def reference_sample_genome(args, threads):
    """Genome size the reference binaries are timed on: the full workload when the host can finish it in
    about a minute per pass (>= 12 threads: ~40 s on 16 and on 128), else a bounded sample.  The reference pays
    a fixed ~7 s per pass for its dense 8 GiB k-mer table, so a small sample understates its throughput."""
    if args.sample_genome > 0:
        return args.sample_genome
    return args.genome if threads >= 12 else max(400_000, args.genome // 8)


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation on the host cores, bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = os.path.join(ROOT, "oracle", "_ref", "oc2cns")
    if not os.path.exists(ref):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries are missing (build with make -C oracle ref)"}))
        return
    sample_genome = reference_sample_genome(args, threads)
    reads = make_reads(sample_genome, args.depth)
    # a CPU process needs no more than one warm-up pass (page cache); more would only burn minutes
    warm = min(args.warmup, 1)
    times = []
    last = None
    # The whole run has to end within a few minutes whatever --steps says: a pass of the full workload takes ~35-40 s, so
    # once the passes done plus one more would go past the budget the remaining ones are skipped (CPU timings of the
    # same binaries on the same files barely move from pass to pass) and the line says how many were executed.
    budget_s, t_begin = 300.0, time.perf_counter()
    for it in range(warm + args.steps):
        with tempfile.TemporaryDirectory() as td:
            last = run_reference(sample_genome, args.depth, threads, td, reads=reads)
        if it >= warm:
            times.append(last["total_s"])
        spent = time.perf_counter() - t_begin
        if times and spent + spent / (it + 1) > budget_s:
            break
    sec = float(np.mean(times))
    val = last["cns_bases"] / sec
    sample = (f"{sample_genome / 1e6:.2f} Mb genome x {args.depth:g} (same generator, seed {SEED}; "
              f"{'the full workload' if sample_genome == args.genome else 'bounded sample'}), oc2pmov+oc2pcan+oc2cns wall, "
              f"{warm} warm-up pass(es) and {len(times)} of {args.steps} timed pass(es) executed (time budget {budget_s:.0f} s)")
    print(json.dumps({
        "impl": "reference", "metric": "corrected_bases_per_sec", "value": val, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"E. coli-size {args.genome / 1e6:.1f} Mb x {args.depth:g} synthetic ONT reads, one correction iteration "
                               "(oc2pmov SENSITIVE -> oc2pcan -> oc2cns -r 0 -f 0)", "bounded_sample": sample},
        "overlaps_per_sec": last["n_candidates"] / sec,
        "stage_s": {k: last[k] for k in ("pmov_s", "pmov_mapping_s", "pcan_s", "cns_s") if k in last},
        "cpu_baseline": {"value": val, "unit": "bases/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--depth", type=float, default=40.0)
    ap.add_argument("--sample-genome", type=int, default=0,
                    help="genome size the CPU reference is timed on (0 = the full workload on hosts with >= 32 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from necat_b200 import capi, formats
    from necat_b200 import dist as nbdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's version banner (printed at NCCL_DEBUG=VERSION and above, WARN
        # included) goes to stderr with the rest of NCCL's log
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    reads = make_reads(args.genome, args.depth)
    vol = reads_to_volume(reads)
    nreads = vol.nseq
    sizes = vol.info["size"].astype(np.int64)
    ctx = capi.Context(local_rank)
    popt = capi.default_map_options(**PM_OPTS)
    copt = capi.default_cns_options(**CNS_OPTS)
    lo, hi = nbdist.shard_range(nreads, rank, world)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stats = {}

    def one_step(dvol):
        """index -> candidates -> exchange -> regroup -> correct.  Returns corrected bases of this rank."""
        w0 = time.perf_counter()
        idx = ctx.build_index(dvol, PM_OPTS["kmer_size"], PM_OPTS["kmer_cnt_cutoff"])
        t = ctx.timings(); stats["index_ms"] = t.index_ms
        w1 = time.perf_counter()
        cands = ctx.map_volume(idx, dvol, 0, dvol, 0, lo, hi, popt)
        w2 = time.perf_counter()
        t = ctx.timings(); stats["map_ms"] = t.map_ms; stats["map_probe_ms"] = t.map_probe_ms
        mc = ctx.map_counters()
        stats["map_counters"] = {k: int(getattr(mc, k)) for k, _ in mc._fields_}
        ctx.free_index(idx)
        packed = capi.pack_candidates(cands)
        # the one exchange step: every rank needs the candidates of all query shards
        packed = nbdist.all_gather_records(packed, dev, world)
        stats["n_candidates"] = int(packed.shape[0])
        # oc2pcan: every record under its subject and (swapped) under its query; one partition (< 100000 reads)
        part = np.concatenate([packed, formats.swap_roles(packed)])
        w3 = time.perf_counter()
        recs, seq = ctx.consensus_job(dvol, part, copt, world, rank)
        w4 = time.perf_counter()
        stats["wall_ms"] = {"index_call": (w1 - w0) * 1e3, "map_call": (w2 - w1) * 1e3, "exchange_regroup": (w3 - w2) * 1e3,
                            "consensus_call": (w4 - w3) * 1e3}
        t = ctx.timings(); stats["extend_ms"] = t.extend_ms; stats["consensus_ms"] = t.consensus_ms
        cc = ctx.cns_counters()
        stats["cns_counters"] = {k: int(getattr(cc, k)) for k, _ in cc._fields_}
        seq.free()
        return int(recs["seq_len"][recs["is_raw"] == 0].sum())

    def timed(n_steps, upload_each):
        dvol = None if upload_each else ctx.upload(vol)
        total_bases = 0
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            flush.fill_(1)                              # flush L2 between iterations
            torch.cuda.synchronize()
            d = ctx.upload(vol) if upload_each else dvol
            total_bases = one_step(d)
            if upload_each:
                ctx.free_volume(d)
        ev1.record()
        barrier()
        # the step mixes kernels on the library's own stream with host logic, all fenced by synchronize():
        # the two events (recorded while the device is idle) bracket the same interval as the host clock
        dt = max(ev0.elapsed_time(ev1) * 1e-3, 0.0) if n_steps else 0.0
        stats["host_clock_s"] = time.perf_counter() - t0
        if not upload_each:
            ctx.free_volume(dvol)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        bb = torch.tensor([total_bases], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        return float(tt.item()), int(bb.item())

    timed(args.warmup, False)
    launches0 = ctx.launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dt, bases = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    launches = (ctx.launches() - launches0) // max(args.steps, 1)
    tm0 = ctx.timings()
    dt_e2e, bases_e2e = timed(args.steps, True)
    tm1 = ctx.timings()
    # One more pass outside the timed region with the extension-round pipeline switched off (one template group,
    # one stream): the extension launches are then as few and as large as they can be, which is the launch the
    # roofline below describes; in the timed passes the rounds of several template groups alternate on the GPU
    # (still one kernel at a time), so there are more, smaller launches with a longer tail each.
    pipelined_stats = {k: stats.get(k) for k in ("extend_ms", "consensus_ms", "wall_ms")}
    keep = os.environ.get("NB_CNS_GROUPS")
    os.environ["NB_CNS_GROUPS"] = "1"
    dv = ctx.upload(vol)
    flush.fill_(1); torch.cuda.synchronize()
    one_step(dv)
    ctx.free_volume(dv)
    if keep is None:
        del os.environ["NB_CNS_GROUPS"]
    else:
        os.environ["NB_CNS_GROUPS"] = keep
    if rank != 0:
        return
    sec = dt / args.steps
    sec_e2e = dt_e2e / args.steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    cc, mc = stats["cns_counters"], stats["map_counters"]
    # dominant kernel = the extension kernel k_extend.  The roofline describes its largest launch (round 0
    # of the extensions, one template group: the extra pass above): algorithmic bytes of that launch = walk-back state
    # written once (8 B per 32-row word per DP column) + packed bases of every block (0.25 B per base, ~2 x 512 per
    # block) + one op byte per alignment column (~1.2 x 512 per block), divided by the launch's duration (CUDA events
    # on the library's stream, around the kernel alone).
    ext_bytes = cc["big_ext_word_updates"] * 8 + cc["big_ext_block_aligns"] * (0.25 * 1024 + 620)
    ext_s = cc["big_ext_usec"] * 1e-6
    ext_gbs = ext_bytes / ext_s / 1e9 if ext_s else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("word_updates") and cc["big_ext_word_updates"]:
            # DRAM bytes of the same launch under ncu --set full, scaled by the launch's word-updates
            traffic = tj["dram_bytes"] * cc["big_ext_word_updates"] / tj["word_updates"]
            traffic_src = tj.get("source")
    out = {
        "metric": "corrected_bases_per_sec", "value": bases / sec, "unit": "bases/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"E. coli-size {args.genome / 1e6:.1f} Mb x {args.depth:g} synthetic ONT reads (15% error, seed {SEED}), "
                               "one correction iteration: oc2pmov SENSITIVE (-j 0) -> oc2pcan -> oc2cns -r 0 -f 0",
                   "reads": int(nreads), "bases": int(vol.db_size), "parallelism": f"reads/templates sharded x{world}",
                   "l2": "512 MB flush write between iterations; index (4 GiB table) and DP scratch exceed L2"},
        "overlaps_per_sec": stats["n_candidates"] / sec,
        "corrected_bases": bases, "candidates": stats["n_candidates"],
        "stage_ms": dict({k: stats[k] for k in ("index_ms", "map_ms", "map_probe_ms", "extend_ms", "consensus_ms")},
                         note="from an extra pass outside the timed region with one template group (fewest, largest launches)",
                         extend_ms_timed_passes=pipelined_stats.get("extend_ms"), consensus_ms_timed_passes=pipelined_stats.get("consensus_ms")),
        "wall_ms": pipelined_stats.get("wall_ms"),
        "wall_ms_one_group": stats.get("wall_ms"),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": bases_e2e / sec_e2e, "unit": "bases/s", "ms_per_step": sec_e2e * 1e3,
                "h2d_bytes_per_step": int((tm1.h2d_bytes - tm0.h2d_bytes) // max(args.steps, 1)),
                "d2h_bytes_per_step": int((tm1.d2h_bytes - tm0.d2h_bytes) // max(args.steps, 1))},
        "roofline": {"kernel": "k_extend", "bound": "hbm", "achieved": ext_gbs, "peak": peak, "unit": "GB/s",
                     "frac": ext_gbs / peak, "frac_of_nominal_8TBs": ext_gbs / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "launch": {"what": "largest k_extend launch (round 0, one template group)", "ms": cc["big_ext_usec"] / 1e3,
                                "tasks": cc["big_ext_tasks"], "block_alignments": cc["big_ext_block_aligns"],
                                "word_updates": cc["big_ext_word_updates"], "algorithmic_bytes": ext_bytes},
                     "word_updates_per_s": cc["big_ext_word_updates"] / ext_s if ext_s else 0.0},
        "counters": {"map": mc, "cns": cc},
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "oc2cns")):
            sg = reference_sample_genome(args, threads)
            with tempfile.TemporaryDirectory() as td:
                r = run_reference(sg, args.depth, threads, td, reads=reads if sg == args.genome else None)
            out["cpu_baseline"] = {"value": r["cns_bases"] / r["total_s"], "unit": "bases/s", "cores": threads, "kind": "reference",
                                   "sample": f"{sg / 1e6:.2f} Mb genome x {args.depth:g} ({'the full workload' if sg == args.genome else 'bounded sample'}), same generator; "
                                             f"oc2pmov {r['pmov_s']:.1f}s (mapping {r.get('pmov_mapping_s', 0):.2f}s) + oc2pcan "
                                             f"{r['pcan_s']:.2f}s + oc2cns {r['cns_s']:.1f}s"}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": threads, "kind": "reference",
                                   "sample": "oracle/_ref missing"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
