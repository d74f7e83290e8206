#!/usr/bin/env python3
"""bench.py -- one correction iteration of NECAT's hot path (oc2pmov -> oc2pcan -> oc2cns)
on N B200s, BASELINE.json config[1]: E. coli-size 4.6 Mb genome, 40x synthetic ONT-like
reads (15 % error), OVLP_SENSITIVE + CNS_SENSITIVE options, -r 0 -f 0.

A "step" is one full pass over the read set: k-mer index build, candidate finding,
candidate regrouping, extension + consensus.  `value` = corrected bases per second
with the packed reads already resident in HBM; `e2e` = the same through the C-ABI
with host buffers (volume upload included; candidates and FASTA records always come
back to the host).  One process per GPU; N > 1 (necat_b200/pipeline.py): the index is built
in N shards of the hash space and all-gathered over NCCL, the query reads are split for
the mapping half, the 28-byte candidate records go through an all-to-all-v, and the
templates of the one partition are corrected id mod N.

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


def workload_name(args):
    """One string for both arms (the driver compares the two lines' config)."""
    return (f"E. coli-size {args.genome / 1e6:.1f} Mb x {args.depth:g} synthetic ONT reads (15% error, seed {SEED}), "
            "one correction iteration: oc2pmov SENSITIVE (-j 0) -> oc2pcan -> oc2cns -r 0 -f 0")


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


def exe_parity(td, threads, ref):
    """Runs necat_b200/bin/{oc2pmov,oc2pcan,oc2cns} on td/vol (where run_reference left the reference's outputs) and
    compares.  -> (parity dict, e2e_exe dict)"""
    from necat_b200 import formats
    import shutil
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    B = os.path.join(ROOT, "necat_b200", "bin")
    wrk = os.path.join(td, "vol")
    cand = os.path.join(td, "g_cand.bin")
    t0 = time.perf_counter()
    subprocess.run([os.path.join(B, "oc2pmov")] + PM_STR.split() + ["-t", str(threads), wrk, "0", cand], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    part = cand + ".part"
    shutil.copy(cand, part)
    subprocess.run([os.path.join(B, "oc2pcan"), "-t", str(threads), wrk, part], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t2 = time.perf_counter()
    subprocess.run([os.path.join(B, "oc2cns")] + CNS_STR.split() + ["-t", str(threads), wrk, part, os.path.join(td, "g_cns.fasta"),
                   os.path.join(td, "g_raw.fasta")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t3 = time.perf_counter()
    gc, rc = formats.read_candidates(cand), formats.read_candidates(os.path.join(td, "cns_candidates.txt"))
    g_cns, r_cns = util.fasta_canon(os.path.join(td, "g_cns.fasta")), util.fasta_canon(os.path.join(td, "cns.fasta"))
    parity = {"candidates": bool(util.same_multiset(gc, rc)), "cns": g_cns == r_cns,
              "raw": util.fasta_canon(os.path.join(td, "g_raw.fasta")) == util.fasta_canon(os.path.join(td, "raw.fasta")),
              "n_candidates": int(gc.shape[0]), "cns_records": len(g_cns),
              "against": "the unmodified reference binaries (oracle/_ref) on the same work directory, canonical order, off the clock"}
    wall = t3 - t0
    bases = sum(len(x) for _, x in g_cns)
    e2e_exe = {"value": bases / wall, "unit": "bases/s", "ms_per_step": wall * 1e3,
               "stage_s": {"oc2pmov": t1 - t0, "oc2pcan": t2 - t1, "oc2cns": t3 - t2},
               "what": "the three drop-in executables as necat.pl runs them (cold processes, files in and out)",
               "speedup_vs_reference_executables": ref["total_s"] / wall}
    # The same three executables with the resident helper (NECAT_B200_HELPER=1, necat_b200/host/nb_helper.h): every process
    # hands its command line to one helper process that keeps the CUDA context and its cached device blocks between calls.
    # Pass 1 starts the helper (one context creation instead of two), pass 2 is what every later call of a necat.pl run sees.
    henv = dict(os.environ, NECAT_B200_HELPER="1", NECAT_B200_HELPER_IDLE="8", NECAT_B200_HELPER_DIR=td)
    passes = []
    same = True
    for k in range(2):
        hc = os.path.join(td, f"h{k}_cand.bin")
        h0 = time.perf_counter()
        subprocess.run([os.path.join(B, "oc2pmov")] + PM_STR.split() + ["-t", str(threads), wrk, "0", hc], check=True, env=henv,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        h1 = time.perf_counter()
        hp = hc + ".part"
        shutil.copy(hc, hp)
        subprocess.run([os.path.join(B, "oc2pcan"), "-t", str(threads), wrk, hp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        h2 = time.perf_counter()
        subprocess.run([os.path.join(B, "oc2cns")] + CNS_STR.split() + ["-t", str(threads), wrk, hp, os.path.join(td, f"h{k}_cns.fasta"),
                       os.path.join(td, f"h{k}_raw.fasta")], check=True, env=henv, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        h3 = time.perf_counter()
        passes.append({"ms_per_step": (h3 - h0) * 1e3, "stage_s": {"oc2pmov": h1 - h0, "oc2pcan": h2 - h1, "oc2cns": h3 - h2}})
        same = same and open(hc, "rb").read() == open(cand, "rb").read() \
            and open(os.path.join(td, f"h{k}_cns.fasta"), "rb").read() == open(os.path.join(td, "g_cns.fasta"), "rb").read() \
            and open(os.path.join(td, f"h{k}_raw.fasta"), "rb").read() == open(os.path.join(td, "g_raw.fasta"), "rb").read()
    e2e_exe["resident_helper"] = {
        "what": "the same executables with NECAT_B200_HELPER=1: pass 1 starts the helper process, pass 2 finds it (context, cached device blocks and pinned buffers kept)",
        "first_pass": passes[0], "later_passes": passes[1], "files_identical_to_stand_alone": bool(same),
        "value": bases / (passes[1]["ms_per_step"] * 1e-3), "unit": "bases/s",
        "speedup_vs_reference_executables": ref["total_s"] / (passes[1]["ms_per_step"] * 1e-3)}
    return parity, e2e_exe


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
        "steps": len(times), "steps_requested": args.steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "bounded_sample": sample},
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
    from necat_b200 import capi, formats, pipeline
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
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stats = {}

    def one_step(dvol):
        """index -> candidates -> exchange -> regroup -> correct (necat_b200/pipeline.py: the index of the one volume is
        built in `world` shards of the hash space and all-gathered, the query reads are split over the ranks, the records
        go through the all-to-all-v, the one partition is shared by all ranks).  Returns corrected bases of this rank."""
        st = {}
        w0 = time.perf_counter()
        res = pipeline.correct_iteration(ctx, [vol], popt, copt, rank, world, dev, part_size=pipeline.PART_SIZE, stats=st,
                                         dvols={0: dvol}, dmerged=dvol)
        bases = 0
        for recs, seq in res:
            bases += int(recs["seq_len"][recs["is_raw"] == 0].sum())
            seq.free()
        stats.update(st)
        stats["wall_ms"] = {"index_call": st["wall_index_ms"], "map_call": st["wall_map_ms"], "exchange_regroup": st["wall_exchange_ms"],
                            "consensus_call": st["wall_consensus_ms"], "step": (time.perf_counter() - w0) * 1e3}
        return bases

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
    # candidates found by all ranks (off the clock)
    nc = torch.tensor([stats["n_candidates_local"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nc, op=dist.ReduceOp.SUM)
    stats["n_candidates"] = int(nc.item())
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
    # Dominant kernel = the extension kernel k_extend.  The roofline describes its largest launch (round 0 of
    # the extensions, one template group: the extra pass above).  Algorithmic bytes of that launch: per DP column the
    # band of (Pv, Ph) words the walk-back needs, written once by the forward pass and read once by the walk (4 words x
    # 8 B each way; the last block of a side keeps 8 words) + the packed bases of every block (0.25 B per base, ~2 x 512
    # per block) + the op bytes of its columns (written to the block's buffer, read, written to the alignment:
    # 3 x ~620 B per block).  Divided by the launch's duration (CUDA events on the library's stream, around the kernel
    # alone).  The kernel is bound by instruction issue, not by HBM ("work, not bytes, is the natural
    # unit"): word-updates per second are reported beside the bytes.
    ext_bytes = cc["big_ext_dp_columns"] * 64 + cc["big_ext_block_aligns"] * (0.25 * 1024 + 3 * 620)
    ext_s = cc["big_ext_usec"] * 1e-6
    ext_gbs = ext_bytes / ext_s / 1e9 if ext_s else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if tj.get("k_extend", {}).get("dp_columns") and cc["big_ext_dp_columns"]:
        # DRAM bytes of the same launch under ncu --set full, scaled by the launch's DP columns
        traffic = tj["k_extend"]["dram_bytes"] * cc["big_ext_dp_columns"] / tj["k_extend"]["dp_columns"]
        traffic_src = tj["k_extend"].get("source")
    # The kernel north_star names: seed probe + vote (k_probe + k_build_blocks).  Algorithmic bytes (with this
    # repo's 4-byte CSR entries): the packed reads once per strand + 8 B of row pointers per probe + 4 B per hit + 12 B
    # per kept seed; DRAM traffic of the two kernels from the ncu capture (random 8-byte gathers fetch whole sectors).
    sv_bytes = mc["query_bases"] * 0.25 + mc["n_probes"] * 8 + mc["n_hits"] * 4 + mc["n_seeds"] * 12
    sv_s = stats["map_probe_ms"] * 1e-3
    sv = {"kernel": "k_probe + k_build_blocks", "bound": "hbm", "achieved": sv_bytes / sv_s / 1e9 if sv_s else 0.0, "peak": peak, "unit": "GB/s",
          "algorithmic_bytes": sv_bytes, "ms": stats["map_probe_ms"], "traffic": None}
    sv["frac"] = sv["achieved"] / peak
    if tj.get("seed_vote", {}).get("n_probes") and mc["n_probes"]:
        sv["traffic"] = tj["seed_vote"]["dram_bytes"] * mc["n_probes"] / tj["seed_vote"]["n_probes"]
        sv["traffic_source"] = tj["seed_vote"].get("source")
        sv["dram_frac_of_peak"] = sv["traffic"] / sv_s / 1e9 / peak if sv_s else None
    out = {
        "metric": "corrected_bases_per_sec", "value": bases / sec, "unit": "bases/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "reads": int(nreads), "bases": int(vol.db_size),
                   "parallelism": f"x{world}: index of the one volume built by every rank (10 ms; volumes of >= 600 Mbp: in {world} shards of the "
                                  "hash space + NCCL all-gather), query reads split, exchange of the 28-byte records (one partition shared by all "
                                  "ranks: all-gather), templates of the one partition id mod N",
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
                                "tasks": cc["big_ext_tasks"], "block_alignments": cc["big_ext_block_aligns"], "dp_columns": cc["big_ext_dp_columns"],
                                "word_updates": cc["big_ext_word_updates"], "algorithmic_bytes": ext_bytes},
                     "word_updates_per_s": cc["big_ext_word_updates"] / ext_s if ext_s else 0.0},
        "roofline_seed_vote": sv,
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
                # Off the clock: this repo's drop-in executables on the SAME work directory, outputs compared with the
                # reference's in canonical order, and their wall clock as necat.pl would see it (cold
                # processes: CUDA context creation included).
                out["parity"], out["e2e_exe"] = exe_parity(td, threads, r)
        else:
            out["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": threads, "kind": "reference",
                                   "sample": "oracle/_ref missing"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
