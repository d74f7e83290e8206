"""One correction iteration of the hot path (oc2pmov -> oc2pcan -> oc2cns) on N GPUs of one box, one process per GPU
(SURVEY.md 8e).  Host-side orchestration over the C-ABI (necat_b200.capi); torch.distributed carries the one exchange.

Layout, following the reference's own job structure:

* mapping half (pm_one_volume/pm_worker.c:372): with V volumes there are V (V + 1) / 2 (index volume i, query volume
  j >= i) pairs.  Index volumes go round-robin over the ranks; a rank builds the index of each volume it owns and
  streams the query volumes j >= i through it -- no traffic between GPUs while mapping.  With fewer volumes than ranks
  (E. coli, yeast: V = 1) the ranks that share an index volume split the query read range, and the index of that volume
  is built ONCE, in shards of the hash space (one per rank of the group), and all-gathered over NCCL.
* the exchange (partition_candidates/pcan.c:58-99): every 28-byte record has to reach the rank(s) that correct its
  subject read and, with the roles of the two reads swapped, the rank(s) that correct its query read: an all-to-all-v
  keyed by the template's partition (id / part_size, part_size = 100 000 in necat.pl's flow).
* correction half (consensus/main.c:71-73): a partition is an independent job.  Partitions go round-robin over the
  ranks; with fewer partitions than ranks a partition is shared by S ranks, rank s of them corrects the templates with
  id % S == s (nb_consensus_job_sharded).  Every rank holds all reads (oc2cns merges the volumes, main.c:66).

The union of the ranks' records is what one GPU -- and the reference -- produce; tests/test_dist_cpu.py checks the
plumbing at world size 2 over gloo, tests/test_gpu_multi.py the whole thing on GPUs.
"""
import os

import numpy as np

from . import formats

PART_SIZE = 100000          # oc2pcan -p (necat.pl never passes it: partition_candidates/main.c default)
INDEX_SHARD_MIN_BASES = 600_000_000     # volumes below this are indexed by every rank that needs them (see correct_iteration)


# ------------------------------------------------------------------------------------------------ who does what
def mapping_plan(n_vol, world):
    """-> per rank a list of (index volume i, query volume j, share, shares): the rank maps the `share`-th of `shares`
    equal parts of the reads of volume j against the index of volume i."""
    plan = [[] for _ in range(world)]
    if n_vol >= world:
        for i in range(n_vol):
            for j in range(i, n_vol):
                plan[i % world].append((i, j, 0, 1))
        return plan
    groups = index_groups(n_vol, world)
    for i, ranks in enumerate(groups):
        for s, r in enumerate(ranks):
            for j in range(i, n_vol):
                plan[r].append((i, j, s, len(ranks)))
    return plan


def index_groups(n_vol, world):
    """Ranks that share index volume i when world > n_vol: rank r works on volume r % n_vol."""
    return [[r for r in range(world) if r % n_vol == i] for i in range(n_vol)]


def correction_plan(n_part, world):
    """-> (S, units per rank): S ranks share a partition when there are fewer partitions than ranks; a unit is
    (partition p, shard s)."""
    S = max(1, world // max(n_part, 1))
    units = [[] for _ in range(world)]
    for p in range(n_part):
        for s in range(S):
            units[(p * S + s) % world].append((p, s))
    return S, units


def record_destinations(template_ids, part_size, S, world):
    """Ranks (S per record) that correct the templates: unit (p, s) lives on rank (p S + s) % world."""
    p = np.asarray(template_ids, np.int64) // part_size
    return [((p * S + s) % world).astype(np.int64) for s in range(S)]


# ------------------------------------------------------------------------------------------------ the exchange
def exchange_records(packed, n_reads, world, rank, device, part_size=PART_SIZE):
    """All-to-all-v of (n, 7) uint32 candidate records.  Every record is sent to the ranks that own the partition of its
    subject read and, role-swapped (gapped_candidate.c:54-69), to those of its query read.  Returns the records this
    rank corrects, grouped nowhere in particular (nb_consensus_job groups by template)."""
    n_part = (int(n_reads) + part_size - 1) // part_size
    S, _ = correction_plan(n_part, world)
    from . import capi
    if world == 1:
        return capi.with_swapped_roles([packed])
    import torch
    import torch.distributed as dist
    if S == world:
        # One partition shared by all ranks (E. coli, yeast: fewer than 100 000 reads): every rank corrects a share of the
        # templates of every record, so every rank needs every record -- the all-to-all-v degenerates into an all-gather
        # of the records as they were found; the role swap (gapped_candidate.c:54-69) happens after the receipt.
        n = int(packed.shape[0])
        cnt = torch.tensor([n], dtype=torch.int64, device=device)
        allc = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allc, cnt)
        counts = [int(x) for x in allc.cpu().tolist()]
        width = max(max(counts), 1)
        send = torch.zeros(width * 7, dtype=torch.int32, device=device)
        if n:
            send[:n * 7] = torch.from_numpy(np.ascontiguousarray(packed).view(np.int32).reshape(-1)).to(device)
        recv = torch.empty(world * width * 7, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(recv, send)
        r = recv.cpu().numpy().view(np.uint32).reshape(world, width, 7)
        return capi.with_swapped_roles([r[i, :counts[i]] for i in range(world)])
    both = np.concatenate([packed, formats.swap_roles(packed)]) if packed.shape[0] else np.zeros((0, 7), np.uint32)
    out_parts, counts = [], np.zeros(world, np.int64)
    dests = record_destinations(both[:, 1], part_size, S, world)
    send = []
    for r in range(world):
        sel = np.zeros(both.shape[0], bool)
        for d in dests:
            sel |= d == r
        rows = both[sel]
        counts[r] = rows.shape[0] * 7
        send.append(rows.reshape(-1))
    send_t = torch.from_numpy(np.concatenate(send).view(np.int32).copy() if counts.sum() else np.zeros(0, np.int32)).to(device)
    cnt_t = torch.from_numpy(counts).to(device)
    rcv_t = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rcv_t, cnt_t)
    rcv = [int(x) for x in rcv_t.cpu().tolist()]
    out_t = torch.empty(sum(rcv), dtype=torch.int32, device=device)
    dist.all_to_all_single(out_t, send_t, output_split_sizes=rcv, input_split_sizes=[int(x) for x in counts])
    return out_t.cpu().numpy().view(np.uint32).reshape(-1, 7)


# ------------------------------------------------------------------------------------------------ one iteration
def merge_volumes(vols):
    """consensus/main.c:66: one read store with global ids.  The packed bases of a volume follow those of the previous one
    without a gap (pdb_merge, packed_db.c:419-448): at a bit offset of 0, 2, 4 or 6, so whole bytes are shifted."""
    if len(vols) == 1:
        return vols[0]
    info = np.concatenate([v.info for v in vols]).copy()
    total = sum(int(v.db_size) for v in vols)
    pac = np.zeros((total + 3) // 4 + 1, np.uint8)
    off, hdr_off, hdr = 0, 0, b""
    k = 0
    for v in vols:
        info["offset"][k:k + v.nseq] = v.info["offset"] + off
        info["hdr_offset"][k:k + v.nseq] = v.info["hdr_offset"] + hdr_off
        n = int(v.db_size)
        nb = (n + 3) // 4
        src = v.pac[:nb]
        if n & 3 and nb:                                    # bits behind the last base of a volume stay out of the store
            src = src.copy()
            src[-1] &= (0xff << (8 - 2 * (n & 3))) & 0xff
        r = 2 * (off & 3)
        o = off >> 2
        if r == 0:
            pac[o:o + nb] = src
        else:
            pac[o:o + nb] |= src >> r
            pac[o + 1:o + nb + 1] |= (src << (8 - r)).astype(np.uint8)
        off += n; hdr_off += len(v.hdr); hdr += v.hdr
        k += v.nseq
    return formats.Volume(pac[:(total + 3) // 4], off, info, hdr)


def correct_iteration(ctx, vols, popt, copt, rank=0, world=1, device=None, part_size=PART_SIZE, stats=None, dvols=None, dmerged=None):
    """vols: the PackedDB volumes (formats.Volume) in id order.  Returns the list of (records, sequence buffer) of the
    correction units of this rank.  dvols / dmerged: device handles of the volumes / of the merged store when the
    caller keeps them resident."""
    from . import capi
    if world > 1 and "NB_HOST_THREADS" not in os.environ:
        # the ranks of a host share its cores: the host phases of the library (replay, request gathering, record filling)
        # get cores / local world size threads each instead of one thread per core in every rank
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        os.environ["NB_HOST_THREADS"] = str(max(2, (os.cpu_count() or 4) // max(1, local_world)))
    n_vol = len(vols)
    start = np.concatenate([[0], np.cumsum([v.nseq for v in vols])]).astype(np.int64)
    n_reads = int(start[-1])
    own_dvols = dvols is None
    if own_dvols:
        dvols = {}
    plan = mapping_plan(n_vol, world)[rank]
    groups = index_groups(n_vol, world) if n_vol < world else None
    import time
    found = []
    cur_i, idx = None, None
    if stats is not None:
        stats.update(index_ms=0.0, map_ms=0.0, map_probe_ms=0.0, wall_index_ms=0.0, wall_map_ms=0.0)
    for (i, j, s, shares) in plan:
        for v in (i, j):
            if v not in dvols:
                dvols[v] = ctx.upload(vols[v])
        if i != cur_i:
            if idx is not None:
                ctx.free_index(idx)
            w0 = time.perf_counter()
            # The ranks that share an index volume either build its index ONCE, in shards of the hash space, and all-gather
            # it (4.3 GB of row pointers + the positions over NVLink: ~7 ms at 2 ranks, ~13 ms at 4), or every one of them
            # builds all of it (no collective; 10 ms for a 184 Mbp volume, proportional to the volume).  Measured on the
            # E. coli-size set the replicated build wins at every N (N = 4: 10.0 ms against 16.5); a volume of the full
            # 2 Gbp takes ~110 ms to index, there the sharded build wins.  NECAT_B200_INDEX=sharded|replicated overrides.
            mode = os.environ.get("NECAT_B200_INDEX", "replicated" if vols[i].db_size < INDEX_SHARD_MIN_BASES else "sharded")
            if groups is not None and len(groups[i]) > 1 and mode != "replicated":
                idx = ctx.build_index_sharded(dvols[i], popt.kmer_size, popt.kmer_cnt_cutoff, groups[i], rank, device,
                                              group=None if len(groups[i]) == world else group_of(groups[i]))
            else:
                idx = ctx.build_index(dvols[i], popt.kmer_size, popt.kmer_cnt_cutoff)
            cur_i = i
            if stats is not None:
                stats["index_ms"] += ctx.timings().index_ms
                stats["wall_index_ms"] += (time.perf_counter() - w0) * 1e3
        n = vols[j].nseq
        lo, hi = (n * s) // shares, (n * (s + 1)) // shares
        w0 = time.perf_counter()
        found.append(ctx.map_volume_packed(idx, dvols[i], int(start[i]), dvols[j], int(start[j]), lo, hi, popt))
        if stats is not None:
            t = ctx.timings()
            stats["map_ms"] += t.map_ms; stats["map_probe_ms"] += t.map_probe_ms
            stats["wall_map_ms"] += (time.perf_counter() - w0) * 1e3
            mc = ctx.map_counters()
            for k_, _ in mc._fields_:
                stats.setdefault("map_counters", {})
                stats["map_counters"][k_] = stats["map_counters"].get(k_, 0) + int(getattr(mc, k_))
    if idx is not None:
        ctx.free_index(idx)
    packed = np.concatenate(found) if found else np.zeros((0, 7), np.uint32)
    w0 = time.perf_counter()
    mine = exchange_records(packed, n_reads, world, rank, device, part_size)
    if stats is not None:
        stats["n_candidates_local"] = int(packed.shape[0])
        stats["wall_exchange_ms"] = (time.perf_counter() - w0) * 1e3
    w0 = time.perf_counter()
    # ---- correction
    n_part = (n_reads + part_size - 1) // part_size
    S, units = correction_plan(n_part, world)
    out = []
    if units[rank]:
        own_merged = dmerged is None
        if own_merged:
            dmerged = dvols[0] if n_vol == 1 and 0 in dvols else ctx.upload(merge_volumes(vols))
        tpl_part = mine[:, 1].astype(np.int64) // part_size
        for (p, s) in units[rank]:
            recs = mine[tpl_part == p]
            out.append(ctx.consensus_job(dmerged, recs, copt, S, s))
            if stats is not None:
                t = ctx.timings()
                stats["extend_ms"] = stats.get("extend_ms", 0.0) + t.extend_ms
                stats["consensus_ms"] = stats.get("consensus_ms", 0.0) + t.consensus_ms
                cc = ctx.cns_counters()
                stats["cns_counters"] = {k_: int(getattr(cc, k_)) for k_, _ in cc._fields_}
        if own_merged and not (n_vol == 1 and 0 in dvols):
            ctx.free_volume(dmerged)
    if stats is not None:
        stats["wall_consensus_ms"] = (time.perf_counter() - w0) * 1e3
    if own_dvols:
        for h in dvols.values():
            ctx.free_volume(h)
    return out


_groups = {}


def group_of(ranks):
    """Process group of a subset of ranks (created once, by every rank, in the same order: see make_groups)."""
    return _groups[tuple(ranks)]


def make_groups(n_vol, world):
    """Call on every rank after init_process_group when world > n_vol > 1: the ranks that share an index volume
    all-gather its index among themselves."""
    import torch.distributed as dist
    if world <= n_vol or n_vol <= 1:
        return
    for ranks in index_groups(n_vol, world):
        if len(ranks) > 1 and tuple(ranks) not in _groups:
            _groups[tuple(ranks)] = dist.new_group(ranks)
