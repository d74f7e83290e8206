"""ctypes binding of libnecat_b200.so (include/necat_b200.h).  The library has no
CPU fallback: creating a context without a B200 raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnecat_b200.so")


class MapOptions(C.Structure):
    _fields_ = [("kmer_size", C.c_int), ("scan_window", C.c_int), ("kmer_cnt_cutoff", C.c_int),
                ("block_size", C.c_int), ("block_score_cutoff", C.c_int), ("num_candidates", C.c_int),
                ("align_size_cutoff", C.c_int), ("ddfs_cutoff", C.c_double), ("error", C.c_double),
                ("num_output", C.c_int), ("num_threads", C.c_int), ("job", C.c_int),
                ("binary_output", C.c_int), ("use_hdr_as_id", C.c_int)]


class CnsOptions(C.Structure):
    _fields_ = [("min_align_size", C.c_int), ("min_cov", C.c_int), ("max_cov", C.c_int), ("min_size", C.c_int),
                ("full_consensus", C.c_int), ("error", C.c_double), ("mapping_ratio", C.c_double),
                ("num_threads", C.c_int), ("rescue_long_indels", C.c_int), ("use_fixed_ident_cutoff", C.c_int),
                ("small_memory", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("index_ms", C.c_float), ("map_ms", C.c_float), ("map_probe_ms", C.c_float),
                ("extend_ms", C.c_float), ("consensus_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class MapCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_units", "n_probes", "n_hits", "n_seeds", "n_blocks", "n_eligible",
                                          "n_candidates", "query_bases")]


class CnsCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_templates", "n_extensions", "n_block_aligns", "n_dp_columns",
                                          "n_dp_word_updates", "n_accepted", "n_tags", "n_cns_records", "cns_bases",
                                          "big_ext_tasks", "big_ext_block_aligns", "big_ext_word_updates", "big_ext_usec", "big_ext_dp_columns")]


CANDIDATE = np.dtype([("qid", "<i4"), ("sid", "<i4"), ("qdir", "<i4"), ("sdir", "<i4"), ("score", "<i4"),
                      ("pad", "<i4"), ("qbeg", "<i8"), ("qend", "<i8"), ("qsize", "<i8"), ("sbeg", "<i8"),
                      ("send", "<i8"), ("ssize", "<i8"), ("qoff", "<i8"), ("soff", "<i8")])
ALIGNMENT = np.dtype([("ok", "<i4"), ("qoff", "<i4"), ("qend", "<i4"), ("toff", "<i4"), ("tend", "<i4"),
                      ("aln_size", "<i4"), ("n_match", "<i4"), ("long_ins", "<i4"), ("ops_offset", "<u8")])
CNS_RECORD = np.dtype([("id", "<i4"), ("left", "<i4"), ("right", "<i4"), ("org_size", "<i4"), ("num_can", "<i4"),
                       ("num_ovlps", "<i4"), ("ident_cutoff", "<f8"), ("seq_offset", "<u8"), ("seq_len", "<i4"),
                       ("is_raw", "<i4")])

EXPORTS = ["nb_ctx_create", "nb_ctx_destroy", "nb_last_error", "nb_version", "nb_launch_count", "nb_get_timings",
           "nb_volume_upload", "nb_volume_free", "nb_index_build", "nb_index_free", "nb_index_stats",
           "nb_index_download", "nb_index_build_shard", "nb_index_shard_export", "nb_index_adopt", "nb_map_options_default", "nb_map_volume", "nb_map_volume_aln", "nb_pack_candidate", "nb_pack_candidates", "nb_swap_candidate_roles", "nb_free",
           "nb_get_map_counters", "nb_cns_options_default", "nb_extend_candidates", "nb_consensus_job",
           "nb_consensus_job_sharded",
           "nb_get_cns_counters", "nb_rescue_local", "nb_rescue_global"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.nb_last_error.restype = C.c_char_p
        L.nb_last_error.argtypes = [C.c_void_p]
        L.nb_version.restype = C.c_char_p
        L.nb_launch_count.restype = C.c_uint64
        L.nb_launch_count.argtypes = [C.c_void_p]
        for f in ("nb_ctx_destroy", "nb_free", "nb_pack_candidates", "nb_swap_candidate_roles"):
            getattr(L, f).restype = None
        L.nb_pack_candidates.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


class NbError(RuntimeError):
    pass


class HostBuffer:
    """A malloc()ed buffer returned by the library: numpy view without copy, nb_free on release."""

    def __init__(self, L, addr, nbytes):
        self._L, self._addr, self.nbytes = L, addr, int(nbytes)

    def array(self):
        if not self.nbytes:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_char * self.nbytes).from_address(self._addr), np.uint8)

    def slice(self, off, n):
        return self.array()[int(off):int(off) + int(n)].tobytes()

    def free(self):
        if self._addr:
            self._L.nb_free(C.c_void_p(self._addr))
            self._addr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class Context:
    """One GPU context (nb_ctx)."""

    def __init__(self, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.nb_ctx_create(C.c_int(device), C.byref(self.h))
        if rc:
            raise NbError(self.L.nb_last_error(None).decode())

    def close(self):
        if self.h:
            self.L.nb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def _check(self, rc):
        if rc:
            raise NbError(self.L.nb_last_error(self.h).decode())

    def upload(self, vol):
        """vol: formats.Volume -> device volume handle"""
        off = np.ascontiguousarray(vol.info["offset"], dtype=np.uint64)
        size = np.ascontiguousarray(vol.info["size"], dtype=np.uint64)
        pac = np.ascontiguousarray(vol.pac, dtype=np.uint8)
        h = C.c_void_p()
        self._check(self.L.nb_volume_upload(self.h, pac.ctypes.data_as(C.c_void_p), C.c_uint64(vol.db_size),
                                            _u64p(off), _u64p(size), C.c_uint64(vol.nseq), C.byref(h)))
        return h

    def free_volume(self, h):
        self.L.nb_volume_free(self.h, h)

    def build_index(self, dvol, k, q):
        h = C.c_void_p()
        self._check(self.L.nb_index_build(self.h, dvol, C.c_int(k), C.c_int(q), C.byref(h)))
        return h

    def free_index(self, h):
        self.L.nb_index_free(self.h, h)
        self._index_keep.pop(h.value, None)

    _index_keep = {}        # torch tensors that back an assembled index stay alive with its handle

    def build_index_sharded(self, dvol, k, q, group_ranks, rank, device, group=None):
        """The index of one volume built ONCE by the ranks of `group_ranks` together: rank number s of them builds the
        CSR of the s-th slice of the hash space (nb_index_build_shard), the slices are all-gathered over NCCL into a
        table every rank holds (lookup_table.c:149-164 builds it serially on one host)."""
        import torch
        import torch.distributed as dist
        n = len(group_ranks)
        s = group_ranks.index(rank)
        sh = C.c_void_p(); n_off = C.c_uint64(); n_dist = C.c_uint64()
        self._check(self.L.nb_index_build_shard(self.h, dvol, C.c_int(k), C.c_int(q), C.c_int(s), C.c_int(n),
                                                C.byref(sh), C.byref(n_off), C.byref(n_dist)))
        n_hash = 1 << (2 * k)
        slice_len = (n_hash + n - 1) // n
        cnt = torch.tensor([n_off.value, n_dist.value], dtype=torch.int64, device=device)
        allc = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(n)]
        dist.all_gather(allc, cnt, group=group)
        offs_n = [int(c[0].item()) for c in allc]
        total, distinct = sum(offs_n), sum(int(c[1].item()) for c in allc)
        base = sum(offs_n[:s])
        tbl = torch.empty(n * slice_len + 8, dtype=torch.int32, device=device)
        offs = torch.empty(total + 16, dtype=torch.int32, device=device)
        width = max(max(offs_n), 1)
        mine = torch.zeros(width, dtype=torch.int32, device=device)
        self._check(self.L.nb_index_shard_export(self.h, sh, C.c_uint64(base), C.c_void_p(tbl.data_ptr() + 4 * s * slice_len),
                                                 C.c_void_p(mine.data_ptr())))
        torch.cuda.synchronize(device)
        dist.all_gather_into_tensor(tbl[:n * slice_len], tbl[s * slice_len:(s + 1) * slice_len], group=group)
        bufs = [torch.empty(width, dtype=torch.int32, device=device) for _ in range(n)]
        dist.all_gather(bufs, mine, group=group)
        pos = 0
        for b, c in zip(bufs, offs_n):
            offs[pos:pos + c] = b[:c]
            pos += c
        tbl[n_hash] = total if total < 2 ** 31 else total - 2 ** 32
        torch.cuda.synchronize(device)
        h = C.c_void_p()
        self._check(self.L.nb_index_adopt(self.h, C.c_int(k), C.c_int(q), C.c_void_p(tbl.data_ptr()), C.c_void_p(offs.data_ptr()),
                                          C.c_uint64(total), C.c_uint64(distinct), C.byref(h)))
        self._index_keep[h.value] = (tbl, offs)
        return h

    def index_arrays(self, idx, k):
        n_off, n_dist = C.c_uint64(), C.c_uint64()
        self._check(self.L.nb_index_stats(self.h, idx, C.byref(n_off), C.byref(n_dist)))
        starts = np.empty((1 << (2 * k)) + 1, np.uint32)
        offs = np.empty(max(n_off.value, 1), np.uint32)
        self._check(self.L.nb_index_download(self.h, idx, starts.ctypes.data_as(C.c_void_p),
                                             offs.ctypes.data_as(C.c_void_p)))
        return starts, offs[:n_off.value], n_dist.value

    def map_volume(self, idx, dref, ref_start, dreads, read_start, lo, hi, opt):
        out = C.c_void_p()
        n = C.c_uint64()
        self._check(self.L.nb_map_volume(self.h, idx, dref, C.c_int(ref_start), dreads, C.c_int(read_start),
                                         C.c_uint64(lo), C.c_uint64(hi), C.byref(opt), C.byref(out), C.byref(n)))
        if n.value:
            buf = (C.c_char * (CANDIDATE.itemsize * n.value)).from_address(out.value)
            arr = np.frombuffer(buf, CANDIDATE).copy()
        else:
            arr = np.zeros(0, CANDIDATE)
        self.L.nb_free(out)
        return arr

    def map_volume_packed(self, idx, dref, ref_start, dreads, read_start, lo, hi, opt):
        """map_volume with the records packed by the library (nb_pack_candidates) straight from its result buffer:
        (n, 7) uint32 PackedGappedCandidate records, what oc2pmov writes in CAN mode."""
        out = C.c_void_p()
        n = C.c_uint64()
        self._check(self.L.nb_map_volume(self.h, idx, dref, C.c_int(ref_start), dreads, C.c_int(read_start),
                                         C.c_uint64(lo), C.c_uint64(hi), C.byref(opt), C.byref(out), C.byref(n)))
        packed = np.empty((n.value, 7), np.uint32)
        if n.value:
            self.L.nb_pack_candidates(out, C.c_uint64(n.value), packed.ctypes.data_as(C.c_void_p))
        self.L.nb_free(out)
        return packed

    def consensus_job(self, dreads, recs, opt, shard_count=1, shard_index=0):
        """recs: (n, 7) uint32 partition records.  Returns (CNS_RECORD array, bytes of all sequences)."""
        recs = np.ascontiguousarray(recs, np.uint32)
        out = C.c_void_p(); n = C.c_uint64(); seq = C.c_void_p(); nb = C.c_uint64()
        self._check(self.L.nb_consensus_job_sharded(self.h, dreads, recs.ctypes.data_as(C.c_void_p),
                                                    C.c_uint64(recs.shape[0]), C.byref(opt), C.c_int(shard_count),
                                                    C.c_int(shard_index), C.byref(out), C.byref(n), C.byref(seq),
                                                    C.byref(nb)))
        if n.value:
            r = np.frombuffer((C.c_char * (CNS_RECORD.itemsize * n.value)).from_address(out.value), CNS_RECORD).copy()
        else:
            r = np.zeros(0, CNS_RECORD)
        self.L.nb_free(out)
        # the sequence buffer (hundreds of MB) is handed over without a copy; it is freed with the returned object
        return r, HostBuffer(self.L, seq.value, nb.value)

    def timings(self):
        t = Timings()
        self.L.nb_get_timings(self.h, C.byref(t))
        return t

    def map_counters(self):
        m = MapCounters()
        self.L.nb_get_map_counters(self.h, C.byref(m))
        return m

    def cns_counters(self):
        m = CnsCounters()
        self.L.nb_get_cns_counters(self.h, C.byref(m))
        return m

    def launches(self):
        return int(self.L.nb_launch_count(self.h))


def default_map_options(**kw):
    o = MapOptions()
    lib().nb_map_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_cns_options(**kw):
    o = CnsOptions()
    lib().nb_cns_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def with_swapped_roles(packed_parts):
    """(n, 7) uint32 records (a list of arrays, concatenated) followed by their role-swapped copies
    (nb_swap_candidate_roles): the 2 n records a correction job takes (pcan.c:86-91)."""
    n = int(sum(p.shape[0] for p in packed_parts))
    out = np.empty((2 * n, 7), np.uint32)
    pos = 0
    for p in packed_parts:
        out[pos:pos + p.shape[0]] = p
        pos += p.shape[0]
    if n:
        L = lib()
        L.nb_swap_candidate_roles(out.ctypes.data_as(C.c_void_p), C.c_uint64(n), C.c_void_p(out.ctypes.data + 28 * n))
    return out


def pack_candidates(cands):
    """CANDIDATE array -> (n, 7) uint32 PackedGappedCandidate records (gapped_candidate.c:13-28)."""
    out = np.zeros((cands.size, 7), np.uint32)
    w = np.minimum(cands["score"], 1000000).astype(np.uint32)
    w |= (cands["sdir"] == 1).astype(np.uint32) << 31
    w |= (cands["qdir"] == 1).astype(np.uint32) << 30
    w |= (cands["qoff"] == cands["qbeg"]).astype(np.uint32) << 29
    out[:, 0] = w
    for i, f in enumerate(("sid", "sbeg", "send", "qid", "qbeg", "qend")):
        out[:, i + 1] = cands[f].astype(np.uint32)
    return out
