"""PackedDB volume files and work-directory layout (host-side mirror of the
reference's formats; numpy only).

* volume file: common/packed_db.c:291-348 (31-byte magic, nseq, db_size,
  32-byte SequenceInfo records, header block, 2-bit bases, first base in the top
  two bits of byte 0)
* work dir: makedb/main.c:29-39,104-117 and common/makedb_aux.c:34-42
"""
import os

import numpy as np

PAC_MAGIC = b"ontcns_pac_header_hofuwhogfuewo"
SEQINFO = np.dtype([("offset", "<u8"), ("size", "<u8"), ("hdr_offset", "<u8"),
                    ("platform", "<i4"), ("pad", "<i4")])


def pack_bases(codes):
    """uint8 codes (0..3) -> PackedDB bytes."""
    n = codes.size
    pad = (-n) % 4
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8)


def unpack_bases(pac, n):
    b = np.asarray(pac, np.uint8)
    out = np.empty((b.size, 4), np.uint8)
    out[:, 0] = b >> 6
    out[:, 1] = (b >> 4) & 3
    out[:, 2] = (b >> 2) & 3
    out[:, 3] = b & 3
    return out.reshape(-1)[:n]


class Volume:
    def __init__(self, pac, db_size, info, hdr):
        self.pac, self.db_size, self.info, self.hdr = pac, int(db_size), info, hdr

    @property
    def nseq(self):
        return self.info.size

    def name(self, i):
        o = int(self.info["hdr_offset"][i])
        return self.hdr[o:self.hdr.index(b"\0", o)].decode()

    def sequence(self, i, strand=0):
        """codes of read i; strand 1 = reverse complement (packed_db.c:255-275)."""
        o, s = int(self.info["offset"][i]), int(self.info["size"][i])
        b0 = o >> 2
        codes = unpack_bases(self.pac[b0:((o + s + 3) >> 2)], (o & 3) + s)[(o & 3):]
        return codes if strand == 0 else (3 - codes[::-1]).astype(np.uint8)


def read_volume(path):
    with open(path, "rb") as f:
        if f.read(len(PAC_MAGIC)) != PAC_MAGIC:
            raise ValueError(f"{path}: not a PackedDB volume")
        nseq, db_size = np.frombuffer(f.read(16), "<u8")
        info = np.frombuffer(f.read(SEQINFO.itemsize * int(nseq)), SEQINFO).copy()
        hdr_size = int(np.frombuffer(f.read(8), "<u8")[0])
        hdr = f.read(hdr_size)
        pac = np.frombuffer(f.read((int(db_size) + 3) >> 2), np.uint8).copy()
    return Volume(pac, db_size, info, hdr)


def write_volume(path, reads, names):
    info = np.zeros(len(reads), SEQINFO)
    hdr = bytearray()
    off = 0
    for i, (r, nm) in enumerate(zip(reads, names)):
        info[i] = (off, r.size, len(hdr), 0, 0)
        hdr += nm.encode() + b"\0"
        off += r.size
    allc = np.concatenate(reads) if reads else np.zeros(0, np.uint8)
    with open(path, "wb") as f:
        f.write(PAC_MAGIC)
        f.write(np.array([len(reads), off], "<u8").tobytes())
        f.write(info.tobytes())
        f.write(np.array([len(hdr)], "<u8").tobytes())
        f.write(bytes(hdr))
        f.write(pack_bases(allc).tobytes())


def write_workdir(wrk_dir, reads, names=None, vol_size=2_000_000_000):
    """What oc2mkdb produces: vol<N>, volume_names.txt, reads_info.txt.  A volume is
    closed once it holds >= vol_size bases (makedb/main.c:29)."""
    os.makedirs(wrk_dir, exist_ok=True)
    if names is None:
        names = [f"r{i}" for i in range(len(reads))]
    vols, cur, cur_names, size, start = [], [], [], 0, 0
    lines = []

    def flush():
        nonlocal cur, cur_names, size, start
        vid = len(vols)
        p = os.path.join(wrk_dir, f"vol{vid}")
        write_volume(p, cur, cur_names)
        lines.append(f"{p}\t{start}\t{len(cur)}\n")
        vols.append(p)
        start += len(cur)
        cur, cur_names, size = [], [], 0

    for r, nm in zip(reads, names):
        cur.append(r)
        cur_names.append(nm)
        size += r.size
        if size >= vol_size:
            flush()
    if cur:
        flush()
    with open(os.path.join(wrk_dir, "volume_names.txt"), "w") as f:
        f.writelines(lines)
    with open(os.path.join(wrk_dir, "reads_info.txt"), "w") as f:
        f.write(f"{len(vols)}\t{len(reads)}\n")
    return vols


def read_candidates(path):
    """28-byte PackedGappedCandidate records (gapped_candidate.h:59-80) as (n, 7) uint32."""
    return np.fromfile(path, dtype="<u4").reshape(-1, 7)


def canonical_records(a):
    """Canonical order for multiset comparison."""
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def read_fasta_records(path):
    """[(header, sequence)] of a 2-line-per-record FASTA as written by oc2cns."""
    recs = []
    with open(path, "rb") as f:
        hdr = None
        for line in f:
            line = line.rstrip(b"\n")
            if line.startswith(b">"):
                hdr = line
            else:
                recs.append((hdr, line))
    return recs


def swap_roles(rec):
    """change_pcan_roles (common/gapped_candidate.c:54-69) on (n, 7) uint32 records."""
    out = rec.copy()
    w = rec[:, 0]
    r = w & np.uint32((1 << 30) - 1)
    r |= ((w >> 31) & 1) << 30
    r |= ((w >> 30) & 1) << 31
    out[:, 0] = r
    out[:, 1:4] = rec[:, 4:7]
    out[:, 4:7] = rec[:, 1:4]
    return out


def normalise_sdir(rec, sizes):
    """normalise_pcan_sdir (gapped_candidate.c:72-93): make the subject forward."""
    out = rec.copy()
    rev = (rec[:, 0] >> 31) == 1
    w = rec[rev, 0]
    r = w & np.uint32((1 << 29) - 1)
    r |= (1 - ((w >> 30) & 1)) << 30
    r |= (1 - ((w >> 29) & 1)) << 29
    out[rev, 0] = r
    qs = sizes[rec[rev, 4]].astype(np.uint32)
    ss = sizes[rec[rev, 1]].astype(np.uint32)
    out[rev, 5] = qs - rec[rev, 6]
    out[rev, 6] = qs - rec[rev, 5]
    out[rev, 2] = ss - rec[rev, 3]
    out[rev, 3] = ss - rec[rev, 2]
    return out


def partition_records(rec, sizes, lo=0, hi=None):
    """What oc2pcan + oc2cns' loader produce for templates [lo, hi): every record under
    its subject and, roles swapped, under its query; subject made forward."""
    if hi is None:
        hi = sizes.size
    a = rec[(rec[:, 1] >= lo) & (rec[:, 1] < hi)]
    b = swap_roles(rec[(rec[:, 4] >= lo) & (rec[:, 4] < hi)])
    allr = np.concatenate([a, b])
    allr = allr[np.argsort(allr[:, 1], kind="stable")]
    return normalise_sdir(allr, sizes)


def cns_fasta_records(recs, seq, names):
    """(cns, raw) lists of (header, sequence) as oc2cns writes them (common/cns_seq.h:23-41) from the records of
    nb_consensus_job: recs = CNS_RECORD array, seq = uint8 array of all sequences, names[id] = read name (bytes)."""
    cns, raw = [], []
    for r in recs:
        name = names[int(r["id"])]
        head = b">" + name + b"_" if name.startswith(b"ontsa_id") else b">ontsa_id_%d_" % int(r["id"])
        head += b"(%d_%d_%d_%d_%d_%d_%f)" % (int(r["left"]), int(r["right"]), int(r["seq_len"]), int(r["org_size"]),
                                               int(r["num_can"]), int(r["num_ovlps"]), float(r["ident_cutoff"]))
        s = bytes(seq[int(r["seq_offset"]):int(r["seq_offset"]) + int(r["seq_len"])])
        (raw if r["is_raw"] else cns).append((head, s))
    return cns, raw
