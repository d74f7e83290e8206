"""Two correction iterations through the files necat.pl moves between the executables (necat.pl:160-365):

    FASTA -> oc2mkdb -> oc2pmov -> oc2pcan -> oc2cns (-r 0 -f 1) -> oc2ReorderCnsReads
          -> oc2mkdb (corrected + uncorrected reads) -> oc2pmov -> oc2pcan -> oc2cns (-r 1 -f 0) -> oc2ReorderCnsReads

CPU (needs oracle/_ref): this repo's ingest / reorder hosts dropped into the reference's flow change nothing, and the
oracle follows the reference through both iterations (names that already start with ontsa_id, -r 1 on corrected reads).
GPU: all five executables of this repo against the oracle's flow."""
import os
import shutil

import numpy as np
import pytest

import util

PM1 = "-k 13 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 2 -j 0 -u 1 -i 1"
PM2 = "-k 13 -z 20 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 2 -j 0 -u 1 -i 1"
CNS1 = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 1"
CNS2 = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"
needs_ref = pytest.mark.skipif(util.ref_bin("oc2pmov") is None, reason="oracle/_ref not built (needs /root/reference)")


def ours(name):
    return os.path.join(util.BIN_DIR, name)


def write_reads(path, seed=71):
    import gen_reads
    _, reads = gen_reads.generate(20_000, 16, seed, err=0.12, long_indel=0.3)
    A = np.frombuffer(b"ACGT", np.uint8)
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">read%d/%d\n" % (i, r.size) + A[r].tobytes() + b"\n")
    return len(reads)


def iteration(d, step, lists, pm, cns, compute, mkdb, reorder):
    """One correction iteration in directory d.  compute: 'ref' | 'oracle' | 'gpu' for oc2pmov/oc2pcan/oc2cns."""
    wrk = "wrk%d" % step
    util.run([mkdb, wrk] + lists, cwd=d)
    cand = "cand%d.bin" % step
    cns_tmp, raw_tmp = "cns%d.tmp.fa" % step, "raw%d.tmp.fa" % step
    if compute == "oracle":
        util.run([util.ORACLE_CLI, "pmov"] + pm.split() + [wrk, "0", cand], cwd=d)
        util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "1", wrk, cand, cns_tmp, raw_tmp], cwd=d)
    else:
        b = util.ref_bin if compute == "ref" else ours
        util.run([b("oc2pmov")] + pm.split() + [wrk, "0", cand], cwd=d)
        util.run([b("oc2pcan"), "-t", "1", wrk, cand], cwd=d)
        util.run([b("oc2cns")] + cns.split() + ["-t", "2", wrk, cand, cns_tmp, raw_tmp], cwd=d)
    cns_out, raw_out = "cns%d.fa" % step, "raw%d.fa" % step
    util.run([reorder, cns_tmp, raw_tmp, cns_out, raw_out], cwd=d)
    return cns_out, raw_out, cns_tmp, raw_tmp


def two_iterations(tmp_path, tag, compute, mkdb, reorder):
    d = str(tmp_path / tag)
    os.makedirs(d)
    shutil.copy(str(tmp_path / "reads.fa"), os.path.join(d, "reads.fa"))
    open(os.path.join(d, "list0"), "w").write("reads.fa\n")
    c1, r1, c1_tmp, r1_tmp = iteration(d, 1, ["list0"], PM1, CNS1, compute, mkdb, reorder)
    open(os.path.join(d, "list1"), "w").write("%s\n%s\n" % (c1, r1))
    c2, r2, _, _ = iteration(d, 2, ["list1"], PM2, CNS2, compute, mkdb, reorder)
    return {k: os.path.join(d, v) for k, v in dict(c1=c1, r1=r1, c2=c2, r2=r2, c1_tmp=c1_tmp, r1_tmp=r1_tmp, vol2="wrk2/vol0").items()}


def ids_of(path):
    import re
    return [int(re.search(rb"\d+", h).group()) for h, _ in __import__("necat_b200.formats", fromlist=["x"]).read_fasta_records(path)]


@needs_ref
def test_two_iterations_cpu(built, tmp_path):
    n = write_reads(str(tmp_path / "reads.fa"))
    ref = two_iterations(tmp_path, "ref", "ref", util.ref_bin("oc2mkdb"), util.ref_bin("oc2ReorderCnsReads"))
    # (1) this repo's oc2mkdb + oc2ReorderCnsReads inside the reference's flow: same files all the way
    mix = two_iterations(tmp_path, "mix", "ref", ours("oc2mkdb"), ours("oc2ReorderCnsReads"))
    # (2) the oracle through the same flow
    ora = two_iterations(tmp_path, "ora", "oracle", ours("oc2mkdb"), ours("oc2ReorderCnsReads"))
    for other in (mix, ora):
        # iteration 1 (-f 1): one record per read, so the reordered files are identical byte for byte, and so is the
        # volume iteration 2 starts from (apart from the reference's uninitialised padding bytes)
        for k in ("c1", "r1"):
            assert open(ref[k], "rb").read() == open(other[k], "rb").read(), k
        from test_host_tools import mask_pads
        assert mask_pads(open(ref["vol2"], "rb").read()) == mask_pads(open(other["vol2"], "rb").read())
        for k in ("c2", "r2"):
            assert util.fasta_canon(ref[k]) == util.fasta_canon(other[k]), k
            assert ids_of(ref[k]) == ids_of(other[k])
    recs = util.fasta_canon(ref["c2"])
    assert len(util.fasta_canon(ref["c1"])) > n // 2 and len(recs) > n // 2
    assert recs[0][0].startswith(b">ontsa_id_") and recs[0][0].count(b"_(") == 2      # the name carried over + new coordinates


@pytest.mark.gpu
def test_two_iterations_gpu(built, tmp_path):
    write_reads(str(tmp_path / "reads.fa"))
    ora = two_iterations(tmp_path, "ora", "oracle", ours("oc2mkdb"), ours("oc2ReorderCnsReads"))
    gpu = two_iterations(tmp_path, "gpu", "gpu", ours("oc2mkdb"), ours("oc2ReorderCnsReads"))
    for k in ("c1", "r1"):
        assert open(ora[k], "rb").read() == open(gpu[k], "rb").read(), k
    for k in ("c2", "r2"):
        assert util.fasta_canon(ora[k]) == util.fasta_canon(gpu[k]), k
    # oc2cns writes its records in read id order, so the reorder step only streams
    for k in ("c1_tmp", "r1_tmp"):
        ids = ids_of(gpu[k])
        assert ids == sorted(ids), k
