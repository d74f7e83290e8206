"""The drop-in claim itself: the reference's UNMODIFIED pipeline driver, `necat.pl correct` (src/pipeline/necat.pl: two
correction iterations -- fsa_rd_tools longest, oc2mkdb, oc2pmov per volume, cat, oc2pcan, oc2cns per partition, cat,
oc2ReorderCnsReads, pigz; iteration 2 with the FAST overlap options and -r 1), run twice on the same reads: once with the
reference's binaries in its bin directory, once with this repo's executables in their place -- the five of the hot path and
the drop-ins of the two helpers around it (fsa_rd_tools longest / n50, pigz).  necat.pl calls its
tools as "$FindBin::RealBin/<tool>" (necat.pl:91), i.e. from the directory necat.pl itself lives in -- so the executables
are put (linked) THERE, which is what INTEGRATION.md tells a user to do.  cns.fasta / raw.fasta of both iterations and
cns_final.fasta must be byte-identical.

CPU: this repo's oc2pmov / oc2pcan / oc2cns come from the emulated build (tests/emu: the CUDA sources compiled for the host),
k = 11 on a 12 kb genome.  GPU (-m gpu): the real executables, necat.pl's default options (k = 15) on a 300 kb genome.
Needs oracle/_ref/pipeline (necat.pl + fsa_rd_tools built by `make -C oracle ref`, which needs /root/reference; it travels to
the GPU box with oracle/_ref) and perl."""
import gzip
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import util

PIPE = os.path.join(util.REF_DIR, "pipeline")
TOOLS = ("oc2mkdb", "oc2pmov", "oc2pcan", "oc2cns", "oc2ReorderCnsReads")
EMU_BIN = os.path.join(util.ROOT, "tests", "emu", "_build", "cudaemu", "bin")

CFG = """PROJECT=prj
ONT_READ_LIST={list}
GENOME_SIZE={genome}
THREADS=4
MIN_READ_LENGTH={min_len}
PREP_OUTPUT_COVERAGE=40
{ovlp}{cns}NUM_ITER=2
CNS_OUTPUT_COVERAGE=30
CLEANUP=0
USE_GRID=false
GRID_NODE=0
SMALL_MEMORY={small}
"""


# what `necat.pl config` writes into a fresh configuration file (necat.pl:22-34); a file without these lines makes necat.pl
# fall back to "-j 1 -u 0" (M4 text) for the overlap step, which oc2pcan cannot read -- with either set of binaries
DEFAULT_OVLP = ("OVLP_FAST_OPTIONS=-n 500 -z 20 -b 2000 -e 0.5 -j 0 -u 1 -a 1000\n"
                "OVLP_SENSITIVE_OPTIONS=-n 500 -z 10 -e 0.5 -j 0 -u 1 -a 1000\n")
DEFAULT_CNS = ("CNS_FAST_OPTIONS=-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0\n"
               "CNS_SENSITIVE_OPTIONS=-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0\n")


def _need_pipeline():
    if not (os.path.exists(os.path.join(PIPE, "necat.pl")) and util.ref_bin("oc2cns")):
        pytest.skip("oracle/_ref/pipeline is not built (make -C oracle ref, needs /root/reference)")
    if not shutil.which("perl"):
        pytest.skip("perl is not installed")


def _bin_dir(path, tool_dirs, own_helpers=False):
    """A NECAT bin directory: the pipeline scripts of oracle/_ref/pipeline, the five tools linked from tool_dirs[tool],
    fsa_rd_tools and pigz either the reference's (pigz: the gzip stand-in of oracle/Makefile.ref) or, own_helpers, this
    repo's drop-ins of the two."""
    os.makedirs(path)
    for f in os.listdir(PIPE):
        src = os.path.join(PIPE, f)
        if own_helpers and f in ("fsa_rd_tools", "pigz"):
            os.symlink(os.path.join(util.BIN_DIR, f), os.path.join(path, f))
        else:
            (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, os.path.join(path, f))
    for t in TOOLS:
        os.symlink(os.path.join(tool_dirs[t], t), os.path.join(path, t))
    return path


def _correct(bin_dir, run_dir, reads, genome, min_len, ovlp=DEFAULT_OVLP, cns=DEFAULT_CNS, small=0, env=None):
    import gen_reads  # noqa: F401  (tools/ on the path through util's conftest)
    os.makedirs(run_dir)
    A = np.frombuffer(b"ACGT", np.uint8)
    with open(os.path.join(run_dir, "reads.fasta"), "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i + A[r].tobytes() + b"\n")
    lst = os.path.join(run_dir, "list.txt")
    open(lst, "w").write(os.path.join(run_dir, "reads.fasta") + "\n")
    open(os.path.join(run_dir, "cfg.txt"), "w").write(CFG.format(list=lst, genome=genome, min_len=min_len, ovlp=ovlp, cns=cns, small=small))
    p = subprocess.run([os.path.join(bin_dir, "necat.pl"), "correct", "cfg.txt"], cwd=run_dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=dict(os.environ, **(env or {})))
    log = p.stdout.decode(errors="replace")
    assert p.returncode == 0 and "Failed to run script" not in log and "[Error]" not in log, log[-3000:]
    out = {}
    for f in ("cns_iter1/cns.fasta.gz", "cns_iter1/raw.fasta.gz", "cns_iter2/cns.fasta.gz", "cns_iter2/raw.fasta.gz", "cns_final.fasta.gz"):
        out[f] = gzip.open(os.path.join(run_dir, "prj", "1-consensus", f)).read()
    return out


def _compare(tmp_path, ours, reads, genome, min_len, ovlp=DEFAULT_OVLP, cns=DEFAULT_CNS, small=0, env=None):
    ref = {t: util.REF_DIR for t in TOOLS}
    a = _correct(_bin_dir(str(tmp_path / "bin_ref"), ref), str(tmp_path / "run_ref"), reads, genome, min_len, ovlp, cns, small)
    b = _correct(_bin_dir(str(tmp_path / "bin_ours"), ours, own_helpers=True), str(tmp_path / "run_ours"), reads, genome, min_len, ovlp, cns, small, env)
    assert a["cns_iter1/cns.fasta.gz"].count(b">") >= 20 and a["cns_final.fasta.gz"].count(b">") >= 20
    for f in a:
        assert a[f] == b[f], f"{f} differs: {len(a[f])} vs {len(b[f])} bytes"


@pytest.mark.parametrize("small,helper", [(0, 0), (1, 1)])
def test_necat_pl_correct_with_the_emulated_executables(built, tmp_path, small, helper):
    """small = 1: SMALL_MEMORY=1 in the configuration, which necat.pl hands to oc2cns as -s 1 (necat.pl:300).
    helper = 1: NECAT_B200_HELPER=1 in necat.pl's environment -- every oc2pmov / oc2cns call of the run is served by one
    resident helper process."""
    import gen_reads
    _need_pipeline()
    subprocess.check_call(["make", "-s", "-C", os.path.join(util.ROOT, "tests", "emu")])
    util.ensure_host_tools()
    _, reads = gen_reads.generate(12000, 15, 5, err=0.10, min_len=1500, max_len=6000, mean_log=7.8, sd_log=0.4)
    ours = {t: (EMU_BIN if t in ("oc2pmov", "oc2pcan", "oc2cns") else util.BIN_DIR) for t in TOOLS}
    sock_dir = tempfile.mkdtemp(prefix="nbh", dir="/tmp")       # (a UNIX socket path holds ~100 characters)
    _compare(tmp_path, ours, reads, 12000, 1000,
             ovlp="OVLP_FAST_OPTIONS=-k 11 -q 100 -n 500 -z 20 -b 2000 -e 0.5 -j 0 -u 1 -a 1000\n"
                  "OVLP_SENSITIVE_OPTIONS=-k 11 -q 100 -n 500 -z 10 -e 0.5 -j 0 -u 1 -a 1000\n",
             cns="CNS_FAST_OPTIONS=-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0\n"
                 "CNS_SENSITIVE_OPTIONS=-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0\n", small=small,
             env={"NECAT_B200_HELPER": "1", "NECAT_B200_HELPER_IDLE": "3", "NECAT_B200_HELPER_DIR": sock_dir} if helper else None)
    shutil.rmtree(sock_dir, ignore_errors=True)


@pytest.mark.gpu
def test_necat_pl_correct_on_the_gpu(built, tmp_path):
    """necat.pl's own default options (k = 15, SENSITIVE then FAST overlap options, -r 1 in iteration 2)."""
    import gen_reads
    _need_pipeline()
    _, reads = gen_reads.generate(300_000, 25, 77)
    _compare(tmp_path, {t: util.BIN_DIR for t in TOOLS}, reads, 300_000, 3000)
