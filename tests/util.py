"""Shared helpers of the test-suite: synthetic data sets, oracle / reference runners,
canonical comparators."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_CLI = os.path.join(ORACLE_DIR, "_build", "oracle_cli")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
BIN_DIR = os.path.join(ROOT, "necat_b200", "bin")

PM_SENSITIVE = "-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"
PM_FAST = "-k 15 -z 20 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"
CNS_SENSITIVE = "-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


def ensure_oracle():
    if not (os.path.exists(ORACLE_CLI) and os.path.exists(ORACLE_LIB)):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def ensure_host_tools():
    """The host-only executables (gcc + zlib, no nvcc needed) are built on demand; the CUDA library and the three
    executables over it come from `make` / __graft_entry__.build()."""
    tools = [os.path.join(BIN_DIR, t) for t in ("oc2mkdb", "oc2ReorderCnsReads", "fsa_rd_tools", "pigz")]
    if not all(os.path.exists(t) for t in tools):
        subprocess.check_call(["make", "-s", "-C", ROOT] + [os.path.relpath(t, ROOT) for t in tools])


def ref_bin(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def make_workdir(path, genome=100_000, depth=20, seed=7, vol_size=2_000_000_000, **gen_kw):
    """Synthetic reads -> PackedDB work dir.  Returns (wrk_dir, reads)."""
    import gen_reads
    from necat_b200 import formats
    _, reads = gen_reads.generate(genome, depth, seed, **gen_kw)
    wrk = os.path.join(str(path), "vol")
    formats.write_workdir(wrk, reads, vol_size=vol_size)
    return wrk, reads


def run(cmd, **kw):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)
    if p.returncode:
        raise RuntimeError("%s exited with %d\nstdout: %s\nstderr: %s" % (" ".join(map(str, cmd)), p.returncode,
                           p.stdout[-2000:].decode(errors="replace"), p.stderr[-2000:].decode(errors="replace")))
    return p


def run_pmov(binary, opts, wrk, vid, out, prefix=()):
    run(list(prefix) + ([binary, "pmov"] if binary == ORACLE_CLI else [binary]) + opts.split() + [wrk, str(vid), out])
    from necat_b200 import formats
    return formats.read_candidates(out)


def canon(a):
    from necat_b200 import formats
    return formats.canonical_records(a)


def same_multiset(a, b):
    return a.shape == b.shape and np.array_equal(canon(a), canon(b))


def fasta_canon(path):
    from necat_b200 import formats
    return sorted(formats.read_fasta_records(path))


GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_workdir(tmp):
    """Copies the committed golden volume into tmp/vol (volume_names.txt holds the relative path vol/vol0);
    commands must run with cwd=tmp."""
    import shutil
    d = os.path.join(str(tmp), "vol")
    os.makedirs(d, exist_ok=True)
    for f in ("vol0", "volume_names.txt", "reads_info.txt"):
        shutil.copy(os.path.join(GOLDEN, f), os.path.join(d, f))
    return "vol"


def golden_workdir_r1(tmp):
    """The long-indel golden set (tests/golden/r1): copies its volume into tmp/vol; run commands with cwd=tmp."""
    import shutil
    d = os.path.join(str(tmp), "vol")
    os.makedirs(d, exist_ok=True)
    for f in ("vol0", "volume_names.txt", "reads_info.txt"):
        shutil.copy(os.path.join(GOLDEN, "r1", f), os.path.join(d, f))
    return "vol"


def golden_fasta(name):
    import gzip
    data = gzip.open(os.path.join(GOLDEN, name + ".fasta.gz")).read().split(b"\n")
    return sorted((data[i], data[i + 1]) for i in range(0, len(data) - 1, 2))


def smoke_cns(td, wrk, cand):
    """Used by __graft_entry__.smoke(): oc2pcan + oc2cns on the GPU against the oracle."""
    import shutil
    c = os.path.join(td, "cand_cns.bin")
    shutil.copy(cand, c)
    run([os.path.join(BIN_DIR, "oc2pcan"), wrk, c])
    run([os.path.join(BIN_DIR, "oc2cns")] + CNS_SENSITIVE.split() + [wrk, c, os.path.join(td, "g_cns.fa"), os.path.join(td, "g_raw.fa")])
    run([ORACLE_CLI, "cns"] + CNS_SENSITIVE.split() + ["-M", "0", wrk, cand, os.path.join(td, "o_cns.fa"), os.path.join(td, "o_raw.fa")])
    assert fasta_canon(os.path.join(td, "g_cns.fa")) == fasta_canon(os.path.join(td, "o_cns.fa")), "cns.fasta mismatch vs oracle"
    assert fasta_canon(os.path.join(td, "g_raw.fa")) == fasta_canon(os.path.join(td, "o_raw.fa")), "raw.fasta mismatch vs oracle"
