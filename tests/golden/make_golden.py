#!/usr/bin/env python3
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference binaries
(oracle/_ref, built from /root/reference by oracle/Makefile.ref) on a small seeded read set.

    python tests/golden/make_golden.py          # needs oracle/_ref (i.e. /root/reference at build time)

Files written:
    vol0, volume_names.txt, reads_info.txt   the PackedDB work dir (reads are committed, not regenerated)
    cand_k13.bin                             oc2pmov -k 13 -z 5 ... -j 0 -u 1   (28-byte records)
    cand_k15.bin                             oc2pmov SENSITIVE options (necat.pl:20,32)
    cns_f0.fasta.gz / raw_f0.fasta.gz        oc2pcan + oc2cns CNS_SENSITIVE -r 0 -f 0 on cand_k15.bin
    cns_f1.fasta.gz / raw_f1.fasta.gz        same with -f 1
    r1/                                      a second, smaller set whose reads carry 300-700 base indels: vol0 + info files,
                                             cand_k13.bin (oc2pmov), cns_r0 / raw_r0 / cns_r1 / raw_r1 .fasta.gz (oc2cns -r 0 / -r 1:
                                             the DALIGNER + global-edlib rescue of correction iterations >= 2 changes most records)
    m4_k13.txt.gz                            oc2pmov -j 1 -u 0 -i 1 (M4 text lines, sorted)
    manifest.json                            options, counts, sha256 of the canonical (sorted) outputs
All outputs are stored in canonical order (records sorted) because the reference's own order varies.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import gen_reads  # noqa: E402
import util  # noqa: E402
from necat_b200 import formats  # noqa: E402

PM13 = "-k 13 -z 5 -q 50 -b 1000 -s 3 -n 500 -a 500 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"
CNS = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0"


def canon_fasta_bytes(path):
    recs = sorted(formats.read_fasta_records(path))
    return b"".join(h + b"\n" + s + b"\n" for h, s in recs)


def main():
    ref = util.REF_DIR
    assert os.path.exists(os.path.join(ref, "oc2cns")), "build oracle/_ref first (make -C oracle ref)"
    _, reads = gen_reads.generate(40_000, 18, 4242, err=0.12, repeat_len=1500, repeat_copies=4, n_repeats=1)
    td = tempfile.mkdtemp()
    man = {"generator": "gen_reads.generate(40000, 18, 4242, err=0.12, repeat_len=1500, repeat_copies=4, n_repeats=1)",
           "reads": len(reads), "bases": int(sum(r.size for r in reads))}
    try:
        os.chdir(td)
        formats.write_workdir("vol", reads)                       # relative volume paths: fixtures stay relocatable
        for name, opts in (("cand_k13.bin", PM13), ("cand_k15.bin", util.PM_SENSITIVE)):
            subprocess.check_call([os.path.join(ref, "oc2pmov")] + opts.split() + ["vol", "0", name], stdout=subprocess.DEVNULL)
            rec = formats.canonical_records(formats.read_candidates(name))
            rec.astype("<u4").tofile(os.path.join(HERE, name))
            man[name] = {"options": opts, "records": int(rec.shape[0]), "sha256": hashlib.sha256(rec.tobytes()).hexdigest()}
        # ALN mode (M4 text, names as ids): sorted lines
        m4opts = "-k 13 -z 10 -q 500 -b 2000 -s 3 -n 100 -a 400 -d 0.25 -e 0.5 -m 500 -t 2 -j 1 -u 0 -i 1"
        subprocess.check_call([os.path.join(ref, "oc2pmov")] + m4opts.split() + ["vol", "0", "m4.txt"], stdout=subprocess.DEVNULL)
        data = b"".join(sorted(open("m4.txt", "rb").read().splitlines(keepends=True)))
        with gzip.GzipFile(os.path.join(HERE, "m4_k13.txt.gz"), "wb", mtime=0) as g:
            g.write(data)
        man["m4_k13"] = {"options": m4opts, "records": data.count(b"\n"), "sha256": hashlib.sha256(data).hexdigest()}
        for f in ("0", "1"):
            shutil.copy("cand_k15.bin", "c.bin")
            subprocess.check_call([os.path.join(ref, "oc2pcan"), "-t", "1", "vol", "c.bin"], stdout=subprocess.DEVNULL)
            subprocess.check_call([os.path.join(ref, "oc2cns")] + (CNS + " -f " + f).split() + ["-t", "2", "vol", "c.bin", "cns.fa", "raw.fa"],
                                  stdout=subprocess.DEVNULL)
            for kind in ("cns", "raw"):
                data = canon_fasta_bytes(kind + ".fa")
                with gzip.GzipFile(os.path.join(HERE, f"{kind}_f{f}.fasta.gz"), "wb", mtime=0) as g:
                    g.write(data)
                man[f"{kind}_f{f}"] = {"options": CNS + " -f " + f, "records": data.count(b">"), "sha256": hashlib.sha256(data).hexdigest()}
        for fn in ("vol0", "volume_names.txt", "reads_info.txt"):
            shutil.copy(os.path.join("vol", fn), os.path.join(HERE, fn))
        # ---- the long-indel set for -r 1 (consensus_aux.c:169-201)
        _, reads2 = gen_reads.generate(20_000, 18, 777, err=0.10, long_indel=0.5)
        man["r1"] = {"generator": "gen_reads.generate(20000, 18, 777, err=0.10, long_indel=0.5)", "reads": len(reads2),
                     "bases": int(sum(r.size for r in reads2))}
        os.makedirs("r1w")
        os.chdir("r1w")
        formats.write_workdir("vol", reads2)
        out2 = os.path.join(HERE, "r1")
        os.makedirs(out2, exist_ok=True)
        pm = "-k 13 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"
        subprocess.check_call([os.path.join(ref, "oc2pmov")] + pm.split() + ["vol", "0", "cand.bin"], stdout=subprocess.DEVNULL)
        rec = formats.canonical_records(formats.read_candidates("cand.bin"))
        rec.astype("<u4").tofile(os.path.join(out2, "cand_k13.bin"))
        man["r1"]["cand_k13.bin"] = {"options": pm, "records": int(rec.shape[0]), "sha256": hashlib.sha256(rec.tobytes()).hexdigest()}
        for r in ("0", "1"):
            shutil.copy("cand.bin", "c.bin")
            subprocess.check_call([os.path.join(ref, "oc2pcan"), "-t", "1", "vol", "c.bin"], stdout=subprocess.DEVNULL)
            opts = CNS.replace("-r 0", "-r " + r) + " -f 0"
            subprocess.check_call([os.path.join(ref, "oc2cns")] + opts.split() + ["-t", "2", "vol", "c.bin", "cns.fa", "raw.fa"], stdout=subprocess.DEVNULL)
            for kind in ("cns", "raw"):
                data = canon_fasta_bytes(kind + ".fa")
                with gzip.GzipFile(os.path.join(out2, f"{kind}_r{r}.fasta.gz"), "wb", mtime=0) as g:
                    g.write(data)
                man["r1"][f"{kind}_r{r}"] = {"options": opts, "records": data.count(b">"), "sha256": hashlib.sha256(data).hexdigest()}
        for fn in ("vol0", "volume_names.txt", "reads_info.txt"):
            shutil.copy(os.path.join("vol", fn), os.path.join(out2, fn))
    finally:
        os.chdir(ROOT)
        shutil.rmtree(td, ignore_errors=True)
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(man, indent=1))


if __name__ == "__main__":
    main()
