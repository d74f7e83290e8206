#!/usr/bin/env python3
"""Random-option soak of the ORACLE against the reference binaries (oracle/_ref): random overlap options (-k -z -q -b -s
-n -a -m), random consensus options (-a -x -y -l -p -u -f -r) and random small read sets (error rate, depth, planted
repeats, long indels, ragged lengths); candidates and both FASTA files must agree in canonical order.  CPU only; test
infrastructure like everything that touches oracle/.

    python tools/soak_oracle.py <seed> <cases>            # oracle against the reference binaries (CPU)
    python tools/soak_oracle.py <seed> <cases> gpu        # this repo's executables against the oracle (needs a B200)
    python tools/soak_oracle.py <seed> <cases> emu        # the EMULATED build of the kernels (tests/emu) against the oracle:
                                                          # CPU only, smaller read sets, ~1 minute per case
"""
import sys, pathlib, tempfile, shutil, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tools')); sys.path.insert(0, ROOT)
import util
rng=np.random.default_rng(int(sys.argv[1]))
EMU = len(sys.argv) > 3 and sys.argv[3] == "emu"
GPU = EMU or (len(sys.argv) > 3 and sys.argv[3] == "gpu")
if EMU:
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
def exe(name):      # the arm that is compared with the oracle
    if EMU:
        return os.path.join(ROOT, "tests", "emu", "_build", "cudaemu", "bin", name)
    return os.path.join(util.BIN_DIR, name) if GPU else util.ref_bin(name)
bad=0
for case in range(int(sys.argv[2])):
    k=int(rng.choice([10,11]) if EMU else rng.choice([11,12,13])); z=int(rng.choice([5,10,20])); q=int(rng.choice([50,100,500])); b=int(rng.choice([1000,2000,4000]))
    s=int(rng.choice([2,3,4])); n=int(rng.choice([5,50,500])); a=int(rng.choice([500,1000,2000])); m=int(rng.choice([20,500]))
    pm=f"-k {k} -z {z} -q {q} -b {b} -s {s} -n {n} -a {a} -d 0.25 -e 0.5 -m {m} -t 4 -j 0 -u 1 -i 1"
    x=int(rng.choice([2,3,4,6])); y=int(rng.choice([8,12,20])); l=int(rng.choice([500,1000,2000])); p=float(rng.choice([0.6,0.8,0.9])); u=int(rng.choice([0,1])); f=int(rng.choice([0,1])); r=int(rng.choice([0,0,1]))
    ca=int(rng.choice([500,1000,2000]))
    cns=f"-a {ca} -x {x} -y {y} -l {l} -e 0.5 -p {p} -u {u} -r {r} -f {f}"
    gen=dict(genome=int(rng.choice([6_000,10_000]) if EMU else rng.choice([20_000,40_000])), depth=int(rng.choice([8,15,25])), seed=int(rng.integers(1,10**6)), err=float(rng.choice([0.06,0.10,0.15])))
    if rng.random()<0.4: gen.update(repeat_len=2000, repeat_copies=int(rng.choice([4,10])), n_repeats=1)
    if rng.random()<0.4: gen.update(long_indel=0.4)
    if rng.random()<0.3: gen.update(min_len=200, max_len=6000, mean_log=7.2, sd_log=0.7)
    t=pathlib.Path(tempfile.mkdtemp()); t0=time.time()
    try:
        wrk,_=util.make_workdir(t, **gen)
        rr=util.run_pmov(exe("oc2pmov"), pm, wrk, 0, str(t/"r.bin"))
        oo=util.run_pmov(util.ORACLE_CLI, pm, wrk, 0, str(t/"o.bin"))
        ok1=util.same_multiset(rr,oo)
        c=str(t/"c.bin"); shutil.copy(str(t/"r.bin"), c)
        util.run([exe("oc2pcan"), "-t","1", wrk, c])
        util.run([exe("oc2cns")]+cns.split()+["-t","4",wrk,c,str(t/"r_cns.fa"),str(t/"r_raw.fa")])
        util.run([util.ORACLE_CLI,"cns"]+cns.split()+["-M","0" if GPU else "1",wrk,str(t/"r.bin"),str(t/"o_cns.fa"),str(t/"o_raw.fa")])
        ok2=util.fasta_canon(str(t/"r_cns.fa"))==util.fasta_canon(str(t/"o_cns.fa")); ok3=util.fasta_canon(str(t/"r_raw.fa"))==util.fasta_canon(str(t/"o_raw.fa"))
        # ALN mode (-j 1): M4 text lines, with names or ids
        ok4 = True
        if rng.random() < 0.5:
            pm4 = pm.replace("-j 0 -u 1 -i 1", "-j 1 -u 0 -i %d" % int(rng.choice([0, 1])))
            util.run([exe("oc2pmov")]+pm4.split()+[wrk,"0",str(t/"r.m4")])
            util.run([util.ORACLE_CLI,"pmov"]+pm4.split()+[wrk,"0",str(t/"o.m4")])
            ok4 = sorted(open(t/"r.m4").readlines()) == sorted(open(t/"o.m4").readlines())
        ok3 = ok3 and ok4
        print(case, "OK" if (ok1 and ok2 and ok3) else "MISMATCH", ok1,ok2,ok3, rr.shape[0], len(util.fasta_canon(str(t/"r_cns.fa"))), "%.0fs"%(time.time()-t0), "|", pm, "|", cns, "|", gen, flush=True)
        bad += not (ok1 and ok2 and ok3)
    except Exception as e:
        print(case, "ERROR", str(e)[:300], "|", pm, "|", cns, "|", gen, flush=True); bad+=1
    shutil.rmtree(t, ignore_errors=True)
print("done, bad =", bad)
