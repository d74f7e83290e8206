#!/usr/bin/env python3
"""Deterministic synthetic ONT-like read generator.

Uniform random ACGT genome; reads sampled uniformly from both strands; lengths
log-normal (N50 ~ 8 kb) clipped to [min_len, max_len]; per-base error `err`
split sub/del/ins = 35/35/30.  ACGT only, names r<i>.  Seeded => reproducible.

Usage: gen_reads.py --genome 4600000 --depth 40 --seed 20260925 --out reads.fasta
"""
import argparse
import numpy as np

COMP = np.array([3, 2, 1, 0], dtype=np.uint8)
ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutate(seq, rng, err, fsub=0.35, fdel=0.35, fins=0.30):
    n = seq.size
    u = rng.random(n)
    is_sub = u < err * fsub
    is_del = (u >= err * fsub) & (u < err * (fsub + fdel))
    is_ins = (u >= err * (fsub + fdel)) & (u < err)
    out = seq.copy()
    ns = int(is_sub.sum())
    if ns:
        out[is_sub] = (out[is_sub] + rng.integers(1, 4, ns, dtype=np.uint8)) & 3
    # build with insertions (a random base after the position) and deletions
    rep = np.ones(n, dtype=np.int64)
    rep[is_del] = 0
    rep[is_ins] = 2
    idx = np.repeat(np.arange(n), rep)
    res = out[idx]
    # second copy of an inserted position becomes a random base
    first = np.ones(idx.size, dtype=bool)
    first[1:] = idx[1:] != idx[:-1]
    nins = int((~first).sum())
    if nins:
        res[~first] = rng.integers(0, 4, nins, dtype=np.uint8)
    return res


def generate(genome_size, depth, seed, err=0.15, min_len=3000, max_len=60000,
             mean_log=8.75, sd_log=0.55, repeat_len=0, repeat_copies=0, n_repeats=0, long_indel=0.0):
    """n_repeats families of `repeat_copies` near-identical (1 % diverged) copies of a
    random repeat_len segment are planted into the genome (exercises the k-mer count
    cutoff, the 40-seed block cap and block zeroing).  long_indel: fraction of reads that
    get one 300-700 base deletion or random insertion in their middle half (what oc2cns -r 1
    rescues: the block-wise extension stops there); 0 draws nothing extra from the RNG."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_size, dtype=np.uint8)
    for _ in range(n_repeats):
        unit = rng.integers(0, 4, repeat_len, dtype=np.uint8)
        for _c in range(repeat_copies):
            cp = unit.copy()
            m = rng.random(repeat_len) < 0.01
            cp[m] = (cp[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
            s0 = int(rng.integers(0, genome_size - repeat_len))
            genome[s0:s0 + repeat_len] = cp
    target = int(genome_size * depth)
    reads = []
    total = 0
    while total < target:
        L = int(np.clip(rng.lognormal(mean_log, sd_log), min_len, max_len))
        L = min(L, genome_size)
        s = int(rng.integers(0, genome_size - L + 1))
        frag = genome[s:s + L]
        if rng.random() < 0.5:
            frag = COMP[frag[::-1]]
        r = mutate(frag, rng, err)
        if long_indel > 0.0 and rng.random() < long_indel and r.size > 4000:
            n = int(rng.integers(300, 700))
            at = int(rng.integers(r.size // 4, 3 * r.size // 4))
            if rng.random() < 0.5:
                r = np.concatenate([r[:at], r[at + n:]])
            else:
                r = np.concatenate([r[:at], rng.integers(0, 4, n, dtype=np.uint8), r[at:]])
        if r.size < min_len:
            continue
        reads.append(r)
        total += r.size
    return genome, reads


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(ALPHA[r].tobytes())
            f.write(b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, required=True)
    ap.add_argument("--depth", type=float, required=True)
    ap.add_argument("--seed", type=int, default=20260925)
    ap.add_argument("--err", type=float, default=0.15)
    ap.add_argument("--min-len", type=int, default=3000)
    ap.add_argument("--max-len", type=int, default=60000)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    _, reads = generate(a.genome, a.depth, a.seed, a.err, a.min_len, a.max_len)
    write_fasta(a.out, reads)
    n = len(reads)
    tot = sum(r.size for r in reads)
    print(f"reads={n} bases={tot}")


if __name__ == "__main__":
    main()
