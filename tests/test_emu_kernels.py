"""CPU: the product's CUDA kernels themselves, compiled for the host (tests/emu/cudaemu: every CUDA thread a fiber, warp
collectives and barriers as rendezvous, bulk copies / cp.async landing at once) from the UNMODIFIED sources under
necat_b200/csrc, driven through the same C-ABI and the same executables, against the oracle.  This is what lets the
CPU-only suite notice a logic regression in a kernel; it says nothing about speed, and the GPU suite (-m gpu) remains the
parity gate on the real device.  Inputs are tiny: the emulation runs a few hundred thousand times slower than a B200."""
import os
import shutil
import subprocess

import pytest

import util

EMU_DIR = os.path.join(util.ROOT, "tests", "emu")
EMU_BIN = os.path.join(EMU_DIR, "_build", "cudaemu", "bin")
PM11 = "-k 11 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"


@pytest.fixture(scope="module")
def emu(built):
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return lambda name: os.path.join(EMU_BIN, name)


def test_emulated_library_is_not_the_product(emu):
    """The emulated library lives under tests/ and the product never loads it."""
    from necat_b200 import capi
    assert os.path.realpath(capi.LIB_PATH).startswith(os.path.join(util.ROOT, "necat_b200"))
    out = subprocess.run(["ldd", emu("oc2cns")], capture_output=True, text=True).stdout
    assert "tests/emu/_build/cudaemu/bin/../lib/libnecat_b200.so" in out and "necat_b200/lib" not in out


def test_index_probe_vote_chain_kernels(emu, tmp_path):
    """k_mark_invalid, k_scan_kmers, the scans, k_sort_*, k_probe (bulk copy + mbarrier), k_build_blocks, k_candidates:
    oc2pmov CAN mode on the golden volume, k = 11."""
    wrk = util.golden_workdir(tmp_path)
    os.chdir(str(tmp_path))
    g = util.run_pmov(emu("oc2pmov"), PM11, wrk, 0, str(tmp_path / "g.bin"))
    o = util.run_pmov(util.ORACLE_CLI, PM11, wrk, 0, str(tmp_path / "o.bin"))
    assert o.shape[0] > 1000 and util.same_multiset(g, o)


def test_long_reads_many_probes_and_unstaged_probe_path(emu, tmp_path):
    """Reads of more than 65 535 probes (-z 1 on ~75 kb reads: the probe number of a hit takes 24 bits of the sort key of
    k_build_blocks) whose packed bases are probed from global memory instead of the shared-memory stage (NB_MAP_STAGE
    lowers the stage so that these reads are 'too long' for it)."""
    wrk, reads = util.make_workdir(tmp_path, genome=90_000, depth=3, seed=91, err=0.10, min_len=70_000, max_len=88_000, mean_log=11.3, sd_log=0.1)
    assert max(r.size for r in reads) > 65_535 + 15
    opts = PM11.replace("-z 10", "-z 1").replace("-k 11", "-k 13")
    g = util.run_pmov(emu("oc2pmov"), opts, wrk, 0, str(tmp_path / "g.bin"), prefix=("env", "NB_MAP_STAGE=4096"))
    o = util.run_pmov(util.ORACLE_CLI, opts, wrk, 0, str(tmp_path / "o.bin"))
    assert o.shape[0] >= 2 and util.same_multiset(g, o)


def test_aln_mode_kernels(emu, tmp_path):
    """oc2pmov -j 1: speculative k_extend + contain-rule replay, M4 text."""
    wrk, _ = util.make_workdir(tmp_path, genome=8000, depth=10, seed=83, err=0.10)
    opts = PM11.replace("-j 0 -u 1", "-j 1 -u 0")
    util.run([emu("oc2pmov")] + opts.split() + [wrk, "0", str(tmp_path / "g.m4")])
    util.run([util.ORACLE_CLI, "pmov"] + opts.split() + [wrk, "0", str(tmp_path / "o.m4")])
    g, o = sorted(open(str(tmp_path / "g.m4")).readlines()), sorted(open(str(tmp_path / "o.m4")).readlines())
    assert len(o) > 20 and g == o


@pytest.mark.parametrize("cns,gen", [
    # k_extend (two candidates per warp), k_cns_links, k_cns_dp (cp.async ring), k_cns_walk, k_gather_ascii
    ("-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0", dict(genome=8000, depth=12, seed=81, err=0.10)),
    # -r 1: k_rescue_local, k_rescue_columns, k_rescue_leaves and the wiring into the extension rounds
    ("-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 1 -f 1", dict(genome=6000, depth=10, seed=82, err=0.10, long_indel=0.5)),
])
def test_extension_and_consensus_kernels(emu, tmp_path, cns, gen):
    wrk, _ = util.make_workdir(tmp_path, **gen)
    cand = str(tmp_path / "o.bin")
    o = util.run_pmov(util.ORACLE_CLI, PM11, wrk, 0, cand)
    assert o.shape[0] > 30
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([emu("oc2pcan"), "-t", "1", wrk, c])
    util.run([emu("oc2cns")] + cns.split() + [wrk, c, str(tmp_path / "g_cns.fa"), str(tmp_path / "g_raw.fa")])
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "0", wrk, cand, str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
    exp = util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert len(exp) >= 5
    assert util.fasta_canon(str(tmp_path / "g_cns.fa")) == exp
    assert util.fasta_canon(str(tmp_path / "g_raw.fa")) == util.fasta_canon(str(tmp_path / "o_raw.fa"))


def test_consensus_over_several_volumes(emu, tmp_path):
    """oc2cns merges the volumes of the work directory into one read store (consensus/main.c:66; nbh_volume_append shifts
    the packed bytes of every further volume to the bit offset the store ends at)."""
    wrk, _ = util.make_workdir(tmp_path, genome=7000, depth=12, seed=85, err=0.10, vol_size=20_001)
    assert int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0]) >= 3
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
    cand = str(tmp_path / "o.bin")
    with open(cand, "wb") as f:
        for vid in range(int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])):
            util.run_pmov(util.ORACLE_CLI, PM11, wrk, vid, str(tmp_path / "v.bin"))
            f.write(open(str(tmp_path / "v.bin"), "rb").read())
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([emu("oc2pcan"), "-t", "1", wrk, c])
    util.run([emu("oc2cns")] + cns.split() + [wrk, c, str(tmp_path / "g_cns.fa"), str(tmp_path / "g_raw.fa")])
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "0", wrk, cand, str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
    exp = util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert len(exp) >= 5
    assert util.fasta_canon(str(tmp_path / "g_cns.fa")) == exp
    assert util.fasta_canon(str(tmp_path / "g_raw.fa")) == util.fasta_canon(str(tmp_path / "o_raw.fa"))


def test_sharded_pipeline_through_the_python_api(emu, tmp_path):
    """bench.py's multi-GPU recipe with the emulated library behind necat_b200.capi (patched in a child
    process only): the reads are mapped in two ranges as two ranks would, the candidate records are concatenated as the
    all-gather does, and each rank corrects the templates it owns; the union of the two ranks' records is what one rank
    alone produces."""
    wrk, _ = util.make_workdir(tmp_path, genome=6000, depth=10, seed=84, err=0.10)
    code = r'''
import sys, os, numpy as np
sys.path.insert(0, %(root)r)
from necat_b200 import capi, formats, dist as nbdist
capi.LIB_PATH = %(lib)r                      # test only: the emulated build (tests/emu), never the product's path
vol = formats.read_volume(os.path.join(%(wrk)r, "vol0"))
ctx = capi.Context(0)
popt = capi.default_map_options(kmer_size=11, scan_window=10, kmer_cnt_cutoff=100, block_size=2000, block_score_cutoff=3,
                                num_candidates=500, align_size_cutoff=1000, ddfs_cutoff=0.25, error=0.5, num_output=500)
copt = capi.default_cns_options(min_align_size=1000, min_cov=3, max_cov=12, min_size=500, error=0.5, mapping_ratio=0.8,
                                rescue_long_indels=0, use_fixed_ident_cutoff=0, full_consensus=0)
dvol = ctx.upload(vol)
idx = ctx.build_index(dvol, 11, 100)
parts = []
for rank in range(2):
    lo, hi = nbdist.shard_range(vol.nseq, rank, 2)
    parts.append(capi.pack_candidates(ctx.map_volume(idx, dvol, 0, dvol, 0, lo, hi, popt)))
packed = np.concatenate(parts)
whole = capi.pack_candidates(ctx.map_volume(idx, dvol, 0, dvol, 0, 0, vol.nseq, popt))
assert formats.canonical_records(packed).tobytes() == formats.canonical_records(whole).tobytes()
part = np.concatenate([packed, formats.swap_roles(packed)])
def records(world, rank):
    recs, seq = ctx.consensus_job(dvol, part, copt, world, rank)
    a = seq.array()
    out = sorted((int(r["id"]), int(r["is_raw"]), int(r["left"]), int(r["right"]), bytes(a[int(r["seq_offset"]):int(r["seq_offset"]) + int(r["seq_len"])])) for r in recs)
    seq.free()
    return out
one = records(1, 0)
two = sorted(records(2, 0) + records(2, 1))
assert len(one) > 5 and one == two, (len(one), len(two))
assert all(nbdist.owns_template([r[0]], 0, 2)[0] for r in records(2, 0))
print("OK", len(packed), len(one))
''' % dict(root=util.ROOT, lib=os.path.join(EMU_DIR, "_build", "cudaemu", "lib", "libnecat_b200.so"), wrk=wrk)
    p = util.run(["python", "-c", code])
    assert p.stdout.startswith(b"OK")


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_results_do_not_depend_on_lane_order(emu, tmp_path, order):
    """NB_EMU_ORDER permutes the order in which the runnable lanes of a warp (and the warps of a block) get their turn.
    Two lanes that communicate through shared or global memory without a __syncwarp() / collective / barrier between them
    would make the result depend on it."""
    env = dict(os.environ, NB_EMU_ORDER=order)
    wrk, _ = util.make_workdir(tmp_path, genome=6000, depth=10, seed=85, err=0.12, long_indel=0.3)
    cand = str(tmp_path / "o.bin")
    o = util.run_pmov(util.ORACLE_CLI, PM11, wrk, 0, cand)
    util.run([emu("oc2pmov")] + PM11.split() + [wrk, "0", str(tmp_path / "g.bin")], env=env)
    from necat_b200 import formats
    assert o.shape[0] > 30 and util.same_multiset(formats.read_candidates(str(tmp_path / "g.bin")), o)
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"
    c = str(tmp_path / "c.bin")
    shutil.copy(cand, c)
    util.run([emu("oc2pcan"), "-t", "1", wrk, c])
    util.run([emu("oc2cns")] + cns.split() + [wrk, c, str(tmp_path / "g_cns.fa"), str(tmp_path / "g_raw.fa")], env=env)
    util.run([util.ORACLE_CLI, "cns"] + cns.split() + ["-M", "0", wrk, cand, str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
    assert util.fasta_canon(str(tmp_path / "g_cns.fa")) == util.fasta_canon(str(tmp_path / "o_cns.fa"))
    assert util.fasta_canon(str(tmp_path / "g_raw.fa")) == util.fasta_canon(str(tmp_path / "o_raw.fa"))


def test_small_memory_mode_matches_reference(emu, tmp_path):
    """oc2cns -s 1 against the reference binary (consensus/main.c:66, consensus_one_read.c:249-255,
    consensus_one_partition.c:173): no dump of the uncorrected reads of the id range."""
    if not util.ref_bin("oc2cns"):
        pytest.skip("oracle/_ref is not built")
    wrk, _ = util.make_workdir(tmp_path, genome=8000, depth=12, seed=86, err=0.10)
    cand = str(tmp_path / "o.bin")
    util.run_pmov(util.ORACLE_CLI, PM11, wrk, 0, cand)
    cns = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0 -s 1"
    for tag, bins in (("g", EMU_BIN), ("r", util.REF_DIR)):
        c = str(tmp_path / f"{tag}.bin")
        shutil.copy(cand, c)
        util.run([os.path.join(bins, "oc2pcan"), "-t", "1", wrk, c])
        util.run([os.path.join(bins, "oc2cns")] + cns.split() + ["-t", "1", wrk, c, str(tmp_path / f"{tag}_cns.fa"), str(tmp_path / f"{tag}_raw.fa")])
    assert len(util.fasta_canon(str(tmp_path / "r_cns.fa"))) >= 5
    assert util.fasta_canon(str(tmp_path / "g_cns.fa")) == util.fasta_canon(str(tmp_path / "r_cns.fa"))
    assert util.fasta_canon(str(tmp_path / "g_raw.fa")) == util.fasta_canon(str(tmp_path / "r_raw.fa"))


def test_mapping_in_pieces(emu, tmp_path):
    """nb_map_volume cuts the read range by probe count and halves a piece whose seed hits exceed the scratch budget."""
    wrk, _ = util.make_workdir(tmp_path, genome=8000, depth=12, seed=87, err=0.10)
    o = util.run_pmov(util.ORACLE_CLI, PM11, wrk, 0, str(tmp_path / "o.bin"))
    from necat_b200 import formats
    for env in ({"NB_MAP_PROBES": "3000"}, {"NB_MAP_HITS": "1500"}, {"NB_MAP_PROBES": "5000", "NB_MAP_HITS": "1100"}):
        util.run([emu("oc2pmov")] + PM11.split() + [wrk, "0", str(tmp_path / "g.bin")], env=dict(os.environ, **env))
        assert o.shape[0] > 30 and util.same_multiset(formats.read_candidates(str(tmp_path / "g.bin")), o), env


def test_sharded_index_build(emu, tmp_path):
    """nb_index_build_shard / nb_index_shard_export / nb_index_adopt: three slices of the hash space assembled by hand
    (on a GPU box the slices come from different ranks and are all-gathered over NCCL, necat_b200/capi.py) give the
    arrays of nb_index_build, and the adopted index maps like it."""
    wrk, _ = util.make_workdir(tmp_path, genome=6000, depth=10, seed=88, err=0.10)
    code = r'''
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, %(root)r)
from necat_b200 import capi, formats
capi.LIB_PATH = %(lib)r                      # test only: the emulated build (tests/emu), never the product's path
vol = formats.read_volume(os.path.join(%(wrk)r, "vol0"))
ctx = capi.Context(0)
L = ctx.L
dvol = ctx.upload(vol)
k, q, n = 11, 100, 3
idx = ctx.build_index(dvol, k, q)
starts, offs, nd = ctx.index_arrays(idx, k)
n_hash = 1 << (2 * k)
slice_len = (n_hash + n - 1) // n
tbl = np.zeros(n * slice_len + 8, np.uint32)
shards, counts = [], []
for s in range(n):
    sh = C.c_void_p(); no = C.c_uint64(); dd = C.c_uint64()
    assert L.nb_index_build_shard(ctx.h, dvol, C.c_int(k), C.c_int(q), C.c_int(s), C.c_int(n), C.byref(sh), C.byref(no), C.byref(dd)) == 0
    shards.append(sh); counts.append((no.value, dd.value))
total = sum(c[0] for c in counts)
of = np.zeros(total + 16, np.uint32)
base = 0
for s in range(n):
    assert L.nb_index_shard_export(ctx.h, shards[s], C.c_uint64(base), C.c_void_p(tbl.ctypes.data + 4 * s * slice_len), C.c_void_p(of.ctypes.data + 4 * base)) == 0
    base += counts[s][0]
tbl[n_hash] = total
assert total == offs.size and sum(c[1] for c in counts) == nd
assert np.array_equal(tbl[:n_hash + 1], starts) and np.array_equal(of[:total], offs)
h = C.c_void_p()
assert L.nb_index_adopt(ctx.h, C.c_int(k), C.c_int(q), C.c_void_p(tbl.ctypes.data), C.c_void_p(of.ctypes.data), C.c_uint64(total), C.c_uint64(nd), C.byref(h)) == 0
popt = capi.default_map_options(kmer_size=11, scan_window=10, kmer_cnt_cutoff=100, block_size=2000, block_score_cutoff=3,
                                num_candidates=500, align_size_cutoff=1000, ddfs_cutoff=0.25, error=0.5, num_output=500)
a = capi.pack_candidates(ctx.map_volume(idx, dvol, 0, dvol, 0, 0, vol.nseq, popt))
b = capi.pack_candidates(ctx.map_volume(h, dvol, 0, dvol, 0, 0, vol.nseq, popt))
assert a.shape[0] > 10 and formats.canonical_records(a).tobytes() == formats.canonical_records(b).tobytes()
ctx.free_index(h); ctx.free_index(idx)
print("OK", total, a.shape[0])
''' % dict(root=util.ROOT, lib=os.path.join(EMU_DIR, "_build", "cudaemu", "lib", "libnecat_b200.so"), wrk=wrk)
    p = util.run(["python", "-c", code])
    assert p.stdout.startswith(b"OK")
