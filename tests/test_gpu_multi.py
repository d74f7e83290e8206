"""Two GPUs, one process each (NCCL): the volume-sharded correction iteration of necat_b200/pipeline.py -- index volumes
round-robin over the ranks, query volumes streamed, the all-to-all-v of candidate records, partitions corrected by their
owners -- against the oracle's single-process run of the same work directory; and the index of a single volume built in
two shards of the hash space and all-gathered.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import pickle

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

PM13 = dict(kmer_size=13, scan_window=10, kmer_cnt_cutoff=100, block_size=2000, block_score_cutoff=3, num_candidates=500,
            align_size_cutoff=1000, ddfs_cutoff=0.25, error=0.5, num_output=500)
PM13_STR = "-k 13 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 2 -j 0 -u 1 -i 1"
CNS = dict(min_align_size=1000, min_cov=3, max_cov=12, min_size=500, full_consensus=0, error=0.5, mapping_ratio=0.8,
           rescue_long_indels=0, use_fixed_ident_cutoff=0)
CNS_STR = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


def _worker(rank, world, port, wrk, part_size, out_dir):
    import torch
    import torch.distributed as dist
    from necat_b200 import capi, formats, pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["NECAT_B200_INDEX"] = "sharded"          # (small volumes would be indexed by every rank: the test wants the sharded build)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    vols = [formats.read_volume(os.path.join(wrk, f"vol{v}")) for v in range(nvol)]
    ctx = capi.Context(rank)
    popt, copt = capi.default_map_options(**PM13), capi.default_cns_options(**CNS)
    stats = {}
    res = pipeline.correct_iteration(ctx, vols, popt, copt, rank, world, dev, part_size=part_size, stats=stats)
    names = [v.name(i).encode() for v in vols for i in range(v.nseq)]
    cns, raw = [], []
    for recs, seq in res:
        c, r = formats.cns_fasta_records(recs, seq.array(), names)
        cns += c; raw += r
        seq.free()
    pickle.dump((cns, raw, stats), open(os.path.join(out_dir, f"rank{rank}.pkl"), "wb"))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


@pytest.mark.parametrize("vol_size,part_size", [(1_400_000, 150), (1_400_000, 100000), (2_000_000_000, 100000)])
def test_two_gpu_iteration_matches_oracle(built, tmp_path, vol_size, part_size):
    """vol_size 1.4 Mb: >= 3 volumes (index volumes round-robin, 6 volume pairs); part_size 150: several partitions per
    rank; 100 000: one partition shared by both ranks; one volume: its index is built in two shards and all-gathered."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    wrk, reads = util.make_workdir(tmp_path, genome=200_000, depth=22, seed=71, err=0.12, vol_size=vol_size)
    nvol = int(open(os.path.join(wrk, "reads_info.txt")).read().split()[0])
    assert nvol >= 3 or vol_size > 10**9
    cand = str(tmp_path / "o_cand.bin")
    with open(cand, "wb") as out:
        for vid in range(nvol):
            p = str(tmp_path / f"pm{vid}.bin")
            util.run_pmov(util.ORACLE_CLI, PM13_STR, wrk, vid, p)
            out.write(open(p, "rb").read())
    util.run([util.ORACLE_CLI, "cns"] + CNS_STR.split() + ["-M", "0", "-P", str(part_size), wrk, cand,
                                                            str(tmp_path / "o_cns.fa"), str(tmp_path / "o_raw.fa")])
    port = 29800 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, wrk, part_size, str(tmp_path)), nprocs=2, join=True)
    cns, raw, ncan = [], [], 0
    for r in range(2):
        c, w, st = pickle.load(open(str(tmp_path / f"rank{r}.pkl"), "rb"))
        cns += c; raw += w; ncan += st["n_candidates_local"]
    from necat_b200 import formats
    assert ncan == formats.read_candidates(cand).shape[0]
    exp_c, exp_r = util.fasta_canon(str(tmp_path / "o_cns.fa")), util.fasta_canon(str(tmp_path / "o_raw.fa"))
    assert len(exp_c) > 100
    assert sorted(cns) == exp_c
    assert sorted(raw) == exp_r
