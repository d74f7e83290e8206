# sweep of the extension-group pipeline on the bench workload (run on the GPU box); NB_TRACE prints the phase times
run() { echo "== $*"; env "$@" NB_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | grep -E "nb_consensus_job|ms_per_step" | tail -3 | cut -c1-460; }
run NB_CNS_GROUPS=1
run NB_CNS_GROUPS=2
run NB_CNS_GROUPS=3
run NB_CNS_GROUPS=4
run NB_CNS_GROUPS=3 NB_CNS_LONG=1
