# sweep of the consensus-kernel tunables on the bench workload (run on the GPU box)
for l in 1 4 8 16 32; do
  echo "LANES=$l"
  NB_CNS_LANES=$l NB_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep nb_consensus_job | tail -1
done
