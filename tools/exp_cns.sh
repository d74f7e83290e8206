for sl in 1 4 8 16; do
  echo "SLICES=$sl"
  NB_CNS_SLICES=$sl NB_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep nb_consensus_job | tail -1
done
