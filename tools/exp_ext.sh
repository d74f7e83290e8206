for c in 3 4; do
  echo "EXT_CTAS=$c"
  NB_EXT_CTAS=$c NB_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 | grep nb_consensus_job | tail -1 | sed 's/.*gpu ms/gpu ms/'
done
