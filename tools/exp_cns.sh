for cfg in "1 1" "1 2" "1 4" "1 8" "0 4"; do
  set -- $cfg
  echo "SORT=$1 LANES=$2"
  NB_CNS_SORT=$1 NB_CNS_LANES=$2 NB_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 | grep nb_consensus_job | tail -1 | sed 's/.*gpu ms/gpu ms/'
done
