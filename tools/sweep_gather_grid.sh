set -x
for cfg in "32 4" "64 16" "128 16" "256 16" "512 16" "256 64" "1024 8"; do
  set -- $cfg
  echo "=== CTAS_PER_SM=$1 MIN_SLICE=$2"
  KPREG_GATHER_CTAS_PER_SM=$1 KPREG_GATHER_MIN_SLICE=$2 python tools/bench_layers.py 64 3 kpconv 2>&1 | tail -14
done
