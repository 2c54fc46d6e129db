for r in 256 0 512 1024 2048; do
  echo "=== KPREG_NORM_ROWS=$r"
  KPREG_NORM_ROWS=$r python tools/bench_layers.py 64 3 segnorm 2>&1 | grep -E "segnorm N=(2653008|943672) |totals"
done
python -m pytest tests/test_gpu_gemm_tc.py -x -q -k "segment_norm" 2>&1 | tail -2
