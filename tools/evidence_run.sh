# One gpurun call that refreshes the evidence under gpurun_out/ (copied to profiles/ by hand): tests, bench line, layer times,
# ncu launch list + DRAM bytes of one 8-pair step, --set full of the aggregate kernel, reference arm.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r3_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r3_smoke.txt
python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err
python tools/bench_layers.py 64 3 > gpurun_out/r3_layer_times.txt 2>&1
KPREG_BENCH_NO_RAMP=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r3_launches.csv python tools/ncu_step.py 8 > gpurun_out/r3_ncu1.log 2>&1
python tools/summarize_launches.py gpurun_out/r3_launches.csv > gpurun_out/r3_launches_summary.md
python tools/summarize_traffic.py gpurun_out/r3_launches.csv 8 > gpurun_out/r3_dram_traffic.json
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_kpconv_gather_mma -c 8 \
    -f -o /tmp/prof_gather python tools/ncu_step.py 8 > gpurun_out/r3_ncu2.log 2>&1
python tools/summarize_ncu_full.py /tmp/prof_gather.ncu-rep > gpurun_out/r3_ncu_full_gather.md
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3_bench_reference_arm.json 2> gpurun_out/r3_bench_reference_arm.err
rm -f gpurun_out/r3_launches.csv.tmp
ls -la gpurun_out | tail -12
