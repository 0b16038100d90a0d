# round-2 evidence run on one B200: sanitizer over the new kernels, launch list of the bench command, ncu of k_is_block
mkdir -p gpurun_out
SEL='test_beam_wide_vs_oracle and (B33 or B64 or B300) or test_beam_wide_batch_ragged or test_fused_block_single_rank and 64-36-20 or test_beam_general_path_vs_oracle and c1-D64-B20 or test_sharded_block_single_rank'
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_r2d_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_r2d_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_r2d_${tool}.log | tail -4
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench_final.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench_final.log
ncu --set full --clock-control none --import-source on -k regex:k_is_block -c 1 -s 1 -o gpurun_out/r2_is_block_c python bench_is.py --reps 1 --cpu-blocks 0 > gpurun_out/r2_is_block_c.log 2>&1
tail -2 gpurun_out/r2_is_block_c.log
