# memcheck over the rest of the -m gpu suite + the cross-stream cache test; ncu --set full of one bench launch at configs[3] size
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 0 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_models.py \
    tests/test_learned_ratios.py tests/test_rejection_sampler.py "tests/test_gpu_parity.py::test_exponent_table_cache_across_streams" -q -m gpu \
    -p no:cacheprovider --deselect tests/test_boxmuller_exhaustive.py > gpurun_out/sanitizer_r2f_memcheck_rest.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_r2f_memcheck_rest.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitizer_r2f_memcheck_rest.log | tail -4
IREC_BENCH_DUMP=gpurun_out/bench_levels_r2f.json ncu --set full --clock-control none --import-source on -k regex:k_beam_encode_tmem -s 30 -c 1 \
    -o gpurun_out/r2_tmem_f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c5 --no-is > gpurun_out/r2_tmem_f.log 2>&1
tail -2 gpurun_out/r2_tmem_f.log | cut -c1-300
