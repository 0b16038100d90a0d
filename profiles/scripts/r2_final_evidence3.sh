# final build: full -m gpu suite, ncu --set full of one configs[3]-size bench launch, the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2
IREC_BENCH_DUMP=gpurun_out/bench_levels_r2g.json ncu --set full --clock-control none --import-source on -k regex:k_beam_encode_tmem -s 30 -c 1 \
    -o gpurun_out/r2_tmem_g python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c5 --no-is > gpurun_out/r2_tmem_g.log 2>&1
tail -1 gpurun_out/r2_tmem_g.log | cut -c1-200
