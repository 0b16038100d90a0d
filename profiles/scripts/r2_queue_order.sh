# queue order in coarse work classes (k_tm_order): parity, DRAM bytes of one configs[3]-size launch, bench values
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q -k "tmem or ragged or batch or default_kernel or streams or compress" 2>&1 | tail -2
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_beam_encode_tmem -s 30 -c 1 --csv \
   --log-file gpurun_out/r2_queue_order_dram.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c5 --no-is > /dev/null 2>&1
grep "k_beam_encode_tmem" gpurun_out/r2_queue_order_dram.csv | sed 's/.*TmemArgs)",//' | cut -c1-200
F="--no-cpu-baseline --no-c5 --no-is --no-e2e --steps 3 --warmup 3"
for n in 1024 128; do python bench.py $F --images-total $n 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print($n, d['value'], d['roofline']['avg_launch_ms'])"; done
