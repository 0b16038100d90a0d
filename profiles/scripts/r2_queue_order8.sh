mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_beam_encode_tmem -s 30 -c 1 --csv \
   --log-file gpurun_out/r2_queue_order8_dram.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c5 --no-is > /dev/null 2>&1
grep "k_beam_encode_tmem" gpurun_out/r2_queue_order8_dram.csv | sed 's/.*TmemArgs)",//' | cut -c1-200
python bench.py --no-cpu-baseline --no-c5 --no-is --no-e2e --steps 3 --warmup 3 --images-total 128 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(128, d['value'], d['roofline']['avg_launch_ms'])"
