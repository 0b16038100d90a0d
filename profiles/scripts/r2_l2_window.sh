# HISTORICAL (round 2): at the time of this run the persisting-L2 window was on by default and IREC_R2_NO_L2_WINDOW=1 switched it off;
# it is now opt-in (IREC_R2_L2_WINDOW=1).
# A/B of the persisting-L2 window on the exponent table: DRAM bytes of one configs[3]-size launch (bench value: identical, 3.046e9)
mkdir -p gpurun_out
for off in 1 0; do
  IREC_R2_NO_L2_WINDOW=$off ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_beam_encode_tmem -s 30 -c 1 --csv \
     --log-file gpurun_out/r2_l2_window_$off.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-c5 --no-is > /dev/null 2>&1
  echo "IREC_R2_NO_L2_WINDOW=$off"; grep "k_beam_encode_tmem" gpurun_out/r2_l2_window_$off.csv | sed 's/.*TmemArgs)",//' | cut -c1-200
done
