# HISTORICAL (round 2): IREC_GP_VARIANT selected compile-time variants of k_gp_fused<20> for this A/B only; variant 1 (no in-place
# bank assignment) is now the only build and the variable is ignored.
# A/B of the scoring variants of k_gp_fused<20> (IREC_GP_VARIANT, experiment) + one full ncu capture of the default
mkdir -p gpurun_out
for v in 0 1 2 3; do
  echo "variant $v"
  IREC_GP_VARIANT=$v python bench_sweep.py --bits 16,18,20 --fused --reps 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['omega_bits'], round(d['ms'], 3), '%.3e' % d['candidates_per_sec'])"
done
ncu --set full --clock-control none --import-source on -k regex:k_gp_fused -c 1 -s 1 -o gpurun_out/r2_gp_fused_a python bench_sweep.py --bits 18 --fused --reps 1 > gpurun_out/r2_gp_fused_a.log 2>&1
tail -2 gpurun_out/r2_gp_fused_a.log
