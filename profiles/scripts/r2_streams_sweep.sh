# HISTORICAL (round 2): run when the cross-stream wait on the exponent table was still unconditional; IREC_R2_XSTREAM_EXPERIMENT
# skipped it for this measurement only.  The library now shares the table across streams safely (r2_tab_acquire) and ignores the variable.
set -x
mkdir -p gpurun_out
export IREC_R2_XSTREAM_EXPERIMENT=1
F="--no-cpu-baseline --no-c5 --no-is --no-e2e --steps 3 --warmup 3"
for k in 1 2 3 4 6; do
  python bench.py $F --images-total 128 --streams $k > gpurun_out/streams_128_$k.json 2> gpurun_out/streams_128_$k.err
done
for k in 1 2 4; do
  python bench.py $F --images-total 256 --streams $k > gpurun_out/streams_256_$k.json 2> gpurun_out/streams_256_$k.err
done
for k in 1 2; do
  python bench.py $F --images-total 1024 --streams $k > gpurun_out/streams_1024_$k.json 2> gpurun_out/streams_1024_$k.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/streams_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
P
