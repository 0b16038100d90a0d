#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over a subset of the -m gpu parity tests that reaches every kernel
# family: resident2 and the tensor-memory kernel (table and in-place exponents), cluster (DSMEM all-gather), the general path
# (k_gp_*) and its fused cooperative version, the importance sampler (k_is_block with and without the candidate table,
# k_is_decode) and the peer-memory exchange kernel in loop-back.  Logs -> gpurun_out/sanitizer_<tool>.log
# usage: profiles/scripts/run_sanitizer.sh [tag] [tools] [extra -k clause]
TAG=${1:-r2}
TOOLS=${2:-"memcheck racecheck synccheck"}      # e.g. "racecheck"
EXTRA=${3:-}                                      # appended to the -k expression, e.g. "and tmem" / "and not tmem"
mkdir -p gpurun_out
SEL='test_beam_resident_vs_oracle and (c1-D64-B20 or c2-D37-B5 or c2-D257-B32) or test_beam_general_path_vs_oracle and c1-D64-B20 or test_gaussian_coder_importance and c1 or test_p2p_exchange_single_rank_loopback or test_beam_ragged_blocks_one_launch or test_sharded_block_single_rank or test_fused_block_single_rank and (64-36-20 or 37-300-5) or test_is_encode_candidate_table_equals_in_place and 300 or test_decode_rejects or test_schedule_export'
if [ -n "$EXTRA" ]; then SEL="($SEL) $EXTRA"; fi
for tool in $TOOLS; do
  echo "=== $tool ===" 
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" -p no:cacheprovider \
      > gpurun_out/sanitizer_${TAG}_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_${TAG}_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_${TAG}_${tool}.log | tail -5
done
