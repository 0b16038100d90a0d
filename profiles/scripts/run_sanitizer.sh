#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over a subset of the -m gpu parity tests that reaches every kernel
# family: resident2 (table and in-place exponents), cluster (DSMEM all-gather), the general path (k_gp_*), the importance
# sampler (k_is_block) and the peer-memory exchange kernel in loop-back.  Logs -> gpurun_out/sanitizer_<tool>.log
# usage: profiles/scripts/run_sanitizer.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
SEL='test_beam_resident_vs_oracle and (c1-D64-B20 or c2-D37-B5 or c2-D257-B32) or test_beam_general_path_vs_oracle and c1-D64-B20 or test_gaussian_coder_importance and c1 or test_p2p_exchange_single_rank_loopback or test_beam_ragged_blocks_one_launch or test_sharded_block_single_rank'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool ===" 
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" -p no:cacheprovider \
      > gpurun_out/sanitizer_${TAG}_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_${TAG}_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_${TAG}_${tool}.log | tail -5
done
