#!/bin/bash
# First GPU call(s) of the next round: validate the kernels that were written after the GPU budget of round 1 ran out, then A/B them.
# Usage (from the repo root, one gpurun call each; all output under gpurun_out/):
#   gpurun --timeout 600 -- 'bash tools/r2_first_calls.sh validate'
#   gpurun --timeout 600 -- 'bash tools/r2_first_calls.sh ab'
set -u
mkdir -p gpurun_out
case "${1:-validate}" in
  validate)
    # One pytest process per kernel family: a device-side trap poisons the CUDA context of its process only, and every family gets
    # its own verdict.  (CTA-pair GEMM, 95 KB GEMM+LN, forked encoder, decode tile widths, cluster split-K.)
    : > gpurun_out/r2_experimental.log
    for fam in "test_linear_pair_bf16" "test_encoder_pair_gemm" "test_prefill_pair_gemm_head_major" "test_linear_add_layernorm_small_footprint" "test_forked_encoder" \
               "test_linear_skinny_tile_width" "test_linear_wide_decode_tiles" "test_linear_cluster_splitk" "test_decode_with_cluster_splitk"; do
      echo "===== $fam" | tee -a gpurun_out/r2_experimental.log
      GSTVD_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q -s -k "$fam" >> gpurun_out/r2_experimental.log 2>&1
      echo "$fam: exit $?" | tee -a gpurun_out/r2_verdicts.log
    done
    cat gpurun_out/r2_verdicts.log
    grep -E "passed|failed|Error|error" gpurun_out/r2_experimental.log | tail -30
    ;;
  ab)
    # same-box A/B of every opt-in switch on the bench workload (3 streams) and on one single-stream round
    for cfg in "X=0" "GSTVD_GEMM_2CTA=1" "GSTVD_GEMM_2CTA=1 GSTVD_GEMM_2CTA_HM=1" "GSTVD_FUSE_LN=16 GSTVD_FUSE_LN_SMALL=1" "GSTVD_GEMM_SPLITK=1" "GSTVD_GEMM_SKINNY_BN=64" "GSTVD_GEMM_SKINNY_BN=128" "GSTVD_GEMM_WIDE_BN=128" "GSTVD_ENC_FORK=1"; do
      echo "== $cfg"
      env $cfg timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
      env $cfg timeout 60 python tools/profile_round.py --hist 150 2>&1 | tail -2
    done | tee gpurun_out/r2_ab.log
    # streams in flight with the default kernels (3 was the optimum before the late round-1 kernels)
    for n in 1 2; do
      echo "== --streams $n"
      timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams $n 2>/dev/null | tail -1
    done | tee -a gpurun_out/r2_ab.log
    ;;
esac
