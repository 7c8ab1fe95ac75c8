#!/bin/bash
# First GPU call(s) of the next round: validate the kernels that were written after the GPU budget of round 1 ran out, then A/B them.
# Usage (from the repo root, one gpurun call each; all output under gpurun_out/):
#   gpurun --timeout 600 -- 'bash tools/r2_first_calls.sh validate'
#   gpurun --timeout 600 -- 'bash tools/r2_first_calls.sh ab'
set -u
mkdir -p gpurun_out
case "${1:-validate}" in
  validate)
    # CTA-pair GEMM (tcgen05 cta_group::2): op-level parity on ragged shapes + whole-encoder comparison; own process, own timeout
    GSTVD_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q -s > gpurun_out/r2_experimental.log 2>&1
    tail -25 gpurun_out/r2_experimental.log
    ;;
  ab)
    # same-box A/B of every opt-in switch on the bench workload (3 streams) and on one single-stream round
    for cfg in "X=0" "GSTVD_GEMM_2CTA=1" "GSTVD_GEMM_2CTA=256" "GSTVD_FUSE_LN=16" "GSTVD_FUSE_LN=16 GSTVD_FUSE_LN_SMALL=1" "GSTVD_GEMM_SPLITK=1" "GSTVD_GEMM_SPLITK=2" "GSTVD_GEMM_SKINNY_BN=64" "GSTVD_GEMM_SKINNY_BN=128" "GSTVD_GEMM_SKINNY_BN=64 GSTVD_GEMM_WIDE_BN=128" "GSTVD_GEMM_WIDE_BN=128" "GSTVD_ENC_FORK=1" "GSTVD_SELF_ANC=0" "GSTVD_SELF_V2=0 GSTVD_SELF_ANC=0"; do
      echo "== $cfg"
      env $cfg timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
      env $cfg timeout 60 python tools/profile_round.py --hist 150 2>&1 | tail -2
    done | tee gpurun_out/r2_ab.log
    # streams in flight with the default kernels (3 was the optimum before the late round-1 kernels)
    for n in 2 4; do
      echo "== --streams $n"
      timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --streams $n 2>/dev/null | tail -1
    done | tee -a gpurun_out/r2_ab.log
    ;;
esac
