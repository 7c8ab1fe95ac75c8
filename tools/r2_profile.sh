#!/bin/bash
# ncu evidence for the round-2 default kernels (one gpurun call, one GPU).  Output: gpurun_out/r2p_* (text summaries; the .ncu-rep
# files are deleted except the small source-annotated one - gpurun copies back at most 64 MiB).  Numbers printed under a profiler
# are never bench values; per-launch times are serialised (no PDL overlap) - compare shares and utilisation.
set -u
mkdir -p gpurun_out
R="python tools/profile_round.py --hist 150"
# 1. every launch of one warm round, caches left warm between launches
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2p_launches_warm.csv $R > gpurun_out/r2p_ncu0.log 2>&1
python tools/agg_launches.py gpurun_out/r2p_launches_warm.csv > gpurun_out/r2p_launches_warm.txt
python tools/agg_launches.py gpurun_out/r2p_launches_warm.csv 340 | tail -340 > gpurun_out/r2p_launch_sequence_head.txt
# 2. --set full: one whole decode step (103 kernels from the CUDA graph), the encoder's first 70 launches, no source import
cap() {  # name skip count
  timeout 900 ncu --profile-from-start off --set full --clock-control none -s "$2" -c "$3" -f -o "gpurun_out/r2p_$1" $R > "gpurun_out/r2p_$1.log" 2>&1
  python tools/ncu_summary.py "gpurun_out/r2p_$1.ncu-rep" > "gpurun_out/r2p_$1_summary.txt" 2>&1
  rm -f "gpurun_out/r2p_$1.ncu-rep"
}
cap decode_step 533 103
cap encoder_head 0 70
# 3. the decode GEMM with the deferred-LayerNorm epilogue, source-annotated (kept)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel" -s 300 -c 3 -f \
    -o gpurun_out/r2p_decode_gemm_src $R > gpurun_out/r2p_decode_gemm_src.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_decode_gemm_src.ncu-rep > gpurun_out/r2p_decode_gemm_src_summary.txt 2>&1
du -sh gpurun_out; ls gpurun_out | head -30
