for v in 2 3; do
GSTVD_CROSS_BUFS=$v python tools/timeline.py --hist 150 --out gpurun_out/tl_crossb$v.txt > /dev/null 2>&1
echo "== cross warp bufs $v (hist 150)"; python tools/stage_times.py gpurun_out/tl_crossb$v.txt; grep "decode:" gpurun_out/tl_crossb$v.txt
GSTVD_CROSS_BUFS=$v python tools/timeline.py --hist 256 --out gpurun_out/tl_crossb${v}_h256.txt > /dev/null 2>&1
echo "== cross warp bufs $v (hist 256)"; python tools/stage_times.py gpurun_out/tl_crossb${v}_h256.txt; grep "decode:" gpurun_out/tl_crossb${v}_h256.txt
done
