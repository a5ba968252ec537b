# Raster-only profiling pass (run under gpurun): one ncu --set full capture of every raster kernel
# of one warm forward+backward at c2, plus the per-source-line instruction counts of the compositors.
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/rep
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_(render|tile_sort|preprocess|scatter|bwd_epilogue|spine)' -s 14 -c 7 -f -o /tmp/rep/${TAG}_raster python tools/raster_timing.py --iters 2 > gpurun_out/${TAG}_raster_ncu.log 2>&1
python tools/ncu_raw.py /tmp/rep/${TAG}_raster.ncu-rep > gpurun_out/${TAG}_raster_ncu_summary.txt 2>&1
for k in 4 5; do
ncu -i /tmp/rep/${TAG}_raster.ncu-rep --page source --csv --kernel-id :::$k 2>/dev/null | gzip > gpurun_out/${TAG}_raster_src_k$k.csv.gz
done
python tools/raster_timing.py > gpurun_out/${TAG}_raster_timing.txt 2>&1
tail -2 gpurun_out/${TAG}_raster_timing.txt
