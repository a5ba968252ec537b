# Round-1 profiling pass (run under gpurun). Reports stay small: <64 MiB comes back.
mkdir -p gpurun_out /tmp/rep
(timeout 300 python tools/gemm_table.py 8 2>&1 | tail -60) > gpurun_out/s4_gemm_table.txt
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_(render|tile_sort|preprocess|scatter|bwd_epilogue|spine)' -s 12 -c 7 -f -o gpurun_out/s4_raster python tools/raster_timing.py --iters 2 > gpurun_out/s4_raster_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flash_attn -s 96 -c 6 -f -o gpurun_out/s4_attn python tools/unet_timing.py --no-graph --iters 1 > gpurun_out/s4_attn_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 672 -c 14 -f -o /tmp/rep/s4_gemm python tools/unet_timing.py --no-graph --iters 1 --dump-shapes gpurun_out/s4_shapes.json > gpurun_out/s4_gemm_ncu.log 2>&1
ncu -i /tmp/rep/s4_gemm.ncu-rep --page raw --csv > gpurun_out/s4_gemm_raw.csv 2>/dev/null
ncu -i /tmp/rep/s4_gemm.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/s4_gemm_src_k1.csv.gz
ncu -i /tmp/rep/s4_gemm.ncu-rep --page source --csv --kernel-id :::3 2>/dev/null | gzip > gpurun_out/s4_gemm_src_k3.csv.gz
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_' -s 1500 -c 520 --csv --log-file gpurun_out/s4_unet_launches_warm.csv python tools/unet_timing.py --no-graph --iters 1 > /dev/null 2>&1
du -sh gpurun_out
