# Round-2 final measurements, single GPU (run under gpurun). Numbers only -- nothing here runs under a profiler.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_final_gputests.txt
timeout -s KILL 600 python bench.py 2> gpurun_out/r02_final_bench.err | tail -1 > gpurun_out/r02_final_bench.json
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_final_bench_reference.json
timeout -s KILL 300 python bench.py --P 500000 --views 8 --res 1024 --phase raster --strong --no-cpu-baseline --no-gpu-reference 2>/dev/null | tail -1 > gpurun_out/r02_final_c5_n1.json
(timeout -s KILL 200 python tools/raster_timing.py --ref 2>&1 | tail -3
 timeout -s KILL 200 python tools/raster_timing.py --P 50000 --views 1 --res 1024 --ref 2>&1 | tail -3
 timeout -s KILL 200 python tools/unet_timing.py --torch --iters 20 2>&1 | tail -2
 timeout -s KILL 200 python tools/vae_timing.py --torch --iters 10 2>&1 | head -2
 timeout -s KILL 200 python tools/attn_timing.py 2>&1 | tail -8
 timeout -s KILL 200 python tools/conv_probe.py 2>&1 | tail -6
 timeout -s KILL 200 python tools/conv_probe.py narrow 2>&1 | tail -8
 timeout -s KILL 200 python tools/gemm_probe.py 2>&1 | tail -30) > gpurun_out/r02_final_timings.txt
timeout -s KILL 200 python tools/vsd_pattern_timing.py 2>/dev/null | tail -1 > gpurun_out/r02_final_c4b_vsd.json
cat gpurun_out/r02_final_gputests.txt; cut -c1-400 gpurun_out/r02_final_bench.json; cat gpurun_out/r02_final_timings.txt | head -20
