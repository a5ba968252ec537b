# compute-sanitizer memcheck + racecheck on the small cases (run under gpurun). Logs -> gpurun_out/r02_sanitizer_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
RASTER='tests/test_raster_gpu.py::test_matches_reference_golden tests/test_raster_gpu.py::test_batched_views_equal_per_view_calls'
UNET='tests/test_unet_ops_gpu.py'
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest $RASTER -q -m gpu -x > gpurun_out/r02_sanitizer_${tool}_raster.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_raster.log
  timeout 1200 $CS --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest $UNET -q -m gpu -x -k "not resize" > gpurun_out/r02_sanitizer_${tool}_unet.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_${tool}_unet.log
done
tail -4 gpurun_out/r02_sanitizer_*.log
