set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_step_profile.py 2>&1 | grep -v DEBUG | head -16 > gpurun_out/step_fin.txt
cat gpurun_out/step_fin.txt
timeout 1500 python -m pytest tests/test_gpu_curvature.py tests/test_gpu_fullsize.py tests/test_gpu_bf16.py -x -q 2>&1 | grep -v DEBUG | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-420
