set -x
mkdir -p gpurun_out
CURV_OP=hessian timeout 600 python tools/gpu_step_profile.py 2>&1 | grep -v DEBUG > gpurun_out/r2_step_profile_hessian.txt
head -34 gpurun_out/r2_step_profile_hessian.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | grep -v DEBUG | tail -8 > gpurun_out/gpu_tests_full.log
cat gpurun_out/gpu_tests_full.log
