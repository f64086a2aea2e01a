set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py -q 2>&1 | grep -v DEBUG | grep -a "assert\|Error\|passed\|failed" | head -12
CURV_MODEL=vit_b_16 CURV_DTYPE=bf16 CURV_B=32 CURV_K=4 timeout 600 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|arn" | head -45 > gpurun_out/c4_step_profile.txt
head -8 gpurun_out/c4_step_profile.txt
