set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v DEBUG | tail -4 > gpurun_out/r2_pytest_full.log; cat gpurun_out/r2_pytest_full.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python bench.py --steps 20 --warmup 5 --config c2-bf16 > gpurun_out/r2_bench_c2bf16.json 2> gpurun_out/r2_bench_c2bf16.err
python bench.py --steps 10 --warmup 3 --config c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
python bench.py --steps 100 --warmup 5 --config c1 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
for f in c2 c2bf16 c3 c1; do cut -c1-200 gpurun_out/r2_bench_$f.json; tail -2 gpurun_out/r2_bench_$f.err; done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_launches_c2.csv python tools/gpu_one_step.py > /dev/null 2>&1
CURV_DTYPE=bf16 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_launches_c2bf16.csv python tools/gpu_one_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_hs" -c 14 -o gpurun_out/r2_hs_fp32 python tools/gpu_one_step.py > /dev/null 2>&1
CURV_DTYPE=bf16 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_hs" -c 14 -o gpurun_out/r2_hs_bf16 python tools/gpu_one_step.py > /dev/null 2>&1
bash tools/ncu_summary.sh gpurun_out/r2_hs_fp32.ncu-rep > gpurun_out/r2_hs_fp32_ncu_full.txt
bash tools/ncu_summary.sh gpurun_out/r2_hs_bf16.ncu-rep > gpurun_out/r2_hs_bf16_ncu_full.txt
rm -f gpurun_out/r2_hs_fp32.ncu-rep gpurun_out/r2_hs_bf16.ncu-rep
CURV_DTYPE=bf16 timeout 300 python tools/gpu_step_profile.py > gpurun_out/r2_step_profile_bf16.txt 2>&1 || true
timeout 300 python tools/gpu_step_profile.py > gpurun_out/r2_step_profile_fp32.txt 2>&1
ls -la gpurun_out | tail -12
