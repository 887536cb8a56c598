#!/bin/bash
# One GPU-box visit: tests, smoke, microbench, bench, ncu launch list (+ optional full capture).
# Usage (under gpurun): bash tools/gpu_check.sh [quick|full]
mode=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" 
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
echo "== microbench"
timeout 120 tools/_bin/microbench > gpurun_out/microbench.json 2>&1; cat gpurun_out/microbench.json
if [ "$mode" = "quick" ]; then exit 0; fi
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "== bench (dfma variant)"
B200BO_PREDICT_IMPL=dfma timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dfma.json 2> gpurun_out/bench_dfma.err; cat gpurun_out/bench_dfma.json; tail -n 5 gpurun_out/bench_dfma.err

echo "== fit / suggest side bench"
timeout 900 python tools/fit_bench.py > gpurun_out/fit_bench.json 2> gpurun_out/fit_bench.err; cat gpurun_out/fit_bench.json; tail -n 3 gpurun_out/fit_bench.err
echo "== bench reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ "$mode" = "bench" ]; then exit 0; fi
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-mode > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
echo "== ncu full (predict kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_acq -s 3 -c 1 -o gpurun_out/prof_predict -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-mode > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
echo "== ncu full (tcgen05 fp32-mode kernel)"
B200BO_PREDICT_IMPL=tf32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_acq_tc -s 3 -c 1 -o gpurun_out/prof_predict_tc -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-mode > gpurun_out/ncu_full_tc.log 2>&1
echo "ncu tc exit $?"; ls -la gpurun_out | tail -n 24
