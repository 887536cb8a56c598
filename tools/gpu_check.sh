#!/bin/bash
# One GPU-box visit.  Usage (under gpurun): bash tools/gpu_check.sh [tests|bench|prof|all] [tag]
mode=${1:-all}
tag=${2:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $out/gpu.txt 2>&1
if [ "$mode" = "tests" ] || [ "$mode" = "all" ]; then
  echo "== pytest gpu"
  timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > $out/pytest_gpu.log 2>&1
  rc=$?; echo "pytest exit $rc"; tail -n 60 $out/pytest_gpu.log
  if [ $rc -ne 0 ]; then
    echo "== pytest gpu, fit-side GEMMs on the 64x64 kernel (A/B)"
    B200BO_GEMM=64 B200BO_POTRF=serial timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider -x > $out/pytest_gpu_gemm64.log 2>&1
    echo "pytest(gemm64) exit $?"; tail -n 15 $out/pytest_gpu_gemm64.log
  fi
  echo "== smoke"
  timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 $out/smoke.log
fi
if [ "$mode" = "bench" ] || [ "$mode" = "all" ]; then
  echo "== bench reference"
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json
  echo "== bench"
  timeout 1200 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; cat $out/bench.json; tail -n 5 $out/bench.err
fi
if [ "$mode" = "bench" ] || [ "$mode" = "all" ] || [ "$mode" = "fit" ]; then
  echo "== LML timings (graph / trailing kernel A/B)"
  python tools/lml_time.py > $out/lml_time.json 2>&1; cat $out/lml_time.json
  echo "== fit / suggest side bench"
  timeout 900 python tools/fit_bench.py > $out/fit_bench.json 2> $out/fit_bench.err; cat $out/fit_bench.json; tail -n 3 $out/fit_bench.err
  B200BO_GEMM=64 B200BO_POTRF=serial timeout 900 python tools/fit_bench.py > $out/fit_bench_gemm64.json 2> $out/fit_bench_gemm64.err; cat $out/fit_bench_gemm64.json
fi
if [ "$mode" = "prof" ]; then
  # ncu reports are summarised ON the box (raw metric page as csv, gzip) and deleted: gpurun_out is capped at 64 MiB
  summarise() { ncu -i $1.ncu-rep --page raw --csv 2>/dev/null | gzip > $1.raw.csv.gz; rm -f $1.ncu-rep; }
  echo "== ncu launch list (default bench step, no extra legs)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/bench_under_ncu.log 2>&1
  echo "ncu exit $?"; wc -l $out/launches.csv; gzip -f $out/launches.csv
  echo "== ncu full (fp64 predict kernels, 16 and 8 warps)"
  timeout 900 ncu --set full --clock-control none -k regex:predict_acq16_kernel -s 3 -c 1 -o $out/prof_predict16 -f \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_full16.log 2>&1
  echo "ncu full exit $?"; summarise $out/prof_predict16
  B200BO_PREDICT_WARPS=8 timeout 900 ncu --set full --clock-control none -k regex:predict_acq_kernel -s 3 -c 1 -o $out/prof_predict8 -f \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_full8.log 2>&1
  echo "ncu full exit $?"; summarise $out/prof_predict8
  echo "== ncu full (fp32-mode kernel)"
  B200BO_PREDICT_IMPL=tf32 timeout 900 ncu --set full --clock-control none -k regex:predict_acq_tc4 -s 3 -c 1 -o $out/prof_predict_tc4 -f \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_full_tc4.log 2>&1
  echo "ncu tc4 exit $?"; summarise $out/prof_predict_tc4
  echo "== launch list + full captures of the fit-side kernels (one fit + one LML+gradient evaluation, N=4096; graph off)"
  B200BO_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_lml.csv python tools/lml_once.py > $out/lml_once.log 2>&1
  gzip -f $out/launches_lml.csv
  B200BO_GRAPH=0 timeout 900 ncu --set full --clock-control none -k regex:dgemm128 -s 130 -c 8 -o $out/prof_fit_gemm -f python tools/lml_once.py > $out/ncu_fit_gemm.log 2>&1
  summarise $out/prof_fit_gemm
  B200BO_GRAPH=0 timeout 900 ncu --set full --clock-control none -k regex:"potrf_diag|lml_grad|kbuild" -s 66 -c 4 -o $out/prof_fit_misc -f python tools/lml_once.py > $out/ncu_fit_misc.log 2>&1
  summarise $out/prof_fit_misc
  echo "ncu fit exit $?"
  echo "== ncu launch list of one suggest() without refit (small-batch kernels)"
  REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_suggest.csv python tools/suggest_once.py > $out/suggest_once.log 2>&1
  tail -2 $out/suggest_once.log; gzip -f $out/launches_suggest.csv
  echo "== timings without profiler: LML (graph on/off), suggest"
  python tools/lml_time.py > $out/lml_time.json 2>&1; cat $out/lml_time.json
  REPS=3 python tools/suggest_once.py > $out/suggest_time.log 2>&1; cat $out/suggest_time.log
fi
if [ "$mode" = "variants" ]; then
  for w in 8 16; do
    B200BO_PREDICT_WARPS=$w timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $out/bench_warps$w.json 2> $out/bench_warps$w.err
    echo "warps=$w exit $?"; cat $out/bench_warps$w.json
  done
fi
ls -la $out | tail -n 24
