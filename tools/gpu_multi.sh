#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): equality tests + the strong/weak scaling bench lines at N ranks.
n=${1:-2}
tag=${2:-multi$n}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
echo "== multi-GPU equality tests"
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider --durations=8 > $out/pytest_multi.log 2>&1
echo "pytest exit $?"; tail -n 30 $out/pytest_multi.log
for w in 1 $n; do
  port=$((29500 + w))
  echo "== bench c5 strong, world=$w"
  if [ $w -eq 1 ]; then
    timeout 600 python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_c5_n$w.json 2> $out/bench_c5_n$w.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $w --config c5 --steps 3 --warmup 3 > $out/bench_c5_n$w.json 2> $out/bench_c5_n$w.err
  fi
  echo "exit $?"; cat $out/bench_c5_n$w.json; tail -n 3 $out/bench_c5_n$w.err
done
echo "== bench c3 weak, world=$n (+ reference arm under torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_c3_n$n.json 2> $out/bench_c3_n$n.err
echo "exit $?"; cat $out/bench_c3_n$n.json; tail -n 3 $out/bench_c3_n$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $n --impl reference --steps 2 --warmup 1 > $out/bench_ref_n$n.json 2> $out/bench_ref_n$n.err
cat $out/bench_ref_n$n.json
