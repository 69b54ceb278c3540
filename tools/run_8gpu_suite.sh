#!/bin/bash
# Round-2 multi-GPU evidence (one 8-GPU box): K8 at 4/8 ranks with NVLink counters, multi-GPU Learn parity test, the
# bench line at 4 and 8 GPUs (with the train block), BASELINE C4 at 8 GPUs, the reduced C5.  Outputs -> gpurun_out/r02_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/r02_8gpu_devices.txt
nvidia-smi topo -m >> gpurun_out/r02_8gpu_devices.txt 2>&1
for n in 4 8; do
  nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_before_$n.txt 2>&1
  timeout 240 $TR --nproc-per-node $n --master-port 2960$n tools/bench_k8.py --batch 256 --iters 10 2>gpurun_out/r02_k8_${n}gpu.err | grep '^{' > gpurun_out/r02_k8_${n}gpu.jsonl
  nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_after_$n.txt 2>&1
  AZ_TRAIN_COLLECTIVE=nccl timeout 240 $TR --nproc-per-node $n --master-port 2961$n tools/bench_k8.py --batch 256 --iters 10 --collective nccl 2>>gpurun_out/r02_k8_${n}gpu.err | grep '^{' >> gpurun_out/r02_k8_${n}gpu.jsonl
  cat gpurun_out/r02_k8_${n}gpu.jsonl
done
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "4 or 8" > gpurun_out/r02_gpu_multi_test.log 2>&1; tail -3 gpurun_out/r02_gpu_multi_test.log
for n in 8 4; do
  timeout 420 $TR --nproc-per-node $n --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_${n}gpu.json'));print($n,'gpus value',d['value'],'e2e',d['e2e']['value'],'train',d.get('train'))"
done
timeout 300 $TR --nproc-per-node 8 --master-port 29640 examples/learn_c4.py --games 4096 --sims 400 --iters 2 --nniters 2 --arena 256 2>gpurun_out/r02_c4_learn_8gpu.err | grep '^{' > gpurun_out/r02_c4_learn_8gpu.jsonl; cat gpurun_out/r02_c4_learn_8gpu.jsonl
timeout 900 $TR --nproc-per-node 8 --master-port 29650 examples/learn_go.py --games 8192 --sims 800 --iters 2 --nniters 1 --arena 1024 --max-moves 3 2>gpurun_out/r02_c5_8gpu.err | grep '^{' > gpurun_out/r02_c5_8gpu.jsonl; cat gpurun_out/r02_c5_8gpu.jsonl; tail -c 600 gpurun_out/r02_c5_8gpu.err
