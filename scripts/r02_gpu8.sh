#!/bin/bash
# Round-2 multi-GPU measurement batch (run under `gpurun --gpus 8`): the NCCL test of ShardedCore, BASELINE.json's C4
# (65,536 envs = 8192 per GPU, with and without the observation all-gather) and the C5 sweep at 2 / 4 / 8 GPUs.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/gpus8.txt
(CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_sharding.py -m gpu -q > $O/pytest_nccl.log 2>&1; echo "rc=$?" >> $O/pytest_nccl.log)
tail -3 $O/pytest_nccl.log
P=29611
for tag in "c4 " "c4_gather_f32 --gather" "c4_gather_f16 --gather --obs-dtype float16"; do
  set -- $tag; name=$1; shift
  P=$((P+1))
  NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/nccl_$name.%h.%p.log timeout 600 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
  tail -1 $O/bench_$name.json | cut -c1-200
done
P=$((P+1)); timeout 600 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --steps 300 --warmup 20 > $O/bench_n8.json 2> $O/bench_n8.err
P=$((P+1)); SWEEP_OUT=$O/sweep_8gpu.json timeout 900 $TR --nproc-per-node 8 --master-port $P scripts/sweep.py > $O/sweep_8gpu.log 2>&1
P=$((P+1)); SWEEP_N=4096,65536,262144 SWEEP_R=16,128,512 SWEEP_OUT=$O/sweep_4gpu.json CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port $P scripts/sweep.py > $O/sweep_4gpu.log 2>&1
P=$((P+1)); SWEEP_N=4096,65536,262144 SWEEP_R=16,128,512 SWEEP_OUT=$O/sweep_2gpu.json CUDA_VISIBLE_DEVICES=0,1 timeout 600 $TR --nproc-per-node 2 --master-port $P scripts/sweep.py > $O/sweep_2gpu.log 2>&1
tail -2 $O/sweep_8gpu.log $O/sweep_4gpu.log $O/sweep_2gpu.log
rm -f $O/nccl_*.log.keep; ls $O | wc -l
