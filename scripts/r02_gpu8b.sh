#!/bin/bash
# Round-2 multi-GPU batch, second part (run under `gpurun --gpus 8`): the benchmark's weak scaling at 1 / 2 / 4 / 8 GPUs
# (every rank a replica of the same shard) and C4 with the 8-bit packed observation all-gather.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29711
timeout 300 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for n in 2 4 8; do
  P=$((P+1)); timeout 400 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --steps 300 --warmup 20 > $O/scale_n$n.json 2> $O/scale_n$n.err
done
for tag in "c4b " "c4b_gather_u8 --gather --obs-dtype uint8" "c4b_gather_f16 --gather --obs-dtype float16"; do
  set -- $tag; name=$1; shift
  P=$((P+1)); timeout 600 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 "$@" > $O/bench_$name.json 2> $O/bench_$name.err
done
for f in scale_n1 scale_n2 scale_n4 scale_n8 bench_c4b bench_c4b_gather_u8 bench_c4b_gather_f16; do tail -n 1 $O/$f.json | cut -c1-160; done
