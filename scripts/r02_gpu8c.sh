#!/bin/bash
# C4 with the packed observation all-gather after msb_pack_obs (one-launch packing): 8-bit, 8-bit with fewer NCCL channels, fp16.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29811
run() { name=$1; shift; P=$((P+1)); timeout 400 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 "$@" > $O/bench_$name.json 2> $O/bench_$name.err; tail -n 1 $O/bench_$name.json | cut -c1-170; }
run c4c_gather_u8 --gather --obs-dtype uint8
NCCL_MAX_NCHANNELS=8 run c4c_gather_u8_8ch --gather --obs-dtype uint8
run c4c_gather_f16 --gather --obs-dtype float16
run c4c
