#!/bin/bash
# C4 with the packed observation all-gather, final forms: 8-bit rows over NCCL and over copy-engine peer copies.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29911
run() { name=$1; shift; P=$((P+1)); timeout 300 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 "$@" > $O/bench_$name.json 2> $O/bench_$name.err; tail -n 1 $O/bench_$name.json | cut -c1-170; }
run c4d_gather_u8_nccl --gather --obs-dtype uint8
run c4d_gather_u8_p2p --gather --obs-dtype uint8 --gather-transport p2p
run c4d_gather_f16_p2p --gather --obs-dtype float16 --gather-transport p2p
run c4d
