#!/bin/bash
# C4 with the 8-bit observation all-gather over copy-engine peer copies, gather warmed up with the step.
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "invalidate" 2>&1 | tail -2) &
timeout 300 $TR --nproc-per-node 8 --master-port 29931 bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 --gather --obs-dtype uint8 --gather-transport p2p > $O/bench_c4e_gather_u8_p2p.json 2> $O/bench_c4e_gather_u8_p2p.err
tail -n 1 $O/bench_c4e_gather_u8_p2p.json | cut -c1-170
timeout 300 $TR --nproc-per-node 8 --master-port 29932 bench.py --gpus 8 --envs 8192 --steps 200 --warmup 10 --gather --obs-dtype float32 --gather-transport p2p > $O/bench_c4e_gather_f32_p2p.json 2> $O/bench_c4e_gather_f32_p2p.err
tail -n 1 $O/bench_c4e_gather_f32_p2p.json | cut -c1-170
wait
