#!/bin/bash
# Round-2 single-GPU measurement batch (run under gpurun): bench lines for every BASELINE config that fits one GPU, both
# arms, the ncu launch list and full captures of the three kernels. Everything lands in gpurun_out/r2/.
O=gpurun_out/r2; mkdir -p $O
timeout 600 python bench.py --steps 1000 --warmup 20 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 400 python bench.py --impl reference --steps 100 --warmup 5 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
for w in explorer explorer-demo deathmatch-demo; do
  timeout 400 python bench.py --workload $w --steps 500 --warmup 20 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
  timeout 400 python bench.py --workload $w --impl reference --steps 50 --warmup 5 > $O/bench_ref_$w.json 2> $O/bench_ref_$w.err
done
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 90 --csv --log-file $O/bench_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph > $O/ncu_launches.log 2>&1
# (gpurun brings back at most 64 MiB: the reports are exported to CSV on the box and only view_kernel's is kept)
export_rep() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > $O/$1_mix.csv 2>/dev/null; }
for kern in view_kernel dyn_kernel physics_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 25 -c 1 -o $O/full_$kern python scripts/ab.py --steps 12 "" > $O/ncu_$kern.log 2>&1
  export_rep full_$kern
  [ $kern != view_kernel ] && rm -f $O/full_$kern.ncu-rep
done
for w in explorer deathmatch-demo; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:view_kernel -s 25 -c 1 -o $O/full_view_$w python scripts/ab.py --workload $w --steps 12 "" > $O/ncu_view_$w.log 2>&1
  export_rep full_view_$w
  rm -f $O/full_view_$w.ncu-rep
done
timeout 500 python scripts/ab.py --envs 16384 --steps 100 "" "persist=1,merge_dyn=2" "persist=1,merge_dyn=2,stages=3" "persist=1" > $O/ab_16k.jsonl 2> $O/ab_16k.err
MEGASTEP_B200_LIB=$PWD/build_variants/lb_128_6.so timeout 300 python scripts/ab.py "nch=4" "nch=2" > $O/ab_lb6.jsonl 2> $O/ab_lb6.err
du -sh $O; ls $O | wc -l
