#!/bin/bash
# A/B timing of library builds (Deathmatch 4096x4x128): every build_variants/*.so through `gpu_exp.py quick`
for lib in build_variants/*.so; do
    echo -n "$lib: "
    MEGASTEP_B200_LIB=$PWD/$lib timeout 300 python scripts/gpu_exp.py quick 2>&1 | tail -1
done
