#!/bin/bash
# The Blackwell-specific instructions of the shipped kernels, from the built library: bulk (TMA) copies UBLKCP, bulk L2
# prefetch UBLKPF, mbarrier SYNCS.*, warp reductions REDUX / CREDUX, programmatic-dependent-launch ACQBULK (griddepcontrol.wait)
# and PREEXIT (launch_dependents).  Usage: scripts/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
LIB=${1:-megastep_b200/libmegastep_b200.so}
echo "# $(date -u +%F) cuobjdump -sass $LIB | grep -E 'UBLKCP|UBLKPF|SYNCS|REDUX|PREEXIT|ACQBULK'  (count x instruction, per kernel)"
for fn in _Z11view_kernelILi2ELb0ELb0EEv5KArgs _Z14physics_kernel5KArgs _Z10dyn_kernelILb0EEv5KArgs _Z17bake_table_kernel5KArgs \
          _Z10vis_kernel5KArgs _Z12table_kernel5KArgs _Z11tick_kernelILi2ELb1ELb0ELi256EEv5KArgs; do
  echo; echo "== $(echo $fn | c++filt)"
  cuobjdump -sass -fun $fn $LIB 2>/dev/null | grep -E "UBLKCP|UBLKPF|SYNCS|REDUX|PREEXIT|ACQBULK" | sed -E 's/^\s*\/\*[0-9a-f]+\*\/\s*//; s/\s*\/\*.*$//; s/\s+;/;/; s/\[[^]]*\]/[..]/g; s/U?R[0-9Z]+/r/g; s/@!?U?P[0-9] //' | sort | uniq -c | sort -rn
done
