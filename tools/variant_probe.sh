#!/bin/bash
# Dev probe: steady-state chain-steps/s of several builds of libpetmh.so (build/variants/*.so, selected through
# PETMH_LIB) on the same tiled golden workload, tuned chains.  Usage: tools/variant_probe.sh [S C SWEEPS WARM]
S=${1:-9472}; C=${2:-16}; SW=${3:-100}; WARM=${4:-1000}
for f in build/variants/*.so; do
  echo "== $f"
  PETMH_LIB=$PWD/$f timeout 300 python tools/perf_probe.py $S $C $SW $WARM 2>&1 | tail -3
done
