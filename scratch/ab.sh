#!/bin/bash
# A/B: old vs new library on the same box. usage: scratch/ab.sh <scale> <hub>
for v in old new; do
  echo "=== $v"
  PPRB200_LIB=$PWD/scratch/libppr_$v.so timeout 200 python scratch/g4.py $1 $2 2>&1 | tail -7 | cut -c1-400
done
