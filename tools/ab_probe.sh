#!/bin/bash
# A/B timing of library variants (fips204_b200/build.py build_variant): tools/ab_probe.sh "<op> <ps> <lanes> <reps>" name...
# "main" = the regular library.  Output: gpurun_out/ab_<name>.txt
args=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset MLDSA_B200_LIB; else export MLDSA_B200_LIB=$PWD/fips204_b200/variants/libmldsa_b200_$v.so; fi
  python tools/op_probe.py $args > gpurun_out/ab_$v.txt 2>&1
  echo "== $v"; tail -n 3 gpurun_out/ab_$v.txt
done
