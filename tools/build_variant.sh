#!/bin/bash
# tools/build_variant.sh <out.so> <file.cu> [-DNAME=VALUE ...]: rebuild ONE translation unit with extra
# defines and link it with the other objects of the regular build into an A/B library (load with USC_LIB).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/ultrasonic-communication_b200/csrc/_build
OUT=$1; SRC=$2; shift 2; V=/tmp/variant_$(basename $OUT .so)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off \
     --expt-relaxed-constexpr -Xptxas -v "$@" -c $ROOT/ultrasonic-communication_b200/csrc/$SRC -o /tmp/variant_$(basename $OUT .so).o 2> /tmp/variant_$(basename $OUT .so).log
OBJS=$(ls $B/*.o | grep -v "/$SRC.o")
nvcc -shared -o $OUT $OBJS /tmp/variant_$(basename $OUT .so).o -Xcompiler -fPIC -cudart static -lm
grep -E "spill|Used" /tmp/variant_$(basename $OUT .so).log | sort | uniq -c | sort -rn | head -4
