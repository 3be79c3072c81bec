#!/bin/bash
# EXPERIMENT helper: build an alternative libdeb200.so with extra nvcc flags applied to ode_dopri.cu (the flagship kernel's
# translation unit); the other objects are the regular build's.  Usage: tools/build_alt.sh NAME -DDEB_VAR_X ...
set -e
NAME=$1; shift
ROOT=$(cd $(dirname $0)/.. && pwd)
mkdir -p $ROOT/build/alt $ROOT/build/alt_obj/$NAME
cd $ROOT/differential-equations_b200/csrc
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c -o $ROOT/build/alt_obj/$NAME/ode_dopri.o ode_dopri.cu
OBJS=$(ls $ROOT/build/obj/*.o | grep -v ode_dopri.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/build/alt/$NAME.so $OBJS $ROOT/build/alt_obj/$NAME/ode_dopri.o -ldl
echo built build/alt/$NAME.so
