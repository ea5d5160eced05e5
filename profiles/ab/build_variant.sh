#!/bin/bash
# Build a variant of libnem_b200.so with extra -D knobs (nem_kernels.cu: JAC_SPT, JAC_SPARSE_MAX,
# FX_CLUSTER, FIXUP_HOIST, CHANGED_ROWS_VEC, ...) for a same-box A/B run.
#   profiles/ab/build_variant.sh NAME [SRC] [-DKNOB=V ...]  ->  scratch/variants/libnem_b200.NAME.so
# scratch/ is git-ignored but travels to the GPU box with gpurun.  The host objects come from the
# regular in-tree build (python -m pangenomenem_b200.build), run that first.
set -e
R=$(cd "$(dirname "$0")/../.." && pwd)
NAME=$1; shift
SRC=$R/pangenomenem_b200/csrc/nem_kernels.cu
if [ -n "$1" ] && [ "${1#-}" = "$1" ]; then SRC=$1; shift; fi
O=$R/pangenomenem_b200/_build
mkdir -p $R/scratch/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
     -I $R/include -I $R/pangenomenem_b200/csrc "$@" -c $SRC -o /tmp/k_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/scratch/variants/libnem_b200.$NAME.so \
     /tmp/k_$NAME.o $O/nem_sub_kernels.cu.o $O/nem_fit.c.o $O/nem_resample.c.o $O/nem_comm.c.o \
     $O/nem_io.c.o $O/nem_api.c.o -lm -ldl -lpthread
echo built $NAME
