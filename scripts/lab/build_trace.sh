#!/bin/bash
# lab build: the library with the symmetric decoder's clock64 trace compiled in → scripts/lab/libdance_b200_trace.so
set -e
cd "$(dirname "$0")/../.."
python -m dance_b200.build > /dev/null
OBJ=dance_b200/lib/obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr --expt-extended-lambda -Xcompiler -fPIC \
     -DB2_BUILDING -DB2_GAE_TRACE -c dance_b200/csrc/gae_sym.cu -o scripts/lab/gae_sym_trace.o
nvcc -shared -o scripts/lab/libdance_b200_trace.so $(ls $OBJ/*.o | grep -v "/gae_sym\.") scripts/lab/gae_sym_trace.o \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart_static -ldl -lrt -lpthread
ls -la scripts/lab/libdance_b200_trace.so
