#!/bin/bash
# Build libsse_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libsse_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr -Xptxas -v"
SRCS="sse_api.cu lstm_simt.cu lstm_tc.cu lstm_cluster.cu search_simt.cu search_tc.cu cnn.cu util.cu train.cu"
mkdir -p build
pids=()
for s in $SRCS; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ sse_common.cuh -nt "$o" ] || [ sse_handle.cuh -nt "$o" ] || [ ../../include/sse_b200.h -nt "$o" ]; then
    ( $NVCC $FLAGS -c "$s" -o "$o" > build/${s%.cu}.log 2>&1 || { cat build/${s%.cu}.log; exit 1; } ) &
    pids+=($!)
  fi
done
CXX=${CXX:-g++}
HOST_SRCS="tsv_io.cpp subword_tok.cpp"
for s in $HOST_SRCS; do
  o=build/${s%.cpp}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ ../../include/sse_b200.h -nt "$o" ] || [ unicode_alnum.inc -nt "$o" ]; then
    ( $CXX -O3 -std=c++17 -fPIC -fvisibility=default -c "$s" -o "$o" > build/${s%.cpp}.log 2>&1 || { cat build/${s%.cpp}.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $(for s in $SRCS; do echo build/${s%.cu}.o; done) $(for s in $HOST_SRCS; do echo build/${s%.cpp}.o; done) -lcudart_static -ldl -lpthread -lrt
echo "built $OUT"
