#!/bin/bash
# Build libsse_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
# An object is reused only if the hash of (compiler version, flags, its source, every header) matches the one
# recorded when it was built: a stale or foreign build/ directory can never be linked.  CLEAN=1 forces a full rebuild.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXX=${CXX:-g++}
OUT=../libsse_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr -Xptxas -v -Wno-deprecated-gpu-targets"
HOST_FLAGS="-O3 -std=c++17 -fPIC -fvisibility=default"
SRCS="sse_api.cu lstm_simt.cu lstm_tc.cu lstm_cluster.cu lstm_gemm.cu search_simt.cu search_tc.cu cnn.cu util.cu train.cu gemm_tc.cu tok_prep.cu"
HOST_SRCS="tsv_io.cpp subword_tok.cpp"
HEADERS="sse_common.cuh sse_handle.cuh ../../include/sse_b200.h unicode_alnum.inc $(ls *.cuh 2>/dev/null | tr '\n' ' ')"
[ "$CLEAN" = "1" ] && rm -rf build
mkdir -p build
hdr_hash=$( (for f in $(echo $HEADERS | tr ' ' '\n' | sort -u); do cat "$f"; done) | sha256sum | cut -d' ' -f1)
nvcc_id=$($NVCC --version | tail -2 | tr '\n' ' ')
cxx_id=$($CXX --version | head -1)
pids=()
build_one() {   # $1 source, $2 compiler id, $3 command prefix
  local s=$1 base=${1%.*}
  local o=build/$base.o want
  want=$( (echo "$2 | $3 | $hdr_hash"; cat "$s") | sha256sum | cut -d' ' -f1)
  if [ -f "$o" ] && [ -f "build/$base.hash" ] && [ "$(cat build/$base.hash)" = "$want" ]; then return 0; fi
  rm -f "$o" "build/$base.hash"
  ( $3 -c "$s" -o "$o" > build/$base.log 2>&1 && echo "$want" > build/$base.hash || { cat build/$base.log; exit 1; } ) &
  pids+=($!)
}
for s in $SRCS; do [ -f "$s" ] && build_one "$s" "$nvcc_id" "$NVCC $FLAGS"; done
for s in $HOST_SRCS; do build_one "$s" "$cxx_id" "$CXX $HOST_FLAGS"; done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
[ $fail = 0 ] || { echo "build failed"; exit 1; }
objs=""
for s in $SRCS $HOST_SRCS; do [ -f "$s" ] && objs="$objs build/${s%.*}.o"; done
$NVCC -shared -o $OUT $objs -lcudart_static -ldl -lpthread -lrt
echo "built $OUT"
