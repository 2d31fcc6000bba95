#!/bin/bash
# build_variant.sh NAME "EXTRA nvcc flags": a second build of libmcgpu.so with other compile-time parameters, for A/B runs on one box
# (MC_B200_LIB=metacherchant_b200/csrc/variants/NAME/libmcgpu.so python bench.py ...).  Objects and library stay under variants/NAME.
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../../metacherchant_b200/csrc"
mkdir -p variants/$NAME
HOSTCXX=$(test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in mc_table mc_bins mc_sk mc_xchg mc_api mc_parse mc_count mc_bfs mc_unitig mc_env mc_classify; do
  /usr/local/cuda/bin/nvcc $ARCH -O3 -lineinfo -std=c++17 -ccbin $HOSTCXX -Xcompiler -fPIC $EXTRA -dc -o variants/$NAME/$f.o $f.cu &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc $ARCH -ccbin $HOSTCXX -shared -o variants/$NAME/libmcgpu.so variants/$NAME/*.o -lcudart_static -ldl -lrt -lpthread -lz
ls -la variants/$NAME/libmcgpu.so
