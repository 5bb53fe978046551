#!/bin/bash
# A/B builds of the library with compile-time tunables changed (selected at run time with GROMA_B200_LIB=...)
#   tools/build_variants.sh name -DFOO=1 [-DBAR=2 ...]     (gemm.cu and attention.cu are recompiled with the flags)
set -e
cd "$(dirname "$0")/.."
name=$1; shift 1
python -m groma_b200.build >/dev/null
mkdir -p groma_b200/lib/variants
for stem in gemm attention; do
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -I groma_b200/csrc -I include "$@" -c groma_b200/csrc/$stem.cu -o /tmp/${stem}_$name.o 2>/dev/null &
done
wait
nvcc -shared -o groma_b200/lib/variants/libgroma_$name.so /tmp/gemm_$name.o /tmp/attention_$name.o $(ls groma_b200/lib/obj/*.o | grep -v "/gemm.o\|/attention.o") -lcudart 2>/dev/null
echo "built groma_b200/lib/variants/libgroma_$name.so"
