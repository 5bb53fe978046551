#!/bin/bash
# A/B of the decode-step knobs on one box
cd "$(dirname "$0")/.."
V=$PWD/groma_b200/lib/variants
run() { env "$@" timeout 200 python tools/decode_step.py 2>&1 | grep -E "decode step|only graph|splits|Error|error" | sed "s|$V/||" ; }
run X=1
run GROMA_GEMM_CTAS_PER_SM=2 GROMA_B200_LIB=$V/libgroma_s4.so
run GROMA_GEMM_CTAS_PER_SM=2 GROMA_B200_LIB=$V/libgroma_s3.so
run GROMA_GEMM_CTAS_PER_SM=3 GROMA_B200_LIB=$V/libgroma_s3.so
run GROMA_B200_LIB=$V/libgroma_s6.so
run GROMA_GEMM_CTAS_PER_SM=2 GROMA_B200_LIB=$V/libgroma_s4.so GROMA_DEC_ATTN_TMA=0
run GROMA_GEMM_CTAS_PER_SM=2 GROMA_B200_LIB=$V/libgroma_s4.so GROMA_DECODE_TILED=1
