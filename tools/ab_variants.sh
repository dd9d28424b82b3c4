#!/bin/bash
# A/B of library variants in ONE gpurun call (boxes differ by a few percent): usage  bash tools/ab_variants.sh "<command>" a.so b.so ...
CMD=$1; shift
for rep in 1 2; do for v in "$@"; do echo "== $v (rep $rep)"; B2F_LIB_PATH=$PWD/opencl_fft_b200/lib/variants/$v bash -c "$CMD"; done; done
