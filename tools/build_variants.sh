#!/bin/bash
# Builds experiment variants of the ERI kernels next to the library:
#   myqc_b200/csrc/variants/libmyqc_eri_<tag>.so   (load with MYQC_LIB=<path> python bench.py ...)
# usage: tools/build_variants.sh tag "-DFLAG=.. -DFLAG=.." [tag flags ...]
set -e
cd "$(dirname "$0")/../myqc_b200/csrc"
mkdir -p variants
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off"
make -s all
while [ $# -ge 2 ]; do
  tag=$1; defs=$2; shift 2
  $NVCC $FLAGS $defs -c eri_kernels.cu -o variants/eri_kernels_$tag.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmyqc_eri_$tag.so variants/eri_kernels_$tag.o \
     eri_api.o fock.o int1e.o ao2mo.o pairs.o qcio.o parse.o hostmem.o -cudart static -lpthread
  rm -f variants/eri_kernels_$tag.o
  echo "built variants/libmyqc_eri_$tag.so ($defs)"
done
