#!/bin/bash
# bash variants/build_variant.sh NAME -DFLAG=... : builds variants/NAME.so with extra nvcc flags (A/B experiments, see profiles/micro/variants.sh)
NAME=$1; shift
cd "$(dirname "$0")/.."
C=rednoise_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -ffp-contract=off -shared -cudart static "$@" -o variants/$NAME.so \
  $C/rn_abi.cu $C/rn_levelsync.cu $C/rn_util.cu $C/rn_fused.cu $C/rn_cells.cu $C/rn_volume.cu $C/rn_rock.cu
