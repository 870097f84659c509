#!/bin/bash
# ncu full capture of the integrate kernel for the default library and variant libraries: gpu_ncu2.sh skip name...
set -u
skip=$1; shift
for name in "$@"; do
  lib=$PWD/memilio_b200/csrc/variants/$name/libmio_cuda.so
  [ "$name" = default ] && lib=$PWD/memilio_b200/csrc/libmio_cuda.so
  MIO_CUDA_LIB=$lib bash tools/gpu_ncu1.sh integrate_nodes $skip integ_$name | tail -1
done
