#!/bin/bash
# Builds the CPU debugging harness: the kernels of psiclass_b200/csrc compiled by g++ as a serial,
# one-lane program (tests/shim/cuda_shim.h) plus the psc_check comparison tool linked against it.
# TEST INFRASTRUCTURE: the output lives in tests/shim/_build and is never loaded by the product.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/shim/_build
g++ -std=c++17 -O2 -g -Wall -Wno-unused-function -Wno-unused-variable -Wno-unknown-pragmas -DPSC_CPU_SHIM -Itests/shim -x c++ -fPIC -shared \
    -o tests/shim/_build/libpsc_shim.so psiclass_b200/csrc/psc_api.cu
g++ -std=c++17 -O2 -g -Wall -o tests/shim/_build/psc_check_shim tools/psc_check.cpp tools/check_core.cpp oracle/psc_oracle.cpp \
    psiclass_b200/csrc/host/bam_io.cpp psiclass_b200/csrc/host/subexon_graph.cpp \
    -Ltests/shim/_build -lpsc_shim -Wl,-rpath,'$ORIGIN' -lz -lpthread
# the `classes` front end linked against the shim, so the host logic (argv, BAM decode, sharding,
# Q8, GTF writer) can be diffed against the reference's GTFs without a GPU
g++ -std=c++17 -O2 -g -Wall -o tests/shim/_build/classes_shim psiclass_b200/csrc/host/classes_main.cpp \
    psiclass_b200/csrc/host/bam_io.cpp psiclass_b200/csrc/host/subexon_graph.cpp psiclass_b200/csrc/host/gtf_out.cpp psiclass_b200/csrc/host/vote.cpp \
    -Ltests/shim/_build -lpsc_shim -Wl,-rpath,'$ORIGIN' -lz -lpthread
