#!/bin/bash
# TEST INFRASTRUCTURE: builds the kernels for the host-side emulator (see cuda_emul.h).
set -e
cd "$(dirname "$0")/../.."
mkdir -p build
g++ -x c++ -std=c++20 -O2 -g -pthread -fPIC -shared -DDMST_EMULATE -DDMST_EMUL_IMPL \
    -Itests/emul -Wno-unused-value -Wno-attributes \
    -o build/libdiffmst_emul.so diffmst_b200/csrc/capi.cu
echo built build/libdiffmst_emul.so
