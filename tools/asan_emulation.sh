#!/bin/bash
# AddressSanitizer over the CPU emulation build: the CUDA sources compiled by g++ (tests/simt), "device" memory = calloc'd
# host memory, so out-of-bounds and use-after-free accesses of kernels and host code alike are caught.  No GPU needed.
cd "$(dirname "$0")/.."
mkdir -p tests/simt/_build
g++ -std=c++17 -O1 -g -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -x c++ -include tests/simt/cpu_simt.h \
    -Iosmo-tetra_b200/csrc osmo-tetra_b200/csrc/tetra_b200.cu tests/simt/cpu_simt.cpp -o tests/simt/_build/libtetra_b200_simt.so || exit 1
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(g++ -print-file-name=libasan.so) \
    python -m pytest tests/test_simt.py tests/test_fuzz.py tests/test_frontend.py tests/test_gsmtap.py tests/test_punct.py -x -q -m "not gpu"
rc=$?
rm -f tests/simt/_build/libtetra_b200_simt.so      # the next test run rebuilds the plain flavour
exit $rc
