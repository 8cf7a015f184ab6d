#!/bin/bash
# ASan + UBSan over the host/device-shared per-element code (csrc/xray_pyramid.h, xray_png.hpp, s2.h, s2_disk.hpp) through the
# sequential test drivers; the CUDA kernels call the same functions.  Restores the normal test libraries afterwards.
set -e
cd "$(dirname "$0")/.."
B=tests/cpu_backend/_build
mkdir -p /tmp/pcv_asan && cp $B/libtbs.so /tmp/pcv_asan/ && cp $B/libtbx.so /tmp/pcv_asan/
trap 'cp /tmp/pcv_asan/libtbs.so /tmp/pcv_asan/libtbx.so '"$B"'/' EXIT
F="-O1 -g -std=c++17 -fPIC -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer -shared"
g++ $F -o $B/libtbs.so tests/cpu_backend/s2_cpu.cpp
g++ $F -o $B/libtbx.so tests/cpu_backend/xray_pyramid_cpu.cpp -lz
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_s2.py tests/test_xray_pyramid.py -x -q -p no:cacheprovider
