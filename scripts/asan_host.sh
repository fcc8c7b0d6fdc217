#!/bin/bash
# Memory / UB check of the native host engines (same randomised harness as scripts/tsan_host.sh).
set -e
OUT=${TMPDIR:-/tmp}/fps_asan_host
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=all -pthread -o "$OUT" \
    tests/native/tsan_mf_main.cpp flink-parameter-server_b200/ops/csrc/fps_host.cpp
timeout 600 "$OUT"
echo "asan/ubsan: clean"
