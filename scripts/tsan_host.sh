#!/bin/bash
# Race detection for the native host engine (worker / server threads over SPSC rings): build the host
# runtime with ThreadSanitizer and run randomised configurations.  Any data race or lock-order report
# makes TSan print a WARNING and exit non-zero.
set -e
OUT=${TMPDIR:-/tmp}/fps_tsan_mf
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -o "$OUT" tests/native/tsan_mf_main.cpp \
    flink-parameter-server_b200/ops/csrc/fps_host.cpp
TSAN_OPTIONS="halt_on_error=1 exitcode=66" timeout 600 "$OUT"
echo "tsan: clean"
