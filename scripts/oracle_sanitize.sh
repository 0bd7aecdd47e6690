#!/bin/bash
# The CPU oracle (test infrastructure) under AddressSanitizer + UndefinedBehaviorSanitizer: builds oracle/_san/libnp_oracle.so and runs
# the oracle's own CPU tests against it (known answers, golden vectors, the conjugate and scalar-noise restatements).  SURVEY 5.
set -e
cd "$(dirname "$0")/.."
make -C oracle sanitize
SAN_CXX=${SAN_CXX:-/usr/bin/g++} # the compiler whose sanitizer runtimes are installed
ASAN=$($SAN_CXX -print-file-name=libasan.so)
UBSAN=$($SAN_CXX -print-file-name=libubsan.so)
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  NP_ORACLE_LIB=$PWD/oracle/_san/libnp_oracle.so \
  python -m pytest tests/test_oracle_golden.py tests/test_oracle_alg2.py tests/test_oracle_scalarnoise.py tests/test_ref_pin.py -q -m "not gpu" -p no:cacheprovider "$@"
