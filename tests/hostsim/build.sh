#!/bin/sh
# Builds the host-side logic check (test infrastructure, see hostsim.cpp).
#   build.sh        -> _hostsim.so
#   build.sh asan   -> _hostsim_asan.so (AddressSanitizer); run the logic tests on it with
#                      LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
#                      PFX_HOSTSIM_SO=tests/hostsim/_hostsim_asan.so python -m pytest tests/test_hostsim_logic.py
# -DPFX_HYBRID=1: the row passes follow the arithmetic of the rounds kernel in use (eight lanes per fit, TSQR merge)
set -e
cd "$(dirname "$0")"
if [ "$1" = "asan" ]; then
  g++ -O1 -g -std=c++17 -ffp-contract=off -DPFX_NO_FMA -DPFX_HYBRID=1 -fPIC -shared -fsanitize=address -fno-omit-frame-pointer \
      -x c++ hostsim.cpp -o _hostsim_asan.so
else
  g++ -O2 -std=c++17 -ffp-contract=off -DPFX_NO_FMA -DPFX_HYBRID=1 -fPIC -shared -x c++ hostsim.cpp -o _hostsim.so
fi
