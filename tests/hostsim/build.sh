#!/bin/sh
# Builds the host-side logic check (test infrastructure, see hostsim.cpp).
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -ffp-contract=off -DPFX_NO_FMA -fPIC -shared -x c++ hostsim.cpp -o _hostsim.so
