#!/bin/sh
# Compiles, from the reference's own sources where they lie under /root/reference, the one part of the hot path that
# builds from a single stand-alone file: MurmurHash3 (src/simkaMin/MurmurHash3.cpp, the hash of the SimkaMin sketch,
# SimkaMinCount.hpp:247-250).  Output: oracle/_ref/libmurmur_ref.so (git-ignored, travels to the GPU box).
# It is a CHECKER for the oracle's restatement (tests/test_oracle_golden.py), never linked into the product.
# The rest of the reference (simka, simkaCount, simkaMerge, simkaMinCore) needs cmake, the vendored HDF5/boost tree and
# a cmake-generated config header, so by the task's rule it is not built here.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SIMKA_REFERENCE:-/root/reference}"
SRC="$REF/src/simkaMin/MurmurHash3.cpp"
if [ ! -f "$SRC" ]; then
    echo "build_ref.sh: $SRC not present, nothing built" >&2
    exit 0
fi
mkdir -p "$HERE/_ref"
g++ -O2 -shared -fPIC -o "$HERE/_ref/libmurmur_ref.so" "$SRC"
