#!/bin/sh
# oracle/_ref: the REFERENCE's own code, compiled from the sources where they lie under /root/reference.  Checker and CPU
# baseline only (tests/, bench.py's cpu_baseline / --impl reference legs); never linked into or called by the product.
#   1. libmurmur_ref.so  <- src/simkaMin/MurmurHash3.cpp (stand-alone file; checks the oracle's hash restatement)
#   2. bin/{simka,simkaCount,simkaCountProcess,simkaMerge,simkaMinCore}: the reference binaries.  GATB-core needs its
#      cmake-generated config header and its vendored HDF5 / boost trees, so this step runs the recipe SURVEY.md 8(c)
#      recorded (cmake is part of this image): an out-of-source build in a scratch directory, then only the five binaries are
#      kept.  About 4 minutes on 8 cores; skipped when the binaries exist (SIMKA_REF_REBUILD=1 forces it) or cmake is absent.
# Outputs are git-ignored (oracle/_ref/) and travel to the GPU box with the snapshot; nothing is copied from the reference.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SIMKA_REFERENCE:-/root/reference}"
SRC="$REF/src/simkaMin/MurmurHash3.cpp"
if [ ! -f "$SRC" ]; then
    echo "build_ref.sh: $SRC not present, nothing built" >&2
    exit 0
fi
mkdir -p "$HERE/_ref/bin"
g++ -O2 -shared -fPIC -o "$HERE/_ref/libmurmur_ref.so" "$SRC"
if [ "$1" = "murmur" ]; then exit 0; fi
if [ -x "$HERE/_ref/bin/simka" ] && [ -x "$HERE/_ref/bin/simkaMinCore" ] && [ -z "$SIMKA_REF_REBUILD" ]; then exit 0; fi
if ! command -v cmake >/dev/null 2>&1; then
    echo "build_ref.sh: cmake not found, the reference binaries are not built" >&2
    exit 0
fi
BUILD="${SIMKA_REF_BUILD_DIR:-/tmp/simka_ref_build}"
mkdir -p "$BUILD"
(cd "$BUILD" && cmake "$REF" -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release > cmake.log 2>&1 \
    && make -j"$(nproc)" > make.log 2>&1) || { echo "build_ref.sh: reference build failed, see $BUILD/*.log" >&2; exit 0; }
cp "$BUILD"/bin/simka "$BUILD"/bin/simkaCount "$BUILD"/bin/simkaCountProcess "$BUILD"/bin/simkaMerge "$BUILD"/bin/simkaMinCore "$HERE/_ref/bin/"
