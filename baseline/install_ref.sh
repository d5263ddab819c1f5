#!/bin/bash
# Build the UNMODIFIED reference (intel/MLSL, /root/reference is read-only -> build in a /tmp copy) and install the
# artefacts bench.py's reference arm needs into baseline/_ref (git-ignored, travels to the GPU box with gpurun):
#   _ref/intel64/{lib/libmlsl.so*,bin/ep_server}  _ref/include/  _ref/mpirt/  _ref/bin/ref_allreduce_bench
# The pip route of the task statement does not apply: the reference ships no setup.py/pyproject.toml
# ("Directory '/root/reference' is not installable"), it is a Makefile project.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
TMP="$(mktemp -d /tmp/mlsl_ref_build.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
cd "$TMP"
make libep MLSL_MODE=process EXTRA_CFLAGS=-w > build_ep.log 2>&1
make libmlsl MLSL_MODE=process EXTRA_CFLAGS=-w > build_mlsl.log 2>&1
rm -rf "$DST"
mkdir -p "$DST/intel64/lib" "$DST/intel64/bin" "$DST/include" "$DST/bin"
cp src/process/libmlsl.so.1.0 "$DST/intel64/lib/"
ln -sf libmlsl.so.1.0 "$DST/intel64/lib/libmlsl.so.1"
ln -sf libmlsl.so.1.0 "$DST/intel64/lib/libmlsl.so"
cp eplib/ep_server "$DST/intel64/bin/"
cp include/mlsl.hpp include/mlsl.h "$DST/include/"
cp -r mpirt "$DST/mpirt"
[ -e "$DST/mpirt/lib/libmpi.so" ] || ln -sf libmpi.so.12 "$DST/mpirt/lib/libmpi.so"
# the harness is OUR source, but it only uses API that exists in the reference and is compiled against the
# reference's own header and library
g++ -O2 -std=c++11 -I"$DST/include" "$HERE/../csrc/tests/mlsl_allreduce_bench.cpp" -o "$DST/bin/ref_allreduce_bench" \
    -L"$DST/intel64/lib" -lmlsl -L"$DST/mpirt/lib" -lmpi -ldl -lrt -lpthread \
    -Wl,-rpath,'$ORIGIN/../intel64/lib' -Wl,-rpath,'$ORIGIN/../mpirt/lib'
g++ -O2 -std=c++11 -I"$DST/include" "$HERE/../csrc/tests/mlsl_sample.cpp" -o "$DST/bin/ref_mlsl_sample" \
    -L"$DST/intel64/lib" -lmlsl -L"$DST/mpirt/lib" -lmpi -ldl -lrt -lpthread \
    -Wl,-rpath,'$ORIGIN/../intel64/lib' -Wl,-rpath,'$ORIGIN/../mpirt/lib'
rm -rf "$TMP"
echo "reference installed into $DST"
