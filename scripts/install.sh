#!/bin/bash
# Non-interactive counterpart of the reference's install.sh: build and install into a prefix.
#   scripts/install.sh [-d <prefix>] [-c (host-only build, no nvcc)]
set -e
prefix="$HOME/mlsl_b200"
extra=""
while getopts "d:ch" o; do
  case $o in
    d) prefix="$OPTARG" ;;
    c) extra="NO_CUDA=1" ;;
    h) sed -n 2,4p "$0"; exit 0 ;;
  esac
done
root="$(cd "$(dirname "$0")/.." && pwd)"
make -C "$root" -j"$(nproc)" $extra
make -C "$root" install PREFIX="$prefix" $extra
echo "Done.  To use:  source $prefix/intel64/bin/mlslvars.sh"
