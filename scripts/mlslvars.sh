#!/bin/bash
# source <root>/scripts/mlslvars.sh [process|thread]          (source tree)
# source <prefix>/intel64/bin/mlslvars.sh [process|thread]    (after `make install PREFIX=<prefix>`)
# Same role as the reference's scripts/mlslvars.sh: MLSL_ROOT, PATH, LD_LIBRARY_PATH, PYTHONPATH (+ CPATH/LIBRARY_PATH
# so that `-lmlsl_b200` and `#include <mlsl.hpp>` just work).  "process" = background progress servers, "thread" =
# launches from the calling thread (the default on GPUs, where a launch is asynchronous anyway).
_here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
if [ -d "$_here/../../intel64/lib" ]; then          # installed layout
  MLSL_ROOT="$(cd "$_here/../.." && pwd)"
  _bin="$MLSL_ROOT/intel64/bin"; _lib="$MLSL_ROOT/intel64/lib"; _inc="$MLSL_ROOT/intel64/include"; _py="$MLSL_ROOT/python"
else                                                # source tree
  MLSL_ROOT="$(cd "$_here/.." && pwd)"
  _bin="$MLSL_ROOT/bin"; _lib="$MLSL_ROOT/mlsl_b200/lib"; _inc="$MLSL_ROOT/include"; _py="$MLSL_ROOT"
fi
export MLSL_ROOT
export PATH="$_bin:$PATH"
export LD_LIBRARY_PATH="$_lib:${LD_LIBRARY_PATH}"
export LIBRARY_PATH="$_lib:${LIBRARY_PATH}"
export CPATH="$_inc:${CPATH}"
export PYTHONPATH="$_py:${PYTHONPATH}"
case "${1:-thread}" in
  process) export MLSL_NUM_SERVERS="${MLSL_NUM_SERVERS:-1}" ;;
  thread) export MLSL_NUM_SERVERS="${MLSL_NUM_SERVERS:-0}" ;;
  *) echo "usage: source mlslvars.sh [process|thread]" ;;
esac
unset _here _bin _lib _inc _py
