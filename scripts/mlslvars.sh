#!/bin/bash
# source scripts/mlslvars.sh   (the reference ships scripts/mlslvars.sh [process|thread]; "process" = background
# progress servers, "thread" = inline launches)
MLSL_B200_ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export MLSL_ROOT="$MLSL_B200_ROOT"
export PATH="$MLSL_B200_ROOT/bin:$PATH"
export LD_LIBRARY_PATH="$MLSL_B200_ROOT/mlsl_b200/lib:${LD_LIBRARY_PATH}"
export PYTHONPATH="$MLSL_B200_ROOT:${PYTHONPATH}"
export CPATH="$MLSL_B200_ROOT/include:${CPATH}"
case "${1:-thread}" in
  process) export MLSL_NUM_SERVERS="${MLSL_NUM_SERVERS:-1}" ;;
  thread) export MLSL_NUM_SERVERS="${MLSL_NUM_SERVERS:-0}" ;;
esac
