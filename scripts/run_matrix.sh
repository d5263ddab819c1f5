#!/bin/bash
# The reference's `make testing`: the functional scenario over its whole matrix on 4 ranks of this machine
# (group_count {1,2,4} x dist_update {0,1} x user_buf {0,1} x use_test {0,1}, quantisation built in and through the sample
# plug-in, the C and the Python twins, and the same over the TCP net backend with two launchers playing two nodes).
# Prints one line per run, "Run FAILED." on any failure (grep FAILED, like the reference), exit code = number of failures.
cd "$(dirname "$0")/.."
export MLSL_BACKEND=${MLSL_BACKEND:-host} MLSL_HEAP_SIZE_GB=${MLSL_HEAP_SIZE_GB:-0.25} MLSL_WATCHDOG_SEC=${MLSL_WATCHDOG_SEC:-60}
fails=0
run() {
  local name="$1"; shift
  if out=$("$@" 2>&1) && ! grep -q ": FAILED" <<<"$out"; then echo "ok      $name"; else echo "Run FAILED. $name"; fails=$((fails + 1)); fi
}
for g in 1 2 4; do for du in 0 1; do for ub in 0 1; do for ut in 0 1; do
  run "c++ groups=$g dist_update=$du user_buf=$ub use_test=$ut" bin/mlslrun -n 4 --timeout 120 bin/mlsl_functional_test $g $du $ub $ut
done; done; done; done
for g in 1 2; do
  run "c++ groups=$g quantised (built-in fp8 blocks)" bin/mlslrun -n 4 --timeout 120 bin/mlsl_functional_test $g 0 0 0 1
  run "c++ groups=$g quantised (sample plug-in)" env MLSL_TEST_QUANT_LIB=$PWD/bin/libmlsl_quant_sample.so bin/mlslrun -n 4 --timeout 120 bin/mlsl_functional_test $g 0 0 0 1
done
for g in 1 2 4; do for du in 0 1; do
  run "c   groups=$g dist_update=$du" bin/mlslrun -n 4 --timeout 120 bin/cmlsl_functional_test $g $du
  run "py  groups=$g dist_update=$du" bin/mlslrun -n 4 --timeout 120 python examples/mlsl_test.py $g $du
done; done
run "c   smoke + samples" bash -c "bin/mlslrun -n 3 bin/cmlsl_smoke_test && bin/mlslrun -n 2 bin/mlsl_sample && bin/mlslrun -n 4 bin/mlsl_example && bin/mlslrun -n 2 bin/cmlsl_eplib_test /tmp/mlsl_matrix_eplib.bin"
port=$((20000 + RANDOM % 20000))
for g in 1 2 4; do
  run "net groups=$g dist_update=1 (2 nodes x 2 ranks over TCP)" bash -c "unset MLSL_BACKEND; (bin/mlslrun -n 2 --nnodes 2 --node-rank 1 --master-addr 127.0.0.1 --master-port $port --timeout 120 bin/mlsl_functional_test $g 1 > /dev/null 2>&1 &); bin/mlslrun -n 2 --nnodes 2 --node-rank 0 --master-addr 127.0.0.1 --master-port $port --timeout 120 bin/mlsl_functional_test $g 1"
  port=$((port + 1))
done
echo "$fails run(s) failed"
exit $fails
