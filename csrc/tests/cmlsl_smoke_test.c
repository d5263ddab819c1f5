/* C binding smoke test: init, every Distribution collective once (fp32), a tiny 1-layer session, finalize. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlsl.h"

#define CHECK(call)                                                             \
  do {                                                                          \
    if ((call) != CMLSL_SUCCESS) {                                              \
      printf("FAILED: %s (%s)\n", #call, mlsl_last_error());                    \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

int main(int argc, char** argv) {
  mlsl_environment env;
  size_t rank, size, i, p;
  int fails = 0, version = 0;
  CHECK(mlsl_environment_get_env(&env));
  CHECK(mlsl_environment_get_version(&version));
  CHECK(mlsl_environment_init(env, &argc, &argv));
  CHECK(mlsl_environment_get_process_idx(env, &rank));
  CHECK(mlsl_environment_get_process_count(env, &size));
  mlsl_distribution dist;
  CHECK(mlsl_environment_create_distribution(env, size, 1, &dist));
  const size_t n = 1000;
  float *a, *b;
  CHECK(mlsl_environment_alloc(env, n * size * sizeof(float), 64, (void**)&a));
  CHECK(mlsl_environment_alloc(env, n * size * sizeof(float), 64, (void**)&b));
  mlsl_comm_req req;

  /* all_reduce */
  for (i = 0; i < n; ++i) a[i] = (float)(rank + i);
  CHECK(mlsl_distribution_all_reduce(dist, a, b, n, DT_FLOAT, RT_SUM, GT_DATA, &req));
  CHECK(mlsl_environment_wait(env, req));
  for (i = 0; i < n; ++i)
    if (fabsf(b[i] - (float)(size * i + size * (size - 1) / 2)) > 1e-3f) fails++;
  /* reduce_scatter */
  for (i = 0; i < n * size; ++i) a[i] = (float)(i + rank);
  CHECK(mlsl_distribution_reduce_scatter(dist, a, b, n, DT_FLOAT, RT_SUM, GT_DATA, &req));
  CHECK(mlsl_environment_wait(env, req));
  for (i = 0; i < n; ++i)
    if (fabsf(b[i] - (float)(size * (rank * n + i) + size * (size - 1) / 2)) > 1e-2f) fails++;
  /* all_gather */
  for (i = 0; i < n; ++i) a[i] = (float)(rank * 1000 + i);
  CHECK(mlsl_distribution_all_gather(dist, a, n, b, DT_FLOAT, GT_DATA, &req));
  CHECK(mlsl_environment_wait(env, req));
  for (p = 0; p < size; ++p)
    for (i = 0; i < n; ++i)
      if (b[p * n + i] != (float)(p * 1000 + i)) fails++;
  /* bcast */
  for (i = 0; i < n; ++i) a[i] = rank == 0 ? (float)i : -1.f;
  CHECK(mlsl_distribution_bcast(dist, a, n, DT_FLOAT, 0, GT_DATA, &req));
  CHECK(mlsl_environment_wait(env, req));
  for (i = 0; i < n; ++i)
    if (a[i] != (float)i) fails++;
  /* all_to_all */
  for (p = 0; p < size; ++p)
    for (i = 0; i < n; ++i) a[p * n + i] = (float)(rank * 100 + p);
  CHECK(mlsl_distribution_all_to_all(dist, a, n, b, DT_FLOAT, GT_DATA, &req));
  CHECK(mlsl_environment_wait(env, req));
  for (p = 0; p < size; ++p)
    for (i = 0; i < n; ++i)
      if (b[p * n + i] != (float)(p * 100 + rank)) fails++;
  /* test() polling + barrier */
  for (i = 0; i < n; ++i) a[i] = 1.f;
  CHECK(mlsl_distribution_all_reduce(dist, a, a, n, DT_FLOAT, RT_MAX, GT_GLOBAL, &req));
  {
    int done = 0;
    while (!done) CHECK(mlsl_environment_test(env, req, &done));
  }
  CHECK(mlsl_distribution_barrier(dist, GT_GLOBAL));

  /* one-layer session through the C API */
  mlsl_session session;
  mlsl_operation_reg_info ri;
  mlsl_operation op;
  mlsl_parameter_set ps;
  size_t op_idx, owned, ksize;
  void* ret;
  CHECK(mlsl_environment_create_session(env, PT_TRAIN, &session));
  CHECK(mlsl_session_set_global_minibatch_size(session, 4 * size));
  CHECK(mlsl_session_create_operation_reg_info(session, OT_CC, &ri));
  CHECK(mlsl_operation_reg_info_set_name(ri, "c_layer"));
  CHECK(mlsl_operation_reg_info_add_input(ri, 8, 4, DT_FLOAT));
  CHECK(mlsl_operation_reg_info_add_output(ri, 8, 4, DT_FLOAT));
  CHECK(mlsl_operation_reg_info_add_parameter_set(ri, 64, 9, DT_FLOAT, 0));
  CHECK(mlsl_session_add_operation_with_distribution(session, ri, dist, &op_idx));
  CHECK(mlsl_session_delete_operation_reg_info(session, ri));
  CHECK(mlsl_session_get_operation(session, op_idx, &op));
  CHECK(mlsl_session_commit(session));
  CHECK(mlsl_operation_get_parameter_set(op, 0, &ps));
  CHECK(mlsl_parameter_set_get_owned_kernel_count(ps, &owned));
  CHECK(mlsl_parameter_set_get_kernel_size(ps, &ksize));
  for (i = 0; i < owned * ksize; ++i) a[i] = (float)i;
  CHECK(mlsl_parameter_set_start_gradient_comm(ps, a));
  CHECK(mlsl_parameter_set_wait_gradient_comm(ps, &ret));
  {
    float* g = ret ? (float*)ret : a;
    for (i = 0; i < owned * ksize; ++i)
      if (fabsf(g[i] - (float)(size * i)) > 1e-3f) fails++;
  }
  CHECK(mlsl_environment_delete_session(env, session));

  /* [ext] a group made by its members only: ranks 0 and size-1 exchange their group-state words (here through a byte
   * all-gather on the world - any rendezvous will do) and build a two-rank distribution; the other ranks do nothing */
  if (size >= 2) {
    unsigned long long mine[2], *all;
    mlsl_comm_req req;
    CHECK(mlsl_environment_alloc(env, 2 * sizeof(unsigned long long) * size, 64, (void**)&all));
    CHECK(mlsl_environment_get_group_state(env, &mine[0], &mine[1]));
    CHECK(mlsl_distribution_all_gather(dist, mine, sizeof(mine), all, DT_BYTE, GT_GLOBAL, &req));
    CHECK(mlsl_environment_wait(env, req));
    if (rank == 0 || rank == size - 1) {
      size_t members[2] = {0, size - 1}, cnt = 0, idx = 99;
      unsigned long long rows = all[0] | all[2 * (size - 1)];
      unsigned long long mark = all[1] > all[2 * (size - 1) + 1] ? all[1] : all[2 * (size - 1) + 1];
      mlsl_distribution pair;
      float* v;
      CHECK(mlsl_environment_create_distribution_from_ranks(env, members, 2, rows, mark, &pair));
      CHECK(mlsl_distribution_get_process_count(pair, GT_DATA, &cnt));
      CHECK(mlsl_distribution_get_process_idx(pair, GT_DATA, &idx));
      if (cnt != 2 || idx != (rank == 0 ? 0u : 1u)) fails++;
      CHECK(mlsl_environment_alloc(env, 16 * sizeof(float), 64, (void**)&v));
      for (i = 0; i < 16; ++i) v[i] = (float)(rank + 1);
      CHECK(mlsl_distribution_all_reduce(pair, v, v, 16, DT_FLOAT, RT_SUM, GT_DATA, &req));
      CHECK(mlsl_environment_wait(env, req));
      for (i = 0; i < 16; ++i)
        if (fabsf(v[i] - (float)(1 + size)) > 1e-5f) fails++;
      CHECK(mlsl_environment_free(env, v));
      CHECK(mlsl_environment_delete_distribution(env, pair));
    }
    CHECK(mlsl_environment_free(env, all));
  }
  CHECK(mlsl_environment_free(env, a));
  CHECK(mlsl_environment_free(env, b));
  CHECK(mlsl_environment_delete_distribution(env, dist));
  CHECK(mlsl_environment_finalize(env));
  printf("[%zu] cmlsl_smoke_test: %s (version %d.%d)\n", rank, fails ? "FAILED" : "PASSED", version >> 16, version & 0xffff);
  return fails ? 1 : 0;
}
