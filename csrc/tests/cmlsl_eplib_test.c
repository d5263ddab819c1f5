/* The EPLIB-style entry points (include/eplib.h) from plain C: init, the allocation family, "is this memory reachable by
 * the servers", suspend / execute, file reads executed by a progress thread, teardown.  Run under bin/mlslrun -n 2 (or alone).
 *   cmlsl_eplib_test <scratch file>   - rank 0 writes the file, every rank reads it back through EPLIB_fread & co. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eplib.h"
#include "mlsl.h"

static int failed = 0;
#define CHECK(cond, what)                                   \
  do {                                                      \
    if (!(cond)) {                                          \
      printf("FAILED: %s (line %d)\n", what, __LINE__);     \
      failed++;                                             \
    }                                                       \
  } while (0)

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/tmp/cmlsl_eplib_test.bin";
  CHECK(EPLIB_init() == 0, "EPLIB_init");
  CHECK(EPLIB_init() == 0, "EPLIB_init twice is harmless");
  mlsl_environment env;
  size_t rank = 0, world = 0;
  CHECK(mlsl_environment_get_env(&env) == 0, "get_env");
  mlsl_environment_get_process_idx(env, &rank);
  mlsl_environment_get_process_count(env, &world);

  /* ---- memory ---- */
  float* a = (float*)EPLIB_malloc(1000 * sizeof(float));
  double* z = (double*)EPLIB_calloc(64, sizeof(double));
  char* al = (char*)EPLIB_memalign(4096, 100);
  int on_stack = 0;
  CHECK(a && z && al, "allocations");
  CHECK(((size_t)al & 4095) == 0, "EPLIB_memalign alignment");
  CHECK(EPLIB_memory_is_shmem(a) == 1 && EPLIB_memory_is_shmem(al) == 1, "library memory is reachable");
  CHECK(EPLIB_memory_is_shmem(NULL) == 0, "NULL is not");
  (void)on_stack;
  int zero = 1;
  for (int i = 0; i < 64; ++i) zero &= z[i] == 0.0;
  CHECK(zero, "EPLIB_calloc clears");
  for (int i = 0; i < 1000; ++i) a[i] = (float)i;
  a = (float*)EPLIB_realloc(a, 5000 * sizeof(float));
  int kept = a != NULL;
  for (int i = 0; kept && i < 1000; ++i) kept &= a[i] == (float)i;
  CHECK(kept, "EPLIB_realloc keeps the contents");
  CHECK(EPLIB_realloc(NULL, 16) != NULL, "realloc(NULL) allocates");

  /* ---- a collective on that memory, with the servers parked and released in between ---- */
  mlsl_distribution dist;
  CHECK(mlsl_environment_create_distribution(env, world, 1, &dist) == 0, "create_distribution");
  for (int i = 0; i < 5000; ++i) a[i] = (float)(rank + 1);
  EPLIB_suspend();
  EPLIB_execute();
  mlsl_comm_req req;
  CHECK(mlsl_distribution_all_reduce(dist, a, a, 5000, DT_FLOAT, RT_SUM, GT_DATA, &req) == 0, "all_reduce");
  CHECK(mlsl_environment_wait(env, req) == 0, "wait");
  CHECK(a[0] == (float)(world * (world + 1) / 2) && a[4999] == a[0], "all-reduce result");

  /* ---- file reads on a progress thread ---- */
  enum { N = 100000 };
  if (rank == 0) {
    FILE* f = fopen(path, "wb");
    for (int i = 0; i < N; ++i) {
      int v = i * 3;
      fwrite(&v, sizeof(v), 1, f);
    }
    fclose(f);
  }
  CHECK(mlsl_distribution_barrier(dist, GT_DATA) == 0, "barrier");
  int* buf = (int*)EPLIB_malloc(N * sizeof(int));
  CHECK(EPLIB_fopen(0, path, "w") == NULL, "write mode is refused");
  EPLIB_FILE s = EPLIB_fopen(0, path, "rb");
  CHECK(s != NULL, "EPLIB_fopen");
  size_t got = EPLIB_fread(0, buf, sizeof(int), 1000, s);                  /* blocking: items 0 .. 999 */
  CHECK(got == 1000 && buf[999] == 999 * 3, "EPLIB_fread");
  EPLIB_Request r[2];
  size_t cnt[2];
  EPLIB_fread_nb(1, buf + 1000, sizeof(int), 9000, s, &r[0]);             /* the stream goes on where it stopped */
  EPLIB_fread_nb(1, buf + 10000, sizeof(int), N, s, &r[1]);               /* asks for more than is left */
  CHECK(EPLIB_fwaitall(2, r, cnt) == 0, "EPLIB_fwaitall");
  CHECK(cnt[0] == 9000 && cnt[1] == N - 10000, "item counts, short read at the end of the file");
  int ok = 1;
  for (int i = 0; i < N; ++i) ok &= buf[i] == i * 3;
  CHECK(ok, "contents");
  CHECK(EPLIB_fclose(0, s) == 0, "EPLIB_fclose");
  memset(buf, 0, N * sizeof(int));
  EPLIB_Request one;
  size_t n1 = 0;
  EPLIB_forc_nb(0, path, "rb", buf, sizeof(int), 500, &one);              /* open + read + close in one command */
  CHECK(EPLIB_fwait(&one, &n1) == 0 && n1 == 500 && buf[499] == 499 * 3, "EPLIB_forc_nb");
  CHECK(EPLIB_fwait(&one, &n1) != 0, "a request can be waited for once");

  CHECK(mlsl_distribution_barrier(dist, GT_DATA) == 0, "barrier");
  if (rank == 0) remove(path);
  EPLIB_free(buf);
  EPLIB_free(a);
  EPLIB_free(z);
  EPLIB_free(al);
  EPLIB_free(NULL);
  mlsl_environment_delete_distribution(env, dist);
  CHECK(EPLIB_finalize() == 0, "EPLIB_finalize");
  CHECK(EPLIB_finalize() == 0, "EPLIB_finalize twice is harmless");
  printf("[%zu] eplib entry points: %s\n", rank, failed ? "FAILED" : "PASSED");
  return failed ? 1 : 0;
}
