/* The functional scenario of csrc/tests/mlsl_functional_test.cpp written against the C binding (the reference ships
 * the same test three times - C++, C, Python; tests/examples/mlsl_test/).
 *
 * Two OT_CC layers (128 -> 256 -> 256 feature maps of 12x12, 3x3 kernels, global minibatch 16), 2 x 3 iterations,
 * index-valued tensors: every exchanged element has a closed-form expected value (see the C++ file for the table).
 *   cmlsl_functional_test <num_groups> [dist_update=0] [use_test=0]        (run under bin/mlslrun -n N) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlsl.h"

#define CHECK(call)                                                                        \
  do {                                                                                     \
    if ((call) != CMLSL_SUCCESS) {                                                         \
      printf("[%zu] FAILED CALL: %s (%s)\n", g_rank, #call, mlsl_last_error());            \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

#define LAYERS 2
static size_t g_rank, g_world;
static long g_passed, g_failed;
static int g_dist_update, g_use_test;
static mlsl_environment g_env;

typedef struct {
  size_t idx;
  mlsl_operation op;
  mlsl_activation in_act, out_act;
  mlsl_parameter_set ps;
  float *in, *in_grad, *out, *out_grad, *w, *dw;
  int got_out_grad;
} layer_t;

static void verdict(int ok, const char* what, size_t layer) {
  if (ok) ++g_passed;
  else ++g_failed;
  printf("[%zu] %s_%zu: %s\n", g_rank, what, layer, ok ? "PASSED" : "FAILED");
}

static float* alloc_f(size_t n) {
  float* p;
  CHECK(mlsl_environment_alloc(g_env, (n ? n : 1) * sizeof(float), 64, (void**)&p));
  memset(p, 0, (n ? n : 1) * sizeof(float));
  return p;
}

/* copy between the packed communication buffer and the [minibatch][feature map][pixel] tensor, block by block */
static void move_blocks(mlsl_activation act, float* comm, float* local, int unpack) {
  size_t lfm, nb, b;
  CHECK(mlsl_activation_get_local_fm_count(act, &lfm));
  if (unpack) CHECK(mlsl_activation_get_unpack_block_count(act, &nb));
  else CHECK(mlsl_activation_get_pack_block_count(act, &nb));
  for (b = 0; b < nb; ++b) {
    mlsl_comm_block_info bi;
    size_t fs, fc, fo, mo, mc, bo, mb, fm;
    if (unpack) CHECK(mlsl_activation_get_unpack_block(act, b, &bi));
    else CHECK(mlsl_activation_get_pack_block(act, b, &bi));
    CHECK(mlsl_comm_block_info_get_fm_size(bi, &fs));
    CHECK(mlsl_comm_block_info_get_fm_count(bi, &fc));
    CHECK(mlsl_comm_block_info_get_fm_offset(bi, &fo));
    CHECK(mlsl_comm_block_info_get_mb_offset(bi, &mo));
    CHECK(mlsl_comm_block_info_get_mb_count(bi, &mc));
    CHECK(mlsl_comm_block_info_get_buf_offset(bi, &bo));
    for (mb = 0; mb < mc; ++mb)
      for (fm = 0; fm < fc; ++fm) {
        float* l = local + ((mb + mo) * lfm + fm + fo) * fs;
        float* q = comm + bo + (mb * fc + fm) * fs;
        if (unpack) memcpy(l, q, fs * sizeof(float));
        else memcpy(q, l, fs * sizeof(float));
      }
  }
}

static size_t model_size(layer_t* L) {
  mlsl_distribution d;
  size_t m;
  CHECK(mlsl_operation_get_distribution(L->op, &d));
  CHECK(mlsl_distribution_get_process_count(d, GT_MODEL, &m));
  return m;
}

static void start_act(mlsl_activation act, float* local) {
  void* comm;
  CHECK(mlsl_activation_get_comm_buf(act, &comm));
  if (comm) {
    move_blocks(act, (float*)comm, local, 0);
    CHECK(mlsl_activation_start_comm(act, comm));
  } else {
    CHECK(mlsl_activation_start_comm(act, local));
  }
}

static void forward(layer_t* L) {
  void* got;
  size_t lmb, lfm, fs, off, i, np, lk, ks, bad = 0;
  CHECK(mlsl_activation_wait_comm(L->in_act, &got));
  if (got) move_blocks(L->in_act, (float*)got, L->in, 1);
  CHECK(mlsl_parameter_set_wait_increment_comm(L->ps, &got));
  CHECK(mlsl_operation_get_local_minibatch_size(L->op, &lmb));
  if (L->idx == 0) {
    CHECK(mlsl_activation_get_local_fm_count(L->out_act, &lfm));
    CHECK(mlsl_activation_get_fm_size(L->out_act, &fs));
    for (i = 0; i < lfm * lmb * fs; ++i) L->out[i] = (float)i;
  } else {
    size_t M = model_size(L), mb, fm, s;
    CHECK(mlsl_activation_get_local_fm_count(L->in_act, &lfm));
    CHECK(mlsl_activation_get_fm_size(L->in_act, &fs));
    CHECK(mlsl_activation_get_global_fm_offset(L->in_act, &off));
    for (mb = 0; mb < lmb; ++mb)
      for (fm = 0; fm < lfm; ++fm)
        for (s = 0; s < fs; ++s) {
          float want = (float)(M * (mb * lfm * fs * M + (off + fm) * fs + s));
          if (fabsf(L->in[(mb * lfm + fm) * fs + s] - want) > 1e-4f) ++bad;
        }
    verdict(bad == 0, "forward_input", L->idx);
  }
  CHECK(mlsl_parameter_set_get_local_kernel_count(L->ps, &lk));
  CHECK(mlsl_parameter_set_get_kernel_size(L->ps, &ks));
  np = lk * ks;
  bad = 0;
  for (i = 0; i < np; ++i)
    if (fabsf(L->w[i] - (float)i) > 1e-4f) ++bad;
  verdict(bad == 0, "forward_param", L->idx);
  start_act(L->out_act, L->out);
  L->got_out_grad = 0;
}

static void fetch_out_grad(layer_t* L) {
  void* got;
  if (L->got_out_grad) return;
  CHECK(mlsl_activation_wait_comm(L->out_act, &got));
  if (got) move_blocks(L->out_act, (float*)got, L->out_grad, 1);
  L->got_out_grad = 1;
}

static void backward(layer_t* L) {
  size_t lmb, lfm, fs, off, i, lk, ks;
  fetch_out_grad(L);
  CHECK(mlsl_operation_get_local_minibatch_size(L->op, &lmb));
  if (L->idx == 0) {
    size_t nub;
    CHECK(mlsl_activation_get_unpack_block_count(L->out_act, &nub));
    if (nub > 0) {
      size_t bad = 0;
      CHECK(mlsl_activation_get_local_fm_count(L->out_act, &lfm));
      CHECK(mlsl_activation_get_fm_size(L->out_act, &fs));
      for (i = 0; i < lfm * lmb * fs; ++i)
        if (fabsf(L->out_grad[i] - (float)i) > 1e-4f) ++bad;
      verdict(bad == 0, "backward_outgrad", L->idx);
    }
  } else {
    size_t M = model_size(L), mb, fm, s;
    CHECK(mlsl_activation_get_local_fm_count(L->in_act, &lfm));
    CHECK(mlsl_activation_get_fm_size(L->in_act, &fs));
    CHECK(mlsl_activation_get_global_fm_offset(L->in_act, &off));
    for (mb = 0; mb < lmb; ++mb)
      for (fm = 0; fm < lfm; ++fm)
        for (s = 0; s < fs; ++s) L->in_grad[(mb * lfm + fm) * fs + s] = (float)(mb * lfm * fs * M + (off + fm) * fs + s);
  }
  start_act(L->in_act, L->in_grad);
  CHECK(mlsl_parameter_set_get_local_kernel_count(L->ps, &lk));
  CHECK(mlsl_parameter_set_get_kernel_size(L->ps, &ks));
  for (i = 0; i < lk * ks; ++i) L->dw[i] = (float)i;
  CHECK(mlsl_parameter_set_start_gradient_comm(L->ps, L->dw));
}

static void update(layer_t* L) {
  void* ret = NULL;
  mlsl_distribution d;
  size_t D, oo, oc, ks, i, bad = 0;
  float* g;
  if (g_use_test) {
    int done = 0;
    while (!done) CHECK(mlsl_parameter_set_test_gradient_comm(L->ps, &done, &ret));
  } else {
    CHECK(mlsl_parameter_set_wait_gradient_comm(L->ps, &ret));
  }
  g = ret ? (float*)ret : L->dw;
  CHECK(mlsl_operation_get_distribution(L->op, &d));
  CHECK(mlsl_distribution_get_process_count(d, GT_DATA, &D));
  CHECK(mlsl_parameter_set_get_owned_kernel_offset(L->ps, &oo));
  CHECK(mlsl_parameter_set_get_owned_kernel_count(L->ps, &oc));
  CHECK(mlsl_parameter_set_get_kernel_size(L->ps, &ks));
  for (i = 0; i < oc * ks; ++i) {
    if (fabsf(g[i] - (float)(D * (oo * ks + i))) > 1e-4f) ++bad;
    L->w[oo * ks + i] = (float)(oo * ks + i);
  }
  verdict(bad == 0, "update_grad", L->idx);
  CHECK(mlsl_parameter_set_start_increment_comm(L->ps, L->w));
}

int main(int argc, char** argv) {
  static const size_t shape[LAYERS][3] = {{128, 256, 12}, {256, 256, 12}}; /* ifm, ofm, map width */
  layer_t layers[LAYERS];
  mlsl_session session;
  mlsl_distribution dist;
  mlsl_statistics stats;
  size_t model_parts, l, it;
  void* drain;
  if (argc < 2) {
    printf("usage: %s <num_groups> [dist_update] [use_test]\n", argv[0]);
    return 2;
  }
  model_parts = (size_t)atoi(argv[1]);
  g_dist_update = argc > 2 ? atoi(argv[2]) : 0;
  g_use_test = argc > 3 ? atoi(argv[3]) : 0;
  CHECK(mlsl_environment_get_env(&g_env));
  CHECK(mlsl_environment_init(g_env, &argc, &argv));
  CHECK(mlsl_environment_get_process_idx(g_env, &g_rank));
  CHECK(mlsl_environment_get_process_count(g_env, &g_world));
  if (model_parts < 1) model_parts = 1;
  if (model_parts > g_world) model_parts = g_world;
  if (g_world % model_parts) {
    if (g_rank == 0) printf("world size %zu not divisible by num_groups %zu\n", g_world, model_parts);
    CHECK(mlsl_environment_finalize(g_env));
    return 2;
  }
  CHECK(mlsl_environment_create_session(g_env, PT_TRAIN, &session));
  CHECK(mlsl_session_set_global_minibatch_size(session, 16));
  CHECK(mlsl_environment_create_distribution(g_env, g_world / model_parts, model_parts, &dist));
  for (l = 0; l < LAYERS; ++l) {
    mlsl_operation_reg_info ri;
    size_t op_idx;
    char name[32];
    snprintf(name, sizeof(name), "layer_%zu", l);
    CHECK(mlsl_session_create_operation_reg_info(session, OT_CC, &ri));
    CHECK(mlsl_operation_reg_info_set_name(ri, name));
    CHECK(mlsl_operation_reg_info_add_input(ri, shape[l][0], shape[l][2] * shape[l][2], DT_FLOAT));
    CHECK(mlsl_operation_reg_info_add_output(ri, shape[l][1], shape[l][2] * shape[l][2], DT_FLOAT));
    CHECK(mlsl_operation_reg_info_add_parameter_set(ri, shape[l][0] * shape[l][1], 9, DT_FLOAT, g_dist_update));
    CHECK(mlsl_session_add_operation_with_distribution(session, ri, dist, &op_idx));
    CHECK(mlsl_session_delete_operation_reg_info(session, ri));
    memset(&layers[l], 0, sizeof(layer_t));
    layers[l].idx = l;
    CHECK(mlsl_session_get_operation(session, op_idx, &layers[l].op));
    if (l > 0) CHECK(mlsl_operation_set_prev(layers[l].op, layers[l - 1].op, 0, 0));
  }
  CHECK(mlsl_session_commit(session));
  for (l = 0; l < LAYERS; ++l) {
    layer_t* L = &layers[l];
    size_t lmb, ilfm, olfm, ifs, ofs, lk, ks, i;
    CHECK(mlsl_operation_get_input(L->op, 0, &L->in_act));
    CHECK(mlsl_operation_get_output(L->op, 0, &L->out_act));
    CHECK(mlsl_operation_get_parameter_set(L->op, 0, &L->ps));
    CHECK(mlsl_operation_get_local_minibatch_size(L->op, &lmb));
    CHECK(mlsl_activation_get_local_fm_count(L->in_act, &ilfm));
    CHECK(mlsl_activation_get_local_fm_count(L->out_act, &olfm));
    CHECK(mlsl_activation_get_fm_size(L->in_act, &ifs));
    CHECK(mlsl_activation_get_fm_size(L->out_act, &ofs));
    CHECK(mlsl_parameter_set_get_local_kernel_count(L->ps, &lk));
    CHECK(mlsl_parameter_set_get_kernel_size(L->ps, &ks));
    L->in = alloc_f(ilfm * lmb * ifs);
    L->in_grad = alloc_f(ilfm * lmb * ifs);
    L->out = alloc_f(olfm * lmb * ofs);
    L->out_grad = alloc_f(olfm * lmb * ofs);
    L->w = alloc_f(lk * ks);
    L->dw = alloc_f(lk * ks);
    for (i = 0; i < lk * ks; ++i) L->w[i] = (float)i;
  }
  /* no exchange between the layers: the consumer reads the producer's tensor directly */
  {
    size_t a, b;
    CHECK(mlsl_activation_get_comm_buf_size(layers[1].in_act, &a));
    CHECK(mlsl_activation_get_comm_buf_size(layers[0].out_act, &b));
    if (a == 0 && b == 0) {
      layers[1].in = layers[0].out;
      layers[0].out_grad = layers[1].in_grad;
    }
  }
  CHECK(mlsl_session_get_stats(session, &stats));
  CHECK(mlsl_statistics_start(stats));
  for (it = 0; it < 6; ++it) {
    for (l = 0; l < LAYERS; ++l) forward(&layers[l]);
    for (l = LAYERS; l-- > 0;) backward(&layers[l]);
    for (l = 0; l < LAYERS; ++l) update(&layers[l]);
  }
  for (l = 0; l < LAYERS; ++l) {
    CHECK(mlsl_parameter_set_wait_increment_comm(layers[l].ps, &drain));
    CHECK(mlsl_activation_wait_comm(layers[l].in_act, &drain));
  }
  CHECK(mlsl_statistics_stop(stats));
  CHECK(mlsl_environment_delete_session(g_env, session));
  CHECK(mlsl_environment_delete_distribution(g_env, dist));
  CHECK(mlsl_environment_finalize(g_env));
  printf("[%zu] summary: %ld PASSED, %ld FAILED\n", g_rank, g_passed, g_failed);
  return g_failed ? 1 : 0;
}
