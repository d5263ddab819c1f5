// Functional test of the Session / Operation graph layer on N ranks.
//
// Scenario (same as the reference's integration test, tests/examples/mlsl_test/mlsl_test.cpp: two OT_CC layers,
// ifm 128 -> ofm 256 -> 256, 12x12 maps, 3x3 kernels, global minibatch 16, 2 epochs x 3 minibatches) expressed
// with index-valued synthetic tensors, so every exchanged element has a closed-form expected value:
//   forward  : layer 0 emits out[i] = i; layer 1 must receive  M * (global index)   (M = model group size)
//   backward : layer 1 emits dIn[i] = global index; layer 0 must receive dOut[i] = i
//   gradients: dW[i] = i must come back as D * (ownedOffset + i)                     (D = data group size)
//   increment: after the all-gather every rank must hold W[i] = i again
// Usage:  mlsl_functional_test <num_groups> [dist_update=0] [user_buf=0] [use_test=0] [quant=0] [--inproc N]
//   num_groups = model parts (1 = data parallel, N = model parallel, between = hybrid)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/mlsl.hpp"

// in-process virtual ranks ([ext] entry points of the C binding; mlsl.h itself is not included because its
// global-scope enumerators would clash with `using namespace MLSL`)
extern "C" {
int mlsl_inproc_world_create(int nranks, int* world_id);
int mlsl_inproc_world_destroy(int world_id);
int mlsl_inproc_bind_thread(int world_id, int rank);
int mlsl_inproc_unbind_thread(void);
}

using namespace MLSL;

namespace {

struct Config {
  size_t modelParts = 1;
  bool distUpdate = false, userBuf = false, useTest = false, quant = false;
};

struct LayerShape {
  size_t ifm, ofm, inW, outW, k;
};
const LayerShape kShapes[2] = {{128, 256, 12, 12, 3}, {256, 256, 12, 12, 3}};
const size_t kGlobalMb = 16, kEpochs = 2, kMbPerEpoch = 3;

struct Net {
  Config cfg;
  size_t rank = 0, world = 1;
  long passed = 0, failed = 0;
  bool aliased = false;
  Environment* env = nullptr;

  void check(bool ok, const char* what, size_t layer) {
    if (ok) {
      ++passed;
      printf("[%zu] %s_%zu: PASSED\n", rank, what, layer);
    } else {
      ++failed;
      printf("[%zu] %s_%zu: FAILED\n", rank, what, layer);
    }
  }

  float* alloc(size_t elems) {
    size_t bytes = (elems ? elems : 1) * sizeof(float);
    float* p = cfg.userBuf ? (float*)malloc(bytes) : (float*)env->Alloc(bytes, 64);
    memset(p, 0, bytes);
    return p;
  }
  void release(float* p) {
    if (cfg.userBuf) free(p);
    else env->Free(p);
  }

  struct Layer {
    size_t idx;
    Operation* op;
    float *in, *inGrad, *out, *outGrad, *w, *dw;
    bool gotOutGrad = false;
  };
  std::vector<Layer> layers;

  static void move_blocks(Activation* a, float* comm, float* local, bool unpack) {
    size_t lfm = a->GetLocalFmCount();
    size_t nb = unpack ? a->GetUnpackBlockCount() : a->GetPackBlockCount();
    for (size_t b = 0; b < nb; ++b) {
      CommBlockInfo* bi = unpack ? a->GetUnpackBlock(b) : a->GetPackBlock(b);
      size_t fs = bi->GetFmSize(), fc = bi->GetFmCount(), fo = bi->GetFmOffset(), mo = bi->GetMbOffset();
      float* c = comm + bi->GetBufOffset();
      for (size_t mb = 0; mb < bi->GetMbCount(); ++mb)
        for (size_t fm = 0; fm < fc; ++fm) {
          float* l = local + ((mb + mo) * lfm + fm + fo) * fs;
          float* q = c + (mb * fc + fm) * fs;
          if (unpack) memcpy(l, q, fs * sizeof(float));
          else memcpy(q, l, fs * sizeof(float));
        }
    }
  }

  void forward(Layer& L) {
    Activation* ia = L.op->GetInput(0);
    float* got = (float*)ia->WaitComm();
    if (got) move_blocks(ia, got, L.in, true);
    ParameterSet* ps = L.op->GetParameterSet(0);
    ps->WaitIncrementComm();
    Activation* oa = L.op->GetOutput(0);
    size_t lmb = L.op->GetLocalMinibatchSize();
    if (L.idx == 0) {
      size_t n = oa->GetLocalFmCount() * lmb * oa->GetFmSize();
      for (size_t i = 0; i < n; ++i) L.out[i] = (float)i;
    } else {
      size_t lfm = ia->GetLocalFmCount(), fs = ia->GetFmSize(), off = ia->GetGlobalFmOffset();
      size_t M = L.op->GetDistribution()->GetProcessCount(GT_MODEL);
      size_t bad = 0;
      for (size_t mb = 0; mb < lmb; ++mb)
        for (size_t fm = 0; fm < lfm; ++fm)
          for (size_t s = 0; s < fs; ++s) {
            float want = (float)(M * (mb * lfm * fs * M + (off + fm) * fs + s));
            if (fabsf(L.in[(mb * lfm + fm) * fs + s] - want) > 1e-4f) ++bad;
          }
      check(bad == 0, "forward_input", L.idx);
    }
    size_t np = ps->GetLocalKernelCount() * ps->GetKernelSize(), bad = 0;
    for (size_t i = 0; i < np; ++i)
      if (fabsf(L.w[i] - (float)i) > 1e-4f) ++bad;
    check(bad == 0, "forward_param", L.idx);
    float* comm = (float*)oa->GetCommBuf();
    if (comm) {
      move_blocks(oa, comm, L.out, false);
      oa->StartComm(comm);
    } else {
      oa->StartComm(L.out);
    }
    L.gotOutGrad = false;
  }

  void fetch_out_grad(Layer& L) {
    if (L.gotOutGrad) return;
    Activation* oa = L.op->GetOutput(0);
    float* got = (float*)oa->WaitComm();
    if (got) move_blocks(oa, got, L.outGrad, true);
    L.gotOutGrad = true;
  }

  void backward_data(Layer& L) {
    fetch_out_grad(L);
    Activation* ia = L.op->GetInput(0);
    Activation* oa = L.op->GetOutput(0);
    size_t lmb = L.op->GetLocalMinibatchSize();
    if (L.idx == 0) {
      if (oa->GetUnpackBlockCount() > 0) {   // only meaningful when something travels backward into this layer
        size_t n = oa->GetLocalFmCount() * lmb * oa->GetFmSize(), bad = 0;
        for (size_t i = 0; i < n; ++i)
          if (fabsf(L.outGrad[i] - (float)i) > 1e-4f) ++bad;
        check(bad == 0, "backward_outgrad", L.idx);
      }
    } else {
      size_t lfm = ia->GetLocalFmCount(), fs = ia->GetFmSize(), off = ia->GetGlobalFmOffset();
      size_t M = L.op->GetDistribution()->GetProcessCount(GT_MODEL);
      for (size_t mb = 0; mb < lmb; ++mb)
        for (size_t fm = 0; fm < lfm; ++fm)
          for (size_t s = 0; s < fs; ++s)
            L.inGrad[(mb * lfm + fm) * fs + s] = (float)(mb * lfm * fs * M + (off + fm) * fs + s);
    }
    float* comm = (float*)ia->GetCommBuf();
    if (comm) {
      move_blocks(ia, comm, L.inGrad, false);
      ia->StartComm(comm);
    } else {
      ia->StartComm(L.inGrad);
    }
  }

  void backward_weights(Layer& L) {
    fetch_out_grad(L);
    ParameterSet* ps = L.op->GetParameterSet(0);
    size_t np = ps->GetLocalKernelCount() * ps->GetKernelSize();
    for (size_t i = 0; i < np; ++i) L.dw[i] = (float)i;
    ps->StartGradientComm(L.dw);
  }

  void update(Layer& L) {
    ParameterSet* ps = L.op->GetParameterSet(0);
    float* g = nullptr;
    if (cfg.useTest) {
      bool done = false;
      while (!done) g = (float*)ps->TestGradientComm(&done);
    } else {
      g = (float*)ps->WaitGradientComm();
    }
    if (!g) g = L.dw;
    size_t D = L.op->GetDistribution()->GetProcessCount(GT_DATA);
    size_t ownOff = ps->GetOwnedKernelOffset() * ps->GetKernelSize();
    size_t own = ps->GetOwnedKernelCount() * ps->GetKernelSize();
    size_t bad = 0;
    double relSum = 0, relMax = 0;
    for (size_t i = 0; i < own; ++i) {
      float want = (float)(D * (ownOff + i));
      float diff = fabsf(g[i] - want);
      if (cfg.quant) {
        double rel = want != 0 ? diff / want : diff;
        relSum += rel;
        if (rel > relMax) relMax = rel;
        // block-scaled fp8 with error feedback: the error is bounded by one quantisation step of the block maximum
        // per quantisation pass (two passes); residuals carried over from earlier iterations add a third
        if (diff > 0.2f * (float)(D * (ownOff + own))) ++bad;
      } else if (diff > 1e-4f) {
        ++bad;
      }
      L.w[ownOff + i] = (float)(ownOff + i);
    }
    if (cfg.quant && rank == 0)
      printf("[%zu] update_%zu: quantised gradient: avg rel err %.4f %%, max rel err %.4f %%\n", rank, L.idx,
             100.0 * relSum / (double)own, 100.0 * relMax);
    check(bad == 0, "update_grad", L.idx);
    ps->StartIncrementComm(L.w);
  }

  int run() {
    env = &Environment::GetEnv();
    env->Init(nullptr, nullptr);
    rank = env->GetProcessIdx();
    world = env->GetProcessCount();
    if (cfg.modelParts < 1) cfg.modelParts = 1;
    if (cfg.modelParts > world) cfg.modelParts = world;
    if (world % cfg.modelParts != 0) {
      if (rank == 0) printf("world size %zu not divisible by num_groups %zu\n", world, cfg.modelParts);
      env->Finalize();
      return 2;
    }
    if (MLSL_VERSION_LT(Environment::GetVersion(), MLSL_VERSION(MLSL_MAJOR_VERSION, MLSL_MINOR_VERSION))) {
      printf("incompatible API version\n");
      return 2;
    }
    if (cfg.quant) {
      QuantParams qp;
      memset(&qp, 0, sizeof(qp));
      // MLSL_TEST_QUANT_LIB=<path to bin/libmlsl_quant_sample.so>: go through the user plug-in interface instead of
      // the built-in fp8 block format (host backend; the CUDA backend always uses its fused kernel)
      if (const char* lib = getenv("MLSL_TEST_QUANT_LIB")) {
        qp.lib_path = (char*)lib;
        qp.quant_buffer_func_name = (char*)"sample_compress";
        qp.dequant_buffer_func_name = (char*)"sample_decompress";
        qp.reduce_sum_func_name = (char*)"sample_reduce_sum";
        qp.block_size = 268;
        qp.elem_in_block = 256;
      }
      env->SetQuantizationParams(&qp);
    }
    Session* session = env->CreateSession(PT_TRAIN);
    session->SetGlobalMinibatchSize(kGlobalMb);
    Distribution* dist = env->CreateDistribution(world / cfg.modelParts, cfg.modelParts);
    if (rank == 0)
      printf("world %zu: data parts %zu, model parts %zu, dist_update %d user_buf %d use_test %d quant %d\n", world,
             world / cfg.modelParts, cfg.modelParts, (int)cfg.distUpdate, (int)cfg.userBuf, (int)cfg.useTest, (int)cfg.quant);
    for (size_t l = 0; l < 2; ++l) {
      const LayerShape& s = kShapes[l];
      OperationRegInfo* ri = session->CreateOperationRegInfo(OT_CC);
      char nm[32];
      snprintf(nm, sizeof(nm), "layer_%zu", l);
      ri->SetName(nm);
      ri->AddInput(s.ifm, s.inW * s.inW, DT_FLOAT);
      ri->AddOutput(s.ofm, s.outW * s.outW, DT_FLOAT);
      ri->AddParameterSet(s.ifm * s.ofm, s.k * s.k, DT_FLOAT, cfg.distUpdate, cfg.quant ? CT_QUANTIZATION : CT_NONE);
      size_t oi = session->AddOperation(ri, dist);
      session->DeleteOperationRegInfo(ri);
      Layer L;
      L.idx = l;
      L.op = session->GetOperation(oi);
      if (l > 0) L.op->SetPrev(layers[l - 1].op, 0, 0);
      layers.push_back(L);
    }
    session->Commit();
    for (Layer& L : layers) {
      Activation* ia = L.op->GetInput(0);
      Activation* oa = L.op->GetOutput(0);
      ParameterSet* ps = L.op->GetParameterSet(0);
      size_t lmb = L.op->GetLocalMinibatchSize();
      size_t ni = ia->GetLocalFmCount() * lmb * ia->GetFmSize(), no = oa->GetLocalFmCount() * lmb * oa->GetFmSize();
      size_t np = ps->GetLocalKernelCount() * ps->GetKernelSize();
      L.in = alloc(ni);
      L.inGrad = alloc(ni);
      L.out = alloc(no);
      L.outGrad = alloc(no);
      L.w = alloc(np);
      L.dw = alloc(np);
      for (size_t i = 0; i < np; ++i) L.w[i] = (float)i;
    }
    // No exchange between two layers (WaitComm returns NULL): the consumer reads the producer's tensor directly.
    for (size_t l = 1; l < layers.size(); ++l)
      if (layers[l].op->GetInput(0)->GetCommBufSize() == 0 && layers[l - 1].op->GetOutput(0)->GetCommBufSize() == 0) {
        release(layers[l].in);
        release(layers[l - 1].outGrad);
        layers[l].in = layers[l - 1].out;
        layers[l - 1].outGrad = layers[l].inGrad;
        aliased = true;
      }
    Statistics* st = session->GetStats();
    st->Start();
    for (size_t e = 0; e < kEpochs; ++e)
      for (size_t it = 0; it < kMbPerEpoch; ++it) {
        for (size_t l = 0; l < layers.size(); ++l) forward(layers[l]);
        for (size_t l = layers.size(); l-- > 0;) {
          // the first layer has no consumer for its input gradient: dX first for the others so its transfer
          // overlaps the dW computation, exactly the ordering the library is designed to overlap
          if (l > 0) backward_data(layers[l]);
          else fetch_out_grad(layers[l]), backward_data(layers[l]);
          backward_weights(layers[l]);
        }
        for (size_t l = 0; l < layers.size(); ++l) update(layers[l]);
      }
    // drain the last parameter all-gathers and backward transfers before tearing down
    for (Layer& L : layers) {
      L.op->GetParameterSet(0)->WaitIncrementComm();
      L.op->GetInput(0)->WaitComm();
    }
    st->Stop();
    if (st->IsEnabled()) st->Print();
    for (Layer& L : layers) {
      if (!(aliased && L.idx > 0)) release(L.in);
      release(L.inGrad);
      release(L.out);
      if (!(aliased && L.idx + 1 < layers.size())) release(L.outGrad);
      release(L.w);
      release(L.dw);
    }
    env->DeleteSession(session);
    env->DeleteDistribution(dist);
    env->Finalize();
    printf("[%zu] summary: %ld PASSED, %ld FAILED\n", rank, passed, failed);
    return failed ? 1 : 0;
  }
};

}  // namespace

int main(int argc, char** argv) {
  Config cfg;
  int inproc = 0;
  std::vector<const char*> pos;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--inproc") && i + 1 < argc) inproc = atoi(argv[++i]);
    else pos.push_back(argv[i]);
  }
  if (pos.empty()) {
    printf("usage: %s <num_groups> [dist_update] [user_buf] [use_test] [quant] [--inproc N]\n", argv[0]);
    return 2;
  }
  cfg.modelParts = (size_t)atoi(pos[0]);
  if (pos.size() > 1) cfg.distUpdate = atoi(pos[1]) != 0;
  if (pos.size() > 2) cfg.userBuf = atoi(pos[2]) != 0;
  if (pos.size() > 3) cfg.useTest = atoi(pos[3]) != 0;
  if (pos.size() > 4) cfg.quant = atoi(pos[4]) != 0;
  if (inproc <= 0) {
    Net net;
    net.cfg = cfg;
    return net.run();
  }
  int world = 0;
  mlsl_inproc_world_create(inproc, &world);
  std::vector<int> rc(inproc, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < inproc; ++r)
    th.emplace_back([&, r] {
      mlsl_inproc_bind_thread(world, r);
      Net net;
      net.cfg = cfg;
      rc[r] = net.run();
      mlsl_inproc_unbind_thread();
    });
  for (auto& t : th) t.join();
  mlsl_inproc_world_destroy(world);
  int bad = 0;
  for (int r : rc) bad |= r;
  printf("%s\n", bad ? "Run FAILED." : "Run PASSED.");
  return bad;
}
