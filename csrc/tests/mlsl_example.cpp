// API tour: the documented user workflow (reference Developer Guide 2.2 / tests/examples/mlsl_example):
// Init -> Session + Distribution -> per layer RegInfo/AddOperation/SetPrev -> Commit -> iterate
// {fwd: WaitComm, compute, StartComm; bwd: WaitComm, compute, StartComm, StartGradientComm;
//  update: WaitGradientComm, optimizer, StartIncrementComm} -> teardown.  No numeric checks; prints the layout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mlsl.hpp"

using namespace MLSL;

int main(int argc, char** argv) {
  Environment& env = Environment::GetEnv();
  env.Init(&argc, &argv);
  size_t rank = env.GetProcessIdx(), world = env.GetProcessCount();
  size_t modelParts = argc > 1 ? (size_t)atoi(argv[1]) : 1;
  if (modelParts < 1 || world % modelParts) modelParts = 1;
  Session* session = env.CreateSession();
  session->SetGlobalMinibatchSize(8 * world);
  Distribution* dist = env.CreateDistribution(world / modelParts, modelParts);
  const size_t layers = 3, fm = 64 * modelParts, fmSize = 16;
  std::vector<Operation*> ops;
  for (size_t l = 0; l < layers; ++l) {
    OperationRegInfo* ri = session->CreateOperationRegInfo(OT_CC);
    char name[32];
    snprintf(name, sizeof(name), "fc_%zu", l);
    ri->SetName(name);
    ri->AddInput(fm, fmSize, DT_FLOAT);
    ri->AddOutput(fm, fmSize, DT_FLOAT);
    ri->AddParameterSet(fm * fm, 1, DT_FLOAT, /*distributedUpdate=*/l % 2 == 1);
    size_t idx = session->AddOperation(ri, dist);
    session->DeleteOperationRegInfo(ri);
    ops.push_back(session->GetOperation(idx));
    if (l) ops[l]->SetPrev(ops[l - 1], 0, 0);
  }
  session->Commit();
  if (rank == 0)
    for (Operation* op : ops) {
      ParameterSet* ps = op->GetParameterSet(0);
      printf("%s: local mb %zu, out fm %zu (pack blocks %zu), kernels local %zu owned %zu @%zu\n", op->GetName(),
             op->GetLocalMinibatchSize(), op->GetOutput(0)->GetLocalFmCount(), op->GetOutput(0)->GetPackBlockCount(),
             ps->GetLocalKernelCount(), ps->GetOwnedKernelCount(), ps->GetOwnedKernelOffset());
    }
  std::vector<float*> grads, weights;
  for (Operation* op : ops) {
    ParameterSet* ps = op->GetParameterSet(0);
    size_t n = ps->GetLocalKernelCount() * ps->GetKernelSize();
    grads.push_back((float*)env.Alloc(n * sizeof(float), 64));
    weights.push_back((float*)env.Alloc(n * sizeof(float), 64));
    memset(grads.back(), 0, n * sizeof(float));
    memset(weights.back(), 0, n * sizeof(float));
  }
  for (int iter = 0; iter < 10; ++iter) {
    for (size_t l = 0; l < layers; ++l) {          // forward
      ops[l]->GetInput(0)->WaitComm();
      ops[l]->GetParameterSet(0)->WaitIncrementComm();
      Activation* out = ops[l]->GetOutput(0);
      if (out->GetCommBuf()) out->StartComm(out->GetCommBuf());
    }
    for (size_t l = layers; l-- > 0;) {            // backward
      ops[l]->GetOutput(0)->WaitComm();
      Activation* in = ops[l]->GetInput(0);
      if (in->GetCommBuf()) in->StartComm(in->GetCommBuf());
      ops[l]->GetParameterSet(0)->StartGradientComm(grads[l]);
    }
    for (size_t l = 0; l < layers; ++l) {          // update
      ops[l]->GetParameterSet(0)->WaitGradientComm();
      ops[l]->GetParameterSet(0)->StartIncrementComm(weights[l]);
    }
  }
  for (size_t l = 0; l < layers; ++l) {
    ops[l]->GetParameterSet(0)->WaitIncrementComm();
    ops[l]->GetInput(0)->WaitComm();
    env.Free(grads[l]);
    env.Free(weights[l]);
  }
  dist->Barrier(GT_GLOBAL);
  if (rank == 0) printf("mlsl_example: done (%zu ranks, %zu model parts)\n", world, modelParts);
  env.DeleteSession(session);
  env.DeleteDistribution(dist);
  env.Finalize();
  return 0;
}
