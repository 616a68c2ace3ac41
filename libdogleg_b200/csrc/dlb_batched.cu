// dlb_batched.cu -- batched dense solves (config C3). Placeholder entry point
// until the batched kernels land; fails loudly.
#include "dogleg_gpu.h"
#include <cstdio>
extern "C" void dlb_set_error(const char* msg);
extern "C" int dogleg_gpu_optimize_dense_batched(double* p, unsigned int Nstate, unsigned int Nmeas,
                                                 unsigned int B, dogleg_gpu_callback_dense_batched_t* f,
                                                 void* cookie, const dogleg_parameters2_t* parameters,
                                                 double* norm2x_out, int* iterations_out)
{
  (void)p; (void)Nstate; (void)Nmeas; (void)B; (void)f; (void)cookie; (void)parameters; (void)norm2x_out; (void)iterations_out;
  dlb_set_error("dogleg_gpu_optimize_dense_batched: not built yet");
  return -1;
}
