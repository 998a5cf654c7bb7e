// Fused eval kernels for D = 1 explored dimensions (klerg_fused.cuh); one translation unit per D so that the
// build compiles them in parallel.
#include "klerg_fused.cuh"

namespace klerg {
template int launch_grad_d<1>(EvalArgs&, int64_t, cudaStream_t);
template int launch_cost_d<1>(EvalArgs&, int64_t, cudaStream_t);
}  // namespace klerg
