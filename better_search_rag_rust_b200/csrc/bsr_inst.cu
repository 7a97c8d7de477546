// One group of scan-kernel instantiations (see bsr_instances.inc); compiled once per group with
// -DBSR_INST_GROUP=<n> by the Makefile.
#define BSR_TEMPLATES_ONLY
#include "bsr_kernels.cuh"

namespace bsr {
#define BSR_INST(...) template __global__ __VA_ARGS__;
#include "bsr_instances.inc"
#undef BSR_INST
}  // namespace bsr
