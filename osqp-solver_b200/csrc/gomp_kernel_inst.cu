// gomp_kernel_inst.cu -- one explicit instantiation of the ADMM kernel per translation unit.
// Built several times with -DGOMP_RPT=.. -DGOMP_CPT=.. -DGOMP_BS=.. so variants compile in parallel.
#include "gomp_kernel.cuh"

#if !defined(GOMP_RPT) || !defined(GOMP_CPT) || !defined(GOMP_BS)
#error "define GOMP_RPT, GOMP_CPT and GOMP_BS"
#endif

namespace gomp_dev {
template cudaError_t prepare_variant<GOMP_RPT, GOMP_CPT, GOMP_BS>(size_t, int *);
template cudaError_t launch_variant<GOMP_RPT, GOMP_CPT, GOMP_BS>(const DevSym &, const DevBatch &, const DevSettings &, int, size_t,
                                                                 cudaStream_t);
} // namespace gomp_dev
