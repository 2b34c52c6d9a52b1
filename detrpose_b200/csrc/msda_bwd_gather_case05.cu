// Gather-form backward, instantiations for 4 lane(s) per pixel x 2 vector(s) per lane, 1024 threads
// (see msda_bwd_gather.cuh; one translation unit per combination so that they compile in parallel).
#include "msda_bwd_gather.cuh"

namespace msda {
template cudaError_t gather_case<4, 2, 1024>(bool, const GatherArgs&);
}  // namespace msda
