// Gather-form backward, instantiations for 4 lane(s) per pixel x 3 vector(s) per lane, 512 threads
// (see msda_bwd_gather.cuh; one translation unit per combination so that they compile in parallel).
#include "msda_bwd_gather.cuh"

namespace msda {
template cudaError_t gather_case<4, 3, 512>(bool, const GatherArgs&);
}  // namespace msda
