// tcgen05 attention kernels (placeholder until the tensor-core kernel lands in this file).
#include "kernels.h"

namespace lmv {
bool attention_tc_supported(const AttnArgs&) { return false; }
int attention_tc_run(const AttnArgs&, cudaStream_t) { return fail(LMV_ERR_UNSUPPORTED, "attention_tc: not built"); }
}  // namespace lmv
