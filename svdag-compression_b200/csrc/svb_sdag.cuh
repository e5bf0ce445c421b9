// svb_sdag.cuh -- SSVDAG mirror-symmetry reduction on the device (svb_sdag.cu).
#pragma once
#include "svb_context.cuh"

namespace svb {
// GeomOctree::toSDAG(false,false): rewrites ctx->out in place, returns sum of level sizes for levels >= 1.
uint64_t to_sdag_device(svb_ctx* ctx);
}
