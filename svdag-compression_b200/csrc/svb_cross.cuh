// svb_cross.cuh -- CSVDAG cross-level merge on the device (svb_cross.cu).
#pragma once
#include "svb_context.cuh"

namespace svb {
// GeomOctree::mergeAcrossAllLevels(): rewrites ctx->out in place (childLevel becomes meaningful).
// Returns the number of nodes removed; *nNodesOut = 1 + surviving nodes of the levels >= 1.
uint64_t cross_merge_device(svb_ctx* ctx, uint64_t* nNodesOut);
}
