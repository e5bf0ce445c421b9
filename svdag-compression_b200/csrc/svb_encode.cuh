// svb_encode.cuh -- GPU-side writers of the reference's file formats (svb_encode.cu).
#pragma once
#include "svb_internal.cuh"

struct svb_ctx;

namespace svb {

// Encodes the context's octree as file kind SVB_FILE_* into ctx->image (pinned host memory); returns the image size.
uint64_t encode_device(svb_ctx* c, int kind);

}  // namespace svb
