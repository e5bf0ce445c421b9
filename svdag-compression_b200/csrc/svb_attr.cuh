// svb_attr.cuh -- material-id leaves and Gray-coded attribute bit-trees (svb_attr.cu).
#pragma once
#include "svb_internal.cuh"

struct svb_ctx;

namespace svb {

// GeomOctree::buildSVO(levels, bbox, false, NULL, putMaterialIdInLeaves = true) (geom_octree.cpp:171-280, :210-211, :252)
void attr_build(svb_ctx* c, uint32_t levels, const double bmin[3], const double bmax[3], const uint32_t* triMaterialHost);
uint64_t attr_leaf_count(const svb_ctx* c);
// the leaf level in the reference's node order: mask[n], mat8[n * 8] (Octree::nullNode for unset voxels)
void attr_download(svb_ctx* c, uint8_t* maskHost, uint32_t* mat8Host);
// node / voxel counts of the bit-tree of attribute bit 0 .. nbits-1 (binary code, or reflected Gray code g = a ^ (a >> 1))
void attr_bit_trees(svb_ctx* c, uint32_t nbits, int gray, uint64_t* nodes, uint64_t* voxels);

}  // namespace svb
