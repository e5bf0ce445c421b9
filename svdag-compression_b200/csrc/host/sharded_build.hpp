// sharded_build.hpp -- GeomOctree::buildDAG over several GPUs of one node from ONE process: one host thread and one
// libsvb context per device, NCCL (one communicator per device, ncclCommInitAll) for the exchanges.
//
// The decomposition is the reference's own (src/symvox/geom_octree.cpp:331-378: sub-octrees are independent; :397-425:
// join + last DAG pass): sub-octrees are dealt over the devices (round-robin in the reference's order, or by top-level
// octant with SVB_SHARD=octant), every device reduces its share, and the levels are merged bottom-up by all-gathering
// the canonical node records -- the svb_shard_* protocol of include/svb.h, driven here exactly like
// svdag-compression_b200/sharded.py drives it from one process per GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "svb.h"

namespace svbhost {

// ctx[r] lives on devices[r]; on success every context holds the same octree in state DAG and *out its stats.
// tris: host triangle soup (9 floats per triangle): every device uploads 1/N of it, NCCL all-gathers the rest.
// Returns false and fills err on failure.
bool build_dag_sharded(const std::vector<int>& devices, const std::vector<svb_ctx*>& ctx, const float* tris, uint64_t ntris,
                       unsigned levels, unsigned step, const double bmin[3], const double bmax[3], svb_stats* out, std::string* err,
                       double* msUpload = nullptr, double* msExchange = nullptr);

}  // namespace svbhost
