// svb_context.cuh -- the opaque svb_ctx behind the C ABI.
#pragma once
#include <memory>

#include "svb_internal.cuh"

struct svb_build_state;   // svb_api.cu: tables and tile list a build keeps between its phases

struct svb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	svb::Pool pool;
	std::string err;
	// Scene triangle soup (scene.hpp:90-91)
	const float* d_tris = nullptr;
	uint64_t T = 0;
	svb::DevBuf<float> trisOwned;
	// GeomOctree state (geom_octree.hpp:100-104)
	int state = SVB_S_EMPTY;
	uint32_t levels = 0;
	std::vector<svb::OutLevel> out;
	svb_stats stats;
	std::vector<uint64_t> svoCounts;
	// instrumentation
	bool profiling = false;
	struct PendingProf { svb_prof_rec rec; cudaEvent_t e0, e1; bool closed = false; };
	std::vector<PendingProf> pending;
	std::vector<svb_prof_rec> prof;
	uint64_t batchBudget = 0;
	// last file image produced by svb_encode (a size query followed by the real call must not encode twice);
	// dropped whenever the octree changes
	std::vector<uint8_t> lastImage;
	int lastImageKind = -1;
	std::shared_ptr<svb_build_state> build;   // non-null between svb_shard_build and svb_shard_finish
	svb_ctx() { memset(&stats, 0, sizeof(stats)); }
};
