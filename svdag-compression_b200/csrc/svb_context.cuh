// svb_context.cuh -- the opaque svb_ctx behind the C ABI.
#pragma once
#include <memory>

#include "svb_internal.cuh"

struct svb_build_state;   // svb_api.cu: tables and tile list a build keeps between its phases
struct svb_attr_state;    // svb_attr.cu: the SVO of an attribute build (material-id leaves, bit-trees)

// page-locked host buffer owned by a context: D2H / H2D copies run at full PCIe rate and without a bounce buffer
struct PinnedBuf {
	uint8_t* p = nullptr;
	size_t cap = 0;
	uint8_t* reserve(size_t bytes) {
		if (bytes > cap) {
			if (p) cudaFreeHost(p);
			p = nullptr;
			cap = 0;
			size_t want = bytes + bytes / 4 + 4096;
			if (cudaMallocHost((void**)&p, want) != cudaSuccess) { cudaGetLastError(); throw svb::Error(SVB_ENOMEM, "cudaMallocHost failed"); }
			cap = want;
		}
		return p;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct svb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool ownsStream = true;   // false: a caller's stream (svb_create_on_stream)
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
	bool profiling = false, profAccumulate = false;
	bool profEmitOnly = false;              // svb_set_profiling(ctx, 3): only the launches of the "emit" family are bracketed
	std::vector<cudaEvent_t> evPool;        // events of resolved records, reused (cudaEventCreate costs more than the record itself)
	struct PendingProf { svb_prof_rec rec; cudaEvent_t e0, e1; bool closed = false; };
	std::vector<PendingProf> pending;
	std::vector<svb_prof_rec> prof;
	uint64_t batchBudget = 0;
	// seeds of the hashed node keys: a detected 64-bit tag collision re-runs the stage with the next seed (svb_api.cu).
	// mergeSeed seeds the tables of the multi-GPU level merge and must be the same on every rank.
	uint64_t hashSeed = 0, mergeSeed = 0;
	uint64_t nHashRetries = 0;
	// last file image produced by svb_encode (a size query followed by the real call must not encode twice);
	// dropped whenever the octree changes
	PinnedBuf image, staging;
	uint64_t imageSize = 0;
	int lastImageKind = -1;
	std::shared_ptr<svb_build_state> build;   // non-null between svb_shard_build and svb_shard_finish
	std::shared_ptr<svb_attr_state> attr;     // non-null after svb_build_svo_materials
	svb_ctx() { memset(&stats, 0, sizeof(stats)); }
};
