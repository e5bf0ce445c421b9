// svb_internal.cuh -- shared declarations of the libsvb.so translation units.
// Hand-written CUDA for sm_100a; no CPU path exists behind any of these.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svb.h"

namespace svb {

// ------------------------------------------------------------------ errors
struct Error : public std::runtime_error {
	int code;
	Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SVB_CUDA(expr)                                                                          \
	do {                                                                                        \
		cudaError_t _e = (expr);                                                                \
		if (_e != cudaSuccess) {                                                                \
			cudaGetLastError();                                                                 \
			throw ::svb::Error(_e == cudaErrorMemoryAllocation ? SVB_ENOMEM : SVB_ECUDA,        \
			                   std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +      \
			                       __FILE__ + ":" + std::to_string(__LINE__) + ")");            \
		}                                                                                       \
	} while (0)

// every kernel launch is followed by this check; it also counts launches (svb_stats::nKernelLaunches)
extern std::atomic<uint64_t> g_launches;
#define SVB_KERNEL_CHECK()            \
	do {                              \
		::svb::g_launches.fetch_add(1, std::memory_order_relaxed); \
		SVB_CUDA(cudaGetLastError()); \
	} while (0)

// ------------------------------------------------------------------ constants
static const uint32_t NULLNODE = SVB_NULL_NODE;   // octree.cpp:22
static const uint32_t NULLREF = 0xFFFFFFFFu;      // "no child / empty child" inside dedup keys
static const uint32_t UNSET = 0xFFFFFFFFu;
static const uint64_t EMPTY_TAG = 0ull;
static const uint64_t MAX_ORDER = 0xFFFFFFFFFFFFFFFFull;

enum LevelKind { KIND_LEAF = 0, KIND_K64 = 1, KIND_INNER = 2 };

// Hashed node keys (inner dedup levels, SDAG class keys, cross-level subtree ids) are 64-bit tags that an exact pass verifies;
// a detected collision makes the stage run again with another seed (svb_api.cu), so it is never a user-visible failure.
// init: start value of the hash chain (a function of the seed); mask: all ones -- except under the test hook
// SVB_TEST_WEAK_HASH=<bits>, which truncates the tags of the FIRST attempt (seed 0) to force the retry path.
struct HashSeed {
	uint64_t init = 0x9E3779B97F4A7C15ull;
	uint64_t mask = ~0ull;
};
HashSeed make_hash_seed(uint64_t seed);   // svb_dedup.cu
__host__ __device__ inline uint64_t finish_tag(uint64_t h, const HashSeed& hs) {
	h &= hs.mask;
	return h ? h : 1ull;
}
// how a dedup kernel reads the refs of the level below
enum ChildMode { CH_MASK_U8 = 0, CH_SLOT_U32 = 1, CH_UID_U32 = 2, CH_MASK_U32 = 3 };

// ------------------------------------------------------------------ device memory
// Slab allocator: a few large cudaMalloc'ed slabs (grown geometrically, kept for the lifetime of the
// context) carved up by an address-ordered first-fit free list with coalescing.  Every buffer of the
// build lives on ONE stream, so a block can be handed out again the moment it is freed (stream order
// serialises the old user before the new one); no driver call sits on the build's critical path once
// the slabs exist.  (cudaMallocAsync's pool was measured to re-map physical memory between builds:
// 2-3x run-to-run variance on the 16K^3 workload.)
struct Pool {
	cudaStream_t stream = nullptr;
	size_t live = 0, peak = 0;
	struct Slab { char* base; size_t size; };
	struct Free { size_t slab; size_t off; size_t size; };
	std::vector<Slab> slabs;
	std::vector<Free> freeList;   // sorted by (slab, off)
	size_t reserved = 0;
	size_t nextSlab = 1ull << 30;
	static size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }

	void* alloc(size_t bytes) {
		bytes = round_up(bytes ? bytes : 16);
		for (int attempt = 0; attempt < 2; ++attempt) {
			for (size_t i = 0; i < freeList.size(); ++i) {
				Free& f = freeList[i];
				if (f.size < bytes) continue;
				char* p = slabs[f.slab].base + f.off;
				if (f.size == bytes) freeList.erase(freeList.begin() + i);
				else { f.off += bytes; f.size -= bytes; }
				live += bytes;
				if (live > peak) peak = live;
				return p;
			}
			grow(bytes);
		}
		throw Error(SVB_ENOMEM, "device slab allocator: out of memory");
	}
	void free(void* ptr, size_t bytes) {
		if (!ptr) return;
		bytes = round_up(bytes ? bytes : 16);
		char* p = (char*)ptr;
		size_t si = 0;
		for (; si < slabs.size(); ++si) if (p >= slabs[si].base && p < slabs[si].base + slabs[si].size) break;
		if (si == slabs.size()) return;
		Free nf{si, (size_t)(p - slabs[si].base), bytes};
		size_t pos = 0;
		while (pos < freeList.size() && (freeList[pos].slab < nf.slab || (freeList[pos].slab == nf.slab && freeList[pos].off < nf.off))) ++pos;
		freeList.insert(freeList.begin() + pos, nf);
		if (pos + 1 < freeList.size() && freeList[pos + 1].slab == nf.slab && nf.off + nf.size == freeList[pos + 1].off) {
			freeList[pos].size += freeList[pos + 1].size;
			freeList.erase(freeList.begin() + pos + 1);
		}
		if (pos > 0 && freeList[pos - 1].slab == nf.slab && freeList[pos - 1].off + freeList[pos - 1].size == freeList[pos].off) {
			freeList[pos - 1].size += freeList[pos].size;
			freeList.erase(freeList.begin() + pos);
		}
		live -= bytes;
	}
	void grow(size_t need) {
		size_t want = nextSlab;
		while (want < need) want <<= 1;
		size_t freeB = 0, totalB = 0;
		cudaMemGetInfo(&freeB, &totalB);
		size_t room = freeB > (1ull << 30) ? freeB - (1ull << 30) : 0;   // leave 1 GiB to the rest of the process
		if (want > room) want = round_up(need) <= room ? (room & ~(size_t)511) : 0;
		if (want < need) throw Error(SVB_ENOMEM, "device slab allocator: cannot grow (requested " + std::to_string(need >> 20) + " MiB)");
		void* p = nullptr;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) {
			cudaGetLastError();
			throw Error(SVB_ENOMEM, std::string("cudaMalloc(slab): ") + cudaGetErrorString(e));
		}
		slabs.push_back(Slab{(char*)p, want});
		reserved += want;
		Free nf{slabs.size() - 1, 0, want};
		freeList.push_back(nf);   // highest slab index: stays sorted
		if (nextSlab < (32ull << 30)) nextSlab <<= 1;
	}
	// bytes that can still be handed out without exhausting the device
	size_t headroom() const {
		size_t freeB = 0, totalB = 0;
		cudaMemGetInfo(&freeB, &totalB);
		size_t room = freeB > (1ull << 30) ? freeB - (1ull << 30) : 0;
		return room + (reserved - live);
	}
	void release_all() {
		for (auto& sl : slabs) cudaFree(sl.base);
		slabs.clear();
		freeList.clear();
		reserved = live = 0;
	}
};

template <class T>
struct DevBuf {
	Pool* pool = nullptr;
	T* p = nullptr;
	size_t n = 0;   // elements
	DevBuf() {}
	DevBuf(Pool& pl, size_t count) { reset(pl, count); }
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	DevBuf(DevBuf&& o) noexcept : pool(o.pool), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
	DevBuf& operator=(DevBuf&& o) noexcept {
		if (this != &o) { release(); pool = o.pool; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
		return *this;
	}
	~DevBuf() { release(); }
	void reset(Pool& pl, size_t count) {
		release();
		pool = &pl;
		n = count;
		p = (T*)pl.alloc(count * sizeof(T));
	}
	void release() {
		if (p && pool) pool->free(p, n * sizeof(T));
		p = nullptr;
		n = 0;
	}
	size_t bytes() const { return n * sizeof(T); }
	void zero() { if (p) SVB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T) ? n * sizeof(T) : 16, pool->stream)); }
	void fill_ff() { if (p) SVB_CUDA(cudaMemsetAsync(p, 0xFF, n * sizeof(T) ? n * sizeof(T) : 16, pool->stream)); }
};

// ------------------------------------------------------------------ geometry of one (sub-)octree
#ifndef SVB_TILEGEOM_DEFINED
#define SVB_TILEGEOM_DEFINED
struct TileGeom {   // host computed, exactly as geom_octree.cpp:177-184,214 / :340-344 would
	double cx, cy, cz;   // root centre (double)
	double rootSide;     // float-rounded max side, widened
};
#endif

// ------------------------------------------------------------------ one level of the current tile batch
struct BatchLevel {
	uint64_t n = 0;
	DevBuf<uint64_t> code;        // (tile_local << 3l) | path   (ascending == Morton order, tile major)
	DevBuf<uint32_t> tstar;       // first-touch triangle
	DevBuf<uint8_t> mask;         // structural child mask / leaf voxel mask (OR'd with 32-bit atomics; padded)
	DevBuf<uint32_t> childBase;   // index of first child in the next level
	DevBuf<uint32_t> ref;         // dedup result (slot / uid / NULLREF)
};

// ------------------------------------------------------------------ persistent per-level dedup table
struct LevelTable {
	int kind = KIND_INNER;
	// open-addressing slots
	uint64_t cap = 0;
	DevBuf<uint64_t> tag;     // K64: the key itself; INNER: 64-bit hash of the 8 child uids; 0 = empty
	DevBuf<uint64_t> minO;    // min order key seen for this slot
	DevBuf<uint32_t> uid;     // dense id (UNSET until the winner pass)
	// dense, indexed by uid
	uint64_t count = 0;       // host copy of *dCount
	uint64_t denseCap = 0;
	DevBuf<uint32_t> dCount;  // device counter (1 element)
	DevBuf<uint64_t> dMinO;
	DevBuf<uint64_t> dKey64;  // KIND_K64: 8 child masks
	DevBuf<uint32_t> dKey8;   // KIND_INNER: 8 child uids (NULLREF = none)
	// KIND_LEAF: 256-entry direct table lives in minO (cap = 256); wide: the order key of this level does not fit 63
	// bits and is kept as (high part = tile_seq | t*, low part = path') in minO[0,256) / minO[256,512)
	bool wide = false;
	// largest tile_seq reduced into this table so far (see DedupArgs::seqLo)
	bool seenAny = false;
	uint32_t maxSeq = 0;
	uint32_t known = 0, lastAdded = 0;   // KIND_LEAF: voxel masks with an entry, and how many of them the last batch brought
	uint64_t lastFresh = 0, lastN = 0;   // KIND_K64 / KIND_INNER: entries the last batch created, out of how many nodes
	// finalize
	DevBuf<uint32_t> rank;    // uid -> final id (LEAF: mask value -> final id)
	uint64_t unique = 0;
	HashSeed hs;              // KIND_INNER: seed of the 64-bit tags
};

// Per-launch timing records from inside the stage drivers (svb_api.cu implements it on top of the context's profile list;
// null = not profiling).  begin() records an event on the stage's stream, end() the closing one plus the launch's units
// and algorithmic bytes.
struct ProfHook {
	virtual int begin(const char* name, uint32_t level, uint64_t n_in) = 0;
	virtual void end(int id, uint64_t n_out, double bytes) = 0;
	virtual ~ProfHook() {}
};

// ------------------------------------------------------------------ final octree (what getNodeData() exposes)
struct OutLevel {
	uint64_t n = 0;
	DevBuf<uint8_t> mask;
	DevBuf<uint32_t> child;      // n*8
	DevBuf<uint8_t> mirror;      // n*3 (zero until toSDAG)
	DevBuf<uint8_t> inv;         // n
	DevBuf<uint32_t> childLevel; // n*8 (only materialised by cross-merge; otherwise lev+1)
	bool hasChildLevel = false;
};

// ------------------------------------------------------------------ primitives (svb_prims.cu)
// exclusive scan of popcount(bytes[i]) -> out[i] (uint32), returns total through *d_total (device, u64)
void scan_popc8(cudaStream_t s, Pool& pool, const uint8_t* bytes, uint64_t n, uint32_t* out, uint64_t* d_total);
// Tile-granular scans (tiles of scan_tile_items() consecutive items): only the exclusive offset of every tile is
// produced; the consumer kernels (k_children / k_emit in svb_voxelize.cu) rebuild the per-item offsets inside the CTA.
uint64_t scan_tile_items();
// rel (optional): tile-relative exclusive offsets at a granularity of 8 items (one u32 per 8 items; the pair variant packs
// A in the low and B in the high half), so that a consumer can place its output warp by warp.
void scan_tiles_popc8(cudaStream_t s, Pool& pool, const uint8_t* bytes, uint64_t n, DevBuf<uint64_t>& tileOffs, uint64_t* d_total, DevBuf<uint32_t>* rel = nullptr);
// A = popcount(hit) of every pair, B = the same restricted to pairs whose flags put their children into the flat stream
void scan_tiles_pairs(cudaStream_t s, Pool& pool, const uint8_t* hit, const uint16_t* flags, uint64_t n,
                      DevBuf<uint64_t>& tileOffsA, DevBuf<uint64_t>& tileOffsB, uint64_t* d_totalA, uint64_t* d_totalB, DevBuf<uint32_t>* rel = nullptr, int flatTri = 0);
// The four tile-granular scans of one voxelizer level in one go (three reduce kernels + ONE launch for the four scans of the
// tile sums): children per node, child pairs of the flat stream, child pairs / flat child pairs of the slow stream.
// d_tot4[0..3] receive the four totals.
void scan_level_tiles(cudaStream_t s, Pool& pool, const uint8_t* nodeMask, uint64_t nNodes, const uint8_t* hitF, uint64_t nF,
                      const uint8_t* hitS, const uint16_t* flagsS, uint64_t nS,
                      DevBuf<uint64_t>& nodeOffs, DevBuf<uint64_t>& offF, DevBuf<uint64_t>& offS, DevBuf<uint64_t>& offSF,
                      DevBuf<uint32_t>* relF, DevBuf<uint32_t>* relS, uint64_t* d_tot4, int flatTri = 0);
// exclusive scan of uint32 values (in place allowed), total to *d_total
void scan_u32(cudaStream_t s, Pool& pool, const uint32_t* in, uint64_t n, uint32_t* out, uint64_t* d_total);
// stable LSD radix sort of (key u64, val u32) pairs on the low `bits` bits of the key.
// Results end up back in keys/vals.
void radix_sort_pairs(cudaStream_t s, Pool& pool, uint64_t* keys, uint32_t* vals, uint64_t n, int bits);

// ------------------------------------------------------------------ device helpers
__host__ __device__ inline uint64_t mix64(uint64_t x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) {
	uint64_t b = (n + threads - 1) / threads;
	if (b == 0) b = 1;
	if (b > 0x7FFFFFFFull) throw Error(SVB_ERANGE, "grid too large");
	return (unsigned)b;
}

}  // namespace svb
