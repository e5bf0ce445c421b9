// svb_voxelize.cu -- conservative triangle voxelization into a sparse voxel octree,
// level-synchronous over (triangle, node) pairs for a whole batch of sub-octrees ("tiles").
//
// Replaces the per-triangle DFS of GeomOctree::buildSVO (src/symvox/geom_octree.cpp:205-261):
// a node exists iff some triangle passed testTriBox at every ancestor, its child mask is the OR
// of the 8-child hit masks of its pairs, and its first-touch triangle t* (which fixes the
// reference's node creation order) is the min triangle id over its pairs.
//
// Layout per level (SoA, nodes in Morton order, tile-major):
//   code[n]  u64  (tile_local << 3l) | path      tstar[n] u32      mask[n] u8      childBase[n] u32
// Pairs: tri[p] u32, node[p] u32, hit[p] u8 -- pairs stay sorted by triangle id at every level.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "svb_internal.cuh"
#include "svb_classify.cuh"
#include "svb_sat.cuh"
#include "svb_voxelize.cuh"

namespace svb {

namespace {

constexpr int VX_THREADS = 256;

// ------------------------------------------------------------------ candidate (triangle, tile) pairs
// A triangle can only pass testTriBox against a child of tile T if its AABB reaches T's cube
// (the box-axis tests are part of the predicate), so an AABB-vs-cube test inflated by a margin
// far above double rounding error is a safe superset of "all triangles" (what the reference
// feeds every sub-octree, geom_octree.cpp:344 -> :205).
struct GridDesc {
	double ox, oy, oz;   // min corner of the root cube
	double inv_cell;     // 1 / tile side
	double margin;       // absolute
	int G;               // tiles per axis
	int lo[3], hi[3];    // grid-cell bounding box of the tiles of interest (the current batch): triangles outside are skipped early
};

__device__ inline void tile_range(double mn, double mx, double o, const GridDesc& g, int& a, int& b) {
	double fa = floor((mn - g.margin - o) * g.inv_cell);
	double fb = floor((mx + g.margin - o) * g.inv_cell);
	a = fa < 0.0 ? 0 : (fa > (double)(g.G - 1) ? g.G : (int)fa);
	b = fb < 0.0 ? -1 : (fb > (double)(g.G - 1) ? g.G - 1 : (int)fb);
}

template <bool EMIT>
__global__ void __launch_bounds__(VX_THREADS) k_candidates(const float* __restrict__ tris, uint64_t T, GridDesc g,
                                                            const int* __restrict__ gridTile, const int* __restrict__ selPos, int posFirst, int posCount, uint32_t* __restrict__ cnt_or_off,
                                                            uint32_t* __restrict__ ptri, uint32_t* __restrict__ pnode) {
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	const float* p = tris + 9 * t;
	double mnx = fmin(fmin((double)p[0], (double)p[3]), (double)p[6]), mxx = fmax(fmax((double)p[0], (double)p[3]), (double)p[6]);
	double mny = fmin(fmin((double)p[1], (double)p[4]), (double)p[7]), mxy = fmax(fmax((double)p[1], (double)p[4]), (double)p[7]);
	double mnz = fmin(fmin((double)p[2], (double)p[5]), (double)p[8]), mxz = fmax(fmax((double)p[2], (double)p[5]), (double)p[8]);
	int ax, bx, ay, by, az, bz;
	tile_range(mnx, mxx, g.ox, g, ax, bx);
	tile_range(mny, mxy, g.oy, g, ay, by);
	tile_range(mnz, mxz, g.oz, g, az, bz);
	ax = max(ax, g.lo[0]); bx = min(bx, g.hi[0]);
	ay = max(ay, g.lo[1]); by = min(by, g.hi[1]);
	az = max(az, g.lo[2]); bz = min(bz, g.hi[2]);
	uint32_t c = 0;
	uint32_t o = EMIT ? cnt_or_off[t] : 0;
	for (int x = ax; x <= bx; ++x)
		for (int y = ay; y <= by; ++y)
			for (int z = az; z <= bz; ++z) {
				int s = gridTile[((size_t)x * g.G + y) * g.G + z];   // global tile_seq of the cell, -1 = no tile
				int loc = (s >= 0) ? selPos[s] - posFirst : -1;      // index inside this batch; outside [0, posCount): other batch / other rank
				if (loc >= posCount) loc = -1;
				if (loc >= 0) {
					if (EMIT) { ptri[o + c] = (uint32_t)t; pnode[o + c] = (uint32_t)loc; }
					++c;
				}
			}
	if (!EMIT) cnt_or_off[t] = c;
}

// ------------------------------------------------------------------ classify: 8 SAT tests per pair
// 8 consecutive lanes share one pair, lane c tests child c (index bits X=4,Y=2,Z=1, octree.hpp:33-42).
// The node centre is rebuilt by replaying the reference's chain centre += (+-k_d) level by level
// (geom_octree.cpp:222-230), so it carries exactly the reference's roundings for any bbox.
__global__ void __launch_bounds__(VX_THREADS) k_classify(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                          const uint64_t* __restrict__ code, int l, const TileGeom* __restrict__ tiles,
                                                          const float* __restrict__ tris, const uint32_t* __restrict__ rootTri, uint8_t* __restrict__ hit, uint8_t* __restrict__ mask, int last,
                                                          uint32_t* __restrict__ lastTri = nullptr) {
	uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t p = gid >> 3;
	int c = (int)(gid & 7);
	bool ok = false;
	uint32_t n = 0, t = 0;
	if (p < P) {
		t = rootTri[ptri[p]];
		n = pnode[p];
		uint64_t cd = code[n];
		uint32_t tile = (uint32_t)(cd >> (3 * l));
		TileGeom tg = tiles[tile];
		double cx = tg.cx, cy = tg.cy, cz = tg.cz;
		double k = tg.rootSide * 0.25;   // getHalfSideD(1) = rootSide / 4 (octree.hpp:115), exact scaling
		for (int d = l - 1; d >= 0; --d) {
			int dig = (int)((cd >> (3 * d)) & 7);
			cx = __dadd_rn(cx, (dig & 4) ? k : -k);
			cy = __dadd_rn(cy, (dig & 2) ? k : -k);
			cz = __dadd_rn(cz, (dig & 1) ? k : -k);
			k *= 0.5;
		}
		cx = __dadd_rn(cx, (c & 4) ? k : -k);
		cy = __dadd_rn(cy, (c & 2) ? k : -k);
		cz = __dadd_rn(cz, (c & 1) ? k : -k);
		ok = tri_box_overlap(cx, cy, cz, k, tris + 9ull * t);
		// leaf level of an attribute build: the LAST triangle (file order) that touches voxel c of leaf node n (1-based; 0 = none)
		if (ok && lastTri) atomicMax(&lastTri[(uint64_t)n * 8 + c], t + 1);
	}
	unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
	int lane = threadIdx.x & 31;
	unsigned m = (b >> (lane & 24)) & 0xFFu;
	if (c == 0 && p < P) {
		if (!last) hit[p] = (uint8_t)m;
		if (m) {
			unsigned cur = mask[n];   // may be stale (L1); bits only ever get set, so a stale value only costs an extra atomic
			if ((cur & m) != m) atomicOr(reinterpret_cast<unsigned*>(mask) + (n >> 2), m << (8 * (n & 3)));
		}
	}
}

// ------------------------------------------------------------------ classify, filtered (default)
// One thread per pair decides all 8 children with classify_pair() (svb_classify.cuh): exact single-axis fast path
// for settled flat triangles, FP64 interval filter + reference-order predicate for the rest.
template <int MINB, bool DIRECT, bool FLATONLY, bool LAST>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_classify_filtered(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                   uint16_t* __restrict__ pflags, const uint64_t* __restrict__ code, int l, double kscale, int last,
                                                                   const TileGeom* __restrict__ tiles, const float* __restrict__ tris, const uint32_t* __restrict__ rootTri,
                                                                   uint8_t* __restrict__ hit, uint8_t* __restrict__ mask, unsigned long long* __restrict__ nExact, int precheck) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	const uint32_t t = rootTri[ptri[p]], n = pnode[p];   // pairs carry the index of their root pair (see make_root_pairs)
	const unsigned fl0 = pflags[p];
	unsigned fl = fl0, nUnsure;
	const uint64_t cd = code[n];
	const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd >> (3 * l)));   // {cx, cy, cz, rootSide}
	const unsigned m = classify_pair<DIRECT, FLATONLY, LAST>(cd, l, tg, kscale, tris + 9ull * t, fl, nUnsure);
	if (nUnsure && nExact) atomicAdd(nExact, (unsigned long long)nUnsure);
	if (!last) {
		hit[p] = (uint8_t)m;
		if (fl != fl0) pflags[p] = (uint16_t)fl;   // inherited by the child pairs (k_emit)
	}
	// precheck (may read a stale value: bits only ever get set, so that only costs an extra atomic)
	if (m && (!precheck || (mask[n] & m) != m)) atomicOr(reinterpret_cast<unsigned*>(mask) + (n >> 2), m << (8 * (n & 3)));
}

// ------------------------------------------------------------------ tile-local scans
// The scans of svb_prims.cu only deliver the exclusive offset of every tile of VX_TILE consecutive items; the two
// expansion kernels below walk their tile in VX_TILE / 256 chunks, rebuild the per-item offsets with a block scan
// and stage their output in shared memory so that it leaves the SM as contiguous, coalesced runs.
constexpr int VX_TILE = 2048;
constexpr int VX_CHUNKS = VX_TILE / VX_THREADS;

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t x, int lane) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
		if (lane >= d) x += y;
	}
	return x;
}
// exclusive prefix of x over the CTA (256 threads); *total = CTA sum.  Two barriers; wsum is caller-provided smem[9].
__device__ __forceinline__ uint32_t cta_excl_scan(uint32_t x, uint32_t* wsum, uint32_t* total) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t inc = warp_incl_scan_u32(x, lane);
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t sv = (lane < VX_THREADS / 32) ? wsum[lane] : 0;
		uint32_t si = warp_incl_scan_u32(sv, lane);
		if (lane < VX_THREADS / 32) wsum[lane] = si - sv;
		if (lane == VX_THREADS / 32 - 1) wsum[8] = si;
	}
	__syncthreads();
	*total = wsum[8];
	return inc - x + wsum[w];
}

// children of node n: contiguous at childBase[n], ascending child index == Morton order.  Writes childBase and the
// child codes of one tile of nodes.
template <bool PIPE>
__global__ void __launch_bounds__(VX_THREADS) k_children(uint64_t N, const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask,
                                                          const uint64_t* __restrict__ tileOffs, uint32_t* __restrict__ childBase, uint64_t* __restrict__ ccode) {
	__shared__ uint64_t s_code[VX_THREADS * 8];
	__shared__ uint32_t wsum[9];
	const uint64_t n0 = (uint64_t)blockIdx.x * VX_TILE;
	uint64_t run = tileOffs[blockIdx.x];
	// the next chunk's node fields are fetched before this chunk's barriers (PIPE): one exposed round trip per tile, not per chunk
	unsigned mNext = 0;
	uint64_t cdNext = 0;
	if (PIPE && n0 + threadIdx.x < N) { mNext = mask[n0 + threadIdx.x]; cdNext = code[n0 + threadIdx.x]; }
	for (int ch = 0; ch < VX_CHUNKS; ++ch) {
		const uint64_t nb = n0 + (uint64_t)ch * VX_THREADS;
		if (nb >= N) break;
		const uint64_t n = nb + threadIdx.x;
		unsigned m = 0;
		uint64_t cd = 0;
		if (PIPE) {
			m = mNext; cd = cdNext << 3;
			const uint64_t nn = n + VX_THREADS;
			mNext = 0; cdNext = 0;
			if (ch + 1 < VX_CHUNKS && nn < N) { mNext = mask[nn]; cdNext = code[nn]; }
		} else if (n < N) { m = mask[n]; cd = code[n] << 3; }
		uint32_t tot;
		uint32_t o = cta_excl_scan(__popc(m), wsum, &tot);
		if (n < N) childBase[n] = (uint32_t)(run + o);
		while (m) {
			const int c = __ffs(m) - 1;
			m &= m - 1;
			s_code[o++] = cd | (uint64_t)c;
		}
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < tot; i += VX_THREADS) ccode[run + i] = s_code[i];
		run += tot;
		__syncthreads();
	}
}

// ------------------------------------------------------------------ k_children with TMA bulk copies (sm_100a; default, SVB_CHILDREN_TMA=0: the kernel above)
// The level buffers this kernel moves are plain contiguous streams: node codes and masks in, child codes out.  Here the
// copy engine moves them: the CTA's 2048-node tile is cut into 4 stages of 512 nodes; one elected thread issues
// cp.async.bulk (global -> shared, completion counted in bytes on an mbarrier) for stage i+1 while the threads work on
// stage i out of shared memory, and the staged child codes of a stage leave with ONE cp.async.bulk (shared -> global) instead
// of a store loop.  Bulk copies want 16-byte aligned addresses and sizes on both sides: the staging buffer is shifted by
// (run & 1) elements so that shared and global addresses agree mod 16, and the at most one leading / trailing 8-byte element
// goes out with a plain store.  childBase is written as one 8-byte store per thread (two consecutive nodes).
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion on an mbarrier (bytes: multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dstGlobal, const void* srcSmem, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smem_u32(srcSmem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tma

constexpr int CT_STAGE = 256;                      // nodes per pipeline stage (1 per thread)
constexpr int CT_STAGES = VX_TILE / CT_STAGE;      // 8 stages per 2048-node tile
__global__ void __launch_bounds__(VX_THREADS, 5) k_children_tma(uint64_t N, const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask,
                                                                 const uint64_t* __restrict__ tileOffs, uint32_t* __restrict__ childBase, uint64_t* __restrict__ ccode) {
	__shared__ __align__(16) uint64_t s_in[2][CT_STAGE];            // node codes of a stage, double buffered (2 KB each)
	__shared__ __align__(16) uint8_t s_mask[2][CT_STAGE];           // node masks
	__shared__ __align__(16) uint64_t s_out[2][CT_STAGE * 8 + 2];   // child codes of a stage (16 KB each, double buffered) + the alignment shift
	__shared__ __align__(8) uint64_t bar[2];
	__shared__ uint32_t wsum[9];
	const uint64_t n0 = (uint64_t)blockIdx.x * VX_TILE;
	const int tid = threadIdx.x;
	auto stage_count = [&](int st) -> uint32_t {   // nodes of stage st that exist
		const uint64_t b = n0 + (uint64_t)st * CT_STAGE;
		return (st >= CT_STAGES || b >= N) ? 0u : (uint32_t)((N - b) < CT_STAGE ? (N - b) : CT_STAGE);
	};
	auto issue = [&](int st) {   // one thread: arm the stage's barrier and start its two bulk loads
		const uint32_t cnt = stage_count(st);
		if (!cnt) return;
		const int buf = st & 1;
		const uint32_t bc = (cnt * 8 + 15) & ~15u, bm = (cnt + 15) & ~15u;   // (rounded-up reads stay inside the pool block)
		tma::mbar_expect_tx(&bar[buf], bc + bm);
		tma::bulk_g2s(s_in[buf], code + n0 + (uint64_t)st * CT_STAGE, bc, &bar[buf]);
		tma::bulk_g2s(s_mask[buf], mask + n0 + (uint64_t)st * CT_STAGE, bm, &bar[buf]);
	};
	if (tid == 0) {
		tma::mbar_init(&bar[0], 1);
		tma::mbar_init(&bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) { issue(0); issue(1); }
	uint64_t run = tileOffs[blockIdx.x];
	for (int st = 0; st < CT_STAGES; ++st) {
		const uint32_t cnt = stage_count(st);
		if (!cnt) break;
		const int buf = st & 1;
		tma::mbar_wait(&bar[buf], (st >> 1) & 1);                     // stage st has landed
		unsigned m = 0;
		uint64_t cd = 0;
		if ((uint32_t)tid < cnt) { m = s_mask[buf][tid]; cd = s_in[buf][tid] << 3; }
		uint32_t tot;
		uint32_t o = cta_excl_scan(__popc(m), wsum, &tot);            // (its barriers also end every thread's reads of the input buffer)
		if (tid == 0) issue(st + 2);                                  // refill the input buffer just consumed
		if ((uint32_t)tid < cnt) childBase[n0 + (uint64_t)st * CT_STAGE + tid] = (uint32_t)(run + o);
		const uint32_t shift = (uint32_t)(run & 1);                   // shared and global addresses agree mod 16
		uint64_t* const so = s_out[buf];                              // (its previous bulk store, two stages ago, has been drained below)
		uint32_t w = shift + o;
		while (m) { const int c = __ffs(m) - 1; m &= m - 1; so[w++] = cd | (uint64_t)c; }
		tma::fence_proxy_async();                                      // the staged codes must be visible to the copy engine
		__syncthreads();
		if (tid == 0) {
			if (tot) {
				uint32_t head = shift ? 1u : 0u;                      // a leading element on an odd global index
				if (head > tot) head = tot;
				const uint32_t body = (tot - head) & ~1u;             // 16-byte granules
				if (head) ccode[run] = so[shift];
				if (body) tma::bulk_s2g(ccode + run + head, &so[shift + head], body * 8);
				if (head + body < tot) ccode[run + tot - 1] = so[shift + tot - 1];
			}
			tma::bulk_commit();                                        // one group per stage (possibly empty)
			asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store of stage st-1 has read its buffer: stage st+1 may overwrite it
		}
		run += tot;
		__syncthreads();
	}
	if (tid == 0) tma::bulk_wait_read0();
}

// ------------------------------------------------------------------ classify, fast stream
// Pairs of the flat stream (pair_is_fast(), svb_classify.cuh): exact box-axis tests only.  Few registers, full
// occupancy: the kernel is a chain of three dependent loads (pair -> node code -> tile geometry / vertex
// coordinates), so resident warps are what hides the latency.
template <bool DIRECT>
__global__ void __launch_bounds__(VX_THREADS, 8) k_classify_fast(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                 uint16_t* __restrict__ pflags, const uint64_t* __restrict__ code, int l, double kscale, int last,
                                                                 const TileGeom* __restrict__ tiles, const float* __restrict__ tris, const uint32_t* __restrict__ rootTri,
                                                                 uint8_t* __restrict__ hit, uint8_t* __restrict__ mask, int precheck) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	const uint32_t t = rootTri[ptri[p]], n = pnode[p];
	const unsigned fl0 = pflags[p];
	unsigned fl = fl0;
	const uint64_t cd = code[n];
	const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd >> (3 * l)));   // {cx, cy, cz, rootSide}
	const unsigned m = classify_pair_flat<DIRECT>(cd, l, tg, kscale, tris + 9ull * t, fl);
	if (!last && fl != fl0) pflags[p] = (uint16_t)fl;   // a box axis settled: inherited by the child pairs
	if (!last) hit[p] = (uint8_t)m;
	if (m && (!precheck || (mask[n] & m) != m)) atomicOr(reinterpret_cast<unsigned*>(mask) + (n >> 2), m << (8 * (n & 3)));
}

// Two pairs per thread (default, SVB_FAST_ILP=1 selects the kernel above): the kernel is a chain of dependent loads, so what a
// warp has in flight decides its rate; the fields of both pairs are fetched before the first is decided.
template <bool DIRECT>
__global__ void __launch_bounds__(VX_THREADS, 6) k_classify_fast2(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                  uint16_t* __restrict__ pflags, const uint64_t* __restrict__ code, int l, double kscale, int last,
                                                                  const TileGeom* __restrict__ tiles, const float* __restrict__ tris, const uint32_t* __restrict__ rootTri,
                                                                  uint8_t* __restrict__ hit, uint8_t* __restrict__ mask, int precheck) {
	const uint64_t pa = (uint64_t)blockIdx.x * (2 * VX_THREADS) + threadIdx.x, pb = pa + VX_THREADS;
	if (pa >= P) return;
	const bool hasB = pb < P;
	const uint32_t qa = ptri[pa], na = pnode[pa];
	const uint32_t qb = hasB ? ptri[pb] : qa, nb = hasB ? pnode[pb] : na;
	const unsigned fla0 = pflags[pa], flb0 = hasB ? pflags[pb] : fla0;
	const uint32_t ta = rootTri[qa], tb = rootTri[qb];
	const uint64_t cda = code[na], cdb = code[nb];
	unsigned fla = fla0, flb = flb0;
	const double* tga = reinterpret_cast<const double*>(tiles + (uint32_t)(cda >> (3 * l)));
	const double* tgb = reinterpret_cast<const double*>(tiles + (uint32_t)(cdb >> (3 * l)));
	const unsigned ma = classify_pair_flat<DIRECT>(cda, l, tga, kscale, tris + 9ull * ta, fla);
	const unsigned mb = hasB ? classify_pair_flat<DIRECT>(cdb, l, tgb, kscale, tris + 9ull * tb, flb) : 0u;
	if (!last) {
		if (fla != fla0) pflags[pa] = (uint16_t)fla;
		hit[pa] = (uint8_t)ma;
		if (hasB) { if (flb != flb0) pflags[pb] = (uint16_t)flb; hit[pb] = (uint8_t)mb; }
	}
	if (ma && (!precheck || (mask[na] & ma) != ma)) atomicOr(reinterpret_cast<unsigned*>(mask) + (na >> 2), ma << (8 * (na & 3)));
	if (mb && (!precheck || (mask[nb] & mb) != mb)) atomicOr(reinterpret_cast<unsigned*>(mask) + (nb >> 2), mb << (8 * (nb & 3)));
}

// ------------------------------------------------------------------ flat stream, last two levels fused
// At the second-to-last level the children of a flat-stream pair are not emitted as pairs: the thread that owns the
// parent pair decides each hit child's 8 voxels on the spot (same exact box-axis tests, one level down) and ORs the
// voxel mask / first-touch triangle straight into the leaf level.  Saves writing and re-reading the bulk of the
// last-level pair list (the interior of every wall), which is the largest array of the whole build.
// STAR: the pair that carries its node's own first-touch index stores the first touch of its children plainly (see k_emit_pipe).
template <bool DIRECT, int MINB, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_flat_leaves(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                               const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                               const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                               int lc, double kscaleParent, const TileGeom* __restrict__ tiles, const float* __restrict__ tris,
                                                               const uint32_t* __restrict__ rootTri, uint8_t* __restrict__ cmask, uint32_t* __restrict__ ctstar, int onlyFlatKids, int precheck) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	// all four pair fields are fetched before the first test: one round trip instead of two (the kernel is latency bound)
	unsigned m = hit[p];
	const unsigned fl = pflags[p];
	const uint32_t t = ptri[p], n = pnode[p];
	if (!m) return;
	if (onlyFlatKids && !pair_is_fast(fl)) return;   // slow-stream parents: only those whose children join the flat stream
	const uint64_t cd = code[n];
	const unsigned nm = mask[n];
	const uint32_t base = childBase[n];
	const bool star = STAR && ctstar && tstar[n] == t;
	const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd >> (3 * (lc - 1))));
	const float* tp = tris + 9ull * rootTri[t];   // t itself (the root pair index) is what orders first touches
	unsigned lohi[3][2];
	flat_leaf_masks<DIRECT>(cd, lc - 1, tg, kscaleParent, tp, fl, lohi);
	// the children of one node are consecutive, so their voxel masks are consecutive bytes: OR them word by word
	unsigned* const words = reinterpret_cast<unsigned*>(cmask);
	uint32_t curWord = 0xFFFFFFFFu;
	unsigned acc = 0;
	while (m) {
		const int c = __ffs(m) - 1;
		m &= m - 1;
		const uint32_t child = base + __popc(nm & ((1u << c) - 1));
		const unsigned mc = lohi[0][(c >> 2) & 1] & lohi[1][(c >> 1) & 1] & lohi[2][c & 1];
		if ((child >> 2) != curWord) {
			if (acc && (!precheck || (words[curWord] & acc) != acc)) atomicOr(words + curWord, acc);
			curWord = child >> 2;
			acc = 0;
		}
		acc |= mc << (8 * (child & 3));
		// precheck: read before the atomic -- pays off only where many pairs share a node (upper levels); at the leaf
		// levels (~1.4 pairs per node) the read is a wasted round trip and the reduction goes out fire-and-forget
		if (!ctstar) continue;   // first touches of the leaf nodes are not tracked
		if (star) ctstar[child] = t;
		else if (!precheck || ctstar[child] > t) atomicMin(&ctstar[child], t);
	}
	if (acc && (!precheck || (words[curWord] & acc) != acc)) atomicOr(words + curWord, acc);
}

// Two parent pairs per thread (SVB_LEAVES_ILP=2; 1 selects the kernel above, the default is k_flat_leaves3 below): the pair fields and the node fields of both
// pairs are in flight before the first pair is decided -- the kernel is a chain of dependent gathers (pair -> node ->
// tile / triangle), and the loads a warp has outstanding decide its rate.
template <bool DIRECT, int MINB, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_flat_leaves2(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                                const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                                int lc, double kscaleParent, const TileGeom* __restrict__ tiles, const float* __restrict__ tris,
                                                                const uint32_t* __restrict__ rootTri, uint8_t* __restrict__ cmask, uint32_t* __restrict__ ctstar, int onlyFlatKids, int precheck) {
	const uint64_t p0 = (uint64_t)blockIdx.x * (2 * VX_THREADS) + threadIdx.x;
	unsigned m[2], fl[2], nm[2];
	uint32_t t[2], n[2], base[2], ts[2], tr[2];
	uint64_t cd[2];
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		const uint64_t p = p0 + (uint64_t)j * VX_THREADS;
		m[j] = 0; fl[j] = 0; t[j] = 0; n[j] = 0;
		if (p < P) { m[j] = hit[p]; fl[j] = pflags[p]; t[j] = ptri[p]; n[j] = pnode[p]; }
		if (onlyFlatKids && !pair_is_fast(fl[j])) m[j] = 0;
	}
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		cd[j] = 0; nm[j] = 0; base[j] = 0; ts[j] = 0; tr[j] = 0;
		if (m[j]) {
			cd[j] = code[n[j]]; nm[j] = mask[n[j]]; base[j] = childBase[n[j]]; tr[j] = rootTri[t[j]];
			if (STAR && ctstar) ts[j] = tstar[n[j]];
		}
	}
	unsigned* const words = reinterpret_cast<unsigned*>(cmask);
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		unsigned mm = m[j];
		if (!mm) continue;
		const bool star = STAR && ctstar && ts[j] == t[j];
		const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd[j] >> (3 * (lc - 1))));
		unsigned lohi[3][2];
		flat_leaf_masks<DIRECT>(cd[j], lc - 1, tg, kscaleParent, tris + 9ull * tr[j], fl[j], lohi);
		uint32_t curWord = 0xFFFFFFFFu;
		unsigned acc = 0;
		while (mm) {
			const int c = __ffs(mm) - 1;
			mm &= mm - 1;
			const uint32_t child = base[j] + __popc(nm[j] & ((1u << c) - 1));
			const unsigned mc = lohi[0][(c >> 2) & 1] & lohi[1][(c >> 1) & 1] & lohi[2][c & 1];
			if ((child >> 2) != curWord) {
				if (acc && (!precheck || (words[curWord] & acc) != acc)) atomicOr(words + curWord, acc);
				curWord = child >> 2;
				acc = 0;
			}
			acc |= mc << (8 * (child & 3));
			if (!ctstar) continue;
			if (star) ctstar[child] = t[j];
			else if (!precheck || ctstar[child] > t[j]) atomicMin(&ctstar[child], t[j]);
		}
		if (acc && (!precheck || (words[curWord] & acc) != acc)) atomicOr(words + curWord, acc);
	}
}

// ------------------------------------------------------------------ leaf-level output of the fused kernels
// vox: byte c = voxel mask of child c.  The children of a node are consecutive (rank = the number of lower children in the
// node's mask nm), so the bytes of the hit children are compacted into one 64-bit value and ORed into the leaf level's mask
// array as up to three aligned words; first touches of the children where they are tracked (star: this pair carries the
// node's own first touch, so it is every child's as well: plain store).
__device__ __forceinline__ void leaf_or_out(const uint64_t vox, unsigned m, const unsigned nm, const uint32_t base, const uint32_t t, const bool star,
                                            unsigned* __restrict__ words, uint32_t* __restrict__ ctstar, const int precheck) {
	uint64_t V = 0;
	while (m) {
		const int c = __ffs(m) - 1;
		m &= m - 1;
		const unsigned r = __popc(nm & ((1u << c) - 1));
		V |= ((vox >> (8 * c)) & 0xFFull) << (8 * r);
		if (!ctstar) continue;   // first touches of the leaf nodes are not tracked
		if (star) ctstar[base + r] = t;
		else if (!precheck || ctstar[base + r] > t) atomicMin(&ctstar[base + r], t);
	}
	const unsigned sh = 8 * (base & 3), lo = (unsigned)V, hi = (unsigned)(V >> 32);
	const unsigned a0 = lo << sh, a1 = __funnelshift_l(lo, hi, sh), a2 = __funnelshift_l(hi, 0u, sh);
	unsigned* const w = words + (base >> 2);
	// (read before the atomic only where many pairs share a node; at the leaf levels -- ~1.4 pairs per node -- the
	// reduction goes out fire-and-forget)
	if (a0 && (!precheck || (w[0] & a0) != a0)) atomicOr(w, a0);
	if (a1 && (!precheck || (w[1] & a1) != a1)) atomicOr(w + 1, a1);
	if (a2 && (!precheck || (w[2] & a2) != a2)) atomicOr(w + 2, a2);
}

// k_flat_leaves2 with the voxels of all children in one 64-bit word (flat_leaf_voxels) and the output above (default,
// SVB_LEAVES_ILP=3): ncu had k_flat_leaves2 issue bound (89 % issue-active), 55 % of its instructions in the per-child loop
// (the per-axis tables indexed by the child's bits compile to compare / select chains).
template <bool DIRECT, int MINB, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_flat_leaves3(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                                const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                                int lc, double kscaleParent, const TileGeom* __restrict__ tiles, const float* __restrict__ tris,
                                                                const uint32_t* __restrict__ rootTri, uint8_t* __restrict__ cmask, uint32_t* __restrict__ ctstar, int onlyFlatKids, int precheck) {
	const uint64_t p0 = (uint64_t)blockIdx.x * (2 * VX_THREADS) + threadIdx.x;
	unsigned m[2], fl[2], nm[2];
	uint32_t t[2], n[2], base[2], ts[2], tr[2];
	uint64_t cd[2];
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		const uint64_t p = p0 + (uint64_t)j * VX_THREADS;
		m[j] = 0; fl[j] = 0; t[j] = 0; n[j] = 0;
		if (p < P) { m[j] = hit[p]; fl[j] = pflags[p]; t[j] = ptri[p]; n[j] = pnode[p]; }
		if (onlyFlatKids && !pair_is_fast(fl[j])) m[j] = 0;
	}
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		cd[j] = 0; nm[j] = 0; base[j] = 0; ts[j] = 0; tr[j] = 0;
		if (m[j]) {
			cd[j] = code[n[j]]; nm[j] = mask[n[j]]; base[j] = childBase[n[j]]; tr[j] = rootTri[t[j]];
			if (STAR && ctstar) ts[j] = tstar[n[j]];
		}
	}
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		if (!m[j]) continue;
		const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd[j] >> (3 * (lc - 1))));
		const uint64_t vox = flat_leaf_voxels<DIRECT>(cd[j], lc - 1, tg, kscaleParent, tris + 9ull * tr[j], fl[j]);
		leaf_or_out(vox, m[j], nm[j], base[j], t[j], STAR && ctstar && ts[j] == t[j], reinterpret_cast<unsigned*>(cmask), ctstar, precheck);
	}
}

// ------------------------------------------------------------------ slow stream of a box mesh, last two levels fused
// Same idea for the pairs that are still in the slow stream at the second-to-last level when every triangle of the scene
// is flat (allFlat): the thread that owns the parent pair decides the 4 x 4 x 4 voxels under its node on the spot
// (slow_leaf_voxels, svb_classify.cuh: box axes exactly, the unsettled in-plane edge axes through the interval filter one
// level further down, the reference-order predicate for voxels within its margin) and ORs the voxel masks of the hit
// children into the leaf level.  The last level then has no pair list at all: no emit at the second-to-last level, no
// classify at the last (SVB_SLOW_LEAVES=0 restores both).
template <bool DIRECT, int MINB, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_slow_leaves(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                               const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                               const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                               int lc, double kscaleParent, const TileGeom* __restrict__ tiles, const float* __restrict__ tris,
                                                               const uint32_t* __restrict__ rootTri, uint8_t* __restrict__ cmask, uint32_t* __restrict__ ctstar,
                                                               unsigned long long* __restrict__ nExact, int precheck) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	unsigned m = hit[p];
	const unsigned fl = pflags[p];
	const uint32_t t = ptri[p], n = pnode[p];
	if (!m || !(fl & (7u << FL_FLAT))) return;   // (mixed scenes: the pairs of general triangles are emitted and classified at the last level)
	const uint64_t cd = code[n];
	const unsigned nm = mask[n];
	const uint32_t base = childBase[n];
	const bool star = STAR && ctstar && tstar[n] == t;
	const double* tg = reinterpret_cast<const double*>(tiles + (uint32_t)(cd >> (3 * (lc - 1))));
	const float* tp = tris + 9ull * rootTri[t];
	uint64_t ask;
	uint64_t vox = slow_leaf_voxels<DIRECT>(cd, lc - 1, tg, kscaleParent, tp, fl, m, ask);
	if (ask) {
		vox |= slow_leaf_exact<DIRECT>(ask, cd, lc - 1, tg, kscaleParent, tp);
		if (nExact) atomicAdd(nExact, (unsigned long long)__popcll(ask));
	}
	leaf_or_out(vox, m, nm, base, t, star, reinterpret_cast<unsigned*>(cmask), ctstar, precheck);
}

// ------------------------------------------------------------------ emit the child pairs
// One CTA per tile of parent pairs.  The pair arrays are kept as two streams: [0, nFlat) flat-stream pairs,
// [slowBase, ...) the others (stable partition: both streams stay sorted by triangle id).  SLOW = false: parents of
// the flat stream (all children flat, tile offsets offsA).  SLOW = true: parents of the other stream; offsA = tile
// offsets over all their children, offsB = over their flat children; flat children go to fastBase + B, the others to
// slowBase + (A - B).  First-touch triangle of the child nodes by atomicMin (pairs are sorted by triangle, so after
// the first touch the pre-check load filters almost every later atomic).
template <bool SLOW>
__global__ void __launch_bounds__(VX_THREADS) k_emit(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                     const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                     const uint64_t* __restrict__ offsA, const uint64_t* __restrict__ offsB, uint64_t fastBase, uint64_t slowBase,
                                                     const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase,
                                                     uint32_t* __restrict__ otri, uint32_t* __restrict__ onode, uint16_t* __restrict__ oflags, uint32_t* __restrict__ ctstar,
                                                     int skipFlat, int precheck) {
	__shared__ uint32_t s_tri[VX_THREADS * 8];
	__shared__ uint32_t s_node[VX_THREADS * 8];
	__shared__ uint16_t s_fl[VX_THREADS * 8];
	__shared__ uint32_t wsum[9];
	const uint64_t p0 = (uint64_t)blockIdx.x * VX_TILE;
	uint64_t runF, runS = 0;
	if (SLOW) { const uint64_t a = offsA[blockIdx.x], b = offsB[blockIdx.x]; runF = fastBase + b; runS = slowBase + (a - b); }
	else runF = fastBase + offsA[blockIdx.x];
	for (int ch = 0; ch < VX_CHUNKS; ++ch) {
		const uint64_t pb = p0 + (uint64_t)ch * VX_THREADS;
		if (pb >= P) break;
		const uint64_t p = pb + threadIdx.x;
		unsigned m = 0;
		uint32_t t = 0, n = 0, fl = 0;
		bool flatKids = true;
		if (p < P) {
			m = hit[p];
			fl = pflags[p]; t = ptri[p]; n = pnode[p];   // fetched unconditionally: one round trip instead of two
			if (m) {
				if (SLOW) flatKids = pair_is_fast(fl);
				if (SLOW && skipFlat && flatKids) m = 0;   // decided in place by k_flat_leaves (second-to-last level)
			}
		}
		const uint32_t cnt = __popc(m);
		uint32_t tot;
		const uint32_t ex = cta_excl_scan(flatKids ? cnt : (cnt << 16), wsum, &tot);   // low half: flat children, high half: others
		const uint32_t nFlat = tot & 0xFFFF, nSlow = tot >> 16;
		if (m) {
			uint32_t o = flatKids ? (ex & 0xFFFF) : nFlat + (ex >> 16);
			const unsigned nm = mask[n];
			const uint32_t base = childBase[n];
			unsigned mm = m;
			while (mm) {
				const int c = __ffs(mm) - 1;
				mm &= mm - 1;
				const uint32_t child = base + __popc(nm & ((1u << c) - 1));
				s_tri[o] = t;
				s_node[o] = child;
				s_fl[o] = (uint16_t)fl;
				++o;
				if (ctstar && (!precheck || ctstar[child] > t)) atomicMin(&ctstar[child], t);   // (ctstar == nullptr: first touches of the children are not tracked)
			}
		}
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < nFlat; i += VX_THREADS) {
			otri[runF + i] = s_tri[i];
			onode[runF + i] = s_node[i];
			oflags[runF + i] = s_fl[i];
		}
		if (SLOW) {
			for (uint32_t i = threadIdx.x; i < nSlow; i += VX_THREADS) {
				otri[runS + i] = s_tri[nFlat + i];
				onode[runS + i] = s_node[nFlat + i];
				oflags[runS + i] = s_fl[nFlat + i];
			}
		}
		runF += nFlat;
		runS += nSlow;
		__syncthreads();
	}
}

// Software-pipelined k_emit (round 1's default, now SVB_EMIT_WARP=0; SVB_EMIT_PIPE=0 selects the kernel above).  k_emit walks its tile in 8 chunks, and
// every chunk used to pay two dependent global round trips in series behind CTA-wide barriers (pair fields, then the
// node's mask / childBase gathered after the scan), which no other warp of the CTA can hide because all of them wait on
// the same barriers.  Here the pair fields are fetched two chunks ahead and the node fields one chunk ahead, so that by
// the time a chunk is scanned and staged everything it needs is already in registers; results are identical.
// REDOUT: the first-touch reductions leave with the coalesced copy-out (entry i = thread i: the children of one parent
// are consecutive nodes, so a warp's 32 atomics fall into a handful of 32-byte sectors) instead of from the staging loop
// (where the 32 lanes of a warp own 32 different parents, i.e. 32 different sectors per instruction).
//
// STAR: exactly one pair of a node carries the node's own first-touch index (t == tstar[n]; the pairs of
// a node belong to distinct triangles).  Every triangle that reaches a child also reaches its parent, so no pair of the
// child can carry a smaller index: the children of that one pair get their first touch with a PLAIN store.  Racing
// atomicMin's of the node's other pairs carry larger values and cannot change the outcome in either order (the word
// ends up t).  At the deep levels ~70 % of the pairs are such pairs (pairs/nodes ~ 1.4), and REDG issues at ~1.3
// cycles per LANE per SM -- the bound of this kernel -- whereas the coalesced stores cost a few wavefronts per warp.
template <bool SLOW, int MINB, bool REDOUT, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_emit_pipe(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                 const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                                 const uint64_t* __restrict__ offsA, const uint64_t* __restrict__ offsB, uint64_t fastBase, uint64_t slowBase,
                                                                 const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                                 uint32_t* __restrict__ otri, uint32_t* __restrict__ onode, uint16_t* __restrict__ oflags, uint32_t* __restrict__ ctstar,
                                                                 int skipFlat, int precheck) {
	__shared__ uint32_t s_tri[VX_THREADS * 8];
	__shared__ uint32_t s_node[VX_THREADS * 8];
	__shared__ uint16_t s_fl[VX_THREADS * 8];
	__shared__ uint8_t s_star[STAR ? VX_THREADS * 8 : 1];
	__shared__ uint32_t wsum[9];
	const uint64_t p0 = (uint64_t)blockIdx.x * VX_TILE;
	uint64_t runF, runS = 0;
	if (SLOW) { const uint64_t a = offsA[blockIdx.x], b = offsB[blockIdx.x]; runF = fastBase + b; runS = slowBase + (a - b); }
	else runF = fastBase + offsA[blockIdx.x];
	struct PairIn { unsigned m; uint32_t t, n, fl; };
	struct NodeIn { unsigned nm; uint32_t base, ts; };
	auto load_pair = [&](int ch) {
		PairIn r = {0u, 0u, 0u, 0u};
		const uint64_t p = p0 + (uint64_t)ch * VX_THREADS + threadIdx.x;
		if (ch < VX_CHUNKS && p < P) { r.m = hit[p]; r.fl = pflags[p]; r.t = ptri[p]; r.n = pnode[p]; }
		return r;
	};
	// drops the parents whose children are decided in place by k_flat_leaves, then gathers the node fields of the rest
	auto load_node = [&](PairIn& f) {
		NodeIn g = {0u, 0u, 0u};
		if (SLOW && skipFlat && pair_is_fast(f.fl)) f.m = 0;
		if (f.m) { g.nm = mask[f.n]; g.base = childBase[f.n]; if (STAR && ctstar) g.ts = tstar[f.n]; }   // (the parents' first touches only serve the children's)
		return g;
	};
	PairIn f0 = load_pair(0), f1 = load_pair(1);
	NodeIn g0 = load_node(f0);
	for (int ch = 0; ch < VX_CHUNKS; ++ch) {
		if (p0 + (uint64_t)ch * VX_THREADS >= P) break;
		const PairIn f2 = load_pair(ch + 2);
		const NodeIn g1 = load_node(f1);
		const unsigned m = f0.m;
		const bool flatKids = SLOW ? pair_is_fast(f0.fl) : true;
		const uint32_t cnt = __popc(m);
		uint32_t tot;
		const uint32_t ex = cta_excl_scan(flatKids ? cnt : (cnt << 16), wsum, &tot);   // low half: flat children, high half: others
		const uint32_t nFlat = tot & 0xFFFF, nSlow = tot >> 16;
		if (m) {
			uint32_t o = flatKids ? (ex & 0xFFFF) : nFlat + (ex >> 16);
			const uint32_t t = f0.t;
			unsigned mm = m;
			while (mm) {
				const int c = __ffs(mm) - 1;
				mm &= mm - 1;
				const uint32_t child = g0.base + __popc(g0.nm & ((1u << c) - 1));
				s_tri[o] = t;
				s_node[o] = child;
				s_fl[o] = (uint16_t)f0.fl;
				if (STAR) s_star[o] = (uint8_t)(g0.ts == t);
				++o;
				if (!REDOUT && ctstar && !(STAR && g0.ts == t) && (!precheck || ctstar[child] > t)) atomicMin(&ctstar[child], t);
			}
		}
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < nFlat; i += VX_THREADS) {
			const uint32_t t = s_tri[i], child = s_node[i];
			otri[runF + i] = t;
			onode[runF + i] = child;
			oflags[runF + i] = s_fl[i];
			if (!ctstar) continue;   // first touches of the children are not tracked (leaf level of a later batch, see voxelize_batch)
			if (STAR && s_star[i]) ctstar[child] = t;
			else if (REDOUT && (!precheck || ctstar[child] > t)) atomicMin(&ctstar[child], t);
		}
		if (SLOW) {
			for (uint32_t i = threadIdx.x; i < nSlow; i += VX_THREADS) {
				const uint32_t t = s_tri[nFlat + i], child = s_node[nFlat + i];
				otri[runS + i] = t;
				onode[runS + i] = child;
				oflags[runS + i] = s_fl[nFlat + i];
				if (!ctstar) continue;
				if (STAR && s_star[nFlat + i]) ctstar[child] = t;
				else if (REDOUT && (!precheck || ctstar[child] > t)) atomicMin(&ctstar[child], t);
			}
		}
		runF += nFlat;
		runS += nSlow;
		__syncthreads();
		f0 = f1; g0 = g1; f1 = f2;
	}
}

// Warp-granular emit (default since round 2; SVB_EMIT_WARP=0 selects k_emit_pipe).  The scans now deliver tile-relative
// offsets for every 8 pairs (rel, svb_prims.cu), so a warp knows where the children of its 32 pairs go without any CTA-wide
// scan: k_emit_pipe spent 4 CTA barriers per 256 pairs (two in the scan, two around the staging buffer), and with every
// warp of the CTA waiting on the same barriers nothing hid the gathers behind them (ncu: 4.8 - 5.6 warps per issue cycle
// stalled on the barrier, 7.4 - 9.0 on the scoreboard).  Here each warp walks its 8 groups of 32 pairs on its own: warp
// scan of the child counts (shuffles), children staged in the warp's private slice of shared memory, written out as
// contiguous runs, __syncwarp only.  Pair fields are fetched two groups ahead, node fields one ahead, as before.
// Output, first-touch handling (star stores / atomics with the copy-out) and the two-stream routing are unchanged.
template <bool SLOW, int MINB, bool STAR>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_emit_warp(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                 const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit,
                                                                 const uint64_t* __restrict__ offsA, const uint64_t* __restrict__ offsB, const uint32_t* __restrict__ rel,
                                                                 uint64_t fastBase, uint64_t slowBase,
                                                                 const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase, const uint32_t* __restrict__ tstar,
                                                                 uint32_t* __restrict__ otri, uint32_t* __restrict__ onode, uint16_t* __restrict__ oflags, uint32_t* __restrict__ ctstar,
                                                                 int skipFlat, int precheck) {
	constexpr int WARPS = VX_THREADS / 32, GROUPS = VX_TILE / VX_THREADS;   // 8 warps x 8 groups of 32 pairs = one 2048-pair tile per CTA
	__shared__ uint32_t s_tri[WARPS][256];
	__shared__ uint32_t s_node[WARPS][256];
	__shared__ uint16_t s_fl[WARPS][256];
	__shared__ uint8_t s_star[STAR ? WARPS : 1][STAR ? 256 : 1];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint64_t tile = blockIdx.x;
	const uint64_t pw = tile * VX_TILE + (uint64_t)w * (32 * GROUPS);   // first pair of this warp
	if (pw >= P) return;
	const uint64_t tA = offsA[tile], tB = SLOW ? offsB[tile] : 0;
	struct PairIn { unsigned m; uint32_t t, n, fl; };
	struct NodeIn { unsigned nm; uint32_t base, ts; };
	auto load_pair = [&](int g) {
		PairIn r = {0u, 0u, 0u, 0u};
		const uint64_t p = pw + (uint64_t)g * 32 + lane;
		if (g < GROUPS && p < P) { r.m = hit[p]; r.fl = pflags[p]; r.t = ptri[p]; r.n = pnode[p]; }
		return r;
	};
	auto load_node = [&](PairIn& f) {
		NodeIn q = {0u, 0u, 0u};
		if (SLOW && skipFlat && (skipFlat == 2 ? (f.fl & (7u << FL_FLAT)) != 0 : pair_is_fast(f.fl))) f.m = 0;   // decided in place by k_flat_leaves / (2: every flat triangle) k_slow_leaves
		if (f.m) { q.nm = mask[f.n]; q.base = childBase[f.n]; if (STAR && ctstar) q.ts = tstar[f.n]; }
		return q;
	};
	PairIn f0 = load_pair(0), f1 = load_pair(1);
	NodeIn g0 = load_node(f0);
	// the 8 relative offsets of this warp's groups, fetched once (lane g holds group g's)
	uint32_t relMine = 0;
	if (lane < GROUPS && pw + (uint64_t)lane * 32 < P) relMine = rel[(pw >> 3) + 4 * lane];
	uint32_t* const wt = s_tri[w];
	uint32_t* const wn = s_node[w];
	uint16_t* const wf = s_fl[w];
	uint8_t* const ws = s_star[STAR ? w : 0];
	for (int g = 0; g < GROUPS; ++g) {
		const uint64_t pg = pw + (uint64_t)g * 32;
		if (pg >= P) break;
		const PairIn f2 = load_pair(g + 2);
		const NodeIn g1 = load_node(f1);
		const uint32_t r = __shfl_sync(0xFFFFFFFFu, relMine, g);   // exclusive offsets (tile-relative) of the group's first pair: low half all children, high half the flat ones
		const unsigned m = f0.m;
		const bool flatKids = SLOW ? pair_is_fast(f0.fl) : true;
		const uint32_t cnt = __popc(m);
		const uint32_t inc = warp_incl_scan_u32(flatKids ? cnt : (cnt << 16), lane);   // low half: flat children, high half: the others (<= 256 each)
		const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
		const uint32_t ex = inc - (flatKids ? cnt : (cnt << 16));
		const uint32_t nFlat = tot & 0xFFFF, nSlow = tot >> 16;
		uint64_t runF, runS = 0;
		if (SLOW) { const uint64_t a = tA + (r & 0xFFFF), b = tB + (r >> 16); runF = fastBase + b; runS = slowBase + (a - b); }
		else runF = fastBase + tA + r;
		if (m) {
			uint32_t o = flatKids ? (ex & 0xFFFF) : nFlat + (ex >> 16);
			const uint32_t t = f0.t;
			unsigned mm = m;
			while (mm) {
				const int c = __ffs(mm) - 1;
				mm &= mm - 1;
				wt[o] = t;
				wn[o] = g0.base + __popc(g0.nm & ((1u << c) - 1));
				wf[o] = (uint16_t)f0.fl;
				if (STAR) ws[o] = (uint8_t)(g0.ts == t);
				++o;
			}
		}
		__syncwarp();
		for (uint32_t i = lane; i < nFlat; i += 32) {
			const uint32_t t = wt[i], child = wn[i];
			otri[runF + i] = t;
			onode[runF + i] = child;
			oflags[runF + i] = wf[i];
			if (!ctstar) continue;   // first touches of the children are not tracked
			if (STAR && ws[i]) ctstar[child] = t;
			else if (!precheck || ctstar[child] > t) atomicMin(&ctstar[child], t);
		}
		if (SLOW) {
			for (uint32_t i = lane; i < nSlow; i += 32) {
				const uint32_t t = wt[nFlat + i], child = wn[nFlat + i];
				otri[runS + i] = t;
				onode[runS + i] = child;
				oflags[runS + i] = wf[nFlat + i];
				if (!ctstar) continue;
				if (STAR && ws[nFlat + i]) ctstar[child] = t;
				else if (!precheck || ctstar[child] > t) atomicMin(&ctstar[child], t);
			}
		}
		__syncwarp();
		f0 = f1; g0 = g1; f1 = f2;
	}
}

// nodes the next level will hold, per tile (weights for cutting an oversized batch)
__global__ void __launch_bounds__(VX_THREADS) k_tile_weights(uint64_t N, const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, int l, uint32_t* __restrict__ w) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	unsigned m = mask[n];
	if (m) atomicAdd(&w[(uint32_t)(code[n] >> (3 * l))], (uint32_t)__popc(m));
}

__global__ void k_init_roots(uint32_t ntiles, uint64_t* code, uint32_t* tstar, const uint32_t* __restrict__ tileStart) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ntiles) { code[i] = i; tstar[i] = tileStart[i]; }
}

// root pairs sorted by (tile, triangle): key = tile_local << 32 | triangle
__global__ void __launch_bounds__(VX_THREADS) k_root_keys(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode, uint64_t* __restrict__ keys) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < P) keys[p] = ((uint64_t)pnode[p] << 32) | ptri[p];
}
__global__ void __launch_bounds__(VX_THREADS) k_root_split(uint64_t P, const uint64_t* __restrict__ keys, uint32_t* __restrict__ rootTri, uint32_t* __restrict__ ptri,
                                                            uint32_t* __restrict__ pnode, uint32_t* __restrict__ tileStart) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	const uint64_t k = keys[p];
	const uint32_t tile = (uint32_t)(k >> 32);
	rootTri[p] = (uint32_t)k;
	ptri[p] = (uint32_t)p;       // from here on a pair is identified by the index of its root pair
	pnode[p] = tile;
	if (p == 0 || (uint32_t)(keys[p - 1] >> 32) != tile) tileStart[tile] = (uint32_t)p;
}
// candidate (triangle, tile) pairs per tile over the WHOLE tile grid: bounds the tile-local triangle rank
__global__ void __launch_bounds__(VX_THREADS) k_tile_hist(const float* __restrict__ tris, uint64_t T, GridDesc g, const int* __restrict__ gridTile, uint32_t* __restrict__ hist) {
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	const float* p = tris + 9 * t;
	double mnx = fmin(fmin((double)p[0], (double)p[3]), (double)p[6]), mxx = fmax(fmax((double)p[0], (double)p[3]), (double)p[6]);
	double mny = fmin(fmin((double)p[1], (double)p[4]), (double)p[7]), mxy = fmax(fmax((double)p[1], (double)p[4]), (double)p[7]);
	double mnz = fmin(fmin((double)p[2], (double)p[5]), (double)p[8]), mxz = fmax(fmax((double)p[2], (double)p[5]), (double)p[8]);
	int ax, bx, ay, by, az, bz;
	tile_range(mnx, mxx, g.ox, g, ax, bx);
	tile_range(mny, mxy, g.oy, g, ay, by);
	tile_range(mnz, mxz, g.oz, g, az, bz);
	for (int x = ax; x <= bx; ++x)
		for (int y = ay; y <= by; ++y)
			for (int z = az; z <= bz; ++z) {
				int s = gridTile[((size_t)x * g.G + y) * g.G + z];
				if (s >= 0) atomicAdd(&hist[s], 1u);
			}
}
__global__ void __launch_bounds__(VX_THREADS) k_max_u32(uint64_t n, const uint32_t* __restrict__ v, uint32_t* __restrict__ out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t x = i < n ? v[i] : 0;
#pragma unroll
	for (int d = 16; d; d >>= 1) x = max(x, __shfl_xor_sync(0xFFFFFFFFu, x, d));
	if ((threadIdx.x & 31) == 0 && x) atomicMax(out, x);
}

// SVB_CLASSIFY=exact selects the unfiltered 8-lanes-per-pair kernel (verification of the filter)
bool classify_exact_only() {
	const char* e = getenv("SVB_CLASSIFY");
	return e && e[0] == 'e';
}

uint64_t read_u64(cudaStream_t s, const uint64_t* d) {
	uint64_t h = 0;
	SVB_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	return h;
}

}  // namespace

// every triangle flat (all three vertices share a coordinate bitwise)?  0 in *notFlat if so
__global__ void __launch_bounds__(VX_THREADS) k_all_flat(const float* __restrict__ tris, uint64_t T, uint32_t* __restrict__ notFlat) {
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	const float* p = tris + 9 * t;
	const bool flat = (p[0] == p[3] && p[3] == p[6]) || (p[1] == p[4] && p[4] == p[7]) || (p[2] == p[5] && p[5] == p[8]);
	if (!flat) *notFlat = 1;
}

// ------------------------------------------------------------------ host drivers
static GridDesc grid_desc(const TileGridHost& grid) {
	GridDesc g;
	g.ox = grid.ox; g.oy = grid.oy; g.oz = grid.oz;
	g.inv_cell = 1.0 / grid.cell;
	// a sub-octree cube is rebuilt from a float-narrowed box (geom_octree.cpp:177-184): it can stick out of its
	// nominal grid cell by up to one float ulp of the largest coordinate; the margin must dominate that
	const double ext = grid.cell * (double)grid.G;
	const double maxAbs = std::max(std::max(std::max(fabs(grid.ox), fabs(grid.ox + ext)), std::max(fabs(grid.oy), fabs(grid.oy + ext))),
	                               std::max(fabs(grid.oz), fabs(grid.oz + ext)));
	g.margin = grid.cell * 1e-6 + maxAbs * 4.8e-7;   // 4.8e-7 = 4 * 2^-23
	g.G = grid.G;
	for (int k = 0; k < 3; ++k) { g.lo[k] = 0; g.hi[k] = grid.G - 1; }
	return g;
}

bool all_triangles_flat(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T) {
	if (!T) return false;
	DevBuf<uint32_t> flag(pool, 1);
	flag.zero();
	k_all_flat<<<blocks_for(T, VX_THREADS), VX_THREADS, 0, s>>>(d_tris, T, flag.p);
	SVB_KERNEL_CHECK();
	uint32_t h = 1;
	SVB_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	return h == 0;
}

uint32_t max_candidates_per_tile(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid, const int* d_gridTile, uint64_t nTiles) {
	if (!T || !nTiles) return 0;
	DevBuf<uint32_t> hist(pool, nTiles), mx(pool, 1);
	hist.zero();
	mx.zero();
	k_tile_hist<<<blocks_for(T, VX_THREADS), VX_THREADS, 0, s>>>(d_tris, T, grid_desc(grid), d_gridTile, hist.p);
	SVB_KERNEL_CHECK();
	k_max_u32<<<blocks_for(nTiles, VX_THREADS), VX_THREADS, 0, s>>>(nTiles, hist.p, mx.p);
	SVB_KERNEL_CHECK();
	uint32_t h = 0;
	SVB_CUDA(cudaMemcpyAsync(&h, mx.p, 4, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	return h;
}

void make_root_pairs(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid,
                     const int* d_gridTile, const int* d_selPos, uint32_t posFirst, uint32_t ntiles, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode,
                     DevBuf<uint32_t>& rootTri, DevBuf<uint32_t>& tileStart, uint64_t& P, const int cellLo[3], const int cellHi[3]) {
	GridDesc g = grid_desc(grid);
	if (cellLo && cellHi) for (int k = 0; k < 3; ++k) { g.lo[k] = std::max(cellLo[k], 0); g.hi[k] = std::min(cellHi[k], grid.G - 1); }
	DevBuf<uint32_t> cnt(pool, T);
	DevBuf<uint64_t> tot(pool, 1);
	unsigned nb = blocks_for(T, VX_THREADS);
	k_candidates<false><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, d_selPos, (int)posFirst, (int)ntiles, cnt.p, nullptr, nullptr);
	SVB_KERNEL_CHECK();
	scan_u32(s, pool, cnt.p, T, cnt.p, tot.p);
	P = read_u64(s, tot.p);
	if (P >= 0xFFFFFFF0ull) throw BatchTooBig();
	ptri.reset(pool, P);
	pnode.reset(pool, P);
	k_candidates<true><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, d_selPos, (int)posFirst, (int)ntiles, cnt.p, ptri.p, pnode.p);
	SVB_KERNEL_CHECK();
	// Sort the root pairs by (tile, triangle).  A pair and all its descendants are then identified by the index q of
	// their root pair: triangle = rootTri[q], and q - tileStart[tile] is the rank of the triangle among the tile's
	// candidates -- the compact, order-preserving stand-in for the triangle id inside order keys (svb_dedup.cu).
	rootTri.reset(pool, P ? P : 1);
	tileStart.reset(pool, ntiles ? ntiles : 1);
	tileStart.zero();
	if (P) {
		DevBuf<uint64_t> keys(pool, P);
		DevBuf<uint32_t> dummy(pool, P);   // payload of the pair sort: unused, zeroed so that the sort never reads uninitialised memory
		dummy.zero();
		unsigned pb = blocks_for(P, VX_THREADS);
		k_root_keys<<<pb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, keys.p);
		SVB_KERNEL_CHECK();
		int tileBits = 1;
		while ((ntiles - 1) >> tileBits) ++tileBits;
		if (ntiles > 1) radix_sort_pairs(s, pool, keys.p, dummy.p, P, 32 + tileBits);   // emitted per triangle: already sorted when there is one tile
		k_root_split<<<pb, VX_THREADS, 0, s>>>(P, keys.p, rootTri.p, ptri.p, pnode.p, tileStart.p);
		SVB_KERNEL_CHECK();
	}
}

// ---- root pairs of ALL sub-octrees of a build, binned once
namespace {
// tileOf is ascending: begin[x] = first q of tile x (an empty tile gets the begin of its successor); begin[nSel] = P
__global__ void __launch_bounds__(VX_THREADS) k_tile_begin(uint64_t P, uint32_t nSel, const uint32_t* __restrict__ tileOf, uint32_t* __restrict__ begin) {
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	const int64_t t = tileOf[p], prev = p ? (int64_t)tileOf[p - 1] : -1;
	for (int64_t x = prev + 1; x <= t; ++x) begin[x] = (uint32_t)p;
	if (p == P - 1) for (int64_t x = t + 1; x <= (int64_t)nSel; ++x) begin[x] = (uint32_t)P;
}
__global__ void __launch_bounds__(VX_THREADS) k_batch_roots(uint64_t Pb, uint32_t a, const uint32_t* __restrict__ tileOf, uint32_t* __restrict__ ptri, uint32_t* __restrict__ pnode) {
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= Pb) return;
	ptri[p] = (uint32_t)p;          // a pair is identified by the index of its root pair inside the batch
	pnode[p] = tileOf[p] - a;
}
__global__ void __launch_bounds__(VX_THREADS) k_batch_tile_start(uint32_t nt, uint32_t base, const uint32_t* __restrict__ begin, uint32_t* __restrict__ tileStart) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nt) tileStart[i] = begin[i] - base;
}
}  // namespace

void make_root_pairs_all(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid, const int* d_gridTile, const int* d_selPos,
                         uint32_t nSel, RootPairsAll& R, const int cellLo[3], const int cellHi[3]) {
	R = RootPairsAll();
	DevBuf<uint32_t> ptri, tileStart;
	make_root_pairs(s, pool, d_tris, T, grid, d_gridTile, d_selPos, 0, nSel, ptri, R.tileOf, R.rootTri, tileStart, R.P, cellLo, cellHi);   // throws BatchTooBig beyond 2^32 pairs
	ptri.release();
	tileStart.release();
	R.nSel = nSel;
	R.tileBegin.reset(pool, (uint64_t)nSel + 1);
	R.tileBegin.zero();
	if (R.P) {
		k_tile_begin<<<blocks_for(R.P, VX_THREADS), VX_THREADS, 0, s>>>(R.P, nSel, R.tileOf.p, R.tileBegin.p);
		SVB_KERNEL_CHECK();
	}
	R.hTileBegin.resize((size_t)nSel + 1);
	SVB_CUDA(cudaMemcpyAsync(R.hTileBegin.data(), R.tileBegin.p, ((uint64_t)nSel + 1) * 4, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	R.valid = true;
}

void batch_root_pairs(cudaStream_t s, Pool& pool, const RootPairsAll& R, uint32_t a, uint32_t nt, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode,
                      const uint32_t*& rootTri, DevBuf<uint32_t>& tileStart, uint64_t& P) {
	const uint32_t base = R.hTileBegin[a];
	P = R.hTileBegin[a + nt] - base;
	ptri.reset(pool, P ? P : 1);
	pnode.reset(pool, P ? P : 1);
	tileStart.reset(pool, nt ? nt : 1);
	rootTri = R.rootTri.p + base;
	if (P) {
		k_batch_roots<<<blocks_for(P, VX_THREADS), VX_THREADS, 0, s>>>(P, a, R.tileOf.p + base, ptri.p, pnode.p);
		SVB_KERNEL_CHECK();
	}
	k_batch_tile_start<<<blocks_for(nt, VX_THREADS), VX_THREADS, 0, s>>>(nt, base, R.tileBegin.p + a, tileStart.p);
	SVB_KERNEL_CHECK();
}

void voxelize_batch(cudaStream_t s, Pool& pool, const float* d_tris, const TileGeom* d_tiles, uint32_t ntiles, int Lt,
                    DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, const uint32_t* rootTri, const uint32_t* tileStart, uint64_t P,
                    uint64_t budget_bytes, uint64_t nodeCap, std::vector<BatchLevel>& lv, uint64_t& pairsTotal, uint64_t* d_nExact, bool directCentre, bool allFlat, int untracked, ProfHook* prof,
                    DevBuf<uint32_t>* leafLastTri) {
	if (const char* e = getenv("SVB_CENTRE")) directCentre = directCentre && e[0] != 'c';   // SVB_CENTRE=chain: always replay the chain
	lv.clear();
	lv.resize(Lt);
	lv[0].n = ntiles;
	lv[0].code.reset(pool, (uint64_t)ntiles + 2);   // (+2: k_children_tma rounds its last bulk load up to 16 bytes)
	lv[0].tstar.reset(pool, ntiles);
	k_init_roots<<<blocks_for(ntiles, 256), 256, 0, s>>>(ntiles, lv[0].code.p, lv[0].tstar.p, tileStart);
	SVB_KERNEL_CHECK();
	DevBuf<uint64_t> tot(pool, 4);   // device totals of the four scans of a level, read back together
	pairsTotal = 0;
	// Pair streams: [0, F) fast pairs, [Fa, Fa + S) the others (Fa = F rounded up to 16 so that both streams start
	// on an aligned address in every pair array); the root pairs are all "slow" (no axis is settled yet).
	uint64_t F = 0, Fa = 0, S = P;
	DevBuf<uint16_t> pflags(pool, P + 16);   // settled-axis flags per pair (svb_classify.cuh)
	pflags.zero();
	const bool exactOnly = classify_exact_only() || leafLastTri != nullptr;   // attribute builds decide every voxel with the reference-order predicate, one lane per voxel
	// k_emit variant: 0 = one chunk at a time, else software-pipelined (read per batch, not cached: A/B inside one process)
	const bool childrenPipe = [] { const char* e = getenv("SVB_CHILDREN_PIPE"); return !(e && e[0] == '0'); }();
	const bool childrenTma = childrenPipe && [] { const char* e = getenv("SVB_CHILDREN_TMA"); return !(e && e[0] == '0'); }();   // cp.async.bulk + mbarrier pipeline (k_children_tma)
	const int emitPipe = [] { const char* e = getenv("SVB_EMIT_PIPE"); return e ? atoi(e) : 8; }();
	const bool emitWarp = emitPipe != 0 && [] { const char* e = getenv("SVB_EMIT_WARP"); return !(e && e[0] == '0'); }();   // warp-granular emit (k_emit_warp)
	const bool starStore = [] { const char* e = getenv("SVB_STAR_STORE"); return !(e && e[0] == '0'); }();   // first touch of the children of a node's own first-touch pair by plain store
	const bool emitRedOut = [] { const char* e = getenv("SVB_EMIT_REDOUT"); return !(e && e[0] == '0'); }();   // the remaining atomics leave with the copy-out too (measured: 667 vs 672 ms voxelize with the star stores; without them the staging loop is the better place)
	static const int occFlat = [] { const char* e = getenv("SVB_VX_OCC_FLAT"); return e ? atoi(e) : 6; }();   // 6 or 8 CTAs/SM (40 / 32 registers, spills) beat 5 on B200: the kernel is latency bound
	static const int occ = [] { const char* e = getenv("SVB_VX_OCC"); return e ? atoi(e) : 5; }();   // CTAs/SM of the slow classify kernel: 5 (48 registers, some spills) measured best on B200
	for (int l = 0; l < Lt; ++l) {
		BatchLevel& L = lv[l];
		if (!L.mask.p) {   // (the leaf level's mask is created one level early, see k_flat_leaves)
			L.mask.reset(pool, (L.n + 3 + 16) & ~3ull);
			L.mask.zero();
		}
		const int last = (l == Lt - 1) ? 1 : 0;
		DevBuf<uint8_t> hit(pool, last ? 16 : Fa + S + 16);
		if (F + S > (1ull << 59)) throw Error(SVB_ERANGE, "too many pairs");
		const double kscale = ldexp(1.0, -(l + 2));
		static const int forcePre = [] { const char* e = getenv("SVB_PRECHECK"); return e ? atoi(e) : -1; }();
		static const uint64_t preRatio10 = [] { const char* e = getenv("SVB_PRECHECK_RATIO10"); return (uint64_t)(e ? atoi(e) : 20); }();
		const int precheck = forcePre >= 0 ? forcePre : (10 * (F + S) > preRatio10 * L.n ? 1 : 0);   // read-before-atomic only where many pairs share a node
		if (F) {
			const int ilp = [] { const char* e = getenv("SVB_FAST_ILP"); return e ? atoi(e) : 2; }();
			if (ilp >= 2) {
				unsigned nb = blocks_for(F, 2 * VX_THREADS);
				if (directCentre) k_classify_fast2<true><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, L.code.p, l, kscale, last, d_tiles, d_tris, rootTri, hit.p, L.mask.p, precheck);
				else k_classify_fast2<false><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, L.code.p, l, kscale, last, d_tiles, d_tris, rootTri, hit.p, L.mask.p, precheck);
			} else {
				unsigned nb = blocks_for(F, VX_THREADS);
				if (directCentre) k_classify_fast<true><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, L.code.p, l, kscale, last, d_tiles, d_tris, rootTri, hit.p, L.mask.p, precheck);
				else k_classify_fast<false><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, L.code.p, l, kscale, last, d_tiles, d_tris, rootTri, hit.p, L.mask.p, precheck);
			}
			SVB_KERNEL_CHECK();
		}
		if (S) {
			const uint32_t* st = ptri.p + Fa; const uint32_t* sn = pnode.p + Fa; uint16_t* sf = pflags.p + Fa; uint8_t* sh = hit.p + (last ? 0 : Fa);
			if (exactOnly) {
				uint32_t* lt = nullptr;
				if (last && leafLastTri) { leafLastTri->reset(pool, L.n * 8 + 8); leafLastTri->zero(); lt = leafLastTri->p; }
				k_classify<<<blocks_for(S * 8, VX_THREADS), VX_THREADS, 0, s>>>(S, st, sn, L.code.p, l, d_tiles, d_tris, rootTri, sh, L.mask.p, last, lt);
			}
			else {
				unsigned nb = blocks_for(S, VX_THREADS);
#define SVB_LAUNCH_CF3(OCC, DIR, FLAT, LST) k_classify_filtered<OCC, DIR, FLAT, LST><<<nb, VX_THREADS, 0, s>>>(S, st, sn, sf, L.code.p, l, kscale, last, d_tiles, d_tris, rootTri, sh, L.mask.p, (unsigned long long*)d_nExact, precheck)
#define SVB_LAUNCH_CF2(OCC, DIR, FLAT) do { if (last) SVB_LAUNCH_CF3(OCC, DIR, FLAT, true); else SVB_LAUNCH_CF3(OCC, DIR, FLAT, false); } while (0)
#define SVB_LAUNCH_CF(OCC, FLAT) do { if (directCentre) SVB_LAUNCH_CF2(OCC, true, FLAT); else SVB_LAUNCH_CF2(OCC, false, FLAT); } while (0)
				if (allFlat) {   // box meshes: the general edge / plane filter is compiled out
					if (occFlat >= 8) SVB_LAUNCH_CF(8, true); else SVB_LAUNCH_CF(6, true);
				} else {
					if (occ >= 6) SVB_LAUNCH_CF(6, false); else if (occ == 5) SVB_LAUNCH_CF(5, false); else SVB_LAUNCH_CF(4, false);
				}
#undef SVB_LAUNCH_CF
#undef SVB_LAUNCH_CF2
#undef SVB_LAUNCH_CF3
			}
			SVB_KERNEL_CHECK();
		}
		pairsTotal += F + S;
		if (getenv("SVB_VX_STATS")) fprintf(stderr, "[vx-stats] tiles %u level %d/%d nodes %llu flat pairs %llu slow pairs %llu\n", ntiles, l, Lt - 1,
		                                    (unsigned long long)L.n, (unsigned long long)F, (unsigned long long)S);
		if (last) break;
		// children of the nodes, child pairs of the pairs: tile-granular scans, one read-back
		DevBuf<uint64_t> nodeOffs, offF, offFB, offS, offSF;
		DevBuf<uint32_t> relF, relS;   // tile-relative offsets per 8 pairs (k_emit_warp)
		const bool scanMulti = [] { const char* e = getenv("SVB_SCAN_MULTI"); return !(e && e[0] == '0') && !(getenv("SVB_SCAN_WIDE") && getenv("SVB_SCAN_WIDE")[0] == '0'); }();
		// second-to-last level: the children of the flat stream are decided in place (k_flat_leaves), and so are the children of
		// the slow-stream pairs of flat triangles (k_slow_leaves; all of them in a box mesh, the flat ones in a mixed scene --
		// there the scans count them apart and the emit drops them)
		const bool slowLeaves = [] { const char* e = getenv("SVB_SLOW_LEAVES"); return !(e && e[0] == '0'); }() && (l == Lt - 2) && !exactOnly && !getenv("SVB_NO_FUSE");
		const bool fuseAll = slowLeaves && allFlat;
		const bool fuseMixed = slowLeaves && !allFlat && emitWarp && S != 0 && [] { const char* e = getenv("SVB_SLOW_LEAVES_MIXED"); return !(e && e[0] == '0'); }();
		if (scanMulti) scan_level_tiles(s, pool, L.mask.p, L.n, hit.p, F, hit.p + Fa, pflags.p + Fa, S, nodeOffs, offF, offS, offSF, emitWarp ? &relF : nullptr, emitWarp ? &relS : nullptr, tot.p, fuseMixed ? 1 : 0);
		else {
			scan_tiles_popc8(s, pool, L.mask.p, L.n, nodeOffs, tot.p + 0);
			scan_tiles_popc8(s, pool, hit.p, F, offF, tot.p + 1, emitWarp ? &relF : nullptr);
			scan_tiles_pairs(s, pool, hit.p + Fa, pflags.p + Fa, S, offS, offSF, tot.p + 2, tot.p + 3, emitWarp ? &relS : nullptr, fuseMixed ? 1 : 0);
		}
		uint64_t h[4];
		SVB_CUDA(cudaMemcpyAsync(h, tot.p, 32, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		const uint64_t Nn = h[0], cF = h[1], cS = h[2], cSF = h[3];
		// second-to-last level: the children of the flat stream are decided in place (k_flat_leaves), not emitted
		static const bool fuseSlowKids = [] { const char* e = getenv("SVB_FUSE_SLOW"); return e ? e[0] != '0' : false; }();   // measured slower on B200 (850 vs 823 ms): off
		// (fuseMixed: cSF counts the children of flat-triangle pairs -- decided in place -- instead of the flat-stream children)
		const bool fuseFlat = (l == Lt - 2) && ((fuseAll || fuseMixed) ? (F + S) != 0 : fuseSlowKids ? (cF + cSF) != 0 : F != 0) && !getenv("SVB_NO_FUSE");
		const bool fuseS = fuseFlat && fuseSlowKids && !fuseAll && !fuseMixed;   // also decide the flat children of slow-stream parents in place
		const uint64_t cFe = fuseFlat ? 0 : cF;
		const uint64_t Fn = (fuseAll || fuseMixed) ? 0 : cFe + (fuseS ? 0 : cSF), Fan = (Fn + 15) & ~15ull, Sn = fuseAll ? 0 : cS - cSF, Pn = Fan + Sn;
		const int precheckKids = forcePre >= 0 ? forcePre : (10 * (cF + cS) > preRatio10 * Nn ? 1 : 0);
		{
			// will this batch fit all the way down?  Surfaces grow ~4x per level; use the observed ratio.
			const int remaining = (Lt - 1) - (l + 1);
			double growth = (l >= 2 && L.n) ? (double)Nn / (double)L.n : 4.0;
			if (growth < 2.0) growth = 2.0;
			if (growth > 5.0) growth = 5.0;
			double estLeaf = (double)Nn;
			for (int r = 0; r < remaining; ++r) estLeaf *= growth;
			bool hard = Nn >= 0xFFFFFFF0ull || Pn >= 0xFFFFFFF0ull || (budget_bytes && pool.live + Nn * 13 + Pn * 9 > budget_bytes);
			bool predicted = ntiles > 1 && l >= 2 && nodeCap && estLeaf > (double)nodeCap;
			if (hard || predicted) {
				BatchTooBig e;
				e.level = l + 1; e.remaining = remaining; e.growth = growth;
				DevBuf<uint32_t> w(pool, ntiles);
				w.zero();
				k_tile_weights<<<blocks_for(L.n, VX_THREADS), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, l, w.p);
				SVB_KERNEL_CHECK();
				e.weight.resize(ntiles);
				SVB_CUDA(cudaMemcpyAsync(e.weight.data(), w.p, ntiles * 4ull, cudaMemcpyDeviceToHost, s));
				SVB_CUDA(cudaStreamSynchronize(s));
				throw e;
			}
		}
		BatchLevel& C = lv[l + 1];
		C.n = Nn;
		C.code.reset(pool, Nn + 2);   // (+2: see lv[0].code)
		// first touches of the leaf nodes only matter for voxel masks the dedup table does not know yet: a later batch
		// goes without them (leafTstar == false; the caller re-runs the batch in the rare case that it does meet one)
		const bool trackKids = (l + 1 < Lt - untracked);   // untracked = 1: the leaf level, 2: the 4^3 level above it as well
		if (trackKids) {
			C.tstar.reset(pool, Nn);
			C.tstar.fill_ff();
		}
		L.childBase.reset(pool, L.n);
		// algorithmic bytes: 9 B read (code, mask) + 4 B childBase written per node, 8 B code written per child node
		const int pidC = prof ? prof->begin("children", (uint32_t)l, L.n) : -1;
		if (childrenTma) k_children_tma<<<blocks_for(L.n, VX_TILE), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, nodeOffs.p, L.childBase.p, C.code.p);
		else if (childrenPipe) k_children<true><<<blocks_for(L.n, VX_TILE), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, nodeOffs.p, L.childBase.p, C.code.p);
		else k_children<false><<<blocks_for(L.n, VX_TILE), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, nodeOffs.p, L.childBase.p, C.code.p);
		SVB_KERNEL_CHECK();
		if (prof) prof->end(pidC, Nn, 13.0 * (double)L.n + 8.0 * (double)Nn);
		DevBuf<uint32_t> ntri(pool, Pn + 16), nnode(pool, Pn + 16);
		DevBuf<uint16_t> nflags(pool, Pn + 16);
		if (fuseFlat) {
			C.mask.reset(pool, (Nn + 3 + 16) & ~3ull);
			C.mask.zero();
			static const int occLeaves = [] { const char* e = getenv("SVB_VX_OCC_LEAVES"); return e ? atoi(e) : 8; }();   // 8 CTAs/SM measured best
#define SVB_LAUNCH_FL(DIR, N, OFF, ONLY) if (occLeaves >= 8) SVB_LAUNCH_FL2(DIR, 8, N, OFF, ONLY); else SVB_LAUNCH_FL2(DIR, 6, N, OFF, ONLY)
#define SVB_LAUNCH_FL2(DIR, MB, N, OFF, ONLY) if (starStore) SVB_LAUNCH_FL3(DIR, MB, true, N, OFF, ONLY); else SVB_LAUNCH_FL3(DIR, MB, false, N, OFF, ONLY)
#define SVB_FL_ARGS(N, OFF, ONLY) N, ptri.p + (OFF), pnode.p + (OFF), pflags.p + (OFF), hit.p + (OFF), \
			L.code.p, L.mask.p, L.childBase.p, L.tstar.p, l + 1, kscale, d_tiles, d_tris, rootTri, C.mask.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), ONLY, precheckKids
			const int leavesIlp = [] { const char* e = getenv("SVB_LEAVES_ILP"); return e ? atoi(e) : 3; }();   // 3: k_flat_leaves3, 2: k_flat_leaves2, 1: k_flat_leaves
#define SVB_LAUNCH_FL3(DIR, MB, ST, N, OFF, ONLY) do { if (leavesIlp >= 3) k_flat_leaves3<DIR, (MB >= 8 ? 6 : 5), ST><<<blocks_for(N, 2 * VX_THREADS), VX_THREADS, 0, s>>>(SVB_FL_ARGS(N, OFF, ONLY)); \
			else if (leavesIlp >= 2) k_flat_leaves2<DIR, (MB >= 8 ? 6 : 5), ST><<<blocks_for(N, 2 * VX_THREADS), VX_THREADS, 0, s>>>(SVB_FL_ARGS(N, OFF, ONLY)); \
			else k_flat_leaves<DIR, MB, ST><<<blocks_for(N, VX_THREADS), VX_THREADS, 0, s>>>(SVB_FL_ARGS(N, OFF, ONLY)); } while (0)
			if (F) { if (directCentre) SVB_LAUNCH_FL(true, F, 0, 0); else SVB_LAUNCH_FL(false, F, 0, 0); SVB_KERNEL_CHECK(); }
			if (fuseS && S && cSF) { if (directCentre) SVB_LAUNCH_FL(true, S, Fa, 1); else SVB_LAUNCH_FL(false, S, Fa, 1); SVB_KERNEL_CHECK(); }
#undef SVB_LAUNCH_FL
#undef SVB_LAUNCH_FL2
#undef SVB_LAUNCH_FL3
#undef SVB_FL_ARGS
			if ((fuseAll || fuseMixed) && S) {
#define SVB_SL_ARGS S, ptri.p + Fa, pnode.p + Fa, pflags.p + Fa, hit.p + Fa, L.code.p, L.mask.p, L.childBase.p, L.tstar.p, l + 1, kscale, d_tiles, d_tris, rootTri, \
			C.mask.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), (unsigned long long*)d_nExact, precheckKids
				const unsigned nb = blocks_for(S, VX_THREADS);
				const int occSL = [] { const char* e = getenv("SVB_VX_OCC_SL"); return e ? atoi(e) : 5; }();   // CTAs/SM: 3 (80 registers), 4 (64), 5 (48), 6 (40); measured on the 16K^3 city: 526.6 / 508.5 / 504.3 ms voxelize
#define SVB_LAUNCH_SL2(DIR, MB) do { if (starStore) k_slow_leaves<DIR, MB, true><<<nb, VX_THREADS, 0, s>>>(SVB_SL_ARGS); else k_slow_leaves<DIR, MB, false><<<nb, VX_THREADS, 0, s>>>(SVB_SL_ARGS); } while (0)
#define SVB_LAUNCH_SL(DIR) do { if (occSL >= 6) SVB_LAUNCH_SL2(DIR, 6); else if (occSL == 5) SVB_LAUNCH_SL2(DIR, 5); else if (occSL == 4) SVB_LAUNCH_SL2(DIR, 4); else SVB_LAUNCH_SL2(DIR, 3); } while (0)
				if (directCentre) SVB_LAUNCH_SL(true); else SVB_LAUNCH_SL(false);
#undef SVB_LAUNCH_SL
#undef SVB_LAUNCH_SL2
#undef SVB_SL_ARGS
				SVB_KERNEL_CHECK();
			}
			pairsTotal += cF + ((fuseS || fuseMixed) ? cSF : 0) + (fuseAll ? cS : 0);   // decided here instead of as pairs of the last level
		} else if (F) {
#define SVB_EMIT_ARGS_F(...) F, ptri.p, pnode.p, pflags.p, hit.p, offF.p, nullptr, 0, 0, L.mask.p, L.childBase.p, __VA_ARGS__ ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), 0, precheckKids
			const unsigned nb = blocks_for(F, VX_TILE);
			// algorithmic bytes: 11 B pair + 9 B node fields (mask, childBase, t*) read per parent pair, 10 B pair written
			// (+ 4 B first touch when tracked) per child pair
			const int pidE = prof ? prof->begin("emit", (uint32_t)l, F) : -1;
			if (emitPipe == 0) k_emit<false><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_F());
			else if (emitWarp && starStore) k_emit_warp<false, 8, true><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, hit.p, offF.p, nullptr, relF.p, 0, 0, L.mask.p, L.childBase.p, L.tstar.p, ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), 0, precheckKids);
			else if (emitWarp) k_emit_warp<false, 8, false><<<nb, VX_THREADS, 0, s>>>(F, ptri.p, pnode.p, pflags.p, hit.p, offF.p, nullptr, relF.p, 0, 0, L.mask.p, L.childBase.p, L.tstar.p, ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), 0, precheckKids);
			else if (!emitRedOut && !starStore) k_emit_pipe<false, 8, false, false><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_F(L.tstar.p,));
			else if (!emitRedOut) k_emit_pipe<false, 8, false, true><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_F(L.tstar.p,));
			else if (!starStore) k_emit_pipe<false, 8, true, false><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_F(L.tstar.p,));
			else k_emit_pipe<false, 8, true, true><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_F(L.tstar.p,));
#undef SVB_EMIT_ARGS_F
			SVB_KERNEL_CHECK();
			if (prof) prof->end(pidE, cF, 20.0 * (double)F + (trackKids ? 14.0 : 10.0) * (double)cF);
		}
		if (S && !fuseAll) {
#define SVB_EMIT_ARGS_S(...) S, ptri.p + Fa, pnode.p + Fa, pflags.p + Fa, hit.p + Fa, offS.p, offSF.p, cFe, Fan, L.mask.p, L.childBase.p, __VA_ARGS__ ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), fuseS ? 1 : 0, precheckKids
			const unsigned nb = blocks_for(S, VX_TILE);
			const uint64_t kids = (fuseS || fuseMixed) ? cS - cSF : cS;
			const int pidE = prof ? prof->begin("emit", (uint32_t)l, S) : -1;
			if (emitPipe == 0) k_emit<true><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_S());
			else if (emitWarp && starStore) k_emit_warp<true, 8, true><<<nb, VX_THREADS, 0, s>>>(S, ptri.p + Fa, pnode.p + Fa, pflags.p + Fa, hit.p + Fa, offS.p, offSF.p, relS.p, cFe, Fan, L.mask.p, L.childBase.p, L.tstar.p, ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), fuseMixed ? 2 : fuseS ? 1 : 0, precheckKids);
			else if (emitWarp) k_emit_warp<true, 8, false><<<nb, VX_THREADS, 0, s>>>(S, ptri.p + Fa, pnode.p + Fa, pflags.p + Fa, hit.p + Fa, offS.p, offSF.p, relS.p, cFe, Fan, L.mask.p, L.childBase.p, L.tstar.p, ntri.p, nnode.p, nflags.p, (trackKids ? C.tstar.p : (uint32_t*)nullptr), fuseMixed ? 2 : fuseS ? 1 : 0, precheckKids);
			else if (!emitRedOut && !starStore) k_emit_pipe<true, 8, false, false><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_S(L.tstar.p,));
			else if (!emitRedOut) k_emit_pipe<true, 8, false, true><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_S(L.tstar.p,));
			else if (!starStore) k_emit_pipe<true, 8, true, false><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_S(L.tstar.p,));
			else k_emit_pipe<true, 8, true, true><<<nb, VX_THREADS, 0, s>>>(SVB_EMIT_ARGS_S(L.tstar.p,));
#undef SVB_EMIT_ARGS_S
			SVB_KERNEL_CHECK();
			if (prof) prof->end(pidE, kids, 20.0 * (double)S + (trackKids ? 14.0 : 10.0) * (double)kids);
		}
		ptri = std::move(ntri);
		pnode = std::move(nnode);
		pflags = std::move(nflags);
		F = Fn; Fa = Fan; S = Sn;
	}
	ptri.release();
	pnode.release();
}

}  // namespace svb
