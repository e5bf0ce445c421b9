// svb_dedup.cu -- per-level bottom-up DAG reduction (unique-node dedup + child remap).
//
// Replaces the std::map<Node,id> loop of GeomOctree::toDAG (src/symvox/geom_octree.cpp:462-548)
// and the "sub-DAGs, join, last DAG pass" of buildDAG (:331-425).  Two nodes are equal iff their
// keys (child mask + child ids, Node::operator< src/symvox/octree_node.cpp:56-86) are equal; the
// reference keeps unique nodes in FIRST-OCCURRENCE order, which here is the order of the smallest
// 64-bit order key  (tile_seq | first-touch triangle | path with the last digit reversed)  among
// the equal nodes (pinned on CPU by oracle/parallel_model.py).
//
// One streaming pass per level over the SoA node arrays; uniqueness is resolved in an
// open-addressing table that is small whenever the level compresses well (the table then lives in
// the 126 MB L2 and the pass runs at HBM speed):
//   KIND_LEAF  (level L-1): key = 8-bit voxel mask -> 256-entry direct table, staged in shared memory.
//   KIND_K64   (level L-2): key = the 8 child masks = one exact 64-bit word (the 4^3 voxel block).
//   KIND_INNER (above)    : key = 8 child uids, tagged by a 64-bit hash, verified exactly afterwards.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "svb_dedup.cuh"
#include "svb_classify.cuh"
#include "svb_sat.cuh"

namespace svb {

namespace {

constexpr int DD_THREADS = 256;
// A ref that still names a table slot (its entry was created by the running launch and has no uid yet) carries this
// bit; final refs are uids < 2^31.  Slot arrays are kept <= 2^30 entries wherever marked refs are used.
constexpr uint32_t REF_MARK = 0x80000000u;

// tstar is the smallest root-pair index q that touches the node (svb_voxelize.cu::make_root_pairs); q - tileStart[tile]
// = rank of the first-touch triangle among the tile's candidate triangles, monotone in the triangle id -- all the order
// needs, in far fewer bits than the id itself (a tile sees thousands of triangles, a scene tens of millions).
__device__ __forceinline__ uint64_t order_key(uint64_t code, uint32_t tstar, const DedupArgs& a) {
	const int l = a.l;
	uint64_t pmask = (l >= 21) ? ~0ull : ((1ull << (3 * l)) - 1);
	uint64_t path = code & pmask;
	uint64_t tile = (l >= 21) ? 0 : (code >> (3 * l));
	if (l > 0) path = (path & ~7ull) | (7ull - (path & 7ull));   // children are created 7 -> 0 (geom_octree.cpp:234)
	return ((uint64_t)a.tileSeq[tile] << (a.tbits + 3 * l)) | ((uint64_t)(tstar - a.tileStart[tile]) << (3 * l)) | path;
}

template <int CHMODE>
__device__ __forceinline__ uint32_t read_child(const void* refs, uint64_t i) {
	if (CHMODE == CH_MASK_U8) return ((const uint8_t*)refs)[i];
	return ((const uint32_t*)refs)[i];
}

// Builds the key of node n.  MASK modes: key64 = child masks (byte c = mask of child c); UID mode: key8.
// Returns false when the node has no non-empty child (cleanEmptyNodes cascade, geom_octree.cpp:437-456).
template <int CHMODE>
__device__ __forceinline__ bool build_key(const DedupArgs& a, uint64_t n, uint32_t key8[8], uint64_t& key64) {
	unsigned m = a.mask[n];
	uint64_t base = a.childBase[n];
	bool any = false;
	key64 = 0;
	int r = 0;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		uint32_t v = NULLREF;
		if ((m >> c) & 1) {
			uint32_t x = read_child<CHMODE>(a.childRefs, base + r);
			++r;
			if (CHMODE == CH_UID_U32) { v = x; any |= (x != NULLREF); }
			else if (x != 0 && x != NULLREF) { v = x; key64 |= (uint64_t)(x & 0xFF) << (8 * c); any = true; }
		}
		key8[c] = v;
	}
	return any;
}

// Nibble c of RANK_SEL[m & 127] = rank of child c among the set bits of child mask m (bit 7 never precedes anyone):
// the __byte_perm selector that moves the packed run of child bytes (refs[childBase ..]) to their child positions.
__device__ const uint32_t RANK_SEL[128] = {
	0x00000000u, 0x11111110u, 0x11111100u, 0x22222210u, 0x11111000u, 0x22222110u, 0x22222100u, 0x33333210u,
	0x11110000u, 0x22221110u, 0x22221100u, 0x33332210u, 0x22221000u, 0x33332110u, 0x33332100u, 0x44443210u,
	0x11100000u, 0x22211110u, 0x22211100u, 0x33322210u, 0x22211000u, 0x33322110u, 0x33322100u, 0x44433210u,
	0x22210000u, 0x33321110u, 0x33321100u, 0x44432210u, 0x33321000u, 0x44432110u, 0x44432100u, 0x55543210u,
	0x11000000u, 0x22111110u, 0x22111100u, 0x33222210u, 0x22111000u, 0x33222110u, 0x33222100u, 0x44333210u,
	0x22110000u, 0x33221110u, 0x33221100u, 0x44332210u, 0x33221000u, 0x44332110u, 0x44332100u, 0x55443210u,
	0x22100000u, 0x33211110u, 0x33211100u, 0x44322210u, 0x33211000u, 0x44322110u, 0x44322100u, 0x55433210u,
	0x33210000u, 0x44321110u, 0x44321100u, 0x55432210u, 0x44321000u, 0x55432110u, 0x55432100u, 0x66543210u,
	0x10000000u, 0x21111110u, 0x21111100u, 0x32222210u, 0x21111000u, 0x32222110u, 0x32222100u, 0x43333210u,
	0x21110000u, 0x32221110u, 0x32221100u, 0x43332210u, 0x32221000u, 0x43332110u, 0x43332100u, 0x54443210u,
	0x21100000u, 0x32211110u, 0x32211100u, 0x43322210u, 0x32211000u, 0x43322110u, 0x43322100u, 0x54433210u,
	0x32210000u, 0x43321110u, 0x43321100u, 0x54432210u, 0x43321000u, 0x54432110u, 0x54432100u, 0x65543210u,
	0x21000000u, 0x32111110u, 0x32111100u, 0x43222210u, 0x32111000u, 0x43222110u, 0x43222100u, 0x54333210u,
	0x32110000u, 0x43221110u, 0x43221100u, 0x54332210u, 0x43221000u, 0x54332110u, 0x54332100u, 0x65443210u,
	0x32100000u, 0x43211110u, 0x43211100u, 0x54322210u, 0x43211000u, 0x54322110u, 0x54322100u, 0x65433210u,
	0x43210000u, 0x54321110u, 0x54321100u, 0x65432210u, 0x54321000u, 0x65432110u, 0x65432100u, 0x76543210u,
};

// CH_MASK_U8 key without the 8-way predicated byte loop: the children of a node are popc(mask) consecutive bytes, so
// three aligned 32-bit loads cover them for any alignment of childBase (the mask array of a level is padded by 16 zeroed
// bytes, svb_voxelize.cu), two funnel shifts pack them, two byte permutes spread them to their child positions and a
// byte mask clears the positions of absent children (the permute fetched a neighbour's byte there).  A child whose
// voxel mask is 0 contributes a zero byte = "no child", as in build_key().  ~25 instructions instead of ~90.
__device__ __forceinline__ bool build_key64_u8(const DedupArgs& a, uint64_t n, uint64_t& key64) {
	const unsigned m = a.mask[n];
	const uint32_t base = a.childBase[n];
	const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a.childRefs) + (base >> 2);
	const uint32_t a0 = w[0], a1 = w[1], a2 = w[2];
	const unsigned sh = (base & 3u) * 8u;
	const uint32_t lo = __funnelshift_r(a0, a1, sh), hi = __funnelshift_r(a1, a2, sh);   // child bytes 0..3 / 4..7 of the run
	const uint32_t sel = __ldg(&RANK_SEL[m & 127u]);
	const uint32_t bmLo = (((m & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;   // 4 mask bits -> 4 byte masks
	const uint32_t bmHi = (((m >> 4) * 0x00204081u) & 0x01010101u) * 0xFFu;
	const uint32_t kLo = __byte_perm(lo, hi, sel & 0xFFFFu) & bmLo;
	const uint32_t kHi = __byte_perm(lo, hi, sel >> 16) & bmHi;
	key64 = ((uint64_t)kHi << 32) | kLo;
	return key64 != 0;
}

__device__ __forceinline__ uint64_t tag_of_key8(const uint32_t k[8], const HashSeed& hs) {
	uint64_t h = hs.init;
#pragma unroll
	for (int c = 0; c < 8; c += 2) h = mix64(h ^ (((uint64_t)k[c + 1] << 32) | k[c])) + 0x9E3779B97F4A7C15ull * (c + 1);
	return finish_tag(h, hs);
}

// ------------------------------------------------------------------ KIND_LEAF
// 4 nodes per thread per trip, 128-bit loads (1 x uchar4 masks, 1 x uint4 t*, 2 x ulonglong2 codes): the kernel
// only streams 13 B/node, so bytes in flight per thread decide how close it gets to the HBM roofline.
// MODE 0: full 64-bit order key.  Wide mode (the order key of the leaf level needs more than 63 bits): MODE 1 takes
// the minimum of the high part (tile_seq | t*) per voxel mask, MODE 2 -- a second pass over the same nodes -- the
// minimum of the low part (path') among the nodes that attain it; (hi, lo) compares like the undivided key.
__device__ __forceinline__ uint64_t order_key_hi(uint64_t code, uint32_t tstar, const DedupArgs& a) {
	uint64_t tile = (a.l >= 21) ? 0 : (code >> (3 * a.l));
	return ((uint64_t)a.tileSeq[tile] << a.tbits) | (uint64_t)(tstar - a.tileStart[tile]);
}
__device__ __forceinline__ uint64_t order_key_lo(uint64_t code, int l) {
	uint64_t pmask = (l >= 21) ? ~0ull : ((1ull << (3 * l)) - 1);
	uint64_t path = code & pmask;
	if (l > 0) path = (path & ~7ull) | (7ull - (path & 7ull));
	return path;
}

template <int MODE>
__global__ void __launch_bounds__(DD_THREADS) k_leaf_min(DedupArgs a, unsigned long long* __restrict__ gmin, unsigned long long* __restrict__ voxels) {
	__shared__ unsigned long long smin[256];
	__shared__ unsigned long long shi[MODE == 2 ? 256 : 1];
	smin[threadIdx.x] = MAX_ORDER;
	if (MODE == 2) shi[threadIdx.x] = gmin[threadIdx.x];
	__syncthreads();
	unsigned vox = 0;
	const uint64_t nq = a.N >> 2;   // full quads
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uchar4* __restrict__ m4 = reinterpret_cast<const uchar4*>(a.mask);
	const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(a.tstar);
	const ulonglong2* __restrict__ c2 = reinterpret_cast<const ulonglong2*>(a.code);
	auto visit = [&](unsigned m, uint32_t ts, unsigned long long cd) {
		if (!m) return;
		unsigned long long O;
		if (MODE == 0) O = order_key(cd, ts, a);
		else if (MODE == 1) O = order_key_hi(cd, ts, a);
		else {
			if (order_key_hi(cd, ts, a) != shi[m]) return;
			O = order_key_lo(cd, a.l);
		}
		if (MODE != 2) vox += __popc(m);
		if (smin[m] > O) atomicMin(&smin[m], O);
	};
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
		const uchar4 mm = m4[q];
		const uint4 tt = t4[q];
		const ulonglong2 ca = c2[2 * q], cb = c2[2 * q + 1];
		visit(mm.x, tt.x, ca.x); visit(mm.y, tt.y, ca.y); visit(mm.z, tt.z, cb.x); visit(mm.w, tt.w, cb.y);
	}
	// tail (< 4 nodes)
	for (uint64_t n = (nq << 2) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += stride) visit(a.mask[n], a.tstar[n], a.code[n]);
	__syncthreads();
	unsigned long long* g = gmin + (MODE == 2 ? 256 : 0);
	unsigned long long v = smin[threadIdx.x];
	if (v != MAX_ORDER && g[threadIdx.x] > v) atomicMin(&g[threadIdx.x], v);
	if (MODE != 2) {
#pragma unroll
		for (int d = 16; d; d >>= 1) vox += __shfl_xor_sync(0xFFFFFFFFu, vox, d);
		if ((threadIdx.x & 31) == 0 && vox) atomicAdd(voxels, (unsigned long long)vox);
	}
}
// Lazy variant of MODE 0 (default when the order key fits one word and tileSeq ascends with the tile-local index).
// Within one batch order keys compare like (t*, path'): t* is a root-pair index, and root pairs are sorted by (tile,
// triangle).  So a node whose t* is larger than the smallest t* this CTA has seen for the same voxel mask cannot hold
// the minimum and its code is never fetched; and when the batch comes after everything the table has seen (`later`), a
// mask that already has an entry is frozen and not even t* is fetched.  What always streams is the mask byte (voxel
// count): 1 B/node instead of 13 for all but the first nodes of each mask.  Loads are skipped per 4-node quad, i.e. per
// 16 B of t* / 32 B of codes, so whole DRAM sectors stay untouched.  Skipping is only ever a filter on nodes that
// provably lose; the 64-bit minimum itself is taken exactly as in MODE 0.
// gminBefore: the table as it was BEFORE this launch.  "Frozen" must be decided on that snapshot, never on the live table: the
// grid does not fit the machine at once (48 registers: 5 CTAs/SM resident out of 8 launched per SM), so late CTAs start
// after early ones have already merged their minima -- a voxel mask that is NEW in this batch would look frozen to them, and
// they would skip nodes that may hold its smallest order key.  (Round-1 bug, found in round 2 when warm-up batches made new
// masks in "later" batches common on terrain-like scenes: tests/golden/size_composite_crop4k.json under low free memory.)
__global__ void __launch_bounds__(DD_THREADS) k_leaf_lazy(DedupArgs a, unsigned long long* __restrict__ gmin, const unsigned long long* __restrict__ gminBefore,
                                                           unsigned long long* __restrict__ voxels, int later) {
	__shared__ unsigned long long smin[256];
	__shared__ uint32_t sq[256];        // smallest t* seen for the mask by this CTA
	__shared__ uint8_t sfrozen[256];
	smin[threadIdx.x] = MAX_ORDER;
	sq[threadIdx.x] = 0xFFFFFFFFu;
	sfrozen[threadIdx.x] = (later && gminBefore[threadIdx.x] != MAX_ORDER) ? 1 : 0;
	if (threadIdx.x == 0) sfrozen[0] = 1;   // empty nodes take no part
	__syncthreads();
	unsigned vox = 0;
	const uint64_t ng = a.N >> 4;   // full groups of 16 nodes = one 128-bit load of masks (bytes in flight per thread decide the rate of a 1 B/node stream)
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint4* __restrict__ m16 = reinterpret_cast<const uint4*>(a.mask);
	const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(a.tstar);
	const ulonglong2* __restrict__ c2 = reinterpret_cast<const ulonglong2*>(a.code);
	auto visit = [&](unsigned m, uint32_t ts, unsigned long long cd) {
		const unsigned long long O = order_key(cd, ts, a);
		if (smin[m] > O) atomicMin(&smin[m], O);
		if (sq[m] > ts) atomicMin(&sq[m], ts);
	};
	auto quad = [&](uint32_t mm, uint64_t q) {   // nodes 4q .. 4q+3
		if (!mm) return;
		const unsigned m0 = mm & 0xFFu, m1 = (mm >> 8) & 0xFFu, m2 = (mm >> 16) & 0xFFu, m3 = mm >> 24;
		const bool l0 = !sfrozen[m0], l1 = !sfrozen[m1], l2 = !sfrozen[m2], l3 = !sfrozen[m3];
		if (!(l0 | l1 | l2 | l3)) return;
		const uint4 tt = t4[q];
		const bool n0 = l0 && tt.x <= sq[m0], n1 = l1 && tt.y <= sq[m1], n2 = l2 && tt.z <= sq[m2], n3 = l3 && tt.w <= sq[m3];
		if (n0 | n1) { const ulonglong2 ca = c2[2 * q]; if (n0) visit(m0, tt.x, ca.x); if (n1) visit(m1, tt.y, ca.y); }
		if (n2 | n3) { const ulonglong2 cb = c2[2 * q + 1]; if (n2) visit(m2, tt.z, cb.x); if (n3) visit(m3, tt.w, cb.y); }
	};
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
		const uint4 mm = m16[g];
		vox += __popc(mm.x) + __popc(mm.y) + __popc(mm.z) + __popc(mm.w);
		quad(mm.x, 4 * g); quad(mm.y, 4 * g + 1); quad(mm.z, 4 * g + 2); quad(mm.w, 4 * g + 3);
	}
	// tail (< 16 nodes)
	for (uint64_t n = (ng << 4) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += stride) {
		const unsigned m = a.mask[n];
		vox += __popc(m);
		if (!sfrozen[m]) { const uint32_t ts = a.tstar[n]; if (ts <= sq[m]) visit(m, ts, a.code[n]); }
	}
	__syncthreads();
	unsigned long long v = smin[threadIdx.x];
	if (v != MAX_ORDER && gmin[threadIdx.x] > v) atomicMin(&gmin[threadIdx.x], v);
#pragma unroll
	for (int d = 16; d; d >>= 1) vox += __shfl_xor_sync(0xFFFFFFFFu, vox, d);
	if ((threadIdx.x & 31) == 0 && vox) atomicAdd(voxels, (unsigned long long)vox);
}

// Leaf level of a later batch whose first touches were not tracked: counts the voxels and checks that every voxel mask
// of the batch already has a (frozen) entry -- then the level is reduced with nothing else to do.  *fail is raised
// otherwise (the caller re-voxelizes the batch with first touches and takes the regular path).
__global__ void __launch_bounds__(DD_THREADS) k_leaf_known(uint64_t N, const uint8_t* __restrict__ mask, const unsigned long long* __restrict__ gmin,
                                                            unsigned long long* __restrict__ voxels, uint32_t* __restrict__ nUnknown, uint32_t* __restrict__ list, uint32_t listCap) {
	__shared__ uint8_t sunknown[256];
	sunknown[threadIdx.x] = (threadIdx.x != 0 && gmin[threadIdx.x] == MAX_ORDER) ? 1 : 0;
	__syncthreads();
	unsigned vox = 0;
	const uint64_t ng = N >> 4;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint4* __restrict__ m16 = reinterpret_cast<const uint4*>(mask);
	auto unknown = [&](uint64_t n) {   // rare: a node whose voxel mask has no entry yet
		const uint32_t k = atomicAdd(nUnknown, 1u);
		if (k < listCap) list[k] = (uint32_t)n;
	};
	auto word = [&](uint32_t w, uint64_t n) {
		vox += __popc(w);
		if (sunknown[w & 0xFFu] | sunknown[(w >> 8) & 0xFFu] | sunknown[(w >> 16) & 0xFFu] | sunknown[w >> 24]) {
#pragma unroll
			for (int j = 0; j < 4; ++j) if (sunknown[(w >> (8 * j)) & 0xFFu]) unknown(n + j);
		}
	};
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
		const uint4 mm = m16[g];
		word(mm.x, 16 * g); word(mm.y, 16 * g + 4); word(mm.z, 16 * g + 8); word(mm.w, 16 * g + 12);
	}
	for (uint64_t n = (ng << 4) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
		const unsigned m = mask[n];
		vox += __popc(m);
		if (sunknown[m]) unknown(n);
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) vox += __shfl_xor_sync(0xFFFFFFFFu, vox, d);
	if ((threadIdx.x & 31) == 0 && vox) atomicAdd(voxels, (unsigned long long)vox);
}
// One warp per listed leaf node: its first touch t* = the smallest root pair q of its tile whose triangle passes
// testTriBox (reference operation order, tri_box_overlap) at every box on the node's path -- which is exactly when the
// level-synchronous build holds a (triangle, node) pair for it (k_classify, svb_voxelize.cu).  Lanes test 32
// consecutive q at a time; the node exists, so the search ends inside its tile's range.
__global__ void __launch_bounds__(DD_THREADS) k_leaf_query(uint32_t cnt, const uint32_t* __restrict__ list, DedupArgs a, LeafQuery lq, unsigned long long* __restrict__ gmin) {
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (w >= cnt) return;
	const uint64_t n = list[w];
	const uint64_t cd = a.code[n];
	const int l = a.l;
	const uint32_t tile = (uint32_t)(cd >> (3 * l));
	const TileGeom tg = reinterpret_cast<const TileGeom*>(lq.tiles)[tile];
	uint32_t found = UNSET;
	for (uint64_t q0 = a.tileStart[tile]; q0 < lq.P && found == UNSET; q0 += 32) {
		const uint64_t q = q0 + lane;
		bool ok = q < lq.P;
		if (ok) {
			const float* tp = lq.tris + 9ull * lq.rootTri[q];
			double cx = tg.cx, cy = tg.cy, cz = tg.cz, k = tg.rootSide * 0.25;
			for (int d = l - 1; d >= 0 && ok; --d) {
				const int dig = (int)((cd >> (3 * d)) & 7);
				cx = __dadd_rn(cx, (dig & 4) ? k : -k);
				cy = __dadd_rn(cy, (dig & 2) ? k : -k);
				cz = __dadd_rn(cz, (dig & 1) ? k : -k);
				ok = tri_box_overlap(cx, cy, cz, k, tp);
				k *= 0.5;
			}
		}
		const unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
		if (b) found = (uint32_t)(q0 + (__ffs(b) - 1));
	}
	if (lane == 0 && found != UNSET) atomicMin(&gmin[a.mask[n]], (unsigned long long)order_key(cd, found, a));
}
__global__ void k_add_u64(unsigned long long* dst, const unsigned long long* src) { *dst += *src; }
__global__ void k_count_known(const unsigned long long* __restrict__ gmin, uint32_t* __restrict__ out) {
	const unsigned n = __syncthreads_count(threadIdx.x != 0 && gmin[threadIdx.x] != MAX_ORDER);
	if (threadIdx.x == 0) *out = n;
}

// wide mode, between the two passes: a mask whose high part improved in this batch forgets its old low part
__global__ void k_leaf_reset_lo(unsigned long long* __restrict__ gmin, const unsigned long long* __restrict__ hiBefore) {
	if (gmin[threadIdx.x] != hiBefore[threadIdx.x]) gmin[256 + threadIdx.x] = MAX_ORDER;
}

// ------------------------------------------------------------------ KIND_K64 / KIND_INNER
struct TableDev {
	unsigned long long* tag;
	unsigned long long* minO;
	uint32_t* uid;
	uint64_t capMask;
	uint64_t countBefore, maxLoad;
	uint32_t* flags;   // [0] overflow, [1] collision, [2] new entries
	HashSeed hs;
	int later;         // every node of this launch has a larger order key than whatever the existing entries hold (DedupArgs::seqLo)
};

__device__ __forceinline__ bool table_find_or_claim(const TableDev& t, uint64_t tag, uint64_t& slot) {
	uint64_t idx = mix64(tag) & t.capMask;
	for (int probe = 0; probe < 8192; ++probe) {
		// the pass is void once the table went over its load limit (it is grown and the pass redone): do not crawl through a
		// nearly full table for nothing
		if ((probe & 63) == 63 && *(volatile uint32_t*)&t.flags[0]) return false;
		unsigned long long cur = t.tag[idx];
		if (cur == tag) { slot = idx; return true; }
		if (cur == EMPTY_TAG) {
			unsigned long long old = atomicCAS(&t.tag[idx], (unsigned long long)EMPTY_TAG, (unsigned long long)tag);
			if (old == EMPTY_TAG) {
				uint32_t nc = atomicAdd(&t.flags[2], 1u) + 1;
				if (t.countBefore + nc > t.maxLoad) t.flags[0] = 1;
				slot = idx;
				return true;
			}
			if (old == tag) { slot = idx; return true; }
		}
		idx = (idx + 1) & t.capMask;
	}
	t.flags[0] = 1;
	return false;
}

// MARKED (inner levels): a node that finds an entry from an earlier launch is finished on the spot -- exact key check
// against the entry's stored key, final uid as its ref -- and only the nodes of entries created by this launch leave a
// marked slot for k_winner / k_convert, which are skipped altogether when the launch created nothing.
// (MARKED) mlist / mcount: the nodes left with a marked ref are also listed (up to mcap), so that k_winner / k_convert visit a
// few thousand nodes instead of streaming the whole level again for the handful of entries a later batch creates.
template <int CHMODE, bool PERM = false, bool MARKED = false>
__global__ void __launch_bounds__(DD_THREADS) k_insert(DedupArgs a, TableDev t, const uint32_t* __restrict__ dKey8,
                                                        uint32_t* __restrict__ mlist = nullptr, uint32_t* __restrict__ mcount = nullptr, uint32_t mcap = 0) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= a.N) return;
	uint32_t k8[8];
	uint64_t k64;
	const bool any = (PERM && CHMODE == CH_MASK_U8) ? build_key64_u8(a, n, k64) : build_key<CHMODE>(a, n, k8, k64);
	if (!any) { a.ref[n] = NULLREF; return; }
	uint64_t tag = (CHMODE == CH_UID_U32) ? tag_of_key8(k8, t.hs) : k64;
	uint64_t slot;
	if (!table_find_or_claim(t, tag, slot)) { a.ref[n] = NULLREF; return; }
	if (MARKED) {
		const uint32_t u = t.uid[slot];   // (a slot claimed in this launch still has uid UNSET)
		if (u < t.countBefore) {
			const uint4 s0 = reinterpret_cast<const uint4*>(dKey8)[(uint64_t)u * 2], s1 = reinterpret_cast<const uint4*>(dKey8)[(uint64_t)u * 2 + 1];   // the stored key: two 128-bit loads
			const bool same = s0.x == k8[0] && s0.y == k8[1] && s0.z == k8[2] && s0.w == k8[3] && s1.x == k8[4] && s1.y == k8[5] && s1.z == k8[6] && s1.w == k8[7];
			if (!same) t.flags[1] = 1;
			a.ref[n] = u;
			if (t.later) return;   // frozen entry
		} else {
			a.ref[n] = REF_MARK | (uint32_t)slot;
			if (mcount) { const uint32_t k = atomicAdd(mcount, 1u); if (k < mcap) mlist[k] = (uint32_t)n; }
		}
	} else {
		a.ref[n] = (uint32_t)slot;
		if (t.later && t.uid[slot] < t.countBefore) return;   // frozen entry
	}
	unsigned long long O = order_key(a.code[n], a.tstar[n], a);
	if (t.minO[slot] > O) atomicMin(&t.minO[slot], O);
}

// ------------------------------------------------------------------ KIND_K64, single pass (default for CH_MASK_U8)
// The 4^3 level compresses to a few 10^5 distinct keys while a tile batch streams 10^8 nodes through it, so almost every
// node finds its key already in the table.  The thread that claims an empty slot draws the dense uid on the spot and
// appends the slot to a list (the dense arrays are filled from that list afterwards: k_assign_list); every other node
// reads the uid together with the slot and writes its final ref at once -- no second pass over the node arrays.  A node
// that finds the slot in the instant between the claim and the claimer's uid store leaves  0x80000000 | slot  as its
// ref and its index in a second list (k_fix_list); if either list overflows, a full pass (k_fix_all / k_assign_k64)
// does the same job.  NPT nodes per thread: their loads and first probes are in flight together.
struct OnePass {
	uint32_t* dCount;      // next dense uid
	uint32_t* newSlots;    // slots claimed by this launch
	uint32_t* unres;       // nodes whose ref is still a marked slot
	uint32_t listCap;
	uint32_t* counters;    // [0] unresolved nodes, [1] nodes of new entries whose first touch has to be queried (UNTRACKED)
	uint2* qlist;          // UNTRACKED: (node, slot) of those nodes
};

// UNTRACKED (only ever with t.later): the level's first touches were not recorded by the voxelizer (a.tstar == nullptr).
// Existing entries are frozen and need none; a node whose entry is NEW is listed and gets its first touch from a direct
// query afterwards (k_k64_query) -- a few thousand nodes out of ~3 * 10^8 per tile batch of the 16K^3 city.
template <int NPT, bool UNTRACKED = false>
__global__ void __launch_bounds__(DD_THREADS) k_insert_k64(DedupArgs a, TableDev t, OnePass op) {
	const uint64_t n0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * NPT;
	if (n0 >= a.N) return;
	uint64_t key[NPT], idx[NPT];
	unsigned long long cur[NPT], O[NPT];
	bool live[NPT];
	uint32_t out[NPT];
#pragma unroll
	for (int j = 0; j < NPT; ++j) {
		const uint64_t n = n0 + j;
		live[j] = n < a.N && build_key64_u8(a, n, key[j]);
		out[j] = NULLREF;
	}
#pragma unroll
	for (int j = 0; j < NPT; ++j) {
		idx[j] = 0; cur[j] = 0; O[j] = 0;
		if (live[j]) {
			idx[j] = mix64(key[j]) & t.capMask;
			cur[j] = t.tag[idx[j]];
			if (!UNTRACKED && !t.later) O[j] = order_key(a.code[n0 + j], a.tstar[n0 + j], a);   // (later: only the rare new entries need it)
		}
	}
#pragma unroll
	for (int j = 0; j < NPT; ++j) {
		if (!live[j]) continue;
		// Neighbours in Morton order are neighbours in space: the interior of a wall is a run of identical 4^3 blocks.  A
		// node with the key of its predecessor takes the predecessor's result -- valid whenever that result is the final uid
		// of a frozen entry (nothing else is owed for such a node: no order key, no list entry).
		if (j > 0 && t.later && live[j - 1] && key[j] == key[j - 1] && out[j - 1] < t.countBefore) { out[j] = out[j - 1]; continue; }
		const uint64_t n = n0 + j, tag = key[j];
		uint64_t i = idx[j];
		unsigned long long c = cur[j];
		bool found = false, claimed = false;
		for (int probe = 0; probe < 8192; ++probe) {
			if (c == tag) { found = true; break; }
			if (c == EMPTY_TAG) {
				const unsigned long long old = atomicCAS(&t.tag[i], (unsigned long long)EMPTY_TAG, (unsigned long long)tag);
				if (old == EMPTY_TAG) { found = claimed = true; break; }
				if (old == tag) { found = true; break; }
			}
			i = (i + 1) & t.capMask;
			c = t.tag[i];
		}
		if (!found) { t.flags[0] = 1; continue; }
		uint32_t u;
		if (claimed) {
			const uint32_t nc = atomicAdd(&t.flags[2], 1u);
			if (t.countBefore + nc + 1 > t.maxLoad) t.flags[0] = 1;
			u = atomicAdd(op.dCount, 1u);
			t.uid[i] = u;
			if (nc < op.listCap) op.newSlots[nc] = (uint32_t)i;
		} else {
			u = __ldcg(&t.uid[i]);
		}
		if (!(t.later && u < t.countBefore)) {   // frozen entries keep their order key
			if (UNTRACKED) {
				const uint32_t k = atomicAdd(&op.counters[1], 1u);
				if (k < op.listCap) op.qlist[k] = make_uint2((uint32_t)n, (uint32_t)i);
			} else {
				if (t.later) O[j] = order_key(a.code[n], a.tstar[n], a);
				if (t.minO[i] > O[j]) atomicMin(&t.minO[i], O[j]);
			}
		}
		if (u == UNSET) {   // claimed a moment ago by another thread: resolved after the launch
			const uint32_t k = atomicAdd(&op.counters[0], 1u);
			if (k < op.listCap) op.unres[k] = (uint32_t)n;
			u = REF_MARK | (uint32_t)i;
		}
		out[j] = u;
	}
	if (NPT == 2 && n0 + 1 < a.N) *reinterpret_cast<uint2*>(a.ref + n0) = make_uint2(out[0], out[NPT - 1]);   // n0 is a multiple of NPT
	else if (NPT == 4 && n0 + 3 < a.N) *reinterpret_cast<uint4*>(a.ref + n0) = make_uint4(out[0], out[1 % NPT], out[2 % NPT], out[3 % NPT]);
	else {
#pragma unroll
		for (int j = 0; j < NPT; ++j) if (n0 + j < a.N) a.ref[n0 + j] = out[j];
	}
}
// First touch of a node whose level was voxelized without first-touch tracking: the smallest root pair q of its tile whose
// triangle passes testTriBox (reference operation order, tri_box_overlap) at every box on the node's path -- which is
// exactly when the level-synchronous build holds a (triangle, node) pair for it.  One warp per node, 32 consecutive q per
// trip.  A triangle that fails the three box-axis tests of the node's OWN box (the last box of the path; the same
// subtractions and comparisons tri_box_overlap starts with) cannot pass, so it is dropped before the chain is walked.
__device__ __forceinline__ uint32_t first_touch_query(const DedupArgs& a, const LeafQuery& lq, uint64_t cd, int lane) {
	const int l = a.l;
	const uint32_t tile = (uint32_t)(cd >> (3 * l));
	const TileGeom tg = reinterpret_cast<const TileGeom*>(lq.tiles)[tile];
	// the node's own box
	double nx = tg.cx, ny = tg.cy, nz = tg.cz, nk = tg.rootSide * 0.25, kl = nk;
	for (int d = l - 1; d >= 0; --d) {
		const int dig = (int)((cd >> (3 * d)) & 7);
		nx = __dadd_rn(nx, (dig & 4) ? nk : -nk);
		ny = __dadd_rn(ny, (dig & 2) ? nk : -nk);
		nz = __dadd_rn(nz, (dig & 1) ? nk : -nk);
		kl = nk;
		nk *= 0.5;
	}
	uint32_t found = UNSET;
	for (uint64_t q0 = a.tileStart[tile]; q0 < lq.P && found == UNSET; q0 += 32) {
		const uint64_t q = q0 + lane;
		bool ok = q < lq.P;
		if (ok) {
			const float* tp = lq.tris + 9ull * lq.rootTri[q];
			if (l > 0) {
				const float x0 = tp[0], x1 = tp[3], x2 = tp[6], y0 = tp[1], y1 = tp[4], y2 = tp[7], z0 = tp[2], z1 = tp[5], z2 = tp[8];
				// fl(t - c) is monotone in t: min / max may be taken on the float inputs
				ok = !(__dsub_rn((double)fminf(x0, fminf(x1, x2)), nx) > kl || __dsub_rn((double)fmaxf(x0, fmaxf(x1, x2)), nx) < -kl) &&
				     !(__dsub_rn((double)fminf(y0, fminf(y1, y2)), ny) > kl || __dsub_rn((double)fmaxf(y0, fmaxf(y1, y2)), ny) < -kl) &&
				     !(__dsub_rn((double)fminf(z0, fminf(z1, z2)), nz) > kl || __dsub_rn((double)fmaxf(z0, fmaxf(z1, z2)), nz) < -kl);
			}
			if (ok) {
				double cx = tg.cx, cy = tg.cy, cz = tg.cz, k = tg.rootSide * 0.25;
				for (int d = l - 1; d >= 0 && ok; --d) {
					const int dig = (int)((cd >> (3 * d)) & 7);
					cx = __dadd_rn(cx, (dig & 4) ? k : -k);
					cy = __dadd_rn(cy, (dig & 2) ? k : -k);
					cz = __dadd_rn(cz, (dig & 1) ? k : -k);
					ok = tri_box_overlap(cx, cy, cz, k, tp);
					k *= 0.5;
				}
			}
		}
		const unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
		if (b) found = (uint32_t)(q0 + (__ffs(b) - 1));
	}
	return found;
}
__global__ void __launch_bounds__(DD_THREADS) k_k64_query(uint32_t cnt, const uint2* __restrict__ qlist, DedupArgs a, LeafQuery lq, unsigned long long* __restrict__ minO, uint32_t* __restrict__ flags) {
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (w >= cnt) return;
	const uint2 e = qlist[w];
	const uint64_t cd = a.code[e.x];
	const uint32_t found = first_touch_query(a, lq, cd, lane);
	if (lane) return;
	if (found == UNSET) { flags[3] = 1; return; }   // cannot happen: the node exists, so some triangle of its tile reaches it
	atomicMin(&minO[e.y], (unsigned long long)order_key(cd, found, a));
}
// (list overflow) the nodes of new entries collected by a pass over the refs: ref = uid >= countBefore, slot found by probing
__global__ void __launch_bounds__(DD_THREADS) k_k64_collect(DedupArgs a, TableDev t, uint32_t cap, uint32_t* __restrict__ counter, uint2* __restrict__ qlist) {
	const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= a.N) return;
	const uint32_t r = a.ref[n];
	if (r == NULLREF || r < t.countBefore) return;
	uint64_t key;
	build_key64_u8(a, n, key);
	uint64_t i = mix64(key) & t.capMask;
	while (t.tag[i] != key) i = (i + 1) & t.capMask;
	const uint32_t k = atomicAdd(counter, 1u);
	if (k < cap) qlist[k] = make_uint2((uint32_t)n, (uint32_t)i);
}

__global__ void __launch_bounds__(DD_THREADS) k_assign_list(uint32_t fresh, const uint32_t* __restrict__ newSlots, const unsigned long long* __restrict__ tag,
                                                             const unsigned long long* __restrict__ minO, const uint32_t* __restrict__ uid,
                                                             uint64_t* __restrict__ dMinO, uint64_t* __restrict__ dKey64) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= fresh) return;
	const uint32_t i = newSlots[k], u = uid[i];
	dMinO[u] = minO[i];
	dKey64[u] = tag[i];
}
// (list overflow) the same from a pass over the table: the new entries are those with uid >= countBefore
__global__ void __launch_bounds__(DD_THREADS) k_assign_scan(uint64_t cap, uint32_t countBefore, const unsigned long long* __restrict__ tag, const unsigned long long* __restrict__ minO,
                                                             const uint32_t* __restrict__ uid, uint64_t* __restrict__ dMinO, uint64_t* __restrict__ dKey64) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cap) return;
	const uint32_t u = uid[i];
	if (u == UNSET || u < countBefore) return;
	dMinO[u] = minO[i];
	dKey64[u] = tag[i];
}
__global__ void __launch_bounds__(DD_THREADS) k_fix_list(uint32_t cnt, const uint32_t* __restrict__ unres, const uint32_t* __restrict__ uid, uint32_t* __restrict__ ref) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= cnt) return;
	const uint32_t n = unres[k];
	ref[n] = uid[ref[n] & ~REF_MARK];
}
__global__ void __launch_bounds__(DD_THREADS) k_fix_all(uint64_t N, const uint32_t* __restrict__ uid, uint32_t* __restrict__ ref) {
	const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	const uint32_t r = ref[n];
	if (r != NULLREF && (r & REF_MARK)) ref[n] = uid[r & ~REF_MARK];
}

// K64: new slots get their dense uid from a pass over the table (the key is the tag itself)
__global__ void __launch_bounds__(DD_THREADS) k_assign_k64(uint64_t cap, const unsigned long long* __restrict__ tag, const unsigned long long* __restrict__ minO,
                                                            uint32_t* __restrict__ uid, uint32_t* __restrict__ dCount,
                                                            uint64_t* __restrict__ dMinO, uint64_t* __restrict__ dKey64) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cap) return;
	unsigned long long tg = tag[i];
	if (tg == EMPTY_TAG || uid[i] != UNSET) return;
	uint32_t u = atomicAdd(dCount, 1u);
	uid[i] = u;
	dMinO[u] = minO[i];
	dKey64[u] = tg;
}

// INNER: the node that holds the minimum order key of a new slot publishes the full key
template <int CHMODE, bool MARKED = false>
__global__ void __launch_bounds__(DD_THREADS) k_winner(DedupArgs a, TableDev t, uint32_t* __restrict__ dCount,
                                                        uint64_t* __restrict__ dMinO, uint32_t* __restrict__ dKey8,
                                                        const uint32_t* __restrict__ list = nullptr, uint64_t count = 0) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (list) { if (n >= count) return; n = list[n]; }
	if (n >= a.N) return;
	uint32_t slot = a.ref[n];
	if (slot == NULLREF) return;
	if (MARKED) { if (!(slot & REF_MARK)) return; slot &= ~REF_MARK; }
	if (t.uid[slot] != UNSET) return;   // the entry has its key already (only this slot's winner sets the uid, and it is unique)
	unsigned long long O = order_key(a.code[n], a.tstar[n], a);
	if (t.minO[slot] != O) return;
	uint32_t k8[8];
	uint64_t k64;
	build_key<CHMODE>(a, n, k8, k64);
	uint32_t u = atomicAdd(dCount, 1u);
	t.uid[slot] = u;
	dMinO[u] = O;
#pragma unroll
	for (int c = 0; c < 8; ++c) dKey8[(uint64_t)u * 8 + c] = k8[c];
}

// slot -> uid for every node (+ exact key check for hashed keys)
template <int CHMODE, bool VERIFY, bool MARKED = false>
__global__ void __launch_bounds__(DD_THREADS) k_convert(DedupArgs a, TableDev t, const uint32_t* __restrict__ dKey8,
                                                         const uint32_t* __restrict__ list = nullptr, uint64_t count = 0) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (list) { if (n >= count) return; n = list[n]; }
	if (n >= a.N) return;
	uint32_t slot = a.ref[n];
	if (slot == NULLREF) return;
	if (MARKED) { if (!(slot & REF_MARK)) return; slot &= ~REF_MARK; }
	uint32_t u = t.uid[slot];
	if (VERIFY) {
		uint32_t k8[8];
		uint64_t k64;
		build_key<CHMODE>(a, n, k8, k64);
		bool same = true;
#pragma unroll
		for (int c = 0; c < 8; ++c) same &= (dKey8[(uint64_t)u * 8 + c] == k8[c]);
		if (!same) t.flags[1] = 1;
	}
	a.ref[n] = u;
}

// re-insert all dense entries into a fresh (larger) slot array
template <bool K64>
__global__ void __launch_bounds__(DD_THREADS) k_rebuild(uint64_t count, const uint64_t* __restrict__ dMinO, const uint64_t* __restrict__ dKey64,
                                                         const uint32_t* __restrict__ dKey8, TableDev t) {
	uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= count) return;
	uint64_t tag;
	if (K64) tag = dKey64[u];
	else {
		uint32_t k8[8];
#pragma unroll
		for (int c = 0; c < 8; ++c) k8[c] = dKey8[u * 8 + c];
		tag = tag_of_key8(k8, t.hs);
	}
	uint64_t idx = mix64(tag) & t.capMask;
	for (;;) {
		unsigned long long old = atomicCAS(&t.tag[idx], (unsigned long long)EMPTY_TAG, (unsigned long long)tag);
		if (old == EMPTY_TAG) break;
		idx = (idx + 1) & t.capMask;
	}
	t.minO[idx] = dMinO[u];
	t.uid[idx] = (uint32_t)u;
}

template <int CHMODE>
__global__ void k_root(DedupArgs a, uint32_t* rootKey8) {
	if (threadIdx.x || blockIdx.x) return;
	uint32_t k8[8];
	uint64_t k64;
	build_key<CHMODE>(a, 0, k8, k64);
	for (int c = 0; c < 8; ++c) rootKey8[c] = k8[c];
}

// ------------------------------------------------------------------ finalize kernels
__global__ void k_iota(uint64_t n, uint32_t* v) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = (uint32_t)i;
}
__global__ void k_rank_scatter(uint64_t n, const uint32_t* __restrict__ sortedUid, uint32_t* __restrict__ rank) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) rank[sortedUid[i]] = (uint32_t)i;
}
__global__ void k_emit_leaf(uint64_t U, const uint32_t* __restrict__ sortedMask, uint8_t* __restrict__ omask, uint32_t* __restrict__ ochild) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= U) return;
	omask[i] = (uint8_t)sortedMask[i];
	for (int c = 0; c < 8; ++c) ochild[i * 8 + c] = NULLNODE;
}
__global__ void k_emit_k64(uint64_t U, const uint32_t* __restrict__ sortedUid, const uint64_t* __restrict__ dKey64,
                           const uint32_t* __restrict__ rankLeaf, uint8_t* __restrict__ omask, uint32_t* __restrict__ ochild) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= U) return;
	uint64_t k = dKey64[sortedUid[i]];
	unsigned m = 0;
	for (int c = 0; c < 8; ++c) {
		unsigned b = (unsigned)((k >> (8 * c)) & 0xFF);
		ochild[i * 8 + c] = b ? rankLeaf[b] : NULLNODE;
		if (b) m |= 1u << c;
	}
	omask[i] = (uint8_t)m;
}
__global__ void k_emit_inner(uint64_t U, const uint32_t* __restrict__ sortedUid, const uint32_t* __restrict__ dKey8,
                             const uint32_t* __restrict__ rankChild, uint8_t* __restrict__ omask, uint32_t* __restrict__ ochild) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= U) return;
	uint64_t u = sortedUid[i];
	unsigned m = 0;
	for (int c = 0; c < 8; ++c) {
		uint32_t r = dKey8[u * 8 + c];
		ochild[i * 8 + c] = (r == NULLREF) ? NULLNODE : rankChild[r];
		if (r != NULLREF) m |= 1u << c;
	}
	omask[i] = (uint8_t)m;
}
__global__ void k_emit_root(const uint32_t* rootKey8, const uint32_t* rankChild, uint8_t* omask, uint32_t* ochild) {
	if (threadIdx.x || blockIdx.x) return;
	unsigned m = 0;
	for (int c = 0; c < 8; ++c) {
		uint32_t r = rootKey8[c];
		ochild[c] = (r == NULLREF) ? NULLNODE : rankChild[r];
		if (r != NULLREF) m |= 1u << c;
	}
	omask[0] = (uint8_t)m;
}

uint64_t next_pow2(uint64_t x) {
	uint64_t p = 1;
	while (p < x) p <<= 1;
	return p;
}

// Does this batch come after everything the table has seen (see DedupArgs::seqLo)?  Records the batch.
bool later_batch(LevelTable& T, const DedupArgs& a) {
	const char* e = getenv("SVB_DEDUP_LAZY");   // 0: always update every entry / read every field (A/B, verification)
	const bool later = T.seenAny && a.seqLo > T.maxSeq && !(e && e[0] == '0');
	// An existing entry's improved order key only reaches the slot array (atomicMin), not the dense copy finalize_levels /
	// k_rebuild / merge_export read: correct as long as batches reach a table in ascending sub-octree order (then an existing
	// entry can never improve).  run_tiles_split guarantees that; anything else must fail loudly, not rank silently wrong.
	if (T.kind != KIND_LEAF && T.seenAny && T.count > 0 && a.seqLo <= T.maxSeq)
		throw Error(SVB_EINVAL, "dedup: tile batches must reach a level table in ascending sub-octree order");
	T.maxSeq = T.seenAny ? std::max(T.maxSeq, a.seqHi) : a.seqHi;
	T.seenAny = true;
	return later;
}

void alloc_slots(cudaStream_t s, Pool& pool, LevelTable& T, uint64_t cap) {
	T.cap = cap;
	T.tag.reset(pool, cap);
	T.minO.reset(pool, cap);
	T.uid.reset(pool, cap);
	T.tag.zero();
	T.minO.fill_ff();
	T.uid.fill_ff();
}

TableDev dev_view(LevelTable& T, uint32_t* flags) {
	TableDev t;
	t.tag = (unsigned long long*)T.tag.p;
	t.minO = (unsigned long long*)T.minO.p;
	t.uid = T.uid.p;
	t.capMask = T.cap - 1;
	t.countBefore = T.count;
	t.maxLoad = T.cap - T.cap / 4;   // 75 %
	t.flags = flags;
	t.hs = T.hs;
	t.later = 0;
	return t;
}

void grow_slots(cudaStream_t s, Pool& pool, LevelTable& T, uint64_t newCap) {
	alloc_slots(s, pool, T, newCap);
	if (T.count == 0) return;
	TableDev t = dev_view(T, nullptr);
	unsigned nb = blocks_for(T.count, DD_THREADS);
	if (T.kind == KIND_K64) k_rebuild<true><<<nb, DD_THREADS, 0, s>>>(T.count, T.dMinO.p, T.dKey64.p, nullptr, t);
	else k_rebuild<false><<<nb, DD_THREADS, 0, s>>>(T.count, T.dMinO.p, nullptr, T.dKey8.p, t);
	SVB_KERNEL_CHECK();
}

template <class Tp>
void grow_copy(cudaStream_t s, Pool& pool, DevBuf<Tp>& b, uint64_t oldElems, uint64_t newElems) {
	DevBuf<Tp> nb(pool, newElems);
	if (oldElems) SVB_CUDA(cudaMemcpyAsync(nb.p, b.p, oldElems * sizeof(Tp), cudaMemcpyDeviceToDevice, s));
	b = std::move(nb);
}

void ensure_dense(cudaStream_t s, Pool& pool, LevelTable& T, uint64_t need) {
	if (need <= T.denseCap) return;
	uint64_t nc = next_pow2(need < 1024 ? 1024 : need);
	grow_copy(s, pool, T.dMinO, T.count, nc);
	if (T.kind == KIND_K64) grow_copy(s, pool, T.dKey64, T.count, nc);
	else grow_copy(s, pool, T.dKey8, T.count * 8, nc * 8);
	T.denseCap = nc;
}

}  // namespace

HashSeed make_hash_seed(uint64_t seed) {
	HashSeed hs;
	hs.init = 0x9E3779B97F4A7C15ull ^ (seed * 0xD1B54A32D192ED03ull);
	if (seed == 0)
		if (const char* e = getenv("SVB_TEST_WEAK_HASH")) { const int b = atoi(e); if (b > 0 && b < 64) hs.mask = (1ull << b) - 1; }
	return hs;
}

void table_init(cudaStream_t s, Pool& pool, LevelTable& T, int kind, uint64_t seed) {
	T.kind = kind;
	T.hs = make_hash_seed(seed);
	T.count = 0;
	T.denseCap = 0;
	T.unique = 0;
	if (kind == KIND_LEAF) {
		T.cap = 256;
		T.minO.reset(pool, 512);   // [0,256): order key (or its high part), [256,512): low part (wide mode only)
		T.minO.fill_ff();
		return;
	}
	T.dCount.reset(pool, 1);
	T.dCount.zero();
	alloc_slots(s, pool, T, 1ull << 16);
}

void dedup_leaf(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a, uint64_t* d_voxels) {
	if (a.N == 0) return;
	unsigned nb = blocks_for(a.N, DD_THREADS * 16);
	if (nb > 148 * 8) nb = 148 * 8;   // persistent-style grid: 8 CTAs of 256 threads per SM, grid-stride
	unsigned long long* g = (unsigned long long*)T.minO.p;
	const bool later = later_batch(T, a);
	if (!T.wide) {
		const char* e = getenv("SVB_LEAF_LAZY");   // 0: stream all 13 B of every node (k_leaf_min<0>)
		DevBuf<uint64_t> snapshot(pool, 256);   // the table before this launch (freed after the read-back below has synchronised)
		if (a.seqMonotone && !(e && e[0] == '0')) {
			SVB_CUDA(cudaMemcpyAsync(snapshot.p, T.minO.p, 256 * 8, cudaMemcpyDeviceToDevice, s));
			k_leaf_lazy<<<nb, DD_THREADS, 0, s>>>(a, g, (const unsigned long long*)snapshot.p, (unsigned long long*)d_voxels, later ? 1 : 0);
		} else k_leaf_min<0><<<nb, DD_THREADS, 0, s>>>(a, g, (unsigned long long*)d_voxels);
		SVB_KERNEL_CHECK();
		// did this batch bring new voxel masks?  (leaf_tstar_needed: the next batch goes without first touches only if not)
		DevBuf<uint32_t> cnt(pool, 1);
		k_count_known<<<1, 256, 0, s>>>((const unsigned long long*)g, cnt.p);
		SVB_KERNEL_CHECK();
		uint32_t h = 0;
		SVB_CUDA(cudaMemcpyAsync(&h, cnt.p, 4, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		T.lastAdded = h - T.known;
		T.known = h;
		return;
	}
	DevBuf<uint64_t> before(pool, 256);
	SVB_CUDA(cudaMemcpyAsync(before.p, T.minO.p, 256 * 8, cudaMemcpyDeviceToDevice, s));
	k_leaf_min<1><<<nb, DD_THREADS, 0, s>>>(a, g, (unsigned long long*)d_voxels);
	SVB_KERNEL_CHECK();
	k_leaf_reset_lo<<<1, 256, 0, s>>>(g, (const unsigned long long*)before.p);
	SVB_KERNEL_CHECK();
	k_leaf_min<2><<<nb, DD_THREADS, 0, s>>>(a, g, nullptr);
	SVB_KERNEL_CHECK();
}

// KIND_K64 over u8 child masks in one pass over the nodes (k_insert_k64)
static void dedup_k64_onepass(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a, int npt, bool later) {
	constexpr uint32_t LIST_CAP = 1u << 20;
	const bool untracked = a.tstar == nullptr;
	if (untracked && (!later || !a.query)) throw Error(SVB_EINVAL, "dedup: a level without first touches needs a later batch and the query data");
	DevBuf<uint32_t> flags(pool, 4), counters(pool, 4), newSlots(pool, LIST_CAP), unres(pool, LIST_CAP);
	DevBuf<uint2> qlist(pool, untracked ? LIST_CAP : 1);
	uint32_t h[4], hc[4];
	for (;;) {
		if (T.cap > (1ull << 30)) throw Error(SVB_ERANGE, "4^3-level table beyond 2^30 slots");
		flags.zero();
		counters.zero();
		TableDev t = dev_view(T, flags.p);
		t.later = later ? 1 : 0;
		OnePass op;
		op.dCount = T.dCount.p; op.newSlots = newSlots.p; op.unres = unres.p; op.listCap = LIST_CAP; op.counters = counters.p; op.qlist = qlist.p;
		if (untracked && npt >= 4) k_insert_k64<4, true><<<blocks_for((a.N + 3) / 4, DD_THREADS), DD_THREADS, 0, s>>>(a, t, op);
		else if (untracked) k_insert_k64<2, true><<<blocks_for((a.N + 1) / 2, DD_THREADS), DD_THREADS, 0, s>>>(a, t, op);
		else if (npt >= 4) k_insert_k64<4><<<blocks_for((a.N + 3) / 4, DD_THREADS), DD_THREADS, 0, s>>>(a, t, op);
		else if (npt == 2) k_insert_k64<2><<<blocks_for((a.N + 1) / 2, DD_THREADS), DD_THREADS, 0, s>>>(a, t, op);
		else k_insert_k64<1><<<blocks_for(a.N, DD_THREADS), DD_THREADS, 0, s>>>(a, t, op);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		if (!h[0]) break;
		// overflow: the claims of this attempt are dropped by the rebuild; take their uids back too
		const uint32_t cnt = (uint32_t)T.count;
		SVB_CUDA(cudaMemcpyAsync(T.dCount.p, &cnt, 4, cudaMemcpyHostToDevice, s));
		SVB_CUDA(cudaStreamSynchronize(s));   // cnt is a stack temporary
		grow_slots(s, pool, T, T.cap * 4);
	}
	const uint64_t fresh = h[2];
	if (T.count + fresh >= 0x7FFFFFF0ull) throw Error(SVB_ERANGE, "more than 2^31 unique nodes in the 4^3 level");
	ensure_dense(s, pool, T, T.count + fresh);
	if (hc[0]) {   // (before the query pass below: k_k64_collect reads final refs)
		if (hc[0] <= LIST_CAP) k_fix_list<<<blocks_for(hc[0], DD_THREADS), DD_THREADS, 0, s>>>(hc[0], unres.p, T.uid.p, a.ref);
		else k_fix_all<<<blocks_for(a.N, DD_THREADS), DD_THREADS, 0, s>>>(a.N, T.uid.p, a.ref);
		SVB_KERNEL_CHECK();
	}
	if (untracked && hc[1]) {   // first touches of the nodes of the new entries, straight from the triangles
		uint32_t cnt = hc[1];
		const uint2* list = qlist.p;
		DevBuf<uint2> big;
		if (cnt > LIST_CAP) {
			big.reset(pool, cnt);
			SVB_CUDA(cudaMemsetAsync(counters.p + 1, 0, 4, s));
			TableDev t = dev_view(T, flags.p);
			k_k64_collect<<<blocks_for(a.N, DD_THREADS), DD_THREADS, 0, s>>>(a, t, cnt, counters.p + 1, big.p);
			SVB_KERNEL_CHECK();
			list = big.p;
		}
		if (getenv("SVB_VX_STATS")) fprintf(stderr, "[vx-stats] 4^3 level without first touches: %u nodes of %u new entries queried directly\n", cnt, h[2]);
		k_k64_query<<<blocks_for((uint64_t)cnt * 32, DD_THREADS), DD_THREADS, 0, s>>>(cnt, list, a, *a.query, (unsigned long long*)T.minO.p, flags.p);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));   // (also keeps `big` alive until the query has run)
		if (h[3]) throw Error(SVB_ECUDA, "dedup: a node of the 4^3 level has no first touch (internal error)");
	}
	T.lastFresh = fresh;
	T.lastN = a.N;
	if (fresh) {
		if (fresh <= LIST_CAP) k_assign_list<<<blocks_for(fresh, DD_THREADS), DD_THREADS, 0, s>>>((uint32_t)fresh, newSlots.p, (const unsigned long long*)T.tag.p, (const unsigned long long*)T.minO.p, T.uid.p, T.dMinO.p, T.dKey64.p);
		else k_assign_scan<<<blocks_for(T.cap, DD_THREADS), DD_THREADS, 0, s>>>(T.cap, (uint32_t)T.count, (const unsigned long long*)T.tag.p, (const unsigned long long*)T.minO.p, T.uid.p, T.dMinO.p, T.dKey64.p);
		SVB_KERNEL_CHECK();
	}
	T.count += fresh;
}

bool leaf_tstar_needed(const LevelTable& T, uint32_t seqLo) {
	const char* e = getenv("SVB_DEDUP_LAZY");
	const char* f = getenv("SVB_LEAF_NOTSTAR");   // 0: always track the first touches of the leaf level
	if ((e && e[0] == '0') || (f && f[0] == '0')) return true;
	// ... and only once a batch has come and gone without bringing a new voxel mask (a batch that does meet one
	// has to be voxelized twice)
	return !(T.kind == KIND_LEAF && !T.wide && T.seenAny && seqLo > T.maxSeq && T.lastAdded == 0);
}

bool k64_tstar_optional(const LevelTable& T, uint32_t seqLo) {
	const char* e = getenv("SVB_DEDUP_LAZY");
	const char* f = getenv("SVB_K64_ONEPASS");
	if ((e && e[0] == '0') || (f && atoi(f) <= 0)) return false;
	// the single-pass insert must be usable for whatever the batch brings (see dedup_level_t): keep well inside its limits
	if (!(T.kind == KIND_K64 && T.seenAny && seqLo > T.maxSeq && T.count > 0 && T.cap <= (1ull << 27) && T.count < (1ull << 29))) return false;
	// every node of a NEW entry costs a direct query (a warp walking the tile's triangles): only worth it where new entries are
	// rare -- box meshes bring a few per 10^8 nodes, a terrain brings one per 10^2 and is better off tracking first touches
	return T.lastN > 0 && (double)T.lastFresh * 50000.0 < (double)T.lastN;
}

bool dedup_leaf_known(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a, const LeafQuery& lq, uint64_t* d_voxels) {
	if (leaf_tstar_needed(T, a.seqLo)) return false;
	T.lastAdded = 0;
	if (a.N) {
		constexpr uint32_t LIST_CAP = 1u << 18;
		DevBuf<uint64_t> tmp(pool, 1);
		DevBuf<uint32_t> cnt(pool, 1), list(pool, LIST_CAP);
		tmp.zero();
		cnt.zero();
		unsigned nb = blocks_for(a.N, DD_THREADS * 64);
		if (nb > 148 * 8) nb = 148 * 8;
		k_leaf_known<<<nb, DD_THREADS, 0, s>>>(a.N, a.mask, (const unsigned long long*)T.minO.p, (unsigned long long*)tmp.p, cnt.p, list.p, LIST_CAP);
		SVB_KERNEL_CHECK();
		uint32_t h = 0;
		SVB_CUDA(cudaMemcpyAsync(&h, cnt.p, 4, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		if (h > LIST_CAP) return false;   // too many for the direct query: the caller voxelizes again with first touches
		if (h) {
			if (getenv("SVB_VX_STATS")) fprintf(stderr, "[vx-stats] leaf level without first touches: %u nodes with a new voxel mask, queried directly\n", h);
			k_leaf_query<<<blocks_for((uint64_t)h * 32, DD_THREADS), DD_THREADS, 0, s>>>(h, list.p, a, lq, (unsigned long long*)T.minO.p);
			SVB_KERNEL_CHECK();
			k_count_known<<<1, 256, 0, s>>>((const unsigned long long*)T.minO.p, cnt.p);
			SVB_KERNEL_CHECK();
			SVB_CUDA(cudaMemcpyAsync(&h, cnt.p, 4, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaStreamSynchronize(s));
			T.lastAdded = h - T.known;
			T.known = h;
		}
		k_add_u64<<<1, 1, 0, s>>>((unsigned long long*)d_voxels, (const unsigned long long*)tmp.p);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaStreamSynchronize(s));   // the temporaries go back to the pool
	}
	later_batch(T, a);   // records the batch
	return true;
}

template <int CHMODE>
static void dedup_level_t(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a) {
	if (a.N == 0) return;
	const bool k64 = (CHMODE != CH_UID_U32);
	if ((T.kind == KIND_K64) != k64) throw Error(SVB_EINVAL, "dedup_level: table kind does not match child mode");
	constexpr uint32_t MLIST_CAP = 1u << 20;
	DevBuf<uint32_t> flags(pool, 4);   // [3]: nodes left with a marked ref (MARKED)
	DevBuf<uint32_t> mlist;
	unsigned nb = blocks_for(a.N, DD_THREADS);
	uint32_t h[4];
	// size the slot array for this level: at least 2x the entries it may end up holding if ~1/8 of the
	// nodes are new; an overflow simply grows x4 and redoes the pass (failed attempts leave no trace:
	// their slots have no uid and are dropped by the rebuild).
	// ... where 1/8 is replaced by what the previous batch showed (its share of new entries, doubled), and by 1/2 for the first
	// batch of a level with 9 or more voxel patterns below it: a terrain level is mostly unique nodes, and a pass that runs the
	// table over its limit costs far more than the pass itself (long probe chains, then everything again).
	uint64_t expectNew = a.N / 8;
	if (T.lastN) expectNew = std::min<uint64_t>(a.N, (uint64_t)(2.0 * (double)T.lastFresh / (double)T.lastN * (double)a.N) + a.N / 64);
	else if (T.count == 0) expectNew = std::max<uint64_t>(a.N / 8, std::min<uint64_t>(a.N / 2, 16ull << 20));
	uint64_t want = next_pow2(2 * (T.count + expectNew + 1024));
	// (the single-pass 4^3 insert needs <= 2^28 slots: do not size it out of reach on a guess -- an overflow grows the table anyway)
	if (k64 && want > (1ull << 28) && T.count < (1ull << 26)) want = 1ull << 28;
	if (want > T.cap) grow_slots(s, pool, T, want);
	const bool later = later_batch(T, a);
	const char* pk = getenv("SVB_K64_PERM");   // 0: the byte-by-byte key builder (A/B, verification)
	const bool permKey = !(pk && pk[0] == '0');
	if (CHMODE == CH_MASK_U8) {
		const char* e = getenv("SVB_K64_ONEPASS");   // 0: insert + table scan + convert passes; 1 / 2: single pass, that many nodes per thread
		const int npt = e ? atoi(e) : 2;
		// marked slots must stay distinguishable from NULLREF and from uids
		if (npt > 0 && T.cap <= (1ull << 28) && T.count + a.N / 8 < (1ull << 30)) { dedup_k64_onepass(s, pool, T, a, npt, later); return; }
	}
	if (!a.tstar) throw Error(SVB_EINVAL, "dedup: a level without first touches can only go through the single-pass 4^3 insert");
	const char* mk = getenv("SVB_INNER_MARKED");   // 0: every node goes through the winner and convert passes
	const bool markedWanted = !k64 && !(mk && mk[0] == '0');
	bool marked = false;
	for (;;) {
		flags.zero();
		TableDev t = dev_view(T, flags.p);
		t.later = later ? 1 : 0;
		marked = markedWanted && T.cap <= (1ull << 30) && T.count + a.N < (1ull << 31);
		if (marked && !mlist.p) mlist.reset(pool, MLIST_CAP);
		if (marked) k_insert<CHMODE, false, CHMODE == CH_UID_U32><<<nb, DD_THREADS, 0, s>>>(a, t, T.dKey8.p, mlist.p, flags.p + 3, MLIST_CAP);
		else if (permKey) k_insert<CHMODE, CHMODE == CH_MASK_U8><<<nb, DD_THREADS, 0, s>>>(a, t, nullptr);
		else k_insert<CHMODE><<<nb, DD_THREADS, 0, s>>>(a, t, nullptr);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		if (!h[0]) break;
		grow_slots(s, pool, T, T.cap * 4);
	}
	uint64_t fresh = h[2];
	if (T.count + fresh >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "more than 2^32 unique nodes in one level");
	ensure_dense(s, pool, T, T.count + fresh);
	TableDev t = dev_view(T, flags.p);
	if (k64) {
		if (fresh) {
			k_assign_k64<<<blocks_for(T.cap, DD_THREADS), DD_THREADS, 0, s>>>(T.cap, t.tag, t.minO, t.uid, T.dCount.p, T.dMinO.p, T.dKey64.p);
			SVB_KERNEL_CHECK();
		}
		k_convert<CHMODE, false><<<nb, DD_THREADS, 0, s>>>(a, t, nullptr);
		SVB_KERNEL_CHECK();
	} else if (marked) {
		if (fresh) {   // only the nodes of the entries this launch created are still open: through their list when it holds them all
			const uint32_t nm = h[3];
			const uint32_t* list = nm <= MLIST_CAP ? mlist.p : nullptr;
			const unsigned nbl = list ? blocks_for(nm, DD_THREADS) : nb;
			k_winner<CHMODE, true><<<nbl, DD_THREADS, 0, s>>>(a, t, T.dCount.p, T.dMinO.p, T.dKey8.p, list, nm);
			SVB_KERNEL_CHECK();
			k_convert<CHMODE, true, true><<<nbl, DD_THREADS, 0, s>>>(a, t, T.dKey8.p, list, nm);
			SVB_KERNEL_CHECK();
			SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaStreamSynchronize(s));
		}
		if (h[1]) throw Error(SVB_ECOLLISION, "64-bit node-key hash collision (exact verify failed)");
	} else {
		if (fresh) {
			k_winner<CHMODE><<<nb, DD_THREADS, 0, s>>>(a, t, T.dCount.p, T.dMinO.p, T.dKey8.p);
			SVB_KERNEL_CHECK();
		}
		k_convert<CHMODE, true><<<nb, DD_THREADS, 0, s>>>(a, t, T.dKey8.p);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		if (h[1]) throw Error(SVB_ECOLLISION, "64-bit node-key hash collision (exact verify failed)");
	}
	T.lastFresh = fresh;
	T.lastN = a.N;
	T.count += fresh;
}

// =================================================================== multi-GPU merge of one level
// Every rank reduced its own sub-octrees into rank-local tables.  Bottom-up, level by level, the ranks
// exchange their unique nodes as records {key with GLOBAL child uids, min order key} (an all-gather
// done by the caller over NCCL), and every rank rebuilds the identical global table from the union:
// equal keys collapse, the smallest order key survives, so the final first-occurrence ranking is the
// one a single GPU would have produced.  Replaces the serial "Joining subtrees" + "Last DAG pass"
// of GeomOctree::buildDAG (src/symvox/geom_octree.cpp:397-425) across devices.
namespace {

struct RecK64 { uint64_t key; uint64_t minO; };
struct RecInner { uint32_t key[8]; uint64_t minO; };

__global__ void __launch_bounds__(DD_THREADS) k_export_k64(uint64_t n, const uint64_t* __restrict__ dKey64, const uint64_t* __restrict__ dMinO, RecK64* __restrict__ out) {
	uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= n) return;
	RecK64 r; r.key = dKey64[u]; r.minO = dMinO[u];
	out[u] = r;
}
__global__ void __launch_bounds__(DD_THREADS) k_export_inner(uint64_t n, const uint32_t* __restrict__ dKey8, const uint64_t* __restrict__ dMinO,
                                                              const uint32_t* __restrict__ l2gChild, RecInner* __restrict__ out) {
	uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= n) return;
	RecInner r;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		uint32_t k = dKey8[u * 8 + c];
		r.key[c] = (k == NULLREF) ? NULLREF : l2gChild[k];
	}
	r.minO = dMinO[u];
	out[u] = r;
}

// entry e of the gathered buffer = (rank e / maxCount, index e % maxCount); valid iff index < counts[rank]
template <bool K64>
__global__ void __launch_bounds__(DD_THREADS) k_import_insert(uint64_t total, uint64_t maxCount, const uint64_t* __restrict__ counts, const char* __restrict__ all,
                                                               uint64_t strideBytes, TableDev t, uint32_t* __restrict__ slotOf) {
	uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= total) return;
	uint64_t r = e / maxCount, i = e % maxCount;
	if (i >= counts[r]) { slotOf[e] = NULLREF; return; }
	uint64_t tag, O;
	if (K64) { const RecK64* p = (const RecK64*)(all + r * strideBytes) + i; tag = p->key; O = p->minO; }
	else { const RecInner* p = (const RecInner*)(all + r * strideBytes) + i; tag = tag_of_key8(p->key, t.hs); O = p->minO; }
	uint64_t slot;
	if (!table_find_or_claim(t, tag, slot)) { slotOf[e] = NULLREF; return; }
	if (t.minO[slot] > O) atomicMin(&t.minO[slot], (unsigned long long)O);
	slotOf[e] = (uint32_t)slot;
}
// The global uid of a merged node must be the SAME on every rank (the next level's keys are built
// from it), so it cannot come from an atomic counter: the winner of a slot (the record carrying the
// slot's minimum order key) is flagged, and uid = exclusive scan of the flags over the gathered
// buffer, which is byte-identical on all ranks.
template <bool K64>
__global__ void __launch_bounds__(DD_THREADS) k_import_flag(uint64_t total, uint64_t maxCount, const char* __restrict__ all, uint64_t strideBytes, TableDev t,
                                                             const uint32_t* __restrict__ slotOf, uint32_t* __restrict__ flag) {
	uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= total) return;
	uint32_t slot = slotOf[e];
	uint32_t f = 0;
	if (slot != NULLREF) {
		uint64_t O;
		if (K64) O = ((const RecK64*)(all + (e / maxCount) * strideBytes) + (e % maxCount))->minO;
		else O = ((const RecInner*)(all + (e / maxCount) * strideBytes) + (e % maxCount))->minO;
		f = (t.minO[slot] == O) ? 1u : 0u;
	}
	flag[e] = f;
}
template <bool K64>
__global__ void __launch_bounds__(DD_THREADS) k_import_assign(uint64_t total, uint64_t maxCount, const char* __restrict__ all, uint64_t strideBytes, TableDev t,
                                                               const uint32_t* __restrict__ slotOf, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                                               uint64_t* __restrict__ dMinO, uint64_t* __restrict__ dKey64, uint32_t* __restrict__ dKey8) {
	uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= total || !flag[e]) return;
	uint32_t slot = slotOf[e];
	uint32_t u = pos[e];
	t.uid[slot] = u;
	if (K64) {
		const RecK64* p = (const RecK64*)(all + (e / maxCount) * strideBytes) + (e % maxCount);
		dMinO[u] = p->minO;
		dKey64[u] = p->key;
	} else {
		const RecInner* p = (const RecInner*)(all + (e / maxCount) * strideBytes) + (e % maxCount);
		dMinO[u] = p->minO;
#pragma unroll
		for (int c = 0; c < 8; ++c) dKey8[(uint64_t)u * 8 + c] = p->key[c];
	}
}
__global__ void __launch_bounds__(DD_THREADS) k_import_verify(uint64_t total, uint64_t maxCount, const char* __restrict__ all, uint64_t strideBytes, TableDev t,
                                                               const uint32_t* __restrict__ slotOf, const uint32_t* __restrict__ dKey8) {
	uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= total) return;
	uint32_t slot = slotOf[e];
	if (slot == NULLREF) return;
	const RecInner* p = (const RecInner*)(all + (e / maxCount) * strideBytes) + (e % maxCount);
	uint32_t u = t.uid[slot];
	bool same = true;
#pragma unroll
	for (int c = 0; c < 8; ++c) same &= (dKey8[(uint64_t)u * 8 + c] == p->key[c]);
	if (!same) t.flags[1] = 1;
}
__global__ void __launch_bounds__(DD_THREADS) k_import_l2g(uint64_t n, const uint32_t* __restrict__ slotOfMine, const uint32_t* __restrict__ uid, uint32_t* __restrict__ l2g) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) l2g[i] = (slotOfMine[i] == NULLREF) ? NULLREF : uid[slotOfMine[i]];
}
__global__ void k_min256(uint32_t world, const unsigned long long* __restrict__ all, uint64_t strideWords, unsigned long long* __restrict__ out, int wide) {
	unsigned i = threadIdx.x;
	unsigned long long m = MAX_ORDER, lo = MAX_ORDER;
	for (uint32_t r = 0; r < world; ++r) {
		unsigned long long v = all[r * strideWords + i];
		unsigned long long w = wide ? all[r * strideWords + 256 + i] : 0;
		if (v < m || (wide && v == m && w < lo)) { m = v; lo = w; }   // lexicographic (hi, lo)
	}
	out[i] = m;
	if (wide) out[256 + i] = lo;
}

}  // namespace

uint32_t merge_rec_bytes(int kind) { return kind == KIND_LEAF ? 8u : (kind == KIND_K64 ? (uint32_t)sizeof(RecK64) : (uint32_t)sizeof(RecInner)); }
uint64_t merge_count(const LevelTable& T) { return T.kind == KIND_LEAF ? (T.wide ? 512 : 256) : T.count; }

void merge_export(cudaStream_t s, Pool& pool, LevelTable& T, const uint32_t* l2gChild, void* d_out) {
	(void)pool;
	if (T.kind == KIND_LEAF) { SVB_CUDA(cudaMemcpyAsync(d_out, T.minO.p, merge_count(T) * 8, cudaMemcpyDeviceToDevice, s)); return; }
	if (T.count == 0) return;
	unsigned nb = blocks_for(T.count, DD_THREADS);
	if (T.kind == KIND_K64) k_export_k64<<<nb, DD_THREADS, 0, s>>>(T.count, T.dKey64.p, T.dMinO.p, (RecK64*)d_out);
	else {
		if (!l2gChild) throw Error(SVB_EINVAL, "merge_export: the level below has not been merged yet");
		k_export_inner<<<nb, DD_THREADS, 0, s>>>(T.count, T.dKey8.p, T.dMinO.p, l2gChild, (RecInner*)d_out);
	}
	SVB_KERNEL_CHECK();
}

// after the import passes: *status = {overflow, collision, claimed slots, winners != claimed, unique nodes}; the unique count also
// becomes the table's device counter
__global__ void k_import_status(const uint32_t* __restrict__ flags, const uint64_t* __restrict__ dTot, uint32_t* __restrict__ dCount, uint32_t* __restrict__ status) {
	const uint32_t fresh = (uint32_t)*dTot;
	status[0] = flags[0];
	status[1] = flags[1];
	status[2] = flags[2];
	status[3] = (fresh != flags[2]) ? 1u : 0u;
	status[4] = fresh;
	*dCount = fresh;
}

// Stream-ordered: no host synchronisation.  Everything is sized from the gathered counts (an upper bound of the number of
// unique nodes); the unique count, overflow / collision flags and the winner check land in d_status (5 x u32, device) and
// are read back ONCE for all levels by merge_resolve() when the build finishes.
void merge_import(cudaStream_t s, Pool& pool, LevelTable& T, const void* d_all, const uint64_t* counts, uint32_t world, uint64_t strideBytes,
                  uint32_t myRank, DevBuf<uint32_t>& l2g, uint32_t* d_status, uint64_t mergeSeed) {
	if (T.kind == KIND_LEAF) {
		k_min256<<<1, 256, 0, s>>>(world, (const unsigned long long*)d_all, strideBytes / 8, (unsigned long long*)T.minO.p, T.wide ? 1 : 0);
		SVB_KERNEL_CHECK();
		return;
	}
	const bool k64 = T.kind == KIND_K64;
	uint64_t maxCount = 0, sum = 0;
	for (uint32_t r = 0; r < world; ++r) { if (counts[r] > maxCount) maxCount = counts[r]; sum += counts[r]; }
	if (sum >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "merge_import: more than 2^32 records in one level");
	const uint64_t myCount = counts[myRank];
	l2g.reset(pool, myCount ? myCount : 1);
	// fresh global table
	LevelTable G;
	G.kind = T.kind;
	G.hs = make_hash_seed(mergeSeed);   // the same on every rank: all of them must reach the same verdict
	G.dCount.reset(pool, 1);
	G.dCount.zero();
	alloc_slots(s, pool, G, next_pow2(2 * sum + 1024));
	if (sum) {
		const uint64_t total = (uint64_t)world * maxCount;
		DevBuf<uint64_t> dCounts(pool, world);
		SVB_CUDA(cudaMemcpyAsync(dCounts.p, counts, world * 8ull, cudaMemcpyHostToDevice, s));   // pageable source: consumed before the call returns
		DevBuf<uint32_t> slotOf(pool, total), flags(pool, 4);
		flags.zero();
		TableDev t = dev_view(G, flags.p);
		unsigned nb = blocks_for(total, DD_THREADS);
		if (k64) k_import_insert<true><<<nb, DD_THREADS, 0, s>>>(total, maxCount, dCounts.p, (const char*)d_all, strideBytes, t, slotOf.p);
		else k_import_insert<false><<<nb, DD_THREADS, 0, s>>>(total, maxCount, dCounts.p, (const char*)d_all, strideBytes, t, slotOf.p);
		SVB_KERNEL_CHECK();
		DevBuf<uint32_t> flag(pool, total), pos(pool, total);
		DevBuf<uint64_t> dTot(pool, 1);
		if (k64) k_import_flag<true><<<nb, DD_THREADS, 0, s>>>(total, maxCount, (const char*)d_all, strideBytes, t, slotOf.p, flag.p);
		else k_import_flag<false><<<nb, DD_THREADS, 0, s>>>(total, maxCount, (const char*)d_all, strideBytes, t, slotOf.p, flag.p);
		SVB_KERNEL_CHECK();
		scan_u32(s, pool, flag.p, total, pos.p, dTot.p);
		ensure_dense(s, pool, G, sum);   // the unique count is only known on the device
		if (k64) k_import_assign<true><<<nb, DD_THREADS, 0, s>>>(total, maxCount, (const char*)d_all, strideBytes, t, slotOf.p, flag.p, pos.p, G.dMinO.p, G.dKey64.p, nullptr);
		else k_import_assign<false><<<nb, DD_THREADS, 0, s>>>(total, maxCount, (const char*)d_all, strideBytes, t, slotOf.p, flag.p, pos.p, G.dMinO.p, nullptr, G.dKey8.p);
		SVB_KERNEL_CHECK();
		if (!k64) {
			k_import_verify<<<nb, DD_THREADS, 0, s>>>(total, maxCount, (const char*)d_all, strideBytes, t, slotOf.p, G.dKey8.p);
			SVB_KERNEL_CHECK();
		}
		k_import_status<<<1, 1, 0, s>>>(flags.p, dTot.p, G.dCount.p, d_status);
		SVB_KERNEL_CHECK();
		if (myCount) {
			k_import_l2g<<<blocks_for(myCount, DD_THREADS), DD_THREADS, 0, s>>>(myCount, slotOf.p + (uint64_t)myRank * maxCount, G.uid.p, l2g.p);
			SVB_KERNEL_CHECK();
		}
		G.count = sum;   // upper bound until merge_resolve()
	} else SVB_CUDA(cudaMemsetAsync(d_status, 0, 5 * sizeof(uint32_t), s));
	T = std::move(G);
}

// host copy of one level's import status (read back by the caller for all levels at once)
void merge_resolve(LevelTable& T, const uint32_t status[5]) {
	if (T.kind == KIND_LEAF) return;
	if (status[0]) throw Error(SVB_ECUDA, "merge_import: table overflow");
	if (status[3]) throw Error(SVB_ECUDA, "merge_import: winner count does not match the number of claimed slots");
	if (status[1]) throw Error(SVB_ECOLLISION, "merge_import: 64-bit node-key hash collision (exact verify failed)");
	T.count = status[4];
}

void dedup_level(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a) {
	switch (a.childMode) {
		case CH_MASK_U8: dedup_level_t<CH_MASK_U8>(s, pool, T, a); break;
		case CH_MASK_U32: dedup_level_t<CH_MASK_U32>(s, pool, T, a); break;
		case CH_UID_U32: dedup_level_t<CH_UID_U32>(s, pool, T, a); break;
		default: throw Error(SVB_EINVAL, "bad child mode");
	}
}

void root_key(cudaStream_t s, const DedupArgs& a, uint32_t* d_rootKey8) {
	switch (a.childMode) {
		case CH_MASK_U8: k_root<CH_MASK_U8><<<1, 32, 0, s>>>(a, d_rootKey8); break;
		case CH_MASK_U32: k_root<CH_MASK_U32><<<1, 32, 0, s>>>(a, d_rootKey8); break;
		default: k_root<CH_UID_U32><<<1, 32, 0, s>>>(a, d_rootKey8); break;
	}
	SVB_KERNEL_CHECK();
}

static void alloc_out(Pool& pool, OutLevel& o, uint64_t n) {
	o.n = n;
	o.mask.reset(pool, n);
	o.child.reset(pool, n * 8);
	o.mirror.reset(pool, n * 3);
	o.inv.reset(pool, n);
	o.mirror.zero();
	o.inv.zero();
	o.hasChildLevel = false;
}

void finalize_levels(cudaStream_t s, Pool& pool, std::vector<LevelTable>& tables, const std::vector<int>& obits,
                     const uint32_t* d_rootKey8, int rootChildMode, std::vector<OutLevel>& out) {
	const int L = (int)tables.size();
	out.clear();
	out.resize(L);
	for (int g = L - 1; g >= 1; --g) {
		LevelTable& T = tables[g];
		if (T.kind == KIND_LEAF) {
			// <= 255 distinct voxel masks: rank them on the host by (order key) or, in wide mode, (high part, low part)
			std::vector<uint64_t> hk(512);
			SVB_CUDA(cudaMemcpyAsync(hk.data(), T.minO.p, 512 * 8, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaStreamSynchronize(s));
			std::vector<uint32_t> ord;
			for (uint32_t m = 0; m < 256; ++m) if (hk[m] != MAX_ORDER) ord.push_back(m);
			std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {
				if (hk[x] != hk[y]) return hk[x] < hk[y];
				return T.wide ? hk[256 + x] < hk[256 + y] : false;
			});
			const uint64_t U = ord.size();
			std::vector<uint32_t> hrank(256, 0);
			for (uint32_t i = 0; i < U; ++i) hrank[ord[i]] = i;
			T.rank.reset(pool, 256);
			SVB_CUDA(cudaMemcpyAsync(T.rank.p, hrank.data(), 256 * 4, cudaMemcpyHostToDevice, s));
			DevBuf<uint32_t> vals(pool, U ? U : 1);
			if (U) SVB_CUDA(cudaMemcpyAsync(vals.p, ord.data(), U * 4, cudaMemcpyHostToDevice, s));
			T.unique = U;
			alloc_out(pool, out[g], U);
			if (U) {
				k_emit_leaf<<<blocks_for(U, 256), 256, 0, s>>>(U, vals.p, out[g].mask.p, out[g].child.p);
				SVB_KERNEL_CHECK();
			}
			SVB_CUDA(cudaStreamSynchronize(s));   // ord / hrank are host temporaries
			continue;
		}
		uint64_t n = T.count;
		DevBuf<uint64_t> keys(pool, n);
		DevBuf<uint32_t> vals(pool, n);
		T.rank.reset(pool, n ? n : 1);
		if (n) {
			SVB_CUDA(cudaMemcpyAsync(keys.p, T.dMinO.p, n * 8, cudaMemcpyDeviceToDevice, s));
			k_iota<<<blocks_for(n, 256), 256, 0, s>>>(n, vals.p);
			SVB_KERNEL_CHECK();
			radix_sort_pairs(s, pool, keys.p, vals.p, n, obits[g]);
			k_rank_scatter<<<blocks_for(n, 256), 256, 0, s>>>(n, vals.p, T.rank.p);
			SVB_KERNEL_CHECK();
		}
		const uint64_t U = n;
		T.unique = U;
		alloc_out(pool, out[g], U);
		if (U) {
			unsigned nb = blocks_for(U, 256);
			if (T.kind == KIND_K64) k_emit_k64<<<nb, 256, 0, s>>>(U, vals.p, T.dKey64.p, tables[g + 1].rank.p, out[g].mask.p, out[g].child.p);
			else k_emit_inner<<<nb, 256, 0, s>>>(U, vals.p, T.dKey8.p, tables[g + 1].rank.p, out[g].mask.p, out[g].child.p);
			SVB_KERNEL_CHECK();
		}
	}
	alloc_out(pool, out[0], 1);
	if (L > 1) {
		(void)rootChildMode;   // the root key already holds uids / mask values that index rank[] of level 1 directly
		k_emit_root<<<1, 32, 0, s>>>(d_rootKey8, tables[1].rank.p, out[0].mask.p, out[0].child.p);
		SVB_KERNEL_CHECK();
	}
}

}  // namespace svb
