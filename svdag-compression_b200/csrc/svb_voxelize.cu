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
#include "svb_internal.cuh"
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
	int seq_lo, seq_hi;  // tile_seq range of this batch [lo, hi)
};

__device__ inline void tile_range(double mn, double mx, double o, const GridDesc& g, int& a, int& b) {
	double fa = floor((mn - g.margin - o) * g.inv_cell);
	double fb = floor((mx + g.margin - o) * g.inv_cell);
	a = fa < 0.0 ? 0 : (fa > (double)(g.G - 1) ? g.G : (int)fa);
	b = fb < 0.0 ? -1 : (fb > (double)(g.G - 1) ? g.G - 1 : (int)fb);
}

template <bool EMIT>
__global__ void __launch_bounds__(VX_THREADS) k_candidates(const float* __restrict__ tris, uint64_t T, GridDesc g,
                                                            const int* __restrict__ gridTile, uint32_t* __restrict__ cnt_or_off,
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
	uint32_t c = 0;
	uint32_t o = EMIT ? cnt_or_off[t] : 0;
	for (int x = ax; x <= bx; ++x)
		for (int y = ay; y <= by; ++y)
			for (int z = az; z <= bz; ++z) {
				int s = gridTile[((size_t)x * g.G + y) * g.G + z];
				if (s >= g.seq_lo && s < g.seq_hi) {
					if (EMIT) { ptri[o + c] = (uint32_t)t; pnode[o + c] = (uint32_t)(s - g.seq_lo); }
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
                                                          const float* __restrict__ tris, uint8_t* __restrict__ hit, uint8_t* __restrict__ mask) {
	uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t p = gid >> 3;
	int c = (int)(gid & 7);
	bool ok = false;
	uint32_t n = 0;
	if (p < P) {
		uint32_t t = ptri[p];
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
	}
	unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
	int lane = threadIdx.x & 31;
	unsigned m = (b >> (lane & 24)) & 0xFFu;
	if (c == 0 && p < P) {
		hit[p] = (uint8_t)m;
		if (m) {
			unsigned cur = mask[n];   // may be stale (L1); bits only ever get set, so a stale value only costs an extra atomic
			if ((cur & m) != m) atomicOr(reinterpret_cast<unsigned*>(mask) + (n >> 2), m << (8 * (n & 3)));
		}
	}
}

// children of node n: contiguous at childBase[n], ascending child index == Morton order
__global__ void __launch_bounds__(VX_THREADS) k_children(uint64_t N, const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask,
                                                          const uint32_t* __restrict__ childBase, uint64_t* __restrict__ ccode) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	unsigned m = mask[n];
	uint64_t cd = code[n] << 3;
	uint32_t o = childBase[n];
	while (m) {
		int c = __ffs(m) - 1;
		m &= m - 1;
		ccode[o++] = cd | (uint64_t)c;
	}
}

// emit the child pairs of every pair; first-touch triangle by atomicMin (pairs are sorted by
// triangle, so after the first touch the pre-check load filters almost every later atomic)
__global__ void __launch_bounds__(VX_THREADS) k_emit(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                      const uint8_t* __restrict__ hit, const uint32_t* __restrict__ poff,
                                                      const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase,
                                                      uint32_t* __restrict__ otri, uint32_t* __restrict__ onode, uint32_t* __restrict__ ctstar) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	unsigned m = hit[p];
	if (!m) return;
	uint32_t t = ptri[p], n = pnode[p];
	unsigned nm = mask[n];
	uint32_t base = childBase[n];
	uint32_t o = poff[p];
	while (m) {
		int c = __ffs(m) - 1;
		m &= m - 1;
		uint32_t child = base + __popc(nm & ((1u << c) - 1));
		otri[o] = t;
		onode[o] = child;
		++o;
		if (ctstar[child] > t) atomicMin(&ctstar[child], t);
	}
}

__global__ void k_init_roots(uint32_t ntiles, uint64_t* code, uint32_t* tstar) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ntiles) { code[i] = i; tstar[i] = 0; }
}

uint64_t read_u64(cudaStream_t s, const uint64_t* d) {
	uint64_t h = 0;
	SVB_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	return h;
}

}  // namespace

// ------------------------------------------------------------------ host drivers
void make_root_pairs(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid,
                     const int* d_gridTile, int seq_lo, int seq_hi, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, uint64_t& P) {
	GridDesc g;
	g.ox = grid.ox; g.oy = grid.oy; g.oz = grid.oz;
	g.inv_cell = 1.0 / grid.cell;
	g.margin = grid.cell * 1e-6;
	g.G = grid.G;
	g.seq_lo = seq_lo; g.seq_hi = seq_hi;
	DevBuf<uint32_t> cnt(pool, T);
	DevBuf<uint64_t> tot(pool, 1);
	unsigned nb = blocks_for(T, VX_THREADS);
	k_candidates<false><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, cnt.p, nullptr, nullptr);
	SVB_KERNEL_CHECK();
	scan_u32(s, pool, cnt.p, T, cnt.p, tot.p);
	P = read_u64(s, tot.p);
	if (P >= 0xFFFFFFF0ull) throw BatchTooBig();
	ptri.reset(pool, P);
	pnode.reset(pool, P);
	k_candidates<true><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, cnt.p, ptri.p, pnode.p);
	SVB_KERNEL_CHECK();
}

void voxelize_batch(cudaStream_t s, Pool& pool, const float* d_tris, const TileGeom* d_tiles, uint32_t ntiles, int Lt,
                    DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, uint64_t P, uint64_t budget_bytes,
                    std::vector<BatchLevel>& lv, uint64_t& pairsTotal) {
	lv.clear();
	lv.resize(Lt);
	lv[0].n = ntiles;
	lv[0].code.reset(pool, ntiles);
	lv[0].tstar.reset(pool, ntiles);
	k_init_roots<<<blocks_for(ntiles, 256), 256, 0, s>>>(ntiles, lv[0].code.p, lv[0].tstar.p);
	SVB_KERNEL_CHECK();
	DevBuf<uint64_t> tot(pool, 1);
	pairsTotal = 0;
	for (int l = 0; l < Lt; ++l) {
		BatchLevel& L = lv[l];
		L.mask.reset(pool, (L.n + 3 + 16) & ~3ull);
		L.mask.zero();
		DevBuf<uint8_t> hit(pool, P + 16);
		if (P) {
			if (P > (1ull << 59)) throw Error(SVB_ERANGE, "too many pairs");
			k_classify<<<blocks_for(P * 8, VX_THREADS), VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p);
			SVB_KERNEL_CHECK();
		}
		pairsTotal += P;
		if (l == Lt - 1) break;
		// children
		L.childBase.reset(pool, L.n);
		scan_popc8(s, pool, L.mask.p, L.n, L.childBase.p, tot.p);
		uint64_t Nn = read_u64(s, tot.p);
		// next pairs
		DevBuf<uint32_t> poff(pool, P);
		scan_popc8(s, pool, hit.p, P, poff.p, tot.p);
		uint64_t Pn = read_u64(s, tot.p);
		if (Nn >= 0xFFFFFFF0ull || Pn >= 0xFFFFFFF0ull) throw BatchTooBig();
		if (budget_bytes && pool.live + Nn * 13 + Pn * 9 > budget_bytes) throw BatchTooBig();
		BatchLevel& C = lv[l + 1];
		C.n = Nn;
		C.code.reset(pool, Nn);
		C.tstar.reset(pool, Nn);
		C.tstar.fill_ff();
		k_children<<<blocks_for(L.n, VX_THREADS), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, L.childBase.p, C.code.p);
		SVB_KERNEL_CHECK();
		DevBuf<uint32_t> ntri(pool, Pn), nnode(pool, Pn);
		if (P) {
			k_emit<<<blocks_for(P, VX_THREADS), VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, hit.p, poff.p, L.mask.p, L.childBase.p, ntri.p, nnode.p, C.tstar.p);
			SVB_KERNEL_CHECK();
		}
		ptri = std::move(ntri);
		pnode = std::move(nnode);
		P = Pn;
	}
	ptri.release();
	pnode.release();
}

}  // namespace svb
