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
#include <cstdlib>

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
};

__device__ inline void tile_range(double mn, double mx, double o, const GridDesc& g, int& a, int& b) {
	double fa = floor((mn - g.margin - o) * g.inv_cell);
	double fb = floor((mx + g.margin - o) * g.inv_cell);
	a = fa < 0.0 ? 0 : (fa > (double)(g.G - 1) ? g.G : (int)fa);
	b = fb < 0.0 ? -1 : (fb > (double)(g.G - 1) ? g.G - 1 : (int)fb);
}

template <bool EMIT>
__global__ void __launch_bounds__(VX_THREADS) k_candidates(const float* __restrict__ tris, uint64_t T, GridDesc g,
                                                            const int* __restrict__ gridTile, const int* __restrict__ localOf, uint32_t* __restrict__ cnt_or_off,
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
				int s = gridTile[((size_t)x * g.G + y) * g.G + z];   // global tile_seq of the cell, -1 = no tile
				int loc = (s >= 0) ? localOf[s] : -1;                // index inside this batch, -1 = other batch / other rank
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

// ------------------------------------------------------------------ classify, filtered (default)
// One thread per pair decides all 8 children.  A cheap FP64 filter, evaluated relative to the PARENT
// centre and shared between the children, decides every child whose 13 separating-axis inequalities
// hold or fail with a margin far above any rounding error (tolerances 2^-40 relative, i.e. >= 4000x
// the worst-case accumulated error of either evaluation order); the few children that sit within
// that margin of a threshold (exact ties such as a wall lying in a voxel face) are re-decided by the
// reference-order predicate tri_box_overlap().  The result is therefore bit-identical to testing all
// 8 children with tri_box_overlap() (k_classify above, kept selectable with SVB_CLASSIFY=exact and
// compared against in tests/test_gpu_parity.py), at ~1/5 of the FP64 work for large triangles:
// interior nodes pass all nine edge axes at the parent level and never evaluate them per child.
__device__ __forceinline__ void node_centre(uint64_t cd, int l, const TileGeom& tg, double& cx, double& cy, double& cz, double& k) {
	cx = tg.cx; cy = tg.cy; cz = tg.cz;
	k = tg.rootSide * 0.25;
	for (int d = l - 1; d >= 0; --d) {
		int dig = (int)((cd >> (3 * d)) & 7);
		cx = __dadd_rn(cx, (dig & 4) ? k : -k);
		cy = __dadd_rn(cy, (dig & 2) ? k : -k);
		cz = __dadd_rn(cz, (dig & 1) ? k : -k);
		k *= 0.5;
	}
}

// children whose index has bit `b` clear / set
#define SVB_LO(b) ((b) == 4 ? 0x0Fu : (b) == 2 ? 0x33u : 0x55u)
#define SVB_HI(b) ((b) == 4 ? 0xF0u : (b) == 2 ? 0xCCu : 0xAAu)

// Per-pair "settled axis" flags, inherited by every descendant pair of the same triangle: bit i set
// means separating axis i can never reject a box that lies inside the pair's node, so it is skipped
// from there on.  Large triangles settle all nine edge axes and the box axes a few levels above the
// leaves; their deep pairs then cost one plane evaluation.
//   bits 0..8  edge axes (edge 0: X,Y,Z; edge 1: X,Y,Z; edge 2: X,Y,Z)     bits 9..11 box axes x,y,z
constexpr unsigned FL_BOX = 9;

// One edge-cross axis: p = ca*v[A] + cb*v[B] on the two vertices the reference projects; the child with
// signs (sA,sB) sees p - k*(ca*sA + cb*sB) against rad = (|ca|+|cb|)*k.  Straight-line over the four
// sign combinations (each shared by two children).
template <unsigned BITA, unsigned BITB>
__device__ __forceinline__ void edge_axis(unsigned flbit, bool degenerate, double ca, double cb, double viA, double viB, double vjA, double vjB,
                                          double k, double tol2, unsigned& alive, unsigned& unsure, unsigned& fl) {
	// `degenerate`: both coefficients are differences of bitwise-equal float inputs, hence exactly 0 for any
	// box centre in the reference-order predicate too: p0 = p1 = +-0, rad = 0, neither "min > rad" nor
	// "max < -rad" can hold -- the axis never separates (axis-aligned edges).
	if (degenerate) { fl |= flbit; return; }
	const double pi = fma(ca, viA, cb * viB), pj = fma(ca, vjA, cb * vjB);
	const bool swap = pj < pi;
	const double mn = swap ? pj : pi, mx = swap ? pi : pj;
	const double rad = (fabs(ca) + fabs(cb)) * k;
	const double r2 = rad + rad;
	// the whole NODE (half side 2k) projects strictly inside the triangle's interval: no box inside it can be
	// separated on this axis, now or at any deeper level
	if (mn + r2 < -tol2 && mx - r2 > tol2) { fl |= flbit; return; }
	// |shift| <= rad for every child: parent centre strictly inside => every child overlaps on this axis
	if (mn < -tol2 && mx > tol2) return;
	if (mn > r2 + tol2 || mx < -r2 - tol2) { alive = 0; return; }
	const double qa = k * ca, qb = k * cb;
	const double R1 = rad + tol2, R0 = rad - tol2;
	const double spp = qa + qb, spm = qa - qb;
#define SVB_COMBO(SH, MASK)                                                       \
	{                                                                             \
		const double lo = mn - (SH), hi = mx - (SH);                              \
		if (lo > R1 || hi < -R1) alive &= ~(MASK);                                \
		else if (!(lo < R0 && hi > -R0)) unsure |= (MASK);                        \
	}
	SVB_COMBO(spp, SVB_HI(BITA) & SVB_HI(BITB))
	SVB_COMBO(spm, SVB_HI(BITA) & SVB_LO(BITB))
	SVB_COMBO(-spm, SVB_LO(BITA) & SVB_HI(BITB))
	SVB_COMBO(-spp, SVB_LO(BITA) & SVB_LO(BITB))
#undef SVB_COMBO
}

// children the filter could not decide: the reference-order predicate decides (kept out of line so that
// its registers do not burden the filter)
__device__ __noinline__ unsigned exact_children(unsigned unsure, double Cx, double Cy, double Cz, double k, const float* __restrict__ tp) {
	float tf[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) tf[i] = tp[i];
	unsigned m = 0;
	while (unsure) {
		int c = __ffs(unsure) - 1;
		unsure &= unsure - 1;
		double cx = __dadd_rn(Cx, (c & 4) ? k : -k), cy = __dadd_rn(Cy, (c & 2) ? k : -k), cz = __dadd_rn(Cz, (c & 1) ? k : -k);
		if (tri_box_overlap(cx, cy, cz, k, tf)) m |= 1u << c;
	}
	return m;
}

template <int MINB>
__global__ void __launch_bounds__(VX_THREADS, MINB) k_classify_filtered(uint64_t P, const uint32_t* __restrict__ ptri, const uint32_t* __restrict__ pnode,
                                                                   uint16_t* __restrict__ pflags, const uint64_t* __restrict__ code, int l,
                                                                   const TileGeom* __restrict__ tiles, const float* __restrict__ tris,
                                                                   uint8_t* __restrict__ hit, uint8_t* __restrict__ mask, unsigned long long* __restrict__ nExact) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	const uint32_t t = ptri[p], n = pnode[p];
	unsigned fl = pflags[p];
	const uint64_t cd = code[n];
	const TileGeom tg = tiles[(uint32_t)(cd >> (3 * l))];
	double Cx, Cy, Cz, k;
	node_centre(cd, l, tg, Cx, Cy, Cz, k);
	const float* tp = tris + 9ull * t;
	float tf[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) tf[i] = tp[i];
	const double v0x = (double)tf[0] - Cx, v0y = (double)tf[1] - Cy, v0z = (double)tf[2] - Cz;
	const double v1x = (double)tf[3] - Cx, v1y = (double)tf[4] - Cy, v1z = (double)tf[5] - Cz;
	const double v2x = (double)tf[6] - Cx, v2y = (double)tf[7] - Cy, v2z = (double)tf[8] - Cz;
	// Axis-aligned ("flat") triangles -- all three vertices share a bitwise-equal coordinate, as every face of
	// a box mesh does.  Say x is shared: the edge x-components are exactly 0 in the reference-order predicate
	// for any box centre, the normal is (nx, +-0, +-0), and
	//   * the six Y-/Z-type edge axes degenerate to  fl(|e|*|vx|) > fl(|e|*h)  (both projected vertices coincide),
	//   * the plane test degenerates to  vx < -h  or  vx > h,
	// all of which are implied false by the box-axis test on x (|vx| <= h; rounding is monotone).  So for such a
	// triangle the predicate IS box-x & box-y & box-z & the three X-type axes: the other seven tests are settled
	// from the start, rigorously (no tolerance involved).
	bool planeImplied = false;
	if (tf[0] == tf[3] && tf[3] == tf[6]) { fl |= 0x1B6u; planeImplied = true; }   // Y,Z-type axes of all edges
	if (tf[1] == tf[4] && tf[4] == tf[7]) { fl |= 0x16Du; planeImplied = true; }   // X,Z-type
	if (tf[2] == tf[5] && tf[5] == tf[8]) { fl |= 0x0DBu; planeImplied = true; }   // X,Y-type
	const double k2 = k + k;
	double M = fmax(fmax(fmax(fabs(v0x), fabs(v0y)), fmax(fabs(v0z), fabs(v1x))), fmax(fmax(fabs(v1y), fabs(v1z)), fmax(fabs(v2x), fmax(fabs(v2y), fabs(v2z))))) + k2;
	const double eps = 9.094947017729282e-13;   // 2^-40
	const double tol1 = M * eps, tol2 = M * tol1, tol3 = M * tol2;
	unsigned alive = 0xFFu, unsure = 0;
	// --- box axes: child with bit clear sits at -k, with bit set at +k
#define SVB_BOX_AXIS(a0, a1, a2, BIT, FLB)                                                     \
	if (!(fl & (FLB))) {                                                                       \
		double mn = fmin(fmin(a0, a1), a2), mx = fmax(fmax(a0, a1), a2);                       \
		if (mn + k2 < -tol1 && mx - k2 > tol1) fl |= (FLB);   /* node strictly inside the triangle's slab */ \
		else {                                                                                 \
			if (mn > tol1 || mx < -k2 - tol1) alive &= ~SVB_LO(BIT);                           \
			else if (!(mn < -tol1 && mx > -k2 + tol1)) unsure |= SVB_LO(BIT);                  \
			if (mn > k2 + tol1 || mx < -tol1) alive &= ~SVB_HI(BIT);                           \
			else if (!(mn < k2 - tol1 && mx > tol1)) unsure |= SVB_HI(BIT);                    \
		}                                                                                      \
	}
	SVB_BOX_AXIS(v0x, v1x, v2x, 4, 1u << (FL_BOX + 0))
	SVB_BOX_AXIS(v0y, v1y, v2y, 2, 1u << (FL_BOX + 1))
	SVB_BOX_AXIS(v0z, v1z, v2z, 1, 1u << (FL_BOX + 2))
#undef SVB_BOX_AXIS
	if (alive) {
		const double e0x = v1x - v0x, e0y = v1y - v0y, e0z = v1z - v0z;
		const double e1x = v2x - v1x, e1y = v2y - v1y, e1z = v2z - v1z;
		// --- plane: overlap <=> |N.v0| <= k*(|Nx|+|Ny|+|Nz|); straight-line over the 8 sign combinations
		const double nx = fma(e0y, e1z, -(e0z * e1y)), ny = fma(e0z, e1x, -(e0x * e1z)), nz = fma(e0x, e1y, -(e0y * e1x));
		const double g = fma(nx, v0x, fma(ny, v0y, nz * v0z));
		const double r = k * (fabs(nx) + fabs(ny) + fabs(nz));
		const double dx = k * nx, dy = k * ny, dz = k * nz;
		if (!planeImplied) {
			const double rp = r + tol3, rm = r - tol3;
			const double g0 = g + dx, g1 = g - dx;                      // x bit clear / set
			const double g00 = g0 + dy, g01 = g0 - dy, g10 = g1 + dy, g11 = g1 - dy;
#define SVB_PLANE(GV, C)                                                          \
			{                                                                     \
				const double a = fabs(GV);                                        \
				if (a > rp) alive &= ~(1u << (C));                                \
				else if (a > rm) unsure |= 1u << (C);                             \
			}
			SVB_PLANE(g00 + dz, 0) SVB_PLANE(g00 - dz, 1) SVB_PLANE(g01 + dz, 2) SVB_PLANE(g01 - dz, 3)
			SVB_PLANE(g10 + dz, 4) SVB_PLANE(g10 - dz, 5) SVB_PLANE(g11 + dz, 6) SVB_PLANE(g11 - dz, 7)
#undef SVB_PLANE
		}
		if (alive && (fl & 0x1FFu) != 0x1FFu) {
			const double e2x = v0x - v2x, e2y = v0y - v2y, e2z = v0z - v2z;
			// bitwise-equal input coordinates => that edge component is exactly zero in either evaluation
			const bool x01 = tf[0] == tf[3], y01 = tf[1] == tf[4], z01 = tf[2] == tf[5];
			const bool x12 = tf[3] == tf[6], y12 = tf[4] == tf[7], z12 = tf[5] == tf[8];
			const bool x20 = tf[6] == tf[0], y20 = tf[7] == tf[1], z20 = tf[8] == tf[2];
			// edge 0: X01(v0,v2)  Y02(v0,v2)  Z12(v1,v2)      p_X = ez*vy - ey*vz, p_Y = -ez*vx + ex*vz, p_Z = ey*vx - ex*vy
			if (!(fl & 0x001u)) edge_axis<2, 1>(0x001u, z01 && y01, e0z, -e0y, v0y, v0z, v2y, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x002u)) edge_axis<4, 1>(0x002u, z01 && x01, -e0z, e0x, v0x, v0z, v2x, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x004u)) edge_axis<4, 2>(0x004u, y01 && x01, e0y, -e0x, v1x, v1y, v2x, v2y, k, tol2, alive, unsure, fl);
			// edge 1: X01(v0,v2)  Y02(v0,v2)  Z0(v0,v1)
			if (alive && !(fl & 0x008u)) edge_axis<2, 1>(0x008u, z12 && y12, e1z, -e1y, v0y, v0z, v2y, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x010u)) edge_axis<4, 1>(0x010u, z12 && x12, -e1z, e1x, v0x, v0z, v2x, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x020u)) edge_axis<4, 2>(0x020u, y12 && x12, e1y, -e1x, v0x, v0y, v1x, v1y, k, tol2, alive, unsure, fl);
			// edge 2: X2(v0,v1)  Y1(v0,v1)  Z12(v1,v2)
			if (alive && !(fl & 0x040u)) edge_axis<2, 1>(0x040u, z20 && y20, e2z, -e2y, v0y, v0z, v1y, v1z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x080u)) edge_axis<4, 1>(0x080u, z20 && x20, -e2z, e2x, v0x, v0z, v1x, v1z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x100u)) edge_axis<4, 2>(0x100u, y20 && x20, e2y, -e2x, v1x, v1y, v2x, v2y, k, tol2, alive, unsure, fl);
		}
	}
	unsure &= alive;
	unsigned m = alive & ~unsure;
	if (unsure) {
		m |= exact_children(unsure, Cx, Cy, Cz, k, tp);
		if (nExact) atomicAdd(nExact, (unsigned long long)__popc(unsure));
	}
	hit[p] = (uint8_t)m;
	pflags[p] = (uint16_t)fl;   // inherited by the child pairs (k_emit)
	if (m) {
		unsigned cur = mask[n];
		if ((cur & m) != m) atomicOr(reinterpret_cast<unsigned*>(mask) + (n >> 2), m << (8 * (n & 3)));
	}
}
#undef SVB_LO
#undef SVB_HI

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
                                                      const uint16_t* __restrict__ pflags, const uint8_t* __restrict__ hit, const uint32_t* __restrict__ poff,
                                                      const uint8_t* __restrict__ mask, const uint32_t* __restrict__ childBase,
                                                      uint32_t* __restrict__ otri, uint32_t* __restrict__ onode, uint16_t* __restrict__ oflags, uint32_t* __restrict__ ctstar) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	unsigned m = hit[p];
	if (!m) return;
	uint32_t t = ptri[p], n = pnode[p];
	const uint16_t fl = pflags[p];
	unsigned nm = mask[n];
	uint32_t base = childBase[n];
	uint32_t o = poff[p];
	while (m) {
		int c = __ffs(m) - 1;
		m &= m - 1;
		uint32_t child = base + __popc(nm & ((1u << c) - 1));
		otri[o] = t;
		onode[o] = child;
		oflags[o] = fl;
		++o;
		if (ctstar[child] > t) atomicMin(&ctstar[child], t);
	}
}

// nodes the next level will hold, per tile (weights for cutting an oversized batch)
__global__ void __launch_bounds__(VX_THREADS) k_tile_weights(uint64_t N, const uint64_t* __restrict__ code, const uint8_t* __restrict__ mask, int l, uint32_t* __restrict__ w) {
	uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	unsigned m = mask[n];
	if (m) atomicAdd(&w[(uint32_t)(code[n] >> (3 * l))], (uint32_t)__popc(m));
}

__global__ void k_init_roots(uint32_t ntiles, uint64_t* code, uint32_t* tstar) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ntiles) { code[i] = i; tstar[i] = 0; }
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

// ------------------------------------------------------------------ host drivers
void make_root_pairs(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid,
                     const int* d_gridTile, const int* d_localOf, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, uint64_t& P) {
	GridDesc g;
	g.ox = grid.ox; g.oy = grid.oy; g.oz = grid.oz;
	g.inv_cell = 1.0 / grid.cell;
	g.margin = grid.cell * 1e-6;
	g.G = grid.G;
	DevBuf<uint32_t> cnt(pool, T);
	DevBuf<uint64_t> tot(pool, 1);
	unsigned nb = blocks_for(T, VX_THREADS);
	k_candidates<false><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, d_localOf, cnt.p, nullptr, nullptr);
	SVB_KERNEL_CHECK();
	scan_u32(s, pool, cnt.p, T, cnt.p, tot.p);
	P = read_u64(s, tot.p);
	if (P >= 0xFFFFFFF0ull) throw BatchTooBig();
	ptri.reset(pool, P);
	pnode.reset(pool, P);
	k_candidates<true><<<nb, VX_THREADS, 0, s>>>(d_tris, T, g, d_gridTile, d_localOf, cnt.p, ptri.p, pnode.p);
	SVB_KERNEL_CHECK();
}

void voxelize_batch(cudaStream_t s, Pool& pool, const float* d_tris, const TileGeom* d_tiles, uint32_t ntiles, int Lt,
                    DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, uint64_t P, uint64_t budget_bytes, uint64_t nodeCap,
                    std::vector<BatchLevel>& lv, uint64_t& pairsTotal, uint64_t* d_nExact) {
	lv.clear();
	lv.resize(Lt);
	lv[0].n = ntiles;
	lv[0].code.reset(pool, ntiles);
	lv[0].tstar.reset(pool, ntiles);
	k_init_roots<<<blocks_for(ntiles, 256), 256, 0, s>>>(ntiles, lv[0].code.p, lv[0].tstar.p);
	SVB_KERNEL_CHECK();
	DevBuf<uint64_t> tot(pool, 1);
	pairsTotal = 0;
	DevBuf<uint16_t> pflags(pool, P + 8);   // settled-axis flags per pair (see k_classify_filtered)
	pflags.zero();
	for (int l = 0; l < Lt; ++l) {
		BatchLevel& L = lv[l];
		L.mask.reset(pool, (L.n + 3 + 16) & ~3ull);
		L.mask.zero();
		DevBuf<uint8_t> hit(pool, P + 16);
		if (P) {
			if (P > (1ull << 59)) throw Error(SVB_ERANGE, "too many pairs");
			if (classify_exact_only())
				k_classify<<<blocks_for(P * 8, VX_THREADS), VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p);
			else {
				static const int occ = [] { const char* e = getenv("SVB_VX_OCC"); return e ? atoi(e) : 4; }();
				unsigned nb = blocks_for(P, VX_THREADS);
				if (occ >= 6) k_classify_filtered<6><<<nb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p, (unsigned long long*)d_nExact);
				else if (occ == 5) k_classify_filtered<5><<<nb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p, (unsigned long long*)d_nExact);
				else if (occ >= 4) k_classify_filtered<4><<<nb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p, (unsigned long long*)d_nExact);
				else if (occ == 3) k_classify_filtered<3><<<nb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p, (unsigned long long*)d_nExact);
				else k_classify_filtered<1><<<nb, VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, L.code.p, l, d_tiles, d_tris, hit.p, L.mask.p, (unsigned long long*)d_nExact);
			}
			SVB_KERNEL_CHECK();
		}
		pairsTotal += P;
		if (getenv("SVB_VX_STATS") && P) {   // debug: how many axes are settled per pair at this level
			uint64_t ns = P < 4000000 ? P : 4000000;
			std::vector<uint16_t> hf(ns);
			std::vector<uint8_t> hh(ns);
			SVB_CUDA(cudaMemcpyAsync(hf.data(), pflags.p + (P - ns) / 2, ns * 2, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaMemcpyAsync(hh.data(), hit.p + (P - ns) / 2, ns, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaStreamSynchronize(s));
			uint64_t hist[13] = {0}, allEdge = 0, allBox = 0, hits = 0;
			for (uint64_t i = 0; i < ns; ++i) { hist[__builtin_popcount(hf[i])]++; allEdge += (hf[i] & 0x1FF) == 0x1FF; allBox += (hf[i] >> 9) == 7; hits += __builtin_popcount(hh[i]); }
			fprintf(stderr, "[vx-stats] level %d P=%llu sample=%llu allEdge=%.3f allBox=%.3f hits/pair=%.2f hist:", l, (unsigned long long)P, (unsigned long long)ns,
			        (double)allEdge / ns, (double)allBox / ns, (double)hits / ns);
			for (int i = 0; i <= 12; ++i) fprintf(stderr, " %.3f", (double)hist[i] / ns);
			fprintf(stderr, "\n");
		}
		if (l == Lt - 1) break;
		// children
		L.childBase.reset(pool, L.n);
		scan_popc8(s, pool, L.mask.p, L.n, L.childBase.p, tot.p);
		uint64_t Nn = read_u64(s, tot.p);
		// next pairs
		DevBuf<uint32_t> poff(pool, P);
		scan_popc8(s, pool, hit.p, P, poff.p, tot.p);
		uint64_t Pn = read_u64(s, tot.p);
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
		C.code.reset(pool, Nn);
		C.tstar.reset(pool, Nn);
		C.tstar.fill_ff();
		k_children<<<blocks_for(L.n, VX_THREADS), VX_THREADS, 0, s>>>(L.n, L.code.p, L.mask.p, L.childBase.p, C.code.p);
		SVB_KERNEL_CHECK();
		DevBuf<uint32_t> ntri(pool, Pn), nnode(pool, Pn);
		DevBuf<uint16_t> nflags(pool, Pn + 8);
		if (P) {
			k_emit<<<blocks_for(P, VX_THREADS), VX_THREADS, 0, s>>>(P, ptri.p, pnode.p, pflags.p, hit.p, poff.p, L.mask.p, L.childBase.p, ntri.p, nnode.p, nflags.p, C.tstar.p);
			SVB_KERNEL_CHECK();
		}
		ptri = std::move(ntri);
		pnode = std::move(nnode);
		pflags = std::move(nflags);
		P = Pn;
	}
	ptri.release();
	pnode.release();
}

}  // namespace svb
