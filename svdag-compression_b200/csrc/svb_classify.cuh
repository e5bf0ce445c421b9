// svb_classify.cuh -- the per-pair decision of the voxelizer: which of the 8 children of a node does a triangle
// overlap, bit-identical with 8 calls of the reference's testTriBox (src/symvox/test_triangle_box.cpp:105-184).
//
// Written as host/device code: libsvb.so only ever runs it inside k_classify_filtered (svb_voxelize.cu);
// tests/classify_harness.cpp compiles the same source with g++ so that the filter can be checked against the
// reference-order predicate on the CPU, pair by pair (test infrastructure, not a fallback).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>

#include "svb_sat.cuh"

namespace svb {

#ifndef SVB_TILEGEOM_DEFINED
#define SVB_TILEGEOM_DEFINED
struct TileGeom {   // host computed, exactly as geom_octree.cpp:177-184,214 / :340-344 would
	double cx, cy, cz;   // root centre (double)
	double rootSide;     // float-rounded max side, widened
};
#endif

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
SVB_HD void node_centre(uint64_t cd, int l, const TileGeom& tg, double& cx, double& cy, double& cz, double& k) {
	cx = tg.cx; cy = tg.cy; cz = tg.cz;
	k = tg.rootSide * 0.25;
	for (int d = l - 1; d >= 0; --d) {
		int dig = (int)((cd >> (3 * d)) & 7);
		cx = SVB_DADD(cx, (dig & 4) ? k : -k);
		cy = SVB_DADD(cy, (dig & 2) ? k : -k);
		cz = SVB_DADD(cz, (dig & 1) ? k : -k);
		k *= 0.5;
	}
}

// Direct form of the same chain.  Every partial sum of the chain is c0 + (integer) * k_finest; when the host
// has verified that all those values are representable doubles for every tile of the batch (centre_chain_exact()),
// each rounding of the chain is the identity and the chain equals  c0 + k * (4X - 2^(l+1) + 2),  X = the node's
// integer coordinate inside the tile (de-interleaved path), evaluated with one exact product and one exact sum.
SVB_HD uint32_t compact3(uint64_t v, int l) {   // bits 0,3,6,... of v -> contiguous
	if (l <= 10) {
		uint32_t x = (uint32_t)v & 0x09249249u;
		x = (x ^ (x >> 2)) & 0x030c30c3u;
		x = (x ^ (x >> 4)) & 0x0300f00fu;
		x = (x ^ (x >> 8)) & 0x030000ffu;
		x = (x ^ (x >> 16)) & 0x000003ffu;
		return x;
	}
	v &= 0x1249249249249249ull;
	v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
	v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
	v = (v ^ (v >> 8)) & 0x001f0000ff0000ffull;
	v = (v ^ (v >> 16)) & 0x001f00000000ffffull;
	v = (v ^ (v >> 32)) & 0x00000000001fffffull;
	return (uint32_t)v;
}
// one coordinate of the node centre; sh = 2 (x), 1 (y), 0 (z); k = child half side of level l
SVB_HD double centre_axis_direct(uint64_t path, int l, int sh, double c0, double k) {
	const int X = (int)compact3(path >> sh, l);
	const int m = 4 * X - (2 << l) + 2;
	return SVB_FMA((double)m, k, c0);
}
SVB_HD double centre_axis_chain(uint64_t cd, int l, int sh, double c0, double rootSide) {
	double c = c0, k = rootSide * 0.25;
	for (int d = l - 1; d >= 0; --d) {
		c = SVB_DADD(c, ((cd >> (3 * d + sh)) & 1) ? k : -k);
		k *= 0.5;
	}
	return c;
}

// children whose index has bit `b` clear / set
#define SVB_LO(b) ((b) == 4 ? 0x0Fu : (b) == 2 ? 0x33u : 0x55u)
#define SVB_HI(b) ((b) == 4 ? 0xF0u : (b) == 2 ? 0xCCu : 0xAAu)

// Per-pair "settled axis" flags, inherited by every descendant pair of the same triangle: bit i set
// means separating axis i can never reject a box that lies inside the pair's node, so it is skipped
// from there on.  Large triangles settle all nine edge axes and the box axes a few levels above the
// leaves; their deep pairs then cost one plane evaluation.
//   bits 0..8  edge axes (edge 0: X,Y,Z; edge 1: X,Y,Z; edge 2: X,Y,Z)     bits 9..11 box axes x,y,z
//   bits 12..14  the triangle is flat on x / y / z (all three vertices share that coordinate bitwise)
//   bit 15       the static analysis of the triangle (flat bits, edge axes implied by box axes) has been done
constexpr unsigned FL_BOX = 9;
constexpr unsigned FL_FLAT = 12;
constexpr unsigned FL_INIT = 1u << 15;

// One edge-cross axis: p = ca*v[A] + cb*v[B] on the two vertices the reference projects; the child with
// signs (sA,sB) sees p - k*(ca*sA + cb*sB) against rad = (|ca|+|cb|)*k.  Straight-line over the four
// sign combinations (each shared by two children).
template <unsigned BITA, unsigned BITB>
SVB_HD void edge_axis(unsigned flbit, double ca, double cb, double viA, double viB, double vjA, double vjB,
                      double k, double tol2, unsigned& alive, unsigned& unsure, unsigned& fl) {
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
SVB_HD_NOINLINE unsigned exact_children(unsigned unsure, double Cx, double Cy, double Cz, double k, const float* __restrict__ tp) {
	float tf[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) tf[i] = tp[i];
	unsigned m = 0;
	while (unsure) {
		int c = SVB_FFS(unsure) - 1;
		unsure &= unsure - 1;
		double cx = SVB_DADD(Cx, (c & 4) ? k : -k), cy = SVB_DADD(Cy, (c & 2) ? k : -k), cz = SVB_DADD(Cz, (c & 1) ? k : -k);
		if (tri_box_overlap(cx, cy, cz, k, tf)) m |= 1u << c;
	}
	return m;
}

// ---- the "flat" pair stream: a flat (axis-aligned) triangle, say on axis a, all of whose nine edge axes are settled.
// The plane test and the six edge axes that involve the a-component are implied by the box test on a (see
// classify_pair), the other edge axes are settled for every box inside the pair's node, so the predicate of a
// child IS the conjunction of the reference's three box-axis tests (test_triangle_box.cpp:165-174):
//   mn = fl(min_i t_i - c), mx = fl(max_i t_i - c);  overlap <=> !(mn > k || mx < -k)      per axis,
// (fl(t - c) is monotone in t, so the min/max may be taken on the float inputs) evaluated here in the reference's
// own operation order at the chain-rounded child centres -- exact, no tolerance.  Box axes that are already settled
// are skipped; an in-plane axis becomes settled once the node lies strictly inside the triangle's slab.  The edge
// flags of such a pair never change, so all its descendants stay in this stream (k_classify_fast).
SVB_HD bool pair_is_fast(unsigned fl) { return (fl & 0x1FFu) == 0x1FFu && (fl & (7u << FL_FLAT)) != 0; }

// children whose index has the bit of axis a (0 = x, 1 = y, 2 = z) clear
SVB_HD unsigned axis_lo_mask(int a) { return (0x55330Fu >> (8 * a)) & 0xFFu; }

// The reference's box test of axis a (test_triangle_box.cpp:165-174) for the two child positions c = fl(C -+ k) of a
// node centred at C (child half side k): children with the axis bit clear / set that pass.  dmin/dmax: the triangle's
// extent on the axis (float inputs widened).  Exact: same operations, same order, same roundings as the reference.
SVB_HD unsigned box_axis_children(const int a, const double C, const double k, const double dmin, const double dmax) {
	const double cLo = SVB_DADD(C, -k), cHi = SVB_DADD(C, k);
	const unsigned lo = axis_lo_mask(a);
	unsigned pass = 0;
	if (!(SVB_DSUB(dmin, cLo) > k || SVB_DSUB(dmax, cLo) < -k)) pass |= lo;
	if (!(SVB_DSUB(dmin, cHi) > k || SVB_DSUB(dmax, cHi) < -k)) pass |= lo ^ 0xFFu;
	return pass;
}
// node (half side 2k) strictly inside the triangle's slab on this axis, with a margin (2^-40 relative) far above
// any rounding: the axis can never reject a box inside this node
SVB_HD bool box_axis_settled(const double C, const double k2, const double dmin, const double dmax) {
	const double mn = dmin - C, mx = dmax - C;
	const double tol = (fmax(fabs(mn), fabs(mx)) + k2) * 9.094947017729282e-13;
	return mn + k2 < -tol && mx - k2 > tol;
}
SVB_HD void axis_extent(const float* __restrict__ tp, const int a, const bool flat, double& dmin, double& dmax) {
	float tmin = tp[a], tmax = tmin;
	if (!flat) {
		const float f1 = tp[3 + a], f2 = tp[6 + a];
		tmin = fminf(tmin, fminf(f1, f2));
		tmax = fmaxf(tmax, fmaxf(f1, f2));
	}
	dmin = (double)tmin; dmax = (double)tmax;
}

// All unsettled box axes of a pair, exactly; settles the axes the node has moved strictly inside of.
// STATIC: three predicated axis blocks (best when most lanes of a warp have the same single unsettled axis: the flat
// stream) instead of a loop over the set bits (best inside the register-hungry general path).
template <bool DIRECT, bool STATIC = false>
SVB_HD unsigned box_axes_exact(const uint64_t cd, const int l, const double* __restrict__ tg4, const double k, const float* __restrict__ tp, unsigned& fl) {
	const double rootSide = tg4[3];
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	unsigned ub = (~fl >> FL_BOX) & 7u;   // unsettled box axes (bit 0 = x)
	unsigned m = 0xFFu;
	if (STATIC) {
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			if (!((ub >> a) & 1u)) continue;
			double dmin, dmax;
			axis_extent(tp, a, ((fl >> (FL_FLAT + a)) & 1u) != 0, dmin, dmax);
			const double C = DIRECT ? centre_axis_direct(path, l, 2 - a, tg4[a], k) : centre_axis_chain(cd, l, 2 - a, tg4[a], rootSide);
			m &= box_axis_children(a, C, k, dmin, dmax);
			if (box_axis_settled(C, k + k, dmin, dmax)) fl |= 1u << (FL_BOX + a);
		}
		return m;
	}
	while (ub) {
		const int a = SVB_FFS(ub) - 1;
		ub &= ub - 1;
		double dmin, dmax;
		axis_extent(tp, a, ((fl >> (FL_FLAT + a)) & 1u) != 0, dmin, dmax);
		const double C = DIRECT ? centre_axis_direct(path, l, 2 - a, tg4[a], k) : centre_axis_chain(cd, l, 2 - a, tg4[a], rootSide);
		m &= box_axis_children(a, C, k, dmin, dmax);
		if (box_axis_settled(C, k + k, dmin, dmax)) fl |= 1u << (FL_BOX + a);
	}
	return m;
}

template <bool DIRECT>
SVB_HD unsigned classify_pair_flat(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp, unsigned& fl) {
	return box_axes_exact<DIRECT, true>(cd, l, tg4, tg4[3] * kscale, tp, fl);
}

// Flat-stream pair at the second-to-last level: the voxel masks of its children (the leaf nodes), without emitting
// child pairs.  Child c sits at fl(C +- k) on every axis (the next step of the centre chain); its voxels at
// fl(c +- k/2).  Per unsettled axis the two possible child positions are tested once; childMask[j] (j = the child's
// bit on that axis) then combine by AND.  out[c] is only meaningful for children the pair hits.
template <bool DIRECT>
SVB_HD void flat_leaf_masks(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp,
                            const unsigned fl, unsigned lohi[3][2]) {
	const double rootSide = tg4[3];
	const double k = rootSide * kscale, kh = k * 0.5;
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	const unsigned ub = (~fl >> FL_BOX) & 7u;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		lohi[a][0] = lohi[a][1] = 0xFFu;
		if (!((ub >> a) & 1u)) continue;
		double dmin, dmax;
		axis_extent(tp, a, ((fl >> (FL_FLAT + a)) & 1u) != 0, dmin, dmax);
		const double C = DIRECT ? centre_axis_direct(path, l, 2 - a, tg4[a], k) : centre_axis_chain(cd, l, 2 - a, tg4[a], rootSide);
		lohi[a][0] = box_axis_children(a, SVB_DADD(C, -k), kh, dmin, dmax);
		lohi[a][1] = box_axis_children(a, SVB_DADD(C, k), kh, dmin, dmax);
	}
}

// Flat triangle (all vertices share coordinate A bitwise) in the slow stream: everything that is left of the predicate
// is (1) the exact 1-D box tests of the unsettled axes and (2) the A-type cross axes of the three edges -- a 2-D
// triangle-vs-square problem in the plane (U, W).  One routine with the axes known at compile time: the six in-plane
// vertex coordinates and the node centre are fetched / computed once and serve both parts.  The edge part is the same
// interval filter as the general path (edge_axis), on 6 instead of 9 vertex offsets and without the plane:
// p = e_W * v_U - e_U * v_W up to a sign the symmetric test does not see (X: ez*vy - ey*vz, Y: -ez*vx + ex*vz,
// Z: ey*vx - ex*vy, test_triangle_box.cpp:60-102); projected vertex pair = {an endpoint, the opposite vertex}.
// LAST: the pair has no children, so nothing needs to be settled for them.
template <bool DIRECT, int A, bool LAST>
SVB_HD unsigned classify_flat_slow(const uint64_t cd, const int l, const double* __restrict__ tg4, const double k, const float* __restrict__ tp,
                                   unsigned& fl, unsigned& nUnsure) {
	constexpr int U = (A == 0) ? 1 : 0, W = (A == 2) ? 1 : 2;
	constexpr unsigned BITU = (U == 0) ? 4u : 2u, BITW = (W == 2) ? 1u : 2u;
	constexpr unsigned E0 = 1u << A, E1 = 8u << A, E2 = 64u << A, EALL = E0 | E1 | E2;
	constexpr unsigned BA = 1u << (FL_BOX + A), BU = 1u << (FL_BOX + U), BW = 1u << (FL_BOX + W);
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	const double k2 = k + k;
	unsigned alive = 0xFFu;
	// ---- flat axis: one coordinate, never settles
	{
		const double CA = DIRECT ? centre_axis_direct(path, l, 2 - A, tg4[A], k) : centre_axis_chain(cd, l, 2 - A, tg4[A], tg4[3]);
		const double ta = (double)tp[A];
		alive &= box_axis_children(A, CA, k, ta, ta);
		if (!alive) return 0;
	}
	const bool needEdges = (fl & EALL) != EALL;
	if (!needEdges && (fl & (BU | BW)) == (BU | BW)) return alive;   // (such a pair belongs to the flat stream; kept for completeness)
	const float fu0 = tp[U], fu1 = tp[3 + U], fu2 = tp[6 + U], fw0 = tp[W], fw1 = tp[3 + W], fw2 = tp[6 + W];
	const double CU = DIRECT ? centre_axis_direct(path, l, 2 - U, tg4[U], k) : centre_axis_chain(cd, l, 2 - U, tg4[U], tg4[3]);
	const double CW = DIRECT ? centre_axis_direct(path, l, 2 - W, tg4[W], k) : centre_axis_chain(cd, l, 2 - W, tg4[W], tg4[3]);
	// ---- in-plane box axes (exact; min / max on the float inputs: fl(t - c) is monotone in t)
	if (!(fl & BU)) {
		const double dmin = (double)fminf(fu0, fminf(fu1, fu2)), dmax = (double)fmaxf(fu0, fmaxf(fu1, fu2));
		alive &= box_axis_children(U, CU, k, dmin, dmax);
		if (!LAST && box_axis_settled(CU, k2, dmin, dmax)) fl |= BU;
	}
	if (!(fl & BW)) {
		const double dmin = (double)fminf(fw0, fminf(fw1, fw2)), dmax = (double)fmaxf(fw0, fmaxf(fw1, fw2));
		alive &= box_axis_children(W, CW, k, dmin, dmax);
		if (!LAST && box_axis_settled(CW, k2, dmin, dmax)) fl |= BW;
	}
	if (!alive || !needEdges) return alive;
	// ---- in-plane edge axes (filter, then the reference-order predicate for children within the tolerance band)
	const double u0 = (double)fu0 - CU, w0 = (double)fw0 - CW;
	const double u1 = (double)fu1 - CU, w1 = (double)fw1 - CW;
	const double u2 = (double)fu2 - CU, w2 = (double)fw2 - CW;
	const double M = fmax(fmax(fmax(fabs(u0), fabs(w0)), fmax(fabs(u1), fabs(w1))), fmax(fabs(u2), fabs(w2))) + k2;
	const double tol2 = M * (M * 9.094947017729282e-13);
	unsigned unsure = 0;
	if (!(fl & E0)) edge_axis<BITU, BITW>(E0, w1 - w0, -(u1 - u0), u0, w0, u2, w2, k, tol2, alive, unsure, fl);                // edge 0: v0 -> v1, opposite v2
	if (alive && !(fl & E1)) edge_axis<BITU, BITW>(E1, w2 - w1, -(u2 - u1), u0, w0, u2, w2, k, tol2, alive, unsure, fl);       // edge 1: v1 -> v2, opposite v0
	if (alive && !(fl & E2)) edge_axis<BITU, BITW>(E2, w0 - w2, -(u0 - u2), u0, w0, u1, w1, k, tol2, alive, unsure, fl);       // edge 2: v2 -> v0, opposite v1
	unsure &= alive;
	unsigned m = alive & ~unsure;
	if (unsure) {
		const double CA = DIRECT ? centre_axis_direct(path, l, 2 - A, tg4[A], k) : centre_axis_chain(cd, l, 2 - A, tg4[A], tg4[3]);
		const double Cx = (A == 0) ? CA : CU, Cy = (A == 1) ? CA : ((A == 0) ? CU : CW), Cz = (A == 2) ? CA : CW;
		m |= exact_children(unsure, Cx, Cy, Cz, k, tp);
		nUnsure = (unsigned)SVB_POPC(unsure);
	}
	return m;
}

// ---- slow-stream pair of a FLAT triangle at the second-to-last level: the voxels of its children, decided in place
// (k_slow_leaves, svb_voxelize.cu) instead of emitting the child pairs and classifying them at the last level.
// Along each axis the 4 x 4 x 4 voxels under the node sit at C + (2i - 3) * kh, i = 2 * (child bit) + (voxel bit),
// kh = k / 2 the voxel half side, reached by the reference through two more steps of the centre chain
// (fl(fl(C +- k) +- kh), geom_octree.cpp:222-230).  All 64 voxels are carried in one 64-bit word, bit 8 c + v =
// voxel v of child c (both indexed X=4, Y=2, Z=1), so that a verdict on a whole slab / column of voxels is one mask.
//   Box axes: the reference's own 1-D tests (test_triangle_box.cpp:165-174) at those centres, one per slab -- exact.
//   Unsettled in-plane edge axes: the interval filter of edge_axis(), one level further down -- 16 in-plane voxel
//   columns per axis instead of 4 child positions, same projections, same 2^-40 margin (the shifts are <= 3 kh < 2 k,
//   inside the bound M the tolerance is built from).  The test "mn - s > rad" is evaluated as "s < mn - rad": another
//   rounding of the same order, far inside the margin.
SVB_HD constexpr uint64_t vox64_children(unsigned bit, bool hi) {   // every voxel of the children whose index has `bit` set / clear
	uint64_t m = 0;
	for (unsigned c = 0; c < 8; ++c)
		if (((c & bit) != 0) == hi) m |= 0xFFull << (8 * c);
	return m;
}
SVB_HD constexpr uint64_t vox64_voxels(unsigned bit, bool hi) {   // in every child, the voxels whose index has `bit` set / clear
	uint64_t m = 0;
	for (unsigned v = 0; v < 64; ++v)
		if (((v & bit) != 0) == hi) m |= 1ull << v;
	return m;
}
SVB_HD constexpr uint64_t vox64_slab(unsigned bit, int i) { return vox64_children(bit, (i >> 1) != 0) & vox64_voxels(bit, (i & 1) != 0); }   // position i = 0..3 along the axis

// box axis a, exactly: clears the slabs of `box` the reference's 1-D test rejects
template <int a>
SVB_HD void box_axis_slabs(const double C, const double k, const double kh, const double dmin, const double dmax, uint64_t& box) {
	constexpr unsigned BIT = 4u >> a;
	const double cLo = SVB_DADD(C, -k), cHi = SVB_DADD(C, k);   // child centres, then the voxel centres: the next two steps of the chain
	const double c0 = SVB_DADD(cLo, -kh), c1 = SVB_DADD(cLo, kh), c2 = SVB_DADD(cHi, -kh), c3 = SVB_DADD(cHi, kh);
	if (SVB_DSUB(dmin, c0) > kh || SVB_DSUB(dmax, c0) < -kh) box &= ~vox64_slab(BIT, 0);
	if (SVB_DSUB(dmin, c1) > kh || SVB_DSUB(dmax, c1) < -kh) box &= ~vox64_slab(BIT, 1);
	if (SVB_DSUB(dmin, c2) > kh || SVB_DSUB(dmax, c2) < -kh) box &= ~vox64_slab(BIT, 2);
	if (SVB_DSUB(dmin, c3) > kh || SVB_DSUB(dmax, c3) < -kh) box &= ~vox64_slab(BIT, 3);
}

// Flat-stream pair at the second-to-last level in the same 64-bit form (k_flat_leaves): only box axes are left of the
// predicate.  Byte c of the result is the voxel mask of child c; bytes of children the pair does not hit are meaningless.
template <bool DIRECT>
SVB_HD uint64_t flat_leaf_voxels(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp, const unsigned fl) {
	const double rootSide = tg4[3];
	const double k = rootSide * kscale, kh = k * 0.5;
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	const unsigned ub = (~fl >> FL_BOX) & 7u;
	uint64_t vox = ~0ull;
	if (ub & 1u) {
		double dmin, dmax;
		axis_extent(tp, 0, ((fl >> (FL_FLAT + 0)) & 1u) != 0, dmin, dmax);
		box_axis_slabs<0>(DIRECT ? centre_axis_direct(path, l, 2, tg4[0], k) : centre_axis_chain(cd, l, 2, tg4[0], rootSide), k, kh, dmin, dmax, vox);
	}
	if (ub & 2u) {
		double dmin, dmax;
		axis_extent(tp, 1, ((fl >> (FL_FLAT + 1)) & 1u) != 0, dmin, dmax);
		box_axis_slabs<1>(DIRECT ? centre_axis_direct(path, l, 1, tg4[1], k) : centre_axis_chain(cd, l, 1, tg4[1], rootSide), k, kh, dmin, dmax, vox);
	}
	if (ub & 4u) {
		double dmin, dmax;
		axis_extent(tp, 2, ((fl >> (FL_FLAT + 2)) & 1u) != 0, dmin, dmax);
		box_axis_slabs<2>(DIRECT ? centre_axis_direct(path, l, 0, tg4[2], k) : centre_axis_chain(cd, l, 0, tg4[2], rootSide), k, kh, dmin, dmax, vox);
	}
	return vox;
}

// one in-plane edge axis over the 16 voxel columns: rej / uns get the columns the axis separates / cannot decide
template <unsigned BITU, unsigned BITW>
SVB_HD void edge_axis16(double ca, double cb, double viA, double viB, double vjA, double vjB,
                        double kh, double tol2, uint64_t& rej, uint64_t& uns) {
	const double pi = fma(ca, viA, cb * viB), pj = fma(ca, vjA, cb * vjB);
	const bool swap = pj < pi;
	const double mn = swap ? pj : pi, mx = swap ? pi : pj;
	const double rad = (fabs(ca) + fabs(cb)) * kh;   // of one voxel
	const double r2 = rad + rad;
	// |shift| <= 3 rad for every voxel: all 16 columns overlap on this axis
	if (mn < -r2 - tol2 && mx > r2 + tol2) return;
	const double qa = kh * ca, qb = kh * cb, qa3 = 3.0 * qa, qb3 = 3.0 * qb;
	const double R1 = rad + tol2, R0 = rad - tol2;
	// column with shift s: separated <=> mn - s > R1 or mx - s < -R1; surely overlapping <=> mn - s < R0 and mx - s > -R0
	const double tLoRej = mn - R1, tHiRej = mx + R1, tLoAcc = mn - R0, tHiAcc = mx + R0;
#pragma unroll
	for (int iu = 0; iu < 4; ++iu) {
		const double sa = (iu == 0) ? -qa3 : (iu == 1) ? -qa : (iu == 2) ? qa : qa3;
#pragma unroll
		for (int iw = 0; iw < 4; ++iw) {
			const double sb = (iw == 0) ? -qb3 : (iw == 1) ? -qb : (iw == 2) ? qb : qb3;
			const double sh = sa + sb;
			const uint64_t col = vox64_slab(BITU, iu) & vox64_slab(BITW, iw);
			if (sh < tLoRej || sh > tHiRej) rej |= col;
			else if (!(sh > tLoAcc && sh < tHiAcc)) uns |= col;
		}
	}
}

// Bit 8 c + v of the result = voxel v of child c, for every child c in `m` (the pair's hit mask); fl = the pair's flags
// AFTER its own classification (edge / box axes settled for the node are settled for every voxel inside it).
template <bool DIRECT, int A>
SVB_HD uint64_t slow_leaf_voxels_axis(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp,
                                      const unsigned fl, const unsigned m, uint64_t& ask) {
	constexpr int U = (A == 0) ? 1 : 0, W = (A == 2) ? 1 : 2;
	constexpr unsigned BITU = 4u >> U, BITW = 4u >> W;
	constexpr unsigned E0 = 1u << A, E1 = 8u << A, E2 = 64u << A, EALL = E0 | E1 | E2;
	constexpr unsigned BU = 1u << (FL_BOX + U), BW = 1u << (FL_BOX + W);
	const double k = tg4[3] * kscale, kh = k * 0.5;
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	// the hit children
	uint64_t vox = 0;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		if ((m >> c) & 1u) vox |= 0xFFull << (8 * c);
	// ---- box axes (flat axis: one coordinate, never settles)
	{
		const double CA = DIRECT ? centre_axis_direct(path, l, 2 - A, tg4[A], k) : centre_axis_chain(cd, l, 2 - A, tg4[A], tg4[3]);
		const double ta = (double)tp[A];
		box_axis_slabs<A>(CA, k, kh, ta, ta, vox);
	}
	const bool needEdges = (fl & EALL) != EALL;
	if ((fl & (BU | BW)) == (BU | BW) && !needEdges) return vox;
	const float fu0 = tp[U], fu1 = tp[3 + U], fu2 = tp[6 + U], fw0 = tp[W], fw1 = tp[3 + W], fw2 = tp[6 + W];
	const double CU = DIRECT ? centre_axis_direct(path, l, 2 - U, tg4[U], k) : centre_axis_chain(cd, l, 2 - U, tg4[U], tg4[3]);
	const double CW = DIRECT ? centre_axis_direct(path, l, 2 - W, tg4[W], k) : centre_axis_chain(cd, l, 2 - W, tg4[W], tg4[3]);
	if (!(fl & BU)) box_axis_slabs<U>(CU, k, kh, (double)fminf(fu0, fminf(fu1, fu2)), (double)fmaxf(fu0, fmaxf(fu1, fu2)), vox);
	if (!(fl & BW)) box_axis_slabs<W>(CW, k, kh, (double)fminf(fw0, fminf(fw1, fw2)), (double)fmaxf(fw0, fmaxf(fw1, fw2)), vox);
	if (!needEdges || !vox) return vox;
	// ---- in-plane edge axes (same coefficient / vertex-pair pattern as classify_flat_slow)
	const double u0 = (double)fu0 - CU, w0 = (double)fw0 - CW;
	const double u1 = (double)fu1 - CU, w1 = (double)fw1 - CW;
	const double u2 = (double)fu2 - CU, w2 = (double)fw2 - CW;
	const double M = fmax(fmax(fmax(fabs(u0), fabs(w0)), fmax(fabs(u1), fabs(w1))), fmax(fabs(u2), fabs(w2))) + (k + k);
	const double tol2 = M * (M * 9.094947017729282e-13);
	uint64_t rej = 0, uns = 0;
	// one copy of the 16-column code for the three edges (a box mesh leaves one edge axis -- the hypotenuse's -- unsettled,
	// which of the three depends on the triangle: lanes with different edges share the instructions, and the kernel stays
	// small enough for the instruction cache)
#pragma unroll 1
	for (unsigned rem = ~fl & EALL; rem; rem &= rem - 1) {
		const unsigned e = rem & (0u - rem);
		// edge 0: v0 -> v1, projected pair (v0, v2); edge 1: v1 -> v2, (v0, v2); edge 2: v2 -> v0, (v0, v1)
		const double ca = (e == E0) ? w1 - w0 : (e == E1) ? w2 - w1 : w0 - w2;
		const double cb = (e == E0) ? -(u1 - u0) : (e == E1) ? -(u2 - u1) : -(u0 - u2);
		const double vjU = (e == E2) ? u1 : u2, vjW = (e == E2) ? w1 : w2;
		edge_axis16<BITU, BITW>(ca, cb, u0, w0, vjU, vjW, kh, tol2, rej, uns);
	}
	vox &= ~rej;
	ask = uns & vox;   // within the margin: slow_leaf_exact() decides (the caller calls it: rare, kept out of this function)
	return vox & ~ask;
}
// fl must carry a flat bit (FL_FLAT..): the triangle is flat on that axis.  ask: the voxels the filter could not decide
// (cleared in the result); the caller hands them to slow_leaf_exact().
template <bool DIRECT>
SVB_HD uint64_t slow_leaf_voxels(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp,
                                 const unsigned fl, const unsigned m, uint64_t& ask) {
	ask = 0;
	if (fl & (1u << (FL_FLAT + 0))) return slow_leaf_voxels_axis<DIRECT, 0>(cd, l, tg4, kscale, tp, fl, m, ask);
	if (fl & (1u << (FL_FLAT + 1))) return slow_leaf_voxels_axis<DIRECT, 1>(cd, l, tg4, kscale, tp, fl, m, ask);
	return slow_leaf_voxels_axis<DIRECT, 2>(cd, l, tg4, kscale, tp, fl, m, ask);
}
// The reference-order predicate (tri_box_overlap) on the voxels in `ask`, at the chain-rounded voxel centres
// fl(fl(C +- k) +- kh); returns the ones that overlap.  Out of line and self-contained (recomputes the node centre).
template <bool DIRECT>
SVB_HD_NOINLINE uint64_t slow_leaf_exact(const uint64_t ask, const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp) {
	const double k = tg4[3] * kscale, kh = k * 0.5;
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	const double Cx = DIRECT ? centre_axis_direct(path, l, 2, tg4[0], k) : centre_axis_chain(cd, l, 2, tg4[0], tg4[3]);
	const double Cy = DIRECT ? centre_axis_direct(path, l, 1, tg4[1], k) : centre_axis_chain(cd, l, 1, tg4[1], tg4[3]);
	const double Cz = DIRECT ? centre_axis_direct(path, l, 0, tg4[2], k) : centre_axis_chain(cd, l, 0, tg4[2], tg4[3]);
	float tf[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) tf[i] = tp[i];
	uint64_t res = 0, rest = ask;
	while (rest) {
		const int b = SVB_FFSLL(rest) - 1;   // bit 8 c + v
		rest &= rest - 1;
		const int c = b >> 3, v = b & 7;
		const double ccx = SVB_DADD(Cx, (c & 4) ? k : -k), ccy = SVB_DADD(Cy, (c & 2) ? k : -k), ccz = SVB_DADD(Cz, (c & 1) ? k : -k);
		const double vx = SVB_DADD(ccx, (v & 4) ? kh : -kh), vy = SVB_DADD(ccy, (v & 2) ? kh : -kh), vz = SVB_DADD(ccz, (v & 1) ? kh : -kh);
		if (tri_box_overlap(vx, vy, vz, kh, tf)) res |= 1ull << b;
	}
	return res;
}

// Decides the 8 children of the node with Morton code `cd` (tile-local level l, tile geometry tg4 = {cx,cy,cz,rootSide}) against the
// triangle tp[0..8].  fl: the pair's settled-axis flags (in: inherited from the parent pair, out: for the child
// pairs).  nUnsure: children that had to be re-decided by the reference-order predicate.  Returns the hit mask.
// FLATONLY: the caller guarantees that every triangle of the scene is flat (a box mesh); the general edge / plane
// filter is then compiled out, which is what lets the kernel fit 40 registers.
template <bool DIRECT, bool FLATONLY = false, bool LAST = false>
SVB_HD unsigned classify_pair(const uint64_t cd, const int l, const double* __restrict__ tg4, const double kscale, const float* __restrict__ tp,
                              unsigned& fl, unsigned& nUnsure) {
	nUnsure = 0;
	const double k = tg4[3] * kscale;   // child half side rootSide / 2^(l+2) (octree.hpp:115), exact scaling
	const uint64_t path = cd & ((1ull << (3 * l)) - 1);
	if (pair_is_fast(fl)) {   // never taken inside k_classify_filtered: such pairs live in the flat stream (k_classify_fast)
		return classify_pair_flat<DIRECT>(cd, l, tg4, kscale, tp, fl);
	}
	float tf[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) tf[i] = tp[i];
	// ---- static analysis of the triangle, once per triangle/pair lineage (FL_INIT), rigorous (no tolerance):
	// (1) An edge whose endpoints share a coordinate q BITWISE has e_q = +0 in the reference-order predicate for any
	//     box centre (fl(t - c) - fl(t - c)).  Its two cross axes that involve e_q then read  p = fl(+-e_r * v_q),
	//     rad = fl(|e_r| * h)  on the vertex pair {one endpoint, opposite vertex}, which spans the triangle's full
	//     q-extent: since rounding is monotone, "min > rad" / "max < -rad" would imply  min_i v_iq > h / max_i v_iq < -h,
	//     i.e. the box test on q rejects too.  Those axes can never change the conjunction: settled from the start.
	//     (test_triangle_box.cpp:60-102 for the vertex pairs, :137-156 for the operand order.)
	// (2) A triangle flat on q (all three vertices share q): (1) settles six edge axes, the normal is (+-0,..,n_q,..)
	//     and the plane test degenerates to  v_q < -h  or  v_q > h  -- implied by the box test on q as well.
	if (!(fl & FL_INIT)) {
		const bool x01 = tf[0] == tf[3], y01 = tf[1] == tf[4], z01 = tf[2] == tf[5];
		const bool x12 = tf[3] == tf[6], y12 = tf[4] == tf[7], z12 = tf[5] == tf[8];
		const bool x20 = tf[6] == tf[0], y20 = tf[7] == tf[1], z20 = tf[8] == tf[2];
		// edge axes: bit 0,1,2 = X,Y,Z of edge 0; 3,4,5 edge 1; 6,7,8 edge 2.   e_x = 0 -> Y,Z;  e_y = 0 -> X,Z;  e_z = 0 -> X,Y
		if (x01) fl |= 0x006u; if (y01) fl |= 0x005u; if (z01) fl |= 0x003u;
		if (x12) fl |= 0x030u; if (y12) fl |= 0x028u; if (z12) fl |= 0x018u;
		if (x20) fl |= 0x180u; if (y20) fl |= 0x140u; if (z20) fl |= 0x0C0u;
		if (x01 && x12) fl |= 1u << (FL_FLAT + 0);
		if (y01 && y12) fl |= 1u << (FL_FLAT + 1);
		if (z01 && z12) fl |= 1u << (FL_FLAT + 2);
		fl |= FL_INIT;
	}
	const bool planeImplied = (fl & (7u << FL_FLAT)) != 0;
	if (planeImplied) {
		// flat triangle: box axes + in-plane edge axes in one routine with compile-time axes.  (A triangle flat on two
		// axes is a segment: all nine edge axes are settled statically, so it is a flat-stream pair -- handled above.)
		if (fl & (1u << (FL_FLAT + 0))) return classify_flat_slow<DIRECT, 0, LAST>(cd, l, tg4, k, tp, fl, nUnsure);
		if (fl & (1u << (FL_FLAT + 1))) return classify_flat_slow<DIRECT, 1, LAST>(cd, l, tg4, k, tp, fl, nUnsure);
		return classify_flat_slow<DIRECT, 2, LAST>(cd, l, tg4, k, tp, fl, nUnsure);
	}
	if (FLATONLY) return 0;   // unreachable under the caller's guarantee
	// ---- box axes: the reference's own 1-D tests at the chain-rounded child centres -- exact, so a wall lying in a
	//      voxel face (a tie on a box axis) never needs the full predicate
	unsigned alive = box_axes_exact<DIRECT>(cd, l, tg4, k, tp, fl);   // (tp, not tf: dynamic axis index)
	unsigned unsure = 0;
	if (!alive) return 0;
	double Cx, Cy, Cz;
	if (DIRECT) {
		Cx = centre_axis_direct(path, l, 2, tg4[0], k);
		Cy = centre_axis_direct(path, l, 1, tg4[1], k);
		Cz = centre_axis_direct(path, l, 0, tg4[2], k);
	} else {
		double kk;
		const TileGeom tg{tg4[0], tg4[1], tg4[2], tg4[3]};
		node_centre(cd, l, tg, Cx, Cy, Cz, kk);
	}
	const double v0x = (double)tf[0] - Cx, v0y = (double)tf[1] - Cy, v0z = (double)tf[2] - Cz;
	const double v1x = (double)tf[3] - Cx, v1y = (double)tf[4] - Cy, v1z = (double)tf[5] - Cz;
	const double v2x = (double)tf[6] - Cx, v2y = (double)tf[7] - Cy, v2z = (double)tf[8] - Cz;
	const double k2 = k + k;
	double M = fmax(fmax(fmax(fabs(v0x), fabs(v0y)), fmax(fabs(v0z), fabs(v1x))), fmax(fmax(fabs(v1y), fabs(v1z)), fmax(fabs(v2x), fmax(fabs(v2y), fabs(v2z))))) + k2;
	const double eps = 9.094947017729282e-13;   // 2^-40
	const double tol2 = M * (M * eps), tol3 = M * tol2;
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
			// edge 0: X01(v0,v2)  Y02(v0,v2)  Z12(v1,v2)      p_X = ez*vy - ey*vz, p_Y = -ez*vx + ex*vz, p_Z = ey*vx - ex*vy
			if (!(fl & 0x001u)) edge_axis<2, 1>(0x001u, e0z, -e0y, v0y, v0z, v2y, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x002u)) edge_axis<4, 1>(0x002u, -e0z, e0x, v0x, v0z, v2x, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x004u)) edge_axis<4, 2>(0x004u, e0y, -e0x, v1x, v1y, v2x, v2y, k, tol2, alive, unsure, fl);
			// edge 1: X01(v0,v2)  Y02(v0,v2)  Z0(v0,v1)
			if (alive && !(fl & 0x008u)) edge_axis<2, 1>(0x008u, e1z, -e1y, v0y, v0z, v2y, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x010u)) edge_axis<4, 1>(0x010u, -e1z, e1x, v0x, v0z, v2x, v2z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x020u)) edge_axis<4, 2>(0x020u, e1y, -e1x, v0x, v0y, v1x, v1y, k, tol2, alive, unsure, fl);
			// edge 2: X2(v0,v1)  Y1(v0,v1)  Z12(v1,v2)
			if (alive && !(fl & 0x040u)) edge_axis<2, 1>(0x040u, e2z, -e2y, v0y, v0z, v1y, v1z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x080u)) edge_axis<4, 1>(0x080u, -e2z, e2x, v0x, v0z, v1x, v1z, k, tol2, alive, unsure, fl);
			if (alive && !(fl & 0x100u)) edge_axis<4, 2>(0x100u, e2y, -e2x, v1x, v1y, v2x, v2y, k, tol2, alive, unsure, fl);
		}
	}
	unsure &= alive;
	unsigned m = alive & ~unsure;
	if (unsure) {
		m |= exact_children(unsure, Cx, Cy, Cz, k, tp);
		nUnsure = (unsigned)SVB_POPC(unsure);
	}
	return m;
}
#undef SVB_LO
#undef SVB_HI

// ------------------------------------------------------------------ host: is the centre chain exact?
// True when every partial sum of the centre chain (geom_octree.cpp:222-230) of a sub-octree with root centre
// (cx,cy,cz), root side `rootSide` and `Lt` levels is a representable double: all partial sums are integer multiples of
// 2^ge (ge = the lowest set bit among the centre coordinates and the finest half side) bounded by |centre| + rootSide,
// hence representable iff that bound is <= 2^53 * 2^ge.  Then every rounding of the chain is the identity and the
// kernels may evaluate node centres in closed form (centre_axis_direct).
inline int low_bit_exp(double x) {   // exponent of the lowest set bit of a finite double (x = odd * 2^e); 0 -> huge
	if (x == 0.0) return 1 << 20;
	int e;
	double m = std::frexp(std::fabs(x), &e);          // x = m * 2^e, m in [0.5,1)
	uint64_t mi = (uint64_t)std::ldexp(m, 53);        // 53-bit integer mantissa
	return e - 53 + __builtin_ctzll(mi);
}
inline bool centre_chain_exact(const TileGeom& g, int Lt) {
	if (!(g.rootSide > 0.0) || !std::isfinite(g.rootSide) || Lt < 1 || Lt > 20) return false;
	const double kFinest = std::ldexp(g.rootSide, -(Lt + 1));   // half side of the deepest children (level Lt-1 nodes test C +- k)
	if (kFinest < 1e-290) return false;
	int ge = low_bit_exp(kFinest);
	ge = std::min(ge, std::min(low_bit_exp(g.cx), std::min(low_bit_exp(g.cy), low_bit_exp(g.cz))));
	const double span = std::max(std::fabs(g.cx), std::max(std::fabs(g.cy), std::fabs(g.cz))) + g.rootSide;
	if (!std::isfinite(span)) return false;
	return std::ldexp(span, -ge) <= 9007199254740992.0;
}

}  // namespace svb
