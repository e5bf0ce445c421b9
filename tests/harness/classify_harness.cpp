// classify_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the voxelizer's per-pair decision
// (svdag-compression_b200/csrc/svb_classify.cuh, the exact text the CUDA kernel k_classify_filtered runs) with g++
// and drives it over a whole hierarchical build on the CPU, comparing every pair's 8-child hit mask with the
// reference-order predicate (tri_box_overlap = testTriBox, src/symvox/test_triangle_box.cpp:105-184) evaluated at
// chain-computed child centres (geom_octree.cpp:222-230).  Must be built with -ffp-contract=off.
// Nothing in the product links or loads this file.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../svdag-compression_b200/csrc/svb_classify.cuh"

namespace {
struct Pair { uint32_t tri; uint64_t code; uint16_t fl; };
}

extern "C" {

// out[0] pairs tested, out[1] mismatching pairs, out[2] fast-path pairs, out[3] children re-decided exactly,
// out[4..9] first mismatch: level, tri, code, got, want, flags-in
int harness_run(const float* tris, uint64_t T, const double centre[3], double rootSide, int Lt, int direct, int chainExactClaim, uint64_t* out) {
	using namespace svb;
	TileGeom tg{centre[0], centre[1], centre[2], rootSide};
	memset(out, 0, 10 * sizeof(uint64_t));
	(void)chainExactClaim;
	std::vector<Pair> cur, nxt;
	cur.reserve(T);
	for (uint64_t t = 0; t < T; ++t) cur.push_back(Pair{(uint32_t)t, 0ull, 0});
	for (int l = 0; l < Lt; ++l) {
		const double kscale = std::ldexp(1.0, -(l + 2));
		nxt.clear();
		for (const Pair& p : cur) {
			unsigned fl = p.fl, nUnsure = 0;
			const float* tp = tris + 9ull * p.tri;
			unsigned m = direct ? classify_pair<true>(p.code, l, tg, kscale, tp, fl, nUnsure) : classify_pair<false>(p.code, l, tg, kscale, tp, fl, nUnsure);
			if (pair_is_fast(p.fl)) out[2]++;
			out[3] += nUnsure;
			// reference: chain centre of the node, then each child centre, then the predicate
			double cx, cy, cz, k;
			node_centre(p.code, l, tg, cx, cy, cz, k);
			unsigned want = 0;
			for (int c = 0; c < 8; ++c) {
				double ccx = cx + ((c & 4) ? k : -k), ccy = cy + ((c & 2) ? k : -k), ccz = cz + ((c & 1) ? k : -k);
				if (tri_box_overlap(ccx, ccy, ccz, k, tp)) want |= 1u << c;
			}
			out[0]++;
			if (m != want) {
				if (out[1] == 0) { out[4] = (uint64_t)l; out[5] = p.tri; out[6] = p.code; out[7] = m; out[8] = want; out[9] = p.fl; }
				out[1]++;
			}
			// descend along the REFERENCE decision so one wrong pair does not hide its subtree
			for (int c = 0; c < 8; ++c)
				if ((want >> c) & 1) nxt.push_back(Pair{p.tri, (p.code << 3) | (uint64_t)c, (uint16_t)fl});
		}
		cur.swap(nxt);
	}
	return 0;
}

}  // extern "C"
