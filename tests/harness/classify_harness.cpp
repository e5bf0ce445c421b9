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
int harness_run(const float* tris, uint64_t T, const double centre[3], double rootSide, int Lt, int direct, int flatOnly, uint64_t* out) {
	using namespace svb;
	TileGeom tg{centre[0], centre[1], centre[2], rootSide};
	memset(out, 0, 10 * sizeof(uint64_t));
	std::vector<Pair> cur, nxt;
	cur.reserve(T);
	for (uint64_t t = 0; t < T; ++t) cur.push_back(Pair{(uint32_t)t, 0ull, 0});
	for (int l = 0; l < Lt; ++l) {
		const double kscale = std::ldexp(1.0, -(l + 2));
		nxt.clear();
		for (const Pair& p : cur) {
			unsigned fl = p.fl, nUnsure = 0;
			const float* tp = tris + 9ull * p.tri;
			const double tgv[4] = {tg.cx, tg.cy, tg.cz, tg.rootSide};
			const bool lastLevel = (l == Lt - 1);
			unsigned m;
			// every template variant the kernels instantiate: closed-form / chain centres x box-mesh variant x last level
			if (flatOnly) {
				if (direct) m = lastLevel ? classify_pair<true, true, true>(p.code, l, tgv, kscale, tp, fl, nUnsure) : classify_pair<true, true, false>(p.code, l, tgv, kscale, tp, fl, nUnsure);
				else m = lastLevel ? classify_pair<false, true, true>(p.code, l, tgv, kscale, tp, fl, nUnsure) : classify_pair<false, true, false>(p.code, l, tgv, kscale, tp, fl, nUnsure);
			} else {
				if (direct) m = lastLevel ? classify_pair<true, false, true>(p.code, l, tgv, kscale, tp, fl, nUnsure) : classify_pair<true, false, false>(p.code, l, tgv, kscale, tp, fl, nUnsure);
				else m = lastLevel ? classify_pair<false, false, true>(p.code, l, tgv, kscale, tp, fl, nUnsure) : classify_pair<false, false, false>(p.code, l, tgv, kscale, tp, fl, nUnsure);
			}
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
			// second-to-last level, pair that joins the flat stream: the fused leaf kernel (k_flat_leaves) decides the
			// voxels of its children with flat_leaf_masks(); compare with the predicate on every voxel
			if (l == Lt - 2 && pair_is_fast(fl)) {
				const double tg4[4] = {tg.cx, tg.cy, tg.cz, tg.rootSide};
				unsigned lohi[3][2];
				if (direct) flat_leaf_masks<true>(p.code, l, tg4, kscale, tp, fl, lohi); else flat_leaf_masks<false>(p.code, l, tg4, kscale, tp, fl, lohi);
				const uint64_t vox64 = direct ? flat_leaf_voxels<true>(p.code, l, tg4, kscale, tp, fl) : flat_leaf_voxels<false>(p.code, l, tg4, kscale, tp, fl);   // (the kernels' default form)
				for (int c = 0; c < 8; ++c) {
					if (!((want >> c) & 1)) continue;
					unsigned got = lohi[0][(c >> 2) & 1] & lohi[1][(c >> 1) & 1] & lohi[2][c & 1];
					if (((unsigned)(vox64 >> (8 * c)) & 0xFFu) != got) got = 0x100u | ((unsigned)(vox64 >> (8 * c)) & 0xFFu);   // the two forms disagree: report the 64-bit one, marked
					double ccx = cx + ((c & 4) ? k : -k), ccy = cy + ((c & 2) ? k : -k), ccz = cz + ((c & 1) ? k : -k);
					const double kh = k * 0.5;
					unsigned wantv = 0;
					for (int q = 0; q < 8; ++q) {
						double vx = ccx + ((q & 4) ? kh : -kh), vy = ccy + ((q & 2) ? kh : -kh), vz = ccz + ((q & 1) ? kh : -kh);
						if (tri_box_overlap(vx, vy, vz, kh, tp)) wantv |= 1u << q;
					}
					out[0]++;
					if (got != wantv) {
						if (out[1] == 0) { out[4] = (uint64_t)l + 100; out[5] = p.tri; out[6] = (p.code << 3) | (uint64_t)c; out[7] = got; out[8] = wantv; out[9] = fl; }
						out[1]++;
					}
				}
			}
			// second-to-last level, flat triangle that stays in the slow stream: the fused kernel k_slow_leaves decides the
			// voxels of its children with slow_leaf_voxels() (box axes exactly, in-plane edge axes through the filter one
			// level further down); compare with the predicate on every voxel.  (Pairs that join the flat stream run
			// through it as well -- the kernel takes every slow-stream parent.)
			if (l == Lt - 2 && (fl & (7u << FL_FLAT)) && want) {
				const double tg4[4] = {tg.cx, tg.cy, tg.cz, tg.rootSide};
				uint64_t ask = 0;
				uint64_t vox = direct ? slow_leaf_voxels<true>(p.code, l, tg4, kscale, tp, fl, want, ask) : slow_leaf_voxels<false>(p.code, l, tg4, kscale, tp, fl, want, ask);
				if (ask) vox |= direct ? slow_leaf_exact<true>(ask, p.code, l, tg4, kscale, tp) : slow_leaf_exact<false>(ask, p.code, l, tg4, kscale, tp);
				out[3] += (uint64_t)__builtin_popcountll(ask);
				for (int c = 0; c < 8; ++c) {
					if (!((want >> c) & 1)) continue;
					const unsigned got = (unsigned)(vox >> (8 * c)) & 0xFFu;
					double ccx = cx + ((c & 4) ? k : -k), ccy = cy + ((c & 2) ? k : -k), ccz = cz + ((c & 1) ? k : -k);
					const double kh = k * 0.5;
					unsigned wantv = 0;
					for (int q = 0; q < 8; ++q) {
						double vx = ccx + ((q & 4) ? kh : -kh), vy = ccy + ((q & 2) ? kh : -kh), vz = ccz + ((q & 1) ? kh : -kh);
						if (tri_box_overlap(vx, vy, vz, kh, tp)) wantv |= 1u << q;
					}
					out[0]++;
					if (got != wantv) {
						if (out[1] == 0) { out[4] = (uint64_t)l + 200; out[5] = p.tri; out[6] = (p.code << 3) | (uint64_t)c; out[7] = got; out[8] = wantv; out[9] = fl; }
						out[1]++;
					}
				}
			}
			// descend along the REFERENCE decision so one wrong pair does not hide its subtree
			for (int c = 0; c < 8; ++c)
				if ((want >> c) & 1) nxt.push_back(Pair{p.tri, (p.code << 3) | (uint64_t)c, (uint16_t)fl});
		}
		cur.swap(nxt);
	}
	return 0;
}

// host-side decision whether the kernels may use closed-form node centres (svb_classify.cuh::centre_chain_exact)
int harness_chain_exact(const double centre[3], double rootSide, int Lt) {
	svb::TileGeom tg{centre[0], centre[1], centre[2], rootSide};
	return svb::centre_chain_exact(tg, Lt) ? 1 : 0;
}

}  // extern "C"
