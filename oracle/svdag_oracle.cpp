// =====================================================================================
//  svdag_oracle.cpp -- CPU restatement of the reference's svbuilder hot path.
//
//  TEST INFRASTRUCTURE ONLY.  Nothing in the product (svdag-compression_b200/, include/)
//  may include, link or call this file; only tests/, __graft_entry__.smoke() and
//  bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or
//  the CPU baseline.
//
//  It is a sequential, deliberately plain restatement (std::map, explicit stacks) of the
//  algorithm in /root/reference/src/symvox (paths below are relative to that directory),
//  written from scratch; every function cites the lines it follows.  Arithmetic of the
//  un-vendored SpaceLand library follows oracle/sl_shim (dot left-to-right, textbook
//  cross, center=(lo+hi)*0.5, half=(hi-lo)*0.5) -- parity w.r.t. the real SL is UNPINNED.
//  Pinned against: oracle/_ref/svbuilder_ref (the unmodified reference compiled against
//  the shim) through tests/golden/* (tests/test_oracle_golden.py, tests/golden/check_oracle_midsize.py).
//
//  Built with -ffp-contract=off: the reference is compiled for baseline x86-64, i.e.
//  without FMA contraction (CMakeLists.txt has no -march).
//
//  C++ (not C) only because the SSVDAG encoder's node order depends on libstdc++'s
//  unstable std::sort (encoded_ssvdag.cpp:273-275); everything else is C-style.
// =====================================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace {

typedef uint32_t id_t32;
const id_t32 NULLNODE = 0xFFFFFFFEu;  // octree.cpp:22

struct Node {  // octree.hpp:57-85 + geom_octree.hpp:60-98
	id_t32 ch[8];
	uint8_t mask;
	uint8_t mir[3];
	uint8_t inv;
	uint32_t chLev[8];
	Node() {
		for (int i = 0; i < 8; ++i) { ch[i] = NULLNODE; chLev[i] = 0xFFFFFFFFu; }
		mask = 0; mir[0] = mir[1] = mir[2] = 0; inv = 0;
	}
	bool hasBit(int c) const { return (mask >> c) & 1; }
	bool hasPtr(int c) const { return ch[c] != NULLNODE; }
};

typedef std::array<uint32_t, 12> Key;  // octree_node.cpp:39-86: (mask, ch[0..7], mir[0..2])
Key keyOf(const Node& n) {
	Key k;
	k[0] = n.mask;
	for (int i = 0; i < 8; ++i) k[1 + i] = n.ch[i];
	for (int i = 0; i < 3; ++i) k[9 + i] = n.mir[i];
	return k;
}

int popc8(unsigned m) { return __builtin_popcount(m & 0xFF); }

uint8_t permBits(uint8_t m, int x) {  // slot i of the result takes slot i^x of the source
	uint8_t r = 0;
	for (int i = 0; i < 8; ++i) if ((m >> (i ^ x)) & 1) r |= uint8_t(1u << i);
	return r;
}

// octree_node.cpp:88-195.  Child index bits: X=4, Y=2, Z=1 (octree.hpp:33-42).
Node mirrorNode(const Node& n, bool x, bool y, bool z, bool applyToChildren = true) {
	int s = (x ? 4 : 0) | (y ? 2 : 0) | (z ? 1 : 0);
	Node r = n;
	r.mask = permBits(n.mask, s);
	for (int a = 0; a < 3; ++a) r.mir[a] = permBits(n.mir[a], s);
	for (int i = 0; i < 8; ++i) { r.ch[i] = n.ch[i ^ s]; r.chLev[i] = n.chLev[i]; }
	if (applyToChildren)
		for (int i = 0; i < 8; ++i)
			if (r.hasPtr(i)) {
				if (x) r.mir[0] ^= uint8_t(1u << i);
				if (y) r.mir[1] ^= uint8_t(1u << i);
				if (z) r.mir[2] ^= uint8_t(1u << i);
			}
	return r;
}

// ---------------------------------------------------------------- SAT triangle/box test
// test_triangle_box.cpp:38-56
bool planeBoxOverlap(const double nrm[3], const double vert[3], double maxbox) {
	double vmin[3], vmax[3];
	for (int q = 0; q < 3; ++q) {
		double v = vert[q];
		if (nrm[q] > 0.0f) { vmin[q] = -maxbox - v; vmax[q] = maxbox - v; }
		else { vmin[q] = maxbox - v; vmax[q] = -maxbox - v; }
	}
	double d0 = nrm[0] * vmin[0]; d0 = d0 + nrm[1] * vmin[1]; d0 = d0 + nrm[2] * vmin[2];
	if (d0 > 0.0) return false;
	double d1 = nrm[0] * vmax[0]; d1 = d1 + nrm[1] * vmax[1]; d1 = d1 + nrm[2] * vmax[2];
	if (d1 >= 0.0) return true;
	return false;
}

inline bool axisReject(double pa, double pb, double rad) {
	double mn, mx;
	if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
	return (mn > rad || mx < -rad);
}

// test_triangle_box.cpp:105-184 (9 edge axes, 3 box axes, plane), same expression order.
bool testTriBox(const double c[3], double h, const float* tri) {
	double v0[3], v1[3], v2[3], e0[3], e1[3], e2[3];
	for (int k = 0; k < 3; ++k) {
		v0[k] = double(tri[k]) - c[k];
		v1[k] = double(tri[3 + k]) - c[k];
		v2[k] = double(tri[6 + k]) - c[k];
	}
	for (int k = 0; k < 3; ++k) { e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v1[k]; e2[k] = v0[k] - v2[k]; }
	double fex, fey, fez;
	// edge 0: X01, Y02, Z12
	fex = std::fabs(e0[0]); fey = std::fabs(e0[1]); fez = std::fabs(e0[2]);
	if (axisReject(e0[2] * v0[1] - e0[1] * v0[2], e0[2] * v2[1] - e0[1] * v2[2], (fez + fey) * h)) return false;
	if (axisReject(-e0[2] * v0[0] + e0[0] * v0[2], -e0[2] * v2[0] + e0[0] * v2[2], (fez + fex) * h)) return false;
	if (axisReject(e0[1] * v1[0] - e0[0] * v1[1], e0[1] * v2[0] - e0[0] * v2[1], (fey + fex) * h)) return false;
	// edge 1: X01, Y02, Z0
	fex = std::fabs(e1[0]); fey = std::fabs(e1[1]); fez = std::fabs(e1[2]);
	if (axisReject(e1[2] * v0[1] - e1[1] * v0[2], e1[2] * v2[1] - e1[1] * v2[2], (fez + fey) * h)) return false;
	if (axisReject(-e1[2] * v0[0] + e1[0] * v0[2], -e1[2] * v2[0] + e1[0] * v2[2], (fez + fex) * h)) return false;
	if (axisReject(e1[1] * v0[0] - e1[0] * v0[1], e1[1] * v1[0] - e1[0] * v1[1], (fey + fex) * h)) return false;
	// edge 2: X2, Y1, Z12
	fex = std::fabs(e2[0]); fey = std::fabs(e2[1]); fez = std::fabs(e2[2]);
	if (axisReject(e2[2] * v0[1] - e2[1] * v0[2], e2[2] * v1[1] - e2[1] * v1[2], (fez + fey) * h)) return false;
	if (axisReject(-e2[2] * v0[0] + e2[0] * v0[2], -e2[2] * v1[0] + e2[0] * v1[2], (fez + fex) * h)) return false;
	if (axisReject(e2[1] * v1[0] - e2[0] * v1[1], e2[1] * v2[0] - e2[0] * v2[1], (fey + fex) * h)) return false;
	// box axes
	for (int k = 0; k < 3; ++k) {
		double mn = v0[k], mx = v0[k];
		if (v1[k] < mn) mn = v1[k];
		if (v1[k] > mx) mx = v1[k];
		if (v2[k] < mn) mn = v2[k];
		if (v2[k] > mx) mx = v2[k];
		if (mn > h || mx < -h) return false;
	}
	double nrm[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
	return planeBoxOverlap(nrm, v0, h);
}

// ---------------------------------------------------------------- octree container
enum State { S_EMPTY = 0, S_SVO = 1, S_DAG = 2, S_SDAG = 3 };  // geom_octree.hpp:35-40

struct Stats {
	uint64_t nTotalVoxels = 0, nNodesSVO = 0, nNodesDAG = 0, nNodesSDAG = 0;
	uint64_t nNodesLastLevSVO = 0, nNodesLastLevDAG = 0, nCrossLevelMerged = 0;
};

struct Oct {
	const float* tris = nullptr;
	uint64_t ntris = 0;
	std::vector<std::vector<Node>> data;
	float bboxF[6] = {0, 0, 0, 0, 0, 0};
	double rootSide = 0;
	unsigned levels = 0;
	uint64_t nVoxels = 0, nNodes = 0;
	int state = S_EMPTY;
	Stats stats;
	std::vector<uint64_t> levelSizesBeforeDag;  // "Reduced level l from A ..." (geom_octree.cpp:509)

	double halfSideD(unsigned level) const { return double(rootSide) / double(1u << (level + 1)); }  // octree.hpp:115

	// geom_octree.cpp:171-280
	// triMat (optional): putMaterialIdInLeaves = true (geom_octree.cpp:210-211, :252): the leaf node's child slot of a voxel holds
	// the material id of the triangle being voxelized, i.e. finally that of the LAST triangle (file order) touching the voxel
	void buildSVO(unsigned lv, const double bmin[3], const double bmax[3], std::vector<std::array<double, 3>>* leavesCenters, const uint32_t* triMat = nullptr) {
		for (int k = 0; k < 3; ++k) { bboxF[k] = float(bmin[k]); bboxF[3 + k] = float(bmax[k]); }
		levels = lv;
		float sides[3];
		for (int k = 0; k < 3; ++k) sides[k] = ((bboxF[3 + k] - bboxF[k]) * 0.5f) * 2.0f;
		rootSide = std::max(std::max(sides[0], sides[1]), sides[2]);
		data.assign(levels, std::vector<Node>());
		data[0].push_back(Node());
		struct Item { id_t32 id; uint8_t level; double c[3]; };
		std::vector<Item> stack;
		double rc[3];
		for (int k = 0; k < 3; ++k) rc[k] = (bmin[k] + bmax[k]) * 0.5;
		for (uint64_t t = 0; t < ntris; ++t) {
			const float* tri = tris + 9 * t;
			Item root; root.id = 0; root.level = 0; root.c[0] = rc[0]; root.c[1] = rc[1]; root.c[2] = rc[2];
			stack.push_back(root);
			while (!stack.empty()) {
				Item qi = stack.back(); stack.pop_back();
				double k = halfSideD(qi.level + 1);
				double cc[8][3];
				for (int i = 0; i < 8; ++i) {
					cc[i][0] = qi.c[0] + ((i & 4) ? +k : -k);
					cc[i][1] = qi.c[1] + ((i & 2) ? +k : -k);
					cc[i][2] = qi.c[2] + ((i & 1) ? +k : -k);
				}
				for (int i = 7; i >= 0; --i) {
					if (!testTriBox(cc[i], k, tri)) continue;
					Node& node = data[qi.level][qi.id];  // re-fetched: vectors may have grown
					node.mask |= uint8_t(1u << i);
					if (!node.hasPtr(i) && (qi.level < levels - 1)) {
						id_t32 nid = id_t32(data[qi.level + 1].size());
						data[qi.level][qi.id].ch[i] = nid;
						data[qi.level + 1].emplace_back();
						nNodes++;
						if (leavesCenters && qi.level == levels - 2) leavesCenters->push_back({cc[i][0], cc[i][1], cc[i][2]});
					}
					if (qi.level + 1u < levels) {
						Item it; it.id = data[qi.level][qi.id].ch[i]; it.level = uint8_t(qi.level + 1);
						it.c[0] = cc[i][0]; it.c[1] = cc[i][1]; it.c[2] = cc[i][2];
						stack.push_back(it);
					} else if (triMat) data[qi.level][qi.id].ch[i] = triMat[t];   // :252
				}
			}
		}
		cleanEmptyNodes();
		for (size_t i = 0; i < data[levels - 1].size(); ++i) nVoxels += popc8(data[levels - 1][i].mask);
		state = S_SVO;
		stats.nNodesSVO = nNodes;
		stats.nNodesLastLevSVO = data[levels - 1].size();
		stats.nTotalVoxels = nVoxels;
	}

	// geom_octree.cpp:437-456
	void cleanEmptyNodes() {
		for (int lev = int(levels) - 1; lev > 0; --lev) {
			std::set<id_t32> empties;
			for (id_t32 i = 0; i < data[lev].size(); ++i) if (data[lev][i].mask == 0) empties.insert(i);
			if (empties.empty()) continue;
			for (auto& p : data[lev - 1])
				for (int j = 0; j < 8; ++j)
					if (empties.count(p.ch[j])) { p.mask &= uint8_t(~(1u << j)); p.ch[j] = NULLNODE; }
		}
	}

	// geom_octree.cpp:462-548
	void toDAG(bool record) {
		if (!(state == S_SVO || state == S_DAG)) return;
		nNodes = 1;
		if (record) levelSizesBeforeDag.assign(levels, 0);
		for (unsigned lev = levels - 1; lev > 0; --lev) {
			size_t old = data[lev].size();
			std::vector<id_t32> corr(old, 0);
			std::map<Key, id_t32> seen;
			std::vector<Node> uniq;
			for (id_t32 i = 0; i < old; ++i) {
				const Node& n = data[lev][i];
				if (n.mask == 0) continue;
				Key k = keyOf(n);
				auto it = seen.find(k);
				if (it != seen.end()) corr[i] = it->second;
				else { seen[k] = id_t32(uniq.size()); corr[i] = id_t32(uniq.size()); uniq.push_back(n); }
			}
			if (record) levelSizesBeforeDag[lev] = old;
			data[lev].swap(uniq);
			nNodes += data[lev].size();
			for (auto& p : data[lev - 1])
				for (int j = 0; j < 8; ++j)
					if (p.hasBit(j)) p.ch[j] = corr[p.ch[j]];
		}
		state = S_DAG;
		stats.nNodesDAG = nNodes;
		stats.nNodesLastLevDAG = data[levels - 1].size();
	}

	// geom_octree.cpp:289-435
	void buildDAG(unsigned lv, unsigned stepLevel, const double bmin[3], const double bmax[3]) {
		std::vector<std::array<double, 3>> leavesCenters;
		unsigned stepLevels = stepLevel + 1;
		buildSVO(stepLevels, bmin, bmax, &leavesCenters);
		double lhs = halfSideD(stepLevels - 1);
		stats.nNodesSVO = nNodes;
		stats.nNodesLastLevSVO = 0;
		std::map<unsigned, Oct*> subs;
		const std::vector<Node>& base = data[levels - 1];
		for (size_t i = 0; i < base.size(); ++i) {
			for (int j = 7; j >= 0; --j) {
				if (!base[i].hasBit(j)) continue;
				Oct* sub = new Oct();
				sub->tris = tris; sub->ntris = ntris;
				double p1[3] = { leavesCenters[i][0], leavesCenters[i][1], leavesCenters[i][2] };
				double p2[3] = { p1[0] + ((j & 4) ? +lhs : -lhs), p1[1] + ((j & 2) ? +lhs : -lhs), p1[2] + ((j & 1) ? +lhs : -lhs) };
				double lo[3], hi[3];
				for (int k = 0; k < 3; ++k) { lo[k] = std::min(p1[k], p2[k]); hi[k] = std::max(p1[k], p2[k]); }
				sub->buildSVO(lv - stepLevels, lo, hi, nullptr);
				stats.nNodesSVO += sub->nNodes;
				stats.nNodesLastLevSVO += sub->data.back().size();
				nVoxels += sub->nVoxels - 1;  // "root doesn't count" (:352-354)
				sub->toDAG(false);
				subs[unsigned(i * 8 + j)] = sub;
			}
		}
		stats.nTotalVoxels = nVoxels;
		data.resize(lv);
		levels = lv;
		for (size_t i = 0; i < data[stepLevels - 1].size(); ++i) {
			for (int j = 7; j >= 0; --j) {
				if (!data[stepLevels - 1][i].hasBit(j)) continue;
				Oct* sub = subs[unsigned(i * 8 + j)];
				data[stepLevels - 1][i].ch[j] = id_t32(data[stepLevels].size());
				for (unsigned k = stepLevels; k < lv; ++k) {
					std::vector<Node>& src = sub->data[k - stepLevels];
					if (k < lv - 1) {
						id_t32 off = id_t32(data[k + 1].size());
						for (auto& n : src) for (int c = 0; c < 8; ++c) if (n.hasPtr(c)) n.ch[c] += off;
					}
					data[k].insert(data[k].end(), src.begin(), src.end());
				}
				delete sub;
			}
		}
		toDAG(true);
		stats.nNodesLastLevDAG = data[levels - 1].size();
	}

	// geom_octree.cpp:822-831
	void invertInvs(Node& n, unsigned lev, bool ix, bool iy, bool iz) const {
		for (int i = 0; i < 8; ++i) {
			if (!n.hasPtr(i)) continue;
			const Node& c = data[lev + 1][n.ch[i]];
			if (ix && (c.inv & 1)) n.mir[0] &= uint8_t(~(1u << i));
			if (iy && (c.inv & 2)) n.mir[1] &= uint8_t(~(1u << i));
			if (iz && (c.inv & 4)) n.mir[2] &= uint8_t(~(1u << i));
		}
	}

	// geom_octree.cpp:551-697 (skipSymmetry == false path, as the CLI calls it)
	void toSDAG() {
		if (!(state == S_SVO || state == S_DAG)) return;
		struct MN { bool x, y, z; id_t32 id; };
		nNodes = 0;
		static const bool AX[8][3] = { {0,0,0}, {1,0,0}, {0,1,0}, {0,0,1}, {1,1,0}, {1,0,1}, {0,1,1}, {1,1,1} };
		for (unsigned lev = levels - 1; lev > 0; --lev) {
			size_t old = data[lev].size();
			std::vector<MN> corr(old, MN{false, false, false, 0});
			std::map<Key, id_t32> seen;
			std::vector<Node> uniq;
			for (id_t32 i = 0; i < old; ++i) {
				Node n = data[lev][i];
				if (n.mask == 0) continue;
				if (keyOf(n) == keyOf(mirrorNode(n, true, false, false))) n.inv |= 1;
				if (keyOf(n) == keyOf(mirrorNode(n, false, true, false))) n.inv |= 2;
				if (keyOf(n) == keyOf(mirrorNode(n, false, false, true))) n.inv |= 4;
				bool found = false;
				for (int v = 0; v < 8 && !found; ++v) {
					Node m = n;
					if (v) {
						// composed one axis at a time exactly like :609-615 (each step toggles child bits)
						if (AX[v][0]) m = mirrorNode(m, true, false, false);
						if (AX[v][1]) m = mirrorNode(m, false, true, false);
						if (AX[v][2]) m = mirrorNode(m, false, false, true);
						if (lev < levels - 1) invertInvs(m, lev, AX[v][0], AX[v][1], AX[v][2]);
					}
					auto it = seen.find(keyOf(m));
					if (it != seen.end()) { corr[i] = MN{AX[v][0], AX[v][1], AX[v][2], it->second}; found = true; }
				}
				if (!found) {
					seen[keyOf(n)] = id_t32(uniq.size());
					corr[i] = MN{false, false, false, id_t32(uniq.size())};
					uniq.push_back(n);
				}
			}
			data[lev].swap(uniq);
			nNodes += data[lev].size();
			for (auto& p : data[lev - 1])
				for (int j = 0; j < 8; ++j)
					if (p.hasPtr(j)) {
						MN mn = corr[p.ch[j]];
						p.ch[j] = mn.id;
						if (mn.x) p.mir[0] |= uint8_t(1u << j);
						if (mn.y) p.mir[1] |= uint8_t(1u << j);
						if (mn.z) p.mir[2] |= uint8_t(1u << j);
					}
		}
		stats.nNodesSDAG = nNodes;
		state = S_SDAG;
	}

	// geom_octree_extension.cpp:21-30
	void initChildLevels() {
		for (unsigned lev = 0; lev < levels; ++lev)
			for (auto& n : data[lev]) for (int c = 0; c < 8; ++c) n.chLev[c] = lev + 1;
	}

	// geom_octree_extension.cpp:238-290
	bool compareSubtrees(unsigned levA, unsigned levB, const Node& nA, const Node& nB,
	                     std::vector<std::map<id_t32, std::pair<unsigned, id_t32>>>& inSub) {
		if (nA.mask != 0 && nB.mask == 0) return false;
		else if (levA == levels - 1) return nA.mask == nB.mask;
		for (int i = 0; i < 8; ++i) {
			if (nA.hasBit(i) != nB.hasBit(i)) return false;
			else if (!nA.hasBit(i) && !nB.hasBit(i)) continue;
			const Node& cA = data[levA + 1][nA.ch[i]];
			const Node& cB = data[levB + 1][nB.ch[i]];
			inSub[levA + 1][nA.ch[i]] = std::make_pair(levB + 1, nB.ch[i]);
			if (!compareSubtrees(levA + 1, levB + 1, cA, cB, inSub)) return false;
		}
		return true;
	}

	// Exact canonical id of the subtree of (lev,idx) truncated to `depth` more levels; used only to
	// enumerate match candidates.  The reference pre-filters candidates with a 64-bit hash
	// (ext.cpp:34-74,105-142) and then runs compareSubtrees on each candidate in (level asc, index
	// asc) order (:1338-1346); false positives of its hash are rejected by compareSubtrees, so
	// enumerating exactly-equal candidates in the same order visits the same first match.
	// NOTE the candidate maps hold the subtree truncated at depth `currentMatchDepth`, and a
	// candidate B only reaches compareSubtrees if that hash matches; since A's subtree has exactly
	// maxMatchDepth = levels-levA-1 levels below it and currentMatchDepth <= maxMatchDepth, an
	// exactly-equal pair always shares the hash, so no true match is filtered out.
	unsigned mergeAcrossAllLevels() {  // geom_octree_extension.cpp:1192-1544
		typedef std::map<id_t32, std::pair<unsigned, id_t32>> CorrMap;
		std::vector<CorrMap> multi(levels), inSub(levels);
		// truncated-subtree ids: tid[d][lev][idx] for depth d
		std::vector<std::vector<std::vector<uint32_t>>> tid;
		{
			unsigned maxD = levels - 1;
			tid.resize(maxD + 1);
			for (unsigned d = 0; d <= maxD; ++d) {
				tid[d].resize(levels);
				std::map<std::array<uint32_t, 9>, uint32_t> intern;
				for (unsigned lev = 0; lev + d < levels; ++lev) {
					tid[d][lev].resize(data[lev].size());
					for (size_t i = 0; i < data[lev].size(); ++i) {
						const Node& n = data[lev][i];
						std::array<uint32_t, 9> k;
						k[0] = n.mask;
						for (int c = 0; c < 8; ++c)
							k[1 + c] = (d > 0 && n.hasBit(c)) ? tid[d - 1][lev + 1][n.ch[c]] : 0xFFFFFFFFu;
						auto it = intern.find(k);
						if (it == intern.end()) it = intern.insert(std::make_pair(k, uint32_t(intern.size()))).first;
						tid[d][lev][i] = it->second;
					}
				}
			}
		}
		std::set<id_t32> cur, next;
		unsigned levA = 1;
		for (id_t32 i = 0; i < data[levA].size(); ++i) cur.insert(i);
		size_t prevNNodes = nNodes;
		nNodes = 1;
		for (; levA < levels; ++levA) {
			unsigned D = levels - levA - 1;
			// candidates per truncated id at depth D: (level asc, index asc), levels 1..levA-1 (levStart=1, geom_octree.hpp:181)
			std::map<uint32_t, std::vector<std::pair<unsigned, id_t32>>> cand;
			for (unsigned levB = 1; levB < levA; ++levB)
				for (id_t32 j = 0; j < data[levB].size(); ++j) cand[tid[D][levB][j]].push_back(std::make_pair(levB, j));
			for (id_t32 idA : cur) {
				Node& nA = data[levA][idA];
				bool found = false;
				auto it = cand.find(tid[D][levA][idA]);
				if (it != cand.end()) {
					for (auto& lb : it->second) {
						for (auto& m : inSub) m.clear();
						if (compareSubtrees(levA, lb.first, nA, data[lb.first][lb.second], inSub)) {
							found = true;
							multi[levA][idA] = lb;
							for (unsigned l = 0; l < levels; ++l) multi[l].insert(inSub[l].begin(), inSub[l].end());
							break;
						}
					}
				}
				if (!found) for (int i = 0; i < 8; ++i) if (nA.hasBit(i)) next.insert(nA.ch[i]);
			}
			cur.clear();
			cur.swap(next);
		}
		for (unsigned lev = levels - 1; lev > 0; --lev) {
			std::vector<Node> uniq;
			std::vector<id_t32> corr(data[lev].size(), 0xFFFFFFFFu);
			for (id_t32 i = 0; i < data[lev].size(); ++i)
				if (multi[lev].count(i) == 0) { corr[i] = id_t32(uniq.size()); uniq.push_back(data[lev][i]); }
			data[lev].swap(uniq);
			nNodes += data[lev].size();
			for (auto& p : data[lev - 1])
				for (int j = 0; j < 8; ++j)
					if (p.hasBit(j)) {
						auto it = multi[lev].find(p.ch[j]);
						if (it != multi[lev].end()) { p.chLev[j] = it->second.first; p.ch[j] = it->second.second; }
						else p.ch[j] = corr[p.ch[j]];
					}
			for (unsigned low = levels - 2; low >= lev; --low)
				for (auto& p : data[low])
					for (int j = 0; j < 8; ++j)
						if (p.hasBit(j) && p.chLev[j] == lev) p.ch[j] = corr[p.ch[j]];
		}
		stats.nNodesDAG = nNodes;
		unsigned total = 0;
		for (unsigned i = 0; i < levels; ++i) total += unsigned(multi[i].size());
		stats.nCrossLevelMerged = total;
		(void)prevNNodes;
		return total;
	}
};

// ---------------------------------------------------------------- encoders
void put(std::vector<uint8_t>& o, const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; o.insert(o.end(), b, b + n); }

void putHeader(std::vector<uint8_t>& o, const Oct& t) {  // encoded_svdag.cpp:88-92
	put(o, t.bboxF, 24);
	float rs = float(t.rootSide);
	put(o, &rs, 4);
	uint32_t lv = t.levels; put(o, &lv, 4);
	uint32_t nn = uint32_t(t.nNodes); put(o, &nn, 4);
}

// encoded_svdag.cpp:105-199 (+save :76-103); ussvdag=true -> encoded_ussvdag.cpp:86-170 (+save :60-84)
bool encodeSVDAG(const Oct& t, bool ussvdag, std::vector<uint8_t>& out) {
	if (!ussvdag && t.state != S_DAG) return false;
	if (ussvdag && t.state != S_SDAG) return false;
	std::vector<uint32_t> truePtrs, words;
	uint32_t counter = 0;
	for (unsigned lev = 0; lev < t.levels; ++lev)
		for (auto& n : t.data[lev]) { truePtrs.push_back(counter); counter += (lev < t.levels - 1) ? popc8(n.mask) + 1 : 1; }
	uint32_t firstLeafPtr = counter;
	std::vector<uint32_t> acc(t.levels);
	uint32_t a = 0;
	for (unsigned lev = 0; lev < t.levels; ++lev) { a += uint32_t(t.data[lev].size()); acc[lev] = a; }
	for (unsigned lev = 0; lev < t.levels; ++lev) {
		for (auto& n : t.data[lev]) {
			uint32_t head = n.mask;
			if (ussvdag) head |= (uint32_t(n.mir[2]) << 24) | (uint32_t(n.mir[1]) << 16) | (uint32_t(n.mir[0]) << 8);
			words.push_back(head);
			if (words.size() < firstLeafPtr)
				for (int k = 7; k >= 0; --k)
					if (n.ch[k] != NULLNODE) {
						size_t off = ussvdag ? acc[lev] : (n.chLev[k] == 0 ? 0 : acc[n.chLev[k] - 1]);
						words.push_back(truePtrs[n.ch[k] + off]);
					}
		}
	}
	putHeader(out, t);
	put(out, &firstLeafPtr, 4);
	uint32_t cnt = uint32_t(words.size()); put(out, &cnt, 4);
	put(out, words.data(), words.size() * 4);
	return true;
}

// encoded_ssvdag.cpp:194-466 (+save :84-117)
bool encodeSSVDAG(const Oct& t, std::vector<uint8_t>& out) {
	if (t.state != S_SDAG && t.state != S_DAG) return false;
	if (t.levels < 3) return false;
	const unsigned L = t.levels;
	std::vector<std::vector<uint16_t>> enc(L - 2);
	std::vector<uint8_t> leaves;
	std::vector<std::pair<id_t32, id_t32>> hist;
	std::vector<id_t32> indirection;
	for (int lev = int(L) - 2; lev >= 0; --lev) {
		const std::vector<Node>& cur = t.data[lev];
		hist.resize(cur.size());
		for (id_t32 i = 0; i < cur.size(); ++i) { hist[i].first = i; hist[i].second = 0; }
		if (lev > 0) {
			for (const Node& n : t.data[lev - 1]) for (int c = 0; c < 8; ++c) if (n.hasPtr(c)) hist[n.ch[c]].second++;
			std::sort(hist.begin(), hist.end(),
			          [](std::pair<id_t32, id_t32> a, std::pair<id_t32, id_t32> b) { return a.second > b.second; });
		}
		std::vector<id_t32> newInd(cur.size());
		if (lev == int(L) - 2) {
			leaves.resize(hist.size() * 8);
			for (id_t32 i = 0; i < hist.size(); ++i) {
				newInd[hist[i].first] = i;
				const Node& n = cur[hist[i].first];
				uint8_t buf[8], nb[8];
				for (int c = 7; c >= 0; --c) {
					if (n.hasPtr(c)) {
						Node cm = mirrorNode(t.data[lev + 1][n.ch[c]], (n.mir[0] >> c) & 1, (n.mir[1] >> c) & 1, (n.mir[2] >> c) & 1);
						buf[c] = cm.mask;
					} else buf[c] = 0;
				}
				memset(nb, 0, 8);
				unsigned off = 0;
				for (unsigned z = 0; z < 4; ++z) for (unsigned y = 0; y < 4; ++y) for (unsigned x = 0; x < 4; ++x) {
					unsigned byteId = x / 2 + (y / 2) * 2 + (z / 2) * 4;  // :119-135
					unsigned bitId = x % 2 + (y % 2) * 2 + (z % 2) * 4;
					unsigned vox = (buf[byteId] >> bitId) & 1;
					nb[off / 8] |= uint8_t(vox << (off % 8));
					++off;
				}
				for (int c = 0; c < 8; ++c) leaves[size_t(i) * 8 + c] = nb[c];
			}
		} else {
			for (id_t32 i = 0; i < hist.size(); ++i) {
				const Node& n = cur[hist[i].first];
				uint16_t tmp[20];
				tmp[0] = 0;
				int sz = 1;
				for (int c = 7; c >= 0; --c) {
					if (!n.hasBit(c)) continue;
					id_t32 addr = indirection[n.ch[c]];
					if (addr < (1u << 13)) {
						uint16_t p = uint16_t(addr);
						tmp[0] |= uint16_t(1u << (2 * c));
						if ((n.mir[0] >> c) & 1) p |= uint16_t(1u << 13);
						if ((n.mir[1] >> c) & 1) p |= uint16_t(1u << 14);
						if ((n.mir[2] >> c) & 1) p |= uint16_t(1u << 15);
						tmp[sz++] = p;
					} else if (addr < (1u << 30)) {
						uint32_t p = addr;
						if (p & (1u << 29)) { tmp[0] |= uint16_t(3u << (2 * c)); p &= ~(1u << 29); }
						else tmp[0] |= uint16_t(2u << (2 * c));
						if ((n.mir[0] >> c) & 1) p |= 1u << 29;
						if ((n.mir[1] >> c) & 1) p |= 1u << 30;
						if ((n.mir[2] >> c) & 1) p |= 1u << 31;
						tmp[sz++] = uint16_t(p >> 16);
						tmp[sz++] = uint16_t(p & 0xFFFF);
					}
				}
				newInd[hist[i].first] = uint32_t(enc[lev].size());
				enc[lev].insert(enc[lev].end(), tmp, tmp + sz);
			}
		}
		indirection.swap(newInd);
	}
	std::vector<uint32_t> offs(L - 2);
	offs[0] = 0;
	for (size_t i = 1; i < offs.size(); ++i) offs[i] = uint32_t(enc[i - 1].size()) + offs[i - 1];
	std::vector<uint16_t> inner;
	for (auto& e : enc) inner.insert(inner.end(), e.begin(), e.end());
	putHeader(out, t);
	uint32_t cnt = uint32_t(inner.size()); put(out, &cnt, 4); put(out, inner.data(), inner.size() * 2);
	cnt = uint32_t(leaves.size()); put(out, &cnt, 4); put(out, leaves.data(), leaves.size());
	cnt = uint32_t(offs.size()); put(out, &cnt, 4); put(out, offs.data(), offs.size() * 4);
	return true;
}

}  // namespace

// ------------------------------------------------------------------------- C API (ctypes)
extern "C" {

void* orc_create(const float* tris9, uint64_t ntris) {
	Oct* o = new Oct();
	o->tris = tris9;  // borrowed
	o->ntris = ntris;
	return o;
}
void orc_destroy(void* h) { delete (Oct*)h; }

// svbuilder/main.cpp:158-175: step 0 -> buildSVO + toDAG, else buildDAG
int orc_build(void* h, unsigned levels, unsigned step, const double bmin[3], const double bmax[3]) {
	Oct* o = (Oct*)h;
	if (levels < 1) return -1;
	if (step == 0) { o->buildSVO(levels, bmin, bmax, nullptr); o->levelSizesBeforeDag.assign(levels, 0); o->toDAG(true); }
	else { if (step + 1 >= levels) return -1; o->buildDAG(levels, step, bmin, bmax); }
	o->initChildLevels();  // main.cpp:203
	return 0;
}
int orc_build_svo_only(void* h, unsigned levels, const double bmin[3], const double bmax[3]) {
	((Oct*)h)->buildSVO(levels, bmin, bmax, nullptr);
	return 0;
}
int orc_build_svo_materials(void* h, unsigned levels, const double bmin[3], const double bmax[3], const uint32_t* triMat) {
	((Oct*)h)->buildSVO(levels, bmin, bmax, nullptr, triMat);
	return 0;
}
int orc_to_dag(void* h) { Oct* o = (Oct*)h; o->toDAG(true); o->initChildLevels(); return 0; }
int orc_to_sdag(void* h) { ((Oct*)h)->toSDAG(); return 0; }
unsigned orc_cross_merge(void* h) { return ((Oct*)h)->mergeAcrossAllLevels(); }
int orc_state(void* h) { return ((Oct*)h)->state; }
unsigned orc_levels(void* h) { return ((Oct*)h)->levels; }
uint64_t orc_level_size(void* h, unsigned lev) { Oct* o = (Oct*)h; return lev < o->data.size() ? o->data[lev].size() : 0; }
uint64_t orc_level_size_before_dag(void* h, unsigned lev) { Oct* o = (Oct*)h; return lev < o->levelSizesBeforeDag.size() ? o->levelSizesBeforeDag[lev] : 0; }
// which: 0 nTotalVoxels 1 nNodesSVO 2 nNodesDAG 3 nNodesSDAG 4 nNodesLastLevSVO 5 nNodesLastLevDAG 6 nCrossLevelMerged 7 nNodes(current)
uint64_t orc_stat(void* h, int which) {
	Oct* o = (Oct*)h;
	switch (which) {
		case 0: return o->stats.nTotalVoxels; case 1: return o->stats.nNodesSVO; case 2: return o->stats.nNodesDAG;
		case 3: return o->stats.nNodesSDAG; case 4: return o->stats.nNodesLastLevSVO; case 5: return o->stats.nNodesLastLevDAG;
		case 6: return o->stats.nCrossLevelMerged; case 7: return o->nNodes;
	}
	return 0;
}
double orc_root_side(void* h) { return ((Oct*)h)->rootSide; }
// The node order of one SSVDAG level exactly as the reference produces it (encoded_ssvdag.cpp:261-276): pairs (index, refs)
// in index order, plain std::sort with the reference's comparator.  order[r] = index of the node at rank r.
void orc_sort_by_refs(uint32_t n, const uint32_t* refs, uint32_t* order) {
	std::vector<std::pair<id_t32, id_t32>> hist(n);
	for (uint32_t i = 0; i < n; ++i) { hist[i].first = i; hist[i].second = refs[i]; }
	std::sort(hist.begin(), hist.end(), [](std::pair<id_t32, id_t32> a, std::pair<id_t32, id_t32> b) { return a.second > b.second; });
	for (uint32_t i = 0; i < n; ++i) order[i] = hist[i].first;
}
// SoA copy-out of one level (any pointer may be NULL)
int orc_get_level(void* h, unsigned lev, uint8_t* mask, uint32_t* child8, uint8_t* mirror3, uint8_t* inv, uint32_t* childLev8) {
	Oct* o = (Oct*)h;
	if (lev >= o->data.size()) return -1;
	const std::vector<Node>& v = o->data[lev];
	for (size_t i = 0; i < v.size(); ++i) {
		if (mask) mask[i] = v[i].mask;
		if (child8) for (int c = 0; c < 8; ++c) child8[i * 8 + c] = v[i].ch[c];
		if (mirror3) for (int a = 0; a < 3; ++a) mirror3[i * 3 + a] = v[i].mir[a];
		if (inv) inv[i] = v[i].inv;
		if (childLev8) for (int c = 0; c < 8; ++c) childLev8[i * 8 + c] = v[i].chLev[c];
	}
	return 0;
}
// kind: 0 .svdag, 1 .ussvdag, 2 .ssvdag/.esvdag.  Returns file size (bytes) or -1; copies if cap suffices.
int64_t orc_encode(void* h, int kind, uint8_t* buf, uint64_t cap) {
	Oct* o = (Oct*)h;
	std::vector<uint8_t> out;
	bool ok = (kind == 0) ? encodeSVDAG(*o, false, out) : (kind == 1) ? encodeSVDAG(*o, true, out) : encodeSSVDAG(*o, out);
	if (!ok) return -1;
	if (buf && cap >= out.size()) memcpy(buf, out.data(), out.size());
	return int64_t(out.size());
}
int orc_test_tri_box(const double center[3], double half, const float tri9[9]) { return testTriBox(center, half, tri9) ? 1 : 0; }

}  // extern "C"
