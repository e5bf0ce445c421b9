// encoders.cpp -- byte-exact writers of the reference's on-disk formats from SoA level arrays.
// Linear host passes (< 1 % of the build), kept on the host like the
// reference does; quirks reproduced on purpose are listed in SURVEY.md Appendix B (5-8).
#include "octree_data.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>

namespace svbhost {

namespace {

const uint32_t kNull = 0xFFFFFFFEu;

// Threads of the encoder loops.  Not OMP_NUM_THREADS: launchers such as torchrun pin that to 1 for every rank, and the
// file is written by one rank only.  SVB_ENCODE_THREADS overrides; default min(8, hardware threads).
int encode_threads() {
	static const int n = [] {
		if (const char* e = getenv("SVB_ENCODE_THREADS")) { int v = atoi(e); if (v > 0) return v; }
		unsigned hw = std::thread::hardware_concurrency();
		return (int)std::max(1u, std::min(8u, hw ? hw : 1u));
	}();
	return n;
}

// ---- the SSVDAG node order: std::sort by reference count, descending (encoded_ssvdag.cpp:261-276) ------------------------
// The reference's std::sort is UNSTABLE, and the file depends on how it happens to leave the ties.  libstdc++'s std::sort is
//     __introsort_loop(first, last, 2*lg(n));  __final_insertion_sort(first, last);
// and the loop is  { cut = __unguarded_partition_pivot(first, last); recurse on [cut, last); last = cut; }  -- after a
// partition the two sides never interact again.  So the same library routines can run the right-hand sides as tasks on
// other threads and produce, comparison for comparison and swap for swap, the permutation the sequential call yields.
// What is called below are libstdc++'s own __unguarded_partition_pivot / __introsort_loop / __partial_sort /
// __final_insertion_sort (the reference binary is built against the same headers); only the recursion is ours.
typedef std::pair<uint32_t, uint32_t> IdxRefs;   // (node index, references from the level above)
struct MoreRefs { bool operator()(const IdxRefs& a, const IdxRefs& b) const { return a.second > b.second; } };

long sort_par_min() {
	static const long v = [] { const char* e = getenv("SVB_SORT_PAR_MIN"); long x = e ? atol(e) : 0; return x > 32 ? x : 16384L; }();
	return v;
}

template <class It, class Comp>
void introsort_tasks(It first, It last, long depth, Comp comp, long parMin) {
	while (last - first > 16) {                                   // _S_threshold
		if (last - first <= parMin) { std::__introsort_loop(first, last, depth, comp); return; }
		if (depth == 0) { std::__partial_sort(first, last, last, comp); return; }
		--depth;
		It cut = std::__unguarded_partition_pivot(first, last, comp);
#pragma omp task default(none) firstprivate(cut, last, depth, comp, parMin)
		introsort_tasks(cut, last, depth, comp, parMin);
		last = cut;
	}
}

// == std::sort(v.begin(), v.end(), MoreRefs()); call from inside an OpenMP parallel region to have the tasks shared out
void sort_by_refs(std::vector<IdxRefs>& v) {
	if (v.size() < 2) return;
	auto comp = __gnu_cxx::__ops::__iter_comp_iter(MoreRefs());
#pragma omp taskgroup
	introsort_tasks(v.begin(), v.end(), (long)std::__lg((long)v.size()) * 2, comp, sort_par_min());
	std::__final_insertion_sort(v.begin(), v.end(), comp);
}

struct ByteSink {
	std::vector<uint8_t>& v;
	template <class T> void pod(const T& x) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&x); v.insert(v.end(), p, p + sizeof(T)); }
	template <class T> void array(const std::vector<T>& a) {
		pod<uint32_t>((uint32_t)a.size());
		const uint8_t* p = reinterpret_cast<const uint8_t*>(a.data());
		v.insert(v.end(), p, p + a.size() * sizeof(T));
	}
};

void common_header(ByteSink& s, const OctreeData& o) {
	for (int k = 0; k < 6; ++k) s.pod<float>(o.bboxF[k]);
	s.pod<float>((float)o.rootSide);                  // Octree::getRootSide()
	s.pod<uint32_t>((uint32_t)o.levels.size());
	s.pod<uint32_t>((uint32_t)o.nNodes);              // low 4 bytes of the size_t
}

inline int popc(unsigned m) { return __builtin_popcount(m & 0xFF); }

// .svdag / .ussvdag: one u32 stream; node = header word, then the child words for k = 7..0
bool pointer_stream(const OctreeData& o, bool withMirrorHeader, std::vector<uint8_t>& out) {
	const size_t L = o.levels.size();
	// absolute word offset of every node, all levels concatenated
	std::vector<uint32_t> levelStart(L + 1, 0);        // node index where a level starts in the concatenation
	size_t totalNodes = 0;
	for (size_t l = 0; l < L; ++l) { levelStart[l] = (uint32_t)totalNodes; totalNodes += o.levels[l].n; }
	levelStart[L] = (uint32_t)totalNodes;
	std::vector<uint32_t> wordOf(totalNodes);
	uint32_t words = 0;
	for (size_t l = 0, g = 0; l < L; ++l)
		for (uint64_t i = 0; i < o.levels[l].n; ++i, ++g) {
			wordOf[g] = words;
			words += (l + 1 < L) ? 1u + (uint32_t)popc(o.levels[l].mask[i]) : 1u;
		}
	const uint32_t firstLeafPtr = words;               // quirk: computed after the leaf level too
	std::vector<uint32_t> data(words);
	for (size_t l = 0; l < L; ++l) {
		const LevelSoA& lv = o.levels[l];
		const bool hasCL = lv.childLevel.size() == lv.n * 8;
		const uint32_t* wo = wordOf.data() + levelStart[l];
#pragma omp parallel for schedule(static) num_threads(encode_threads()) if (lv.n > 4096)
		for (int64_t ii = 0; ii < (int64_t)lv.n; ++ii) {
			const uint64_t i = (uint64_t)ii;
			uint32_t head = lv.mask[i];
			if (withMirrorHeader) head |= ((uint32_t)lv.mirror[i * 3 + 2] << 24) | ((uint32_t)lv.mirror[i * 3 + 1] << 16) | ((uint32_t)lv.mirror[i * 3] << 8);
			uint32_t w = wo[i];
			data[w++] = head;
			if (l + 1 >= L) continue;                     // leaf level: mask word only
			for (int k = 7; k >= 0; --k) {
				uint32_t c = lv.child[i * 8 + k];
				if (c == kNull) continue;
				// USSVDAG always points into the next level; SVDAG honours childLevels (cross-level merge)
				size_t tl = withMirrorHeader ? l + 1 : (hasCL ? lv.childLevel[i * 8 + k] : l + 1);
				data[w++] = wordOf[(size_t)c + levelStart[tl]];
			}
		}
	}
	ByteSink s{out};
	common_header(s, o);
	s.pod<uint32_t>(firstLeafPtr);
	s.array(data);
	return true;
}

uint8_t mirror_mask(uint8_t m, int s) {   // Node::mirror on a leaf mask: slot i <- slot i^s
	uint8_t r = 0;
	for (int i = 0; i < 8; ++i) if ((m >> (i ^ s)) & 1) r |= (uint8_t)(1u << i);
	return r;
}

// lookup tables for the 4^3 leaf bricks: MIRROR[s][m] = mirror_mask(m, s); BRICK[c][m] = the 64-bit brick bits that the
// 2^3 voxel mask m of sub-block c contributes (bits re-ordered x-fastest, encoded_ssvdag.cpp:119-135, :280-351)
struct LeafTables {
	uint8_t mirror[8][256];
	uint64_t brick[8][256];
	LeafTables() {
		for (int s = 0; s < 8; ++s) for (int m = 0; m < 256; ++m) mirror[s][m] = mirror_mask((uint8_t)m, s);
		for (int c = 0; c < 8; ++c)
			for (int m = 0; m < 256; ++m) {
				uint64_t b = 0;
				for (unsigned bit = 0; bit < 64; ++bit) {
					unsigned x = bit & 3, y = (bit >> 2) & 3, z = bit >> 4;
					unsigned byteId = (x >> 1) + ((y >> 1) << 1) + ((z >> 1) << 2);
					unsigned bitId = (x & 1) + ((y & 1) << 1) + ((z & 1) << 2);
					if ((int)byteId == c && ((m >> bitId) & 1)) b |= 1ull << bit;
				}
				brick[c][m] = b;
			}
	}
};

bool ssvdag(const OctreeData& o, std::vector<uint8_t>& out, std::string* err) {
	const int L = (int)o.levels.size();
	if (L < 3) { if (err) *err = "SSVDAG needs at least 3 levels"; return false; }
	std::vector<std::vector<uint16_t>> inner(L - 2);
	std::vector<uint8_t> leaves;
	std::vector<uint32_t> addr, nextAddr;   // node index -> address inside its encoded level
	for (int lev = 0; lev <= L - 2; ++lev)
		if (o.levels[lev].n > (1ull << 30)) { if (err) *err = "level too big for 30-bit pointers"; return false; }
	// Phase 1: the node order of every level (most-referenced first) only depends on the pointers of the level above,
	// so the levels are counted and sorted concurrently, one thread per level.  The reference uses the unstable
	// std::sort, so must we (same libstdc++, same initial sequence, same comparator => same permutation).
	std::vector<std::vector<IdxRefs>> orders(L - 1);
#pragma omp parallel for schedule(dynamic, 1) num_threads(encode_threads())
	for (int lev = L - 2; lev >= 0; --lev) {
		const LevelSoA& cur = o.levels[lev];
		std::vector<IdxRefs>& order = orders[lev];
		order.resize(cur.n);
		for (uint32_t i = 0; i < cur.n; ++i) order[i] = IdxRefs(i, 0);
		if (lev > 0) {
			const LevelSoA& up = o.levels[lev - 1];
			for (uint64_t q = 0; q < up.n * 8; ++q) if (up.child[q] != kNull) order[up.child[q]].second++;
			sort_by_refs(order);
		}
	}
	// Phase 2: bottom-up, a level's pointers need the addresses of the level below
	for (int lev = L - 2; lev >= 0; --lev) {
		const LevelSoA& cur = o.levels[lev];
		const std::vector<IdxRefs>& order = orders[lev];
		nextAddr.assign(cur.n, 0);
		if (lev == L - 2) {
			// two deepest levels fused into 4^3 bit bricks, bits re-ordered x-fastest
			static const LeafTables lut;
			const LevelSoA& leaf = o.levels[lev + 1];
			leaves.assign(cur.n * 8, 0);
#pragma omp parallel for schedule(static) num_threads(encode_threads()) if (order.size() > 2048)
			for (int64_t rr = 0; rr < (int64_t)order.size(); ++rr) {
				const uint32_t r = (uint32_t)rr;
				uint32_t i = order[r].first;
				nextAddr[i] = r;
				uint64_t b = 0;
				for (int c = 0; c < 8; ++c) {
					uint32_t ch = cur.child[(uint64_t)i * 8 + c];
					if (ch == kNull) continue;
					int sft = (((cur.mirror[(uint64_t)i * 3] >> c) & 1) << 2) | (((cur.mirror[(uint64_t)i * 3 + 1] >> c) & 1) << 1) | ((cur.mirror[(uint64_t)i * 3 + 2] >> c) & 1);
					b |= lut.brick[c][lut.mirror[sft][leaf.mask[ch]]];
				}
				memcpy(&leaves[(uint64_t)r * 8], &b, 8);   // little endian: byte j holds brick bits 8j..8j+7
			}
		} else {
			// pass 1: encoded size of every node (1 header + 1 or 2 shorts per child), pass 2: fill at the prefix offsets
			std::vector<uint16_t>& enc = inner[lev];
			std::vector<uint32_t> sz(order.size());
#pragma omp parallel for schedule(static) num_threads(encode_threads()) if (order.size() > 2048)
			for (int64_t rr = 0; rr < (int64_t)order.size(); ++rr) {
				uint32_t i = order[rr].first, n = 1;
				for (int c = 7; c >= 0; --c) {
					if (!((cur.mask[i] >> c) & 1)) continue;
					uint32_t a = addr[cur.child[(uint64_t)i * 8 + c]];
					n += (a < (1u << 13)) ? 1u : ((a < (1u << 30)) ? 2u : 0u);
				}
				sz[rr] = n;
			}
			uint64_t total = 0;
			for (size_t r = 0; r < order.size(); ++r) { uint32_t n = sz[r]; sz[r] = (uint32_t)total; total += n; }
			enc.assign(total, 0);
#pragma omp parallel for schedule(static) num_threads(encode_threads()) if (order.size() > 2048)
			for (int64_t rr = 0; rr < (int64_t)order.size(); ++rr) {
				uint32_t i = order[rr].first;
				nextAddr[i] = sz[rr];
				size_t w = sz[rr];
				const size_t headAt = w++;
				uint16_t head = 0;
				for (int c = 7; c >= 0; --c) {
					if (!((cur.mask[i] >> c) & 1)) continue;
					uint32_t a = addr[cur.child[(uint64_t)i * 8 + c]];
					unsigned mx = (cur.mirror[(uint64_t)i * 3] >> c) & 1, my = (cur.mirror[(uint64_t)i * 3 + 1] >> c) & 1, mz = (cur.mirror[(uint64_t)i * 3 + 2] >> c) & 1;
					if (a < (1u << 13)) {
						head |= (uint16_t)(1u << (2 * c));
						enc[w++] = (uint16_t)(a | (mx << 13) | (my << 14) | (mz << 15));
					} else if (a < (1u << 30)) {
						uint32_t pp = a;
						if (pp & (1u << 29)) { head |= (uint16_t)(3u << (2 * c)); pp &= ~(1u << 29); }
						else head |= (uint16_t)(2u << (2 * c));
						pp |= (mx << 29) | (my << 30) | (mz << 31);
						enc[w++] = (uint16_t)(pp >> 16);
						enc[w++] = (uint16_t)(pp & 0xFFFF);
					}
				}
				enc[headAt] = head;
			}
		}
		addr.swap(nextAddr);
	}
	std::vector<uint32_t> levelOffsets(L - 2, 0);
	for (int i = 1; i < L - 2; ++i) levelOffsets[i] = levelOffsets[i - 1] + (uint32_t)inner[i - 1].size();
	std::vector<uint16_t> all;
	for (auto& e : inner) all.insert(all.end(), e.begin(), e.end());
	ByteSink s{out};
	common_header(s, o);
	s.array(all);
	s.array(leaves);
	s.array(levelOffsets);
	return true;
}

}  // namespace

// The SSVDAG node order of every level from its reference counts (the device encoder computes the counts, svb_encode.cu):
// refs = counts of levels 0 .. L-2 concatenated, start[l] = first entry of level l; order (same layout) receives the node
// indices, most referenced first, ties exactly as the reference's std::sort leaves them.
void ssvdag_order_from_refs(const uint32_t* refs, const uint32_t* start, int nLevels, uint32_t* order) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(encode_threads())
	for (int lev = nLevels - 1; lev >= 0; --lev) {
		const uint32_t n = start[lev + 1] - start[lev];
		std::vector<IdxRefs> v(n);
		for (uint32_t i = 0; i < n; ++i) v[i] = IdxRefs(i, refs[start[lev] + i]);
		if (lev > 0) sort_by_refs(v);
		for (uint32_t i = 0; i < n; ++i) order[start[lev] + i] = v[i].first;
	}
}

// EncodedSVDAG::load + decode (encoded_svdag.cpp:43-74, :200-270): one u32 stream, levels stored one after the other;
// a level ends where the smallest child pointer of its nodes points.  Nodes keep file order; child pointers (absolute
// word offsets in the file) become indices into the next level.  Cross-level (-multi) files are not decodable, as in
// the reference.
bool decode_svdag(const uint8_t* file, uint64_t size, OctreeData& o, std::string* err) {
	if (size < 44) { if (err) *err = "not an SVDAG file"; return false; }
	uint32_t levels, nNodes, firstLeafPtr, count;
	float rootSide;
	memcpy(o.bboxF, file, 24);
	memcpy(&rootSide, file + 24, 4);
	memcpy(&levels, file + 28, 4);
	memcpy(&nNodes, file + 32, 4);
	memcpy(&firstLeafPtr, file + 36, 4);
	memcpy(&count, file + 40, 4);
	if (levels < 2 || levels > 32 || size < 44 + 4ull * count || count == 0) { if (err) *err = "corrupt SVDAG header"; return false; }
	const uint32_t* data = reinterpret_cast<const uint32_t*>(file + 44);
	o.rootSide = rootSide;
	o.nNodes = nNodes;
	o.nVoxels = 0;                                    // EncodedSVDAG::load does not know it either
	o.state = 2;                                      // S_DAG
	o.levels.assign(levels, LevelSoA());
	std::vector<uint32_t> indexOfWord(count, 0);      // word offset of a node -> its index inside its level
	std::vector<uint8_t> levelOfWord(count, 0xFF);    // level of the node whose header sits at that word (0xFF: not a header)
	std::vector<uint32_t> levStart(levels, 0xFFFFFFFFu);
	levStart[0] = 0;
	uint32_t lev = 0;
	for (uint32_t i = 0; i < count; ++i) {
		if (lev + 1 < levels && i == levStart[lev + 1]) ++lev;
		LevelSoA& L = o.levels[lev];
		const uint8_t m = (uint8_t)data[i];
		indexOfWord[i] = (uint32_t)L.n;
		levelOfWord[i] = (uint8_t)lev;
		L.mask.push_back(m);
		L.child.insert(L.child.end(), 8, kNull);
		if (lev + 1 < levels) {
			uint32_t c = 0;
			for (int k = 7; k >= 0; --k) {
				if (!((m >> k) & 1)) continue;
				if (i + 1 + c >= count) { if (err) *err = "corrupt SVDAG stream"; return false; }
				const uint32_t ptr = data[i + 1 + c];
				++c;
				L.child[L.n * 8 + k] = ptr;               // absolute for now
				levStart[lev + 1] = std::min(levStart[lev + 1], ptr);
			}
			i += c;
		}
		L.n++;
	}
	for (uint32_t l = 0; l + 1 < levels; ++l)
		for (uint32_t& c : o.levels[l].child)
			if (c != kNull) {
				if (c >= count) { if (err) *err = "SVDAG pointer out of range"; return false; }
				// a pointer must land on the header word of a node of the next level (a -multi file, or a corrupt one, does not)
				if (levelOfWord[c] != l + 1) { if (err) *err = "SVDAG pointer does not address a node of the next level (cross-level or corrupt file)"; return false; }
				c = indexOfWord[c];
			}
	for (auto& L : o.levels) { L.mirror.assign(L.n * 3, 0); L.inv.assign(L.n, 0); }
	return true;
}

bool encode_file(const OctreeData& o, int kind, std::vector<uint8_t>& out, std::string* err) {
	out.clear();
	const int S_DAG = 2, S_SDAG = 3;
	if (kind == 0) {
		if (o.state != S_DAG) { if (err) *err = "FAILED! Octree is not in DAG state"; return false; }      // encoded_svdag.cpp:109-112
		return pointer_stream(o, false, out);
	}
	if (kind == 1) {
		if (o.state != S_SDAG) { if (err) *err = "FAILED! Octree is not in SDAG state"; return false; }    // encoded_ussvdag.cpp:90-93
		return pointer_stream(o, true, out);
	}
	if (kind == 2) {
		if (o.state != S_DAG && o.state != S_SDAG) { if (err) *err = "FAILED! Octree is not in SDAG state"; return false; }   // encoded_ssvdag.cpp:218-221
		return ssvdag(o, out, err);
	}
	if (err) *err = "unknown encoding";
	return false;
}

}  // namespace svbhost
