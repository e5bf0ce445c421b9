// svb_cross.cu -- CSVDAG: merge identical subtrees ACROSS levels (GeomOctree::mergeAcrossAllLevels,
// src/symvox/geom_octree_extension.cpp:1192-1544, compareSubtrees :238-290).
//
// The reference walks the DAG top-down and, for a node A at level a whose subtree spans D = L-1-a
// levels, searches the levels [1, a) for the first node B (lowest level, then lowest index) whose
// subtree TRUNCATED to D levels has the same child masks; A and everything below it is then dropped
// and pointers to A become (level(B), B).  Closed form used here (derivation in DESIGN.md §4, pinned
// against the sequential oracle and the reference's -multi.svdag bytes):
//     n at level l is removed  <=>  some node at a level in [1, l) has the same depth-D(l) truncated-subtree id;
//     its replacement is the minimum (level, index) such node, which always survives;
//     survivors keep their relative order.
// Truncated-subtree ids are computed bottom-up in the depth d = 0 .. L-2: tid_0 = child mask,
// tid_d(n) = intern(mask(n), tid_{d-1}(children)), interned JOINTLY over the levels 1 .. L-1-d in one
// open-addressing table (64-bit tag + exact verification), whose slots also hold the minimum
// (level, index) of the nodes that share the id.
#include "svb_cross.cuh"

namespace svb {

namespace {

constexpr int CM_THREADS = 256;
constexpr int MAXL = 22;

struct LevelView {
	uint64_t n;
	const uint8_t* mask;
	const uint32_t* child;     // n*8, NULLNODE = none
	const uint32_t* tidPrev;   // depth d-1 ids of THIS level's nodes
	uint32_t* slotOf;          // depth d ids (table slots) being computed
};
struct Views { LevelView lv[MAXL]; int L; };

struct CKey { uint32_t w[9]; };

__device__ __forceinline__ void make_key(const Views& V, int l, uint64_t i, int d, CKey& k) {
	const LevelView& X = V.lv[l];
	k.w[0] = X.mask[i];
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		uint32_t ch = X.child[i * 8 + c];
		k.w[1 + c] = (d == 0 || ch == NULLNODE) ? NULLREF : V.lv[l + 1].tidPrev[ch];
	}
}
__device__ __forceinline__ uint64_t key_tag(const CKey& k, const HashSeed& hs) {
	uint64_t h = hs.init ^ k.w[0];
#pragma unroll
	for (int c = 1; c < 9; c += 2) h = mix64(h ^ (((uint64_t)k.w[c + 1] << 32) | k.w[c])) + 0x9E3779B97F4A7C15ull * c;
	return finish_tag(h, hs);
}

__global__ void __launch_bounds__(CM_THREADS) k_cm_insert(Views V, int l, int d, unsigned long long* __restrict__ tag, unsigned long long* __restrict__ minLoc,
                                                           uint64_t capMask, uint32_t* __restrict__ flags, HashSeed hs) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= V.lv[l].n) return;
	CKey k;
	make_key(V, l, i, d, k);
	uint64_t h = key_tag(k, hs);
	uint64_t idx = mix64(h) & capMask;
	bool found = false;
	for (int probe = 0; probe < 8192; ++probe) {
		unsigned long long cur = tag[idx];
		if (cur == h) { found = true; break; }
		if (cur == 0ull) {
			unsigned long long old = atomicCAS(&tag[idx], 0ull, (unsigned long long)h);
			if (old == 0ull || old == h) { found = true; break; }
		}
		idx = (idx + 1) & capMask;
	}
	if (!found) { flags[0] = 1; V.lv[l].slotOf[i] = UNSET; return; }
	unsigned long long loc = ((unsigned long long)l << 32) | (unsigned long long)i;
	if (minLoc[idx] > loc) atomicMin(&minLoc[idx], loc);
	V.lv[l].slotOf[i] = (uint32_t)idx;
}

// exact check: a node's key must equal the key of the slot's first (level, index) node
__global__ void __launch_bounds__(CM_THREADS) k_cm_verify(Views V, int l, int d, const unsigned long long* __restrict__ minLoc, uint32_t* __restrict__ flags) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= V.lv[l].n) return;
	uint32_t s = V.lv[l].slotOf[i];
	if (s == UNSET) return;
	unsigned long long loc = minLoc[s];
	int lw = (int)(loc >> 32);
	uint64_t iw = loc & 0xFFFFFFFFull;
	if (lw == l && iw == i) return;
	CKey a, b;
	make_key(V, l, i, d, a);
	make_key(V, lw, iw, d, b);
	bool same = true;
#pragma unroll
	for (int c = 0; c < 9; ++c) same &= (a.w[c] == b.w[c]);
	if (!same) flags[1] = 1;
}

// level l* = L-1-d: removed iff the id's first node lives on a higher-up level
__global__ void __launch_bounds__(CM_THREADS) k_cm_decide(uint64_t n, int l, const uint32_t* __restrict__ slotOf, const unsigned long long* __restrict__ minLoc,
                                                           uint32_t* __restrict__ keep, unsigned long long* __restrict__ target) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned long long loc = minLoc[slotOf[i]];
	bool removed = (int)(loc >> 32) < l;
	keep[i] = removed ? 0u : 1u;
	target[i] = removed ? loc : (((unsigned long long)l << 32) | i);
}

struct OutView {
	const uint32_t* keep;              // per level
	const uint32_t* newIndex;
	const unsigned long long* target;
};
struct OutViews { OutView lv[MAXL]; };

__global__ void __launch_bounds__(CM_THREADS) k_cm_compact(uint64_t n, int l, int L, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ child, OutViews O,
                                                            uint8_t* __restrict__ omask, uint32_t* __restrict__ ochild, uint32_t* __restrict__ olevel) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !O.lv[l].keep[i]) return;
	uint64_t o = O.lv[l].newIndex[i];
	omask[o] = mask[i];
	for (int c = 0; c < 8; ++c) {
		uint32_t ch = child[i * 8 + c];
		uint32_t nl = (uint32_t)(l + 1), ni = NULLNODE;
		if (ch != NULLNODE && l + 1 < L) {
			unsigned long long t = O.lv[l + 1].target[ch];   // itself if kept, else its replacement (level, index)
			nl = (uint32_t)(t >> 32);
			ni = O.lv[nl].newIndex[t & 0xFFFFFFFFull];
		}
		ochild[o * 8 + c] = ni;
		olevel[o * 8 + c] = nl;
	}
}

__global__ void k_fill_u32(uint64_t n, uint32_t v, uint32_t* p) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}
__global__ void k_self_target(uint64_t n, int l, unsigned long long* t) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) t[i] = ((unsigned long long)l << 32) | i;
}

}  // namespace

uint64_t cross_merge_device(svb_ctx* c, uint64_t* nNodesOut) {
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	const int L = (int)c->levels;
	if (L > MAXL) throw Error(SVB_ERANGE, "too many levels");
	std::vector<DevBuf<uint32_t>> tidPrev(L), slotOf(L), keep(L), newIndex(L);
	std::vector<DevBuf<uint64_t>> target(L);
	for (int l = 0; l < L; ++l) {
		uint64_t n = c->out[l].n;
		if (n >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "level too large");
		tidPrev[l].reset(pool, n ? n : 1);
		slotOf[l].reset(pool, n ? n : 1);
		keep[l].reset(pool, n ? n : 1);
		newIndex[l].reset(pool, n ? n : 1);
		target[l].reset(pool, n ? n : 1);
		if (n) {
			k_fill_u32<<<blocks_for(n, 256), 256, 0, s>>>(n, 1u, keep[l].p);
			SVB_KERNEL_CHECK();
			k_self_target<<<blocks_for(n, 256), 256, 0, s>>>(n, l, (unsigned long long*)target[l].p);
			SVB_KERNEL_CHECK();
		}
	}
	DevBuf<uint32_t> flags(pool, 4);
	for (int d = 0; d <= L - 2; ++d) {
		const int lstar = L - 1 - d;
		if (lstar < 2) break;   // level-1 nodes have no level in [1,1) to match against; level 0 is never a target (geom_octree.hpp:181)
		uint64_t total = 0;
		for (int l = 1; l <= lstar; ++l) total += c->out[l].n;
		uint64_t cap = 1024;
		while (cap < 2 * total) cap <<= 1;
		DevBuf<uint64_t> tag(pool, cap), minLoc(pool, cap);
		tag.zero();
		minLoc.fill_ff();
		flags.zero();
		Views V;
		V.L = L;
		for (int l = 0; l < L; ++l) {
			V.lv[l].n = c->out[l].n; V.lv[l].mask = c->out[l].mask.p; V.lv[l].child = c->out[l].child.p;
			V.lv[l].tidPrev = tidPrev[l].p; V.lv[l].slotOf = slotOf[l].p;
		}
		for (int l = 1; l <= lstar; ++l) {
			if (!c->out[l].n) continue;
			k_cm_insert<<<blocks_for(c->out[l].n, CM_THREADS), CM_THREADS, 0, s>>>(V, l, d, (unsigned long long*)tag.p, (unsigned long long*)minLoc.p, cap - 1, flags.p, make_hash_seed(c->hashSeed));
			SVB_KERNEL_CHECK();
		}
		for (int l = 1; l <= lstar; ++l) {
			if (!c->out[l].n) continue;
			k_cm_verify<<<blocks_for(c->out[l].n, CM_THREADS), CM_THREADS, 0, s>>>(V, l, d, (const unsigned long long*)minLoc.p, flags.p);
			SVB_KERNEL_CHECK();
		}
		if (c->out[lstar].n) {
			k_cm_decide<<<blocks_for(c->out[lstar].n, CM_THREADS), CM_THREADS, 0, s>>>(c->out[lstar].n, lstar, slotOf[lstar].p, (const unsigned long long*)minLoc.p,
			                                                                          keep[lstar].p, (unsigned long long*)target[lstar].p);
			SVB_KERNEL_CHECK();
		}
		uint32_t h[4];
		SVB_CUDA(cudaMemcpyAsync(h, flags.p, 16, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		if (h[0]) throw Error(SVB_ECUDA, "cross-level merge: hash table overflow");
		if (h[1]) throw Error(SVB_ECOLLISION, "cross-level merge: 64-bit subtree-id hash collision");
		for (int l = 1; l <= lstar; ++l) std::swap(tidPrev[l], slotOf[l]);   // depth-d ids feed depth d+1
	}
	// survivors keep their relative order (ext.cpp:1434-1440)
	DevBuf<uint64_t> tot(pool, 1);
	std::vector<uint64_t> newN(L, 0);
	for (int l = 0; l < L; ++l) {
		scan_u32(s, pool, keep[l].p, c->out[l].n, newIndex[l].p, tot.p);
		SVB_CUDA(cudaMemcpyAsync(&newN[l], tot.p, 8, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
	}
	OutViews O;
	for (int l = 0; l < L; ++l) { O.lv[l].keep = keep[l].p; O.lv[l].newIndex = newIndex[l].p; O.lv[l].target = (const unsigned long long*)target[l].p; }
	for (int l = L; l < MAXL; ++l) { O.lv[l].keep = nullptr; O.lv[l].newIndex = nullptr; O.lv[l].target = nullptr; }
	std::vector<OutLevel> res(L);
	uint64_t removed = 0, nn = 1;
	for (int l = 0; l < L; ++l) {
		OutLevel& X = c->out[l];
		OutLevel& Y = res[l];
		uint64_t m = newN[l];
		Y.n = m;
		Y.mask.reset(pool, m); Y.child.reset(pool, m * 8); Y.mirror.reset(pool, m * 3); Y.inv.reset(pool, m); Y.childLevel.reset(pool, m * 8);
		Y.mirror.zero(); Y.inv.zero();
		Y.hasChildLevel = true;
		if (X.n) {
			k_cm_compact<<<blocks_for(X.n, CM_THREADS), CM_THREADS, 0, s>>>(X.n, l, L, X.mask.p, X.child.p, O, Y.mask.p, Y.child.p, Y.childLevel.p);
			SVB_KERNEL_CHECK();
		}
		removed += X.n - m;
		if (l >= 1) nn += m;
	}
	SVB_CUDA(cudaStreamSynchronize(s));
	c->out = std::move(res);
	if (nNodesOut) *nNodesOut = nn;   // ext.cpp:1292,1448: 1 + sum of the surviving levels >= 1
	return removed;
}

}  // namespace svb
