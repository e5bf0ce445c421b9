// svb_sdag.cu -- SSVDAG mirror-symmetry canonicalisation on the device.
//
// Replaces the sequential loop of GeomOctree::toSDAG (src/symvox/geom_octree.cpp:551-697): per
// level, bottom-up, every node is matched against the 8 mirror images (identity, X, Y, Z, XY, XZ,
// YZ, XYZ -- in that lookup priority, :630-653) of the unique nodes seen so far.  Parallel
// restatement (pinned on CPU by oracle/parallel_model.py::to_sdag against the sequential oracle):
//   class(n)  = lexicographic min over the 8 variant keys K_s(n)            (orbit representative key)
//   rep       = smallest index in the class  -> keeps its identity key, new id = rank among reps
//   flags(n)  = first s in priority order with K_s(n) == K_id(rep)
// K_s(n) = Node::mirror (src/symvox/octree_node.cpp:88-195: slots i <-> i^s in mask, children and
// the three mirror masks, then toggle the mirrored axes' bits of every existing child) followed by
// invertInvs (:822-831: clear the bit where the child is invariant on that axis) for inner levels.
// The invariant bits themselves come from the RAW mirror (no invertInvs), as in :604-606.
#include "svb_sdag.cuh"

namespace svb {

namespace {

constexpr int SD_THREADS = 128;
// lookup priority of the reference as axis masks (X=4, Y=2, Z=1)
__constant__ int c_priority[8] = {0, 4, 2, 1, 6, 5, 3, 7};

struct SKey { uint32_t w[12]; };   // mask, child[0..7], mirror x,y,z  (Node::operator< order)

__device__ __forceinline__ unsigned perm_bits(unsigned m, int s) {   // bit i of result = bit (i^s) of m
	if (s & 4) m = ((m << 4) | (m >> 4)) & 0xFF;
	if (s & 2) m = ((m & 0x33) << 2) | ((m >> 2) & 0x33);
	if (s & 1) m = ((m & 0x55) << 1) | ((m >> 1) & 0x55);
	return m;
}

struct NodeIn {
	unsigned mask;
	uint32_t ch[8];
	unsigned mir[3];
};

__device__ __forceinline__ void load_node(const uint8_t* mask, const uint32_t* child, const uint8_t* mirror, uint64_t i, NodeIn& n) {
	n.mask = mask[i];
	const uint4* c4 = reinterpret_cast<const uint4*>(child + i * 8);
	uint4 a = c4[0], b = c4[1];
	n.ch[0] = a.x; n.ch[1] = a.y; n.ch[2] = a.z; n.ch[3] = a.w; n.ch[4] = b.x; n.ch[5] = b.y; n.ch[6] = b.z; n.ch[7] = b.w;
	n.mir[0] = mirror[i * 3]; n.mir[1] = mirror[i * 3 + 1]; n.mir[2] = mirror[i * 3 + 2];
}

// childInv == nullptr: raw mirror only (leaf level, or the invariance test)
__device__ __forceinline__ void variant(const NodeIn& n, int s, const uint8_t* __restrict__ childInv, SKey& k) {
	k.w[0] = perm_bits(n.mask, s);
	unsigned has = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		uint32_t c = n.ch[i ^ s];
		k.w[1 + i] = c;
		if (c != NULLNODE) has |= 1u << i;
	}
	unsigned mx = perm_bits(n.mir[0], s), my = perm_bits(n.mir[1], s), mz = perm_bits(n.mir[2], s);
	if (s & 4) mx ^= has;
	if (s & 2) my ^= has;
	if (s & 1) mz ^= has;
	if (childInv && s) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (!((has >> i) & 1)) continue;
			unsigned inv = childInv[k.w[1 + i]];
			if ((s & 4) && (inv & 1)) mx &= ~(1u << i);
			if ((s & 2) && (inv & 2)) my &= ~(1u << i);
			if ((s & 1) && (inv & 4)) mz &= ~(1u << i);
		}
	}
	k.w[9] = mx; k.w[10] = my; k.w[11] = mz;
}

__device__ __forceinline__ bool key_eq(const SKey& a, const SKey& b) {
	bool e = true;
#pragma unroll
	for (int i = 0; i < 12; ++i) e &= (a.w[i] == b.w[i]);
	return e;
}
__device__ __forceinline__ bool key_lt(const SKey& a, const SKey& b) {
#pragma unroll
	for (int i = 0; i < 12; ++i) {
		if (a.w[i] < b.w[i]) return true;
		if (a.w[i] > b.w[i]) return false;
	}
	return false;
}
__device__ __forceinline__ uint64_t key_hash(const SKey& k, const HashSeed& hs) {
	uint64_t h = hs.init;
#pragma unroll
	for (int i = 0; i < 12; i += 2) h = mix64(h ^ (((uint64_t)k.w[i + 1] << 32) | k.w[i])) + 0x9E3779B97F4A7C15ull * (i + 1);
	return finish_tag(h, hs);
}

struct LevelIn {
	uint64_t n;
	const uint8_t* mask;
	const uint32_t* child;
	const uint8_t* mirror;
	const uint8_t* childInv;   // inv of the (already reduced) level below; nullptr at the leaf level
};

// pass 1: invariant bits, class key (as argmin variant), hash-table insert with atomicMin(index)
__global__ void __launch_bounds__(SD_THREADS) k_sdag_class(LevelIn L, uint8_t* __restrict__ inv, uint8_t* __restrict__ clsVar, uint32_t* __restrict__ slotOf,
                                                            unsigned long long* __restrict__ tag, uint32_t* __restrict__ minIdx, uint64_t capMask, uint32_t* __restrict__ flags, HashSeed hs) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L.n) return;
	NodeIn n;
	load_node(L.mask, L.child, L.mirror, i, n);
	if (n.mask == 0) { slotOf[i] = UNSET; inv[i] = 0; clsVar[i] = 0; return; }
	SKey k0, best, k;
	variant(n, 0, nullptr, k0);
	unsigned iv = 0;
	variant(n, 4, nullptr, k); if (key_eq(k, k0)) iv |= 1;
	variant(n, 2, nullptr, k); if (key_eq(k, k0)) iv |= 2;
	variant(n, 1, nullptr, k); if (key_eq(k, k0)) iv |= 4;
	inv[i] = (uint8_t)iv;
	best = k0;
	int bs = 0;
	for (int s = 1; s < 8; ++s) {
		variant(n, s, L.childInv, k);
		if (key_lt(k, best)) { best = k; bs = s; }
	}
	clsVar[i] = (uint8_t)bs;
	uint64_t h = key_hash(best, hs);
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
	if (!found) { flags[0] = 1; slotOf[i] = UNSET; return; }
	if (minIdx[idx] > (uint32_t)i) atomicMin(&minIdx[idx], (uint32_t)i);
	slotOf[i] = (uint32_t)idx;
}

// pass 2: representative, exact class check, mirror flags of non-representatives
__global__ void __launch_bounds__(SD_THREADS) k_sdag_resolve(LevelIn L, const uint8_t* __restrict__ clsVar, const uint32_t* __restrict__ slotOf,
                                                              const uint32_t* __restrict__ minIdx, uint32_t* __restrict__ rep, uint8_t* __restrict__ flag,
                                                              uint32_t* __restrict__ isRep, uint32_t* __restrict__ flags) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L.n) return;
	uint32_t sl = slotOf[i];
	if (sl == UNSET) { rep[i] = UNSET; flag[i] = 0; isRep[i] = 0; return; }
	uint32_t r = minIdx[sl];
	rep[i] = r;
	if (r == (uint32_t)i) { flag[i] = 0; isRep[i] = 1; return; }
	isRep[i] = 0;
	NodeIn n, nr;
	load_node(L.mask, L.child, L.mirror, i, n);
	load_node(L.mask, L.child, L.mirror, r, nr);
	SKey a, b;
	variant(n, clsVar[i], clsVar[i] ? L.childInv : nullptr, a);
	variant(nr, clsVar[r], clsVar[r] ? L.childInv : nullptr, b);
	if (!key_eq(a, b)) { flags[1] = 1; flag[i] = 0; return; }
	variant(nr, 0, nullptr, b);   // identity key of the representative
	int f = -1;
	for (int q = 0; q < 8; ++q) {
		int s = c_priority[q];
		variant(n, s, s ? L.childInv : nullptr, a);
		if (key_eq(a, b)) { f = s; break; }
	}
	if (f < 0) { flags[2] = 1; f = 0; }   // orbit-min formulation violated (never observed; reported as an error)
	flag[i] = (uint8_t)f;
}

__global__ void __launch_bounds__(256) k_sdag_compact(LevelIn L, const uint32_t* __restrict__ isRep, const uint32_t* __restrict__ newId, const uint8_t* __restrict__ inv,
                                                       uint8_t* __restrict__ omask, uint32_t* __restrict__ ochild, uint8_t* __restrict__ omirror, uint8_t* __restrict__ oinv) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L.n || !isRep[i]) return;
	uint64_t o = newId[i];
	omask[o] = L.mask[i];
	const uint4* c4 = reinterpret_cast<const uint4*>(L.child + i * 8);
	uint4* o4 = reinterpret_cast<uint4*>(ochild + o * 8);
	o4[0] = c4[0]; o4[1] = c4[1];
	omirror[o * 3] = L.mirror[i * 3]; omirror[o * 3 + 1] = L.mirror[i * 3 + 1]; omirror[o * 3 + 2] = L.mirror[i * 3 + 2];
	oinv[o] = inv[i];
}

// parents: child -> new id of its representative, OR in the mirror flags (geom_octree.cpp:667-679)
__global__ void __launch_bounds__(256) k_sdag_parents(uint64_t np, uint32_t* __restrict__ pchild, uint8_t* __restrict__ pmirror,
                                                       const uint32_t* __restrict__ rep, const uint32_t* __restrict__ newId, const uint8_t* __restrict__ flag) {
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= np) return;
	unsigned mx = pmirror[p * 3], my = pmirror[p * 3 + 1], mz = pmirror[p * 3 + 2];
	for (int j = 0; j < 8; ++j) {
		uint32_t c = pchild[p * 8 + j];
		if (c == NULLNODE) continue;
		uint32_t r = rep[c];
		unsigned f = flag[c];
		pchild[p * 8 + j] = (r == UNSET) ? 0u : newId[r];
		if (f & 4) mx |= 1u << j;
		if (f & 2) my |= 1u << j;
		if (f & 1) mz |= 1u << j;
	}
	pmirror[p * 3] = (uint8_t)mx; pmirror[p * 3 + 1] = (uint8_t)my; pmirror[p * 3 + 2] = (uint8_t)mz;
}

}  // namespace

uint64_t to_sdag_device(svb_ctx* c) {
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	const int L = (int)c->levels;
	uint64_t total = 0;
	DevBuf<uint32_t> flags(pool, 4);
	DevBuf<uint64_t> tot(pool, 1);
	for (int lev = L - 1; lev >= 1; --lev) {
		OutLevel& X = c->out[lev];
		const uint64_t n = X.n;
		if (n == 0) continue;
		if (n >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "level too large");
		LevelIn in;
		in.n = n; in.mask = X.mask.p; in.child = X.child.p; in.mirror = X.mirror.p;
		in.childInv = (lev < L - 1) ? c->out[lev + 1].inv.p : nullptr;
		uint64_t cap = 1024;
		while (cap < 2 * n) cap <<= 1;
		DevBuf<uint64_t> tag(pool, cap);
		DevBuf<uint32_t> minIdx(pool, cap);
		DevBuf<uint8_t> inv(pool, n), clsVar(pool, n), flag(pool, n);
		DevBuf<uint32_t> slotOf(pool, n), rep(pool, n), isRep(pool, n), newId(pool, n);
		unsigned nb = blocks_for(n, SD_THREADS);
		uint64_t hU = 0;
		// A level is only replaced once its pass came out clean, so a detected tag collision (two class keys behind one 64-bit
		// tag: the exact resolve pass sees a node that matches no mirror image of "its" representative) simply repeats the
		// level's pass with another seed.
		for (int attempt = 0;; ++attempt) {
			tag.zero();
			minIdx.fill_ff();
			flags.zero();
			k_sdag_class<<<nb, SD_THREADS, 0, s>>>(in, inv.p, clsVar.p, slotOf.p, (unsigned long long*)tag.p, minIdx.p, cap - 1, flags.p, make_hash_seed(c->hashSeed + (uint64_t)attempt));
			SVB_KERNEL_CHECK();
			k_sdag_resolve<<<nb, SD_THREADS, 0, s>>>(in, clsVar.p, slotOf.p, minIdx.p, rep.p, flag.p, isRep.p, flags.p);
			SVB_KERNEL_CHECK();
			scan_u32(s, pool, isRep.p, n, newId.p, tot.p);
			uint32_t hf[4];
			SVB_CUDA(cudaMemcpyAsync(&hU, tot.p, 8, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaMemcpyAsync(hf, flags.p, 16, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaStreamSynchronize(s));
			if (hf[0]) throw Error(SVB_ECUDA, "toSDAG: hash table overflow");
			if (!hf[1] && !hf[2]) break;
			c->nHashRetries++;
			if (attempt >= 4) throw Error(SVB_ECOLLISION, "toSDAG: 64-bit class-key hash collisions under five different seeds");
		}
		OutLevel Y;
		Y.n = hU;
		Y.mask.reset(pool, hU); Y.child.reset(pool, hU * 8); Y.mirror.reset(pool, hU * 3); Y.inv.reset(pool, hU);
		k_sdag_compact<<<blocks_for(n, 256), 256, 0, s>>>(in, isRep.p, newId.p, inv.p, Y.mask.p, Y.child.p, Y.mirror.p, Y.inv.p);
		SVB_KERNEL_CHECK();
		OutLevel& P = c->out[lev - 1];
		k_sdag_parents<<<blocks_for(P.n, 256), 256, 0, s>>>(P.n, P.child.p, P.mirror.p, rep.p, newId.p, flag.p);
		SVB_KERNEL_CHECK();
		SVB_CUDA(cudaStreamSynchronize(s));
		X = std::move(Y);
		total += hU;
	}
	return total;
}

}  // namespace svb
