// svb_attr.cu -- material-id leaves and Gray-coded attribute bit-trees (BASELINE.json configs[3]).
//
// (1) Material ids in the leaves.  The reference has this hook inside the hot path's first stage:
//         GeomOctree::buildSVO(levels, bbox, internalCall, leavesCenters, putMaterialIdInLeaves = true)
//     stores, while voxelizing triangle iTri, `triMatId = _scene->getTriangleMaterialId(iTri)` into the leaf node's child slot
//     of every voxel the triangle touches (src/symvox/geom_octree.cpp:210-211, :252).  Triangles are processed in file order,
//     so a voxel ends up with the material of the LAST triangle that touches it.  Its command line never switches the hook
//     on; oracle/ref_attr_driver.cpp calls it on the unmodified sources, and svb_build_svo_materials() reproduces its leaf
//     level node for node: same node order (the SVO creation order = rank of (first-touch triangle, path')), same masks,
//     same 8 child slots (material id, or Octree::nullNode for an unset voxel).
//     On the GPU: the exact classifier decides every voxel in its own lane (svb_voxelize.cu::k_classify) and keeps, per
//     voxel, the largest triangle index that passed -- an atomicMax instead of the reference's overwrite in file order.
//
// (2) Gray code + bit-trees.  The reference's readme names "an attribute compression enhancement for the bit-tree
//     representation, by encoding attributes as Gray codes instead of a conventional binary encoding" (readme.md:8), with no
//     code behind it.  SELF-SPECIFIED here: the attribute of a voxel is its material id a (or its Gray code
//     g = a ^ (a >> 1)); bit-tree b is the sparse voxel DAG of the voxels whose code has bit b set, reduced bottom-up exactly
//     like the geometry (same per-level dedup kernels); svb_attribute_bit_trees() reports the node and voxel count of every
//     bit-tree.  Codes of neighbouring ids differ in one bit under Gray coding, so the bit-trees share more subtrees.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "svb_context.cuh"
#include "svb_attr.cuh"
#include "svb_dedup.cuh"
#include "svb_voxelize.cuh"

struct svb_attr_state {
	uint32_t L = 0;
	int tbits = 1;
	std::vector<svb::BatchLevel> lv;          // the SVO of the attribute build, levels 0 .. L-1 (Morton order per level)
	svb::DevBuf<uint32_t> lastTri;            // per leaf node and voxel: 1 + last triangle touching it (0 = none)
	svb::DevBuf<uint32_t> triMat;             // per triangle: material id
	svb::DevBuf<uint32_t> order;              // rank in the reference's SVO order -> leaf node (Morton index)
	svb::DevBuf<uint32_t> dSeq, dTileStart;   // the single root tile
	uint64_t n = 0;                           // leaf nodes
};

namespace svb {

namespace {

constexpr int AT_THREADS = 256;

int bits_for(uint64_t maxval) {
	int b = 1;
	while (b < 64 && (maxval >> b)) ++b;
	return b;
}
int kind_of(uint32_t g, uint32_t L) { return g == L - 1 ? KIND_LEAF : (g == L - 2 ? KIND_K64 : KIND_INNER); }

// creation order of the leaf nodes inside the one octree: (first-touch triangle, path with the last digit reversed)
__global__ void __launch_bounds__(AT_THREADS) k_attr_keys(uint64_t n, const uint64_t* __restrict__ code, const uint32_t* __restrict__ tstar, uint32_t tileStart0, int l,
                                                          uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t path = code[i] & ((l >= 21) ? ~0ull : ((1ull << (3 * l)) - 1));
	if (l > 0) path = (path & ~7ull) | (7ull - (path & 7ull));   // children are created 7 -> 0 (geom_octree.cpp:234)
	keys[i] = ((uint64_t)(tstar[i] - tileStart0) << (3 * l)) | path;
	vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(AT_THREADS) k_attr_gather(uint64_t n, const uint32_t* __restrict__ order, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ lastTri,
                                                            const uint32_t* __restrict__ triMat, uint8_t* __restrict__ omask, uint32_t* __restrict__ omat) {
	const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t r = gid >> 3;
	const int v = (int)(gid & 7);
	if (r >= n) return;
	const uint32_t node = order[r];
	const unsigned m = mask[node];
	if (v == 0) omask[r] = (uint8_t)m;
	const uint32_t lt = lastTri[(uint64_t)node * 8 + v];
	omat[r * 8 + v] = ((m >> v) & 1u) && lt ? triMat[lt - 1] : NULLNODE;   // an unset voxel keeps Octree::nullNode in its child slot
}
// voxel mask of bit-tree `bit`: the voxels of the node whose attribute code has that bit set
__global__ void __launch_bounds__(AT_THREADS) k_attr_plane(uint64_t n, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ lastTri, const uint32_t* __restrict__ triMat,
                                                           int bit, int gray, uint8_t* __restrict__ pmask) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned m = mask[i];
	unsigned pm = 0;
#pragma unroll
	for (int v = 0; v < 8; ++v) {
		const uint32_t lt = lastTri[i * 8 + v];
		if (!((m >> v) & 1u) || !lt) continue;
		uint32_t a = triMat[lt - 1];
		if (gray) a ^= a >> 1;   // reflected binary Gray code
		if ((a >> bit) & 1u) pm |= 1u << v;
	}
	pmask[i] = (uint8_t)pm;
}

}  // namespace

void attr_build(svb_ctx* c, uint32_t L, const double bmin[3], const double bmax[3], const uint32_t* triMatHost) {
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	if (L < 2 || L > 16) throw Error(SVB_EINVAL, "attribute build: levels must be in [2,16]");
	if (!c->d_tris || !c->T) throw Error(SVB_EINVAL, "no triangles set");
	if (!triMatHost) throw Error(SVB_EINVAL, "null material array");
	c->attr.reset(new svb_attr_state());
	svb_attr_state& A = *c->attr;
	A.L = L;
	A.tbits = bits_for(c->T - 1);
	if (A.tbits + 3 * (int)(L - 1) > 63) throw Error(SVB_ERANGE, "attribute build: order key exceeds 63 bits");
	A.triMat.reset(pool, c->T);
	SVB_CUDA(cudaMemcpyAsync(A.triMat.p, triMatHost, c->T * 4ull, cudaMemcpyHostToDevice, s));
	// root cube: geom_octree.cpp:177-184 (float-narrowed box, max float side), :214 (double centre)
	float side = 0.f;
	for (int k = 0; k < 3; ++k) {
		const float lo = (float)bmin[k], hi = (float)bmax[k];
		const float sd = ((hi - lo) * 0.5f) * 2.0f;
		if (k == 0 || sd > side) side = sd;
	}
	TileGeom rootG;
	rootG.cx = (bmin[0] + bmax[0]) * 0.5; rootG.cy = (bmin[1] + bmax[1]) * 0.5; rootG.cz = (bmin[2] + bmax[2]) * 0.5;
	rootG.rootSide = (double)side;
	TileGridHost grid1;
	grid1.G = 1; grid1.cell = side > 0 ? (double)side : 1.0;
	grid1.ox = rootG.cx - side * 0.5; grid1.oy = rootG.cy - side * 0.5; grid1.oz = rootG.cz - side * 0.5;
	DevBuf<int> dGrid1(pool, 1), dLocal1(pool, 1);
	dGrid1.zero(); dLocal1.zero();
	DevBuf<TileGeom> dTiles(pool, 1);
	SVB_CUDA(cudaMemcpyAsync(dTiles.p, &rootG, sizeof(TileGeom), cudaMemcpyHostToDevice, s));
	A.dSeq.reset(pool, 1); A.dSeq.zero();
	DevBuf<uint32_t> ptri, pnode, rootTri;
	uint64_t P = 0, pairs = 0;
	make_root_pairs(s, pool, c->d_tris, c->T, grid1, dGrid1.p, dLocal1.p, 0, 1, ptri, pnode, rootTri, A.dTileStart, P);
	DevBuf<uint64_t> dExact(pool, 1);
	dExact.zero();
	const uint64_t budget = pool.live + (uint64_t)(0.80 * (double)pool.headroom());
	try {
		voxelize_batch(s, pool, c->d_tris, dTiles.p, 1, (int)L, ptri, pnode, rootTri.p, A.dTileStart.p, P, budget, 0, A.lv, pairs, dExact.p,
		               centre_chain_exact(rootG, (int)L), false, 0, nullptr, &A.lastTri);
	} catch (const BatchTooBig&) {
		throw Error(SVB_ENOMEM, "attribute build: the octree does not fit device memory in one piece");
	}
	// the reference's leaf order
	BatchLevel& X = A.lv[L - 1];
	A.n = X.n;
	A.order.reset(pool, X.n ? X.n : 1);
	if (X.n) {
		if (X.n >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "attribute build: too many leaf nodes");
		uint32_t ts0 = 0;
		SVB_CUDA(cudaMemcpyAsync(&ts0, A.dTileStart.p, 4, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		DevBuf<uint64_t> keys(pool, X.n);
		k_attr_keys<<<blocks_for(X.n, AT_THREADS), AT_THREADS, 0, s>>>(X.n, X.code.p, X.tstar.p, ts0, (int)(L - 1), keys.p, A.order.p);
		SVB_KERNEL_CHECK();
		radix_sort_pairs(s, pool, keys.p, A.order.p, X.n, A.tbits + 3 * (int)(L - 1));
	}
	SVB_CUDA(cudaStreamSynchronize(s));
}

uint64_t attr_leaf_count(const svb_ctx* c) { return c->attr ? c->attr->n : 0; }

void attr_download(svb_ctx* c, uint8_t* maskHost, uint32_t* mat8Host) {
	if (!c->attr) throw Error(SVB_EINVAL, "no attribute build in this context");
	svb_attr_state& A = *c->attr;
	cudaStream_t s = c->stream;
	if (!A.n) return;
	DevBuf<uint8_t> omask(c->pool, A.n);
	DevBuf<uint32_t> omat(c->pool, A.n * 8);
	k_attr_gather<<<blocks_for(A.n * 8, AT_THREADS), AT_THREADS, 0, s>>>(A.n, A.order.p, A.lv[A.L - 1].mask.p, A.lastTri.p, A.triMat.p, omask.p, omat.p);
	SVB_KERNEL_CHECK();
	if (maskHost) SVB_CUDA(cudaMemcpyAsync(maskHost, omask.p, A.n, cudaMemcpyDeviceToHost, s));
	if (mat8Host) SVB_CUDA(cudaMemcpyAsync(mat8Host, omat.p, A.n * 32, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
}

// node / voxel count of the bit-tree of every attribute bit: the leaf masks are restricted to the voxels whose code has the bit
// set and the octree is reduced bottom-up by the geometry's own dedup passes (empty subtrees vanish: svb_dedup.cu::build_key)
void attr_bit_trees(svb_ctx* c, uint32_t nbits, int gray, uint64_t* nodes, uint64_t* voxels) {
	if (!c->attr) throw Error(SVB_EINVAL, "no attribute build in this context");
	if (nbits == 0 || nbits > 32) throw Error(SVB_EINVAL, "nbits must be in [1,32]");
	svb_attr_state& A = *c->attr;
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	const uint32_t L = A.L;
	const std::vector<uint32_t> hseq(1, 0);
	for (uint32_t b = 0; b < nbits; ++b) {
		nodes[b] = 0;
		if (voxels) voxels[b] = 0;
		if (!A.n) continue;
		BatchLevel& X = A.lv[L - 1];
		DevBuf<uint8_t> pmask(pool, (X.n + 3 + 16) & ~3ull);   // padded like a level's mask array (the 4^3 key builder reads aligned words)
		pmask.zero();
		k_attr_plane<<<blocks_for(X.n, AT_THREADS), AT_THREADS, 0, s>>>(X.n, X.mask.p, A.lastTri.p, A.triMat.p, (int)b, gray, pmask.p);
		SVB_KERNEL_CHECK();
		std::vector<LevelTable> tables(L);
		for (uint32_t g = 1; g < L; ++g) table_init(s, pool, tables[g], kind_of(g, L), c->hashSeed);
		DevBuf<uint64_t> dVox(pool, 1);
		dVox.zero();
		std::vector<DevBuf<uint32_t>> refs(L);
		for (int l = (int)L - 1; l >= 1; --l) {
			BatchLevel& Y = A.lv[l];
			DedupArgs a;
			a.N = Y.n; a.code = Y.code.p; a.tstar = Y.tstar.p; a.childBase = Y.childBase.p;
			a.mask = (l == (int)L - 1) ? pmask.p : Y.mask.p;
			a.l = l; a.tbits = A.tbits; a.tileSeq = A.dSeq.p; a.tileStart = A.dTileStart.p;
			a.seqLo = a.seqHi = 0; a.seqMonotone = true;
			LevelTable& T = tables[l];
			if (T.kind == KIND_LEAF) { dedup_leaf(s, pool, T, a, dVox.p); continue; }
			const bool leafBelow = kind_of((uint32_t)l + 1, L) == KIND_LEAF;
			a.childMode = leafBelow ? CH_MASK_U8 : CH_UID_U32;
			a.childRefs = leafBelow ? (const void*)pmask.p : (const void*)refs[l + 1].p;
			refs[l].reset(pool, Y.n ? Y.n : 1);
			a.ref = refs[l].p;
			dedup_level(s, pool, T, a);
		}
		uint64_t hv = 0;
		SVB_CUDA(cudaMemcpyAsync(&hv, dVox.p, 8, cudaMemcpyDeviceToHost, s));
		SVB_CUDA(cudaStreamSynchronize(s));
		uint64_t nn = 0;
		for (uint32_t g = 1; g < L; ++g) nn += tables[g].kind == KIND_LEAF ? tables[g].known : tables[g].count;
		nodes[b] = hv ? nn + 1 : 0;   // + the root (level 0 is never reduced, geom_octree.cpp:483); an empty bit-tree has no nodes
		if (voxels) voxels[b] = hv;
	}
}

}  // namespace svb
