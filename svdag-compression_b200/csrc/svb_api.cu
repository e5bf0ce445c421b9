// svb_api.cu -- context, build orchestration and the extern "C" boundary of libsvb.so.
//
// Host-side control flow mirrors what src/svbuilder/main.cpp:147-203 asks of GeomOctree:
//   step == 0 : buildSVO(levels) + toDAG()                    (geom_octree.cpp:171-280, 462-548)
//   step  > 0 : buildDAG(levels, step)                         (geom_octree.cpp:289-435)
// but every per-voxel / per-node operation runs in the CUDA kernels of svb_voxelize.cu,
// svb_dedup.cu, svb_sdag.cu and svb_prims.cu.  The host only sizes buffers, enumerates the
// sub-octree ("tile") list of step mode from the (tiny) base octree, and sequences launches.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "svb_context.cuh"
#include "svb_attr.cuh"
#include "svb_cross.cuh"
#include "host/octree_data.hpp"
#include "svb_dedup.cuh"
#include "svb_encode.cuh"
#include "svb_sdag.cuh"
#include "svb_voxelize.cuh"

using namespace svb;

namespace {

int bits_for(uint64_t maxval) {   // bits needed to store values in [0, maxval]
	int b = 1;
	while (b < 64 && (maxval >> b)) ++b;
	return b;
}

__global__ void k_gather_u32(uint64_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_scatter_u32(uint64_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[idx[i]] = src[i];
}
__global__ void k_scatter_u8(uint64_t n, const uint32_t* __restrict__ idx, const uint8_t* __restrict__ src, uint32_t* __restrict__ dst) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[idx[i]] = src[i];
}
// sub-octree roots after the merge: local uid -> global uid (entries of other ranks stay UNSET)
__global__ void k_remap_refs(uint64_t n, uint32_t* __restrict__ refs, const uint32_t* __restrict__ l2g) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && refs[i] != UNSET) refs[i] = l2g[refs[i]];
}
__global__ void k_min_u32_rows(uint64_t n, uint32_t rows, const uint32_t* __restrict__ all, uint32_t* __restrict__ out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t m = UNSET;
	for (uint32_t r = 0; r < rows; ++r) { uint32_t v = all[(uint64_t)r * n + i]; if (v < m) m = v; }
	out[i] = m;
}

static cudaEvent_t prof_event(svb_ctx* c) {
	if (!c->evPool.empty()) { cudaEvent_t e = c->evPool.back(); c->evPool.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}
static bool prof_wanted(const svb_ctx* c, const char* name) { return c->profiling && (!c->profEmitOnly || strcmp(name, "emit") == 0); }

struct ProfScope {
	svb_ctx* c;
	int idx = -1;
	ProfScope(svb_ctx* ctx, const char* name, uint32_t level, uint64_t n_in) : c(ctx) {
		if (!prof_wanted(c, name)) return;
		svb_ctx::PendingProf p;
		memset(&p.rec, 0, sizeof(p.rec));
		snprintf(p.rec.name, sizeof(p.rec.name), "%s", name);
		p.rec.level = level;
		p.rec.n_in = n_in;
		p.e0 = prof_event(c);
		p.e1 = prof_event(c);
		cudaEventRecord(p.e0, c->stream);
		c->pending.push_back(p);
		idx = (int)c->pending.size() - 1;
	}
	void done(uint64_t n_out, double bytes) {
		if (idx < 0) return;
		cudaEventRecord(c->pending[idx].e1, c->stream);
		c->pending[idx].closed = true;
		c->pending[idx].rec.n_out = n_out;
		c->pending[idx].rec.bytes = bytes;
	}
};

struct CtxProfHook : ProfHook {
	svb_ctx* c;
	explicit CtxProfHook(svb_ctx* ctx) : c(ctx) {}
	int begin(const char* name, uint32_t level, uint64_t n_in) override {
		if (!prof_wanted(c, name)) return -1;
		svb_ctx::PendingProf p;
		memset(&p.rec, 0, sizeof(p.rec));
		snprintf(p.rec.name, sizeof(p.rec.name), "%s", name);
		p.rec.level = level;
		p.rec.n_in = n_in;
		p.e0 = prof_event(c);
		p.e1 = prof_event(c);
		cudaEventRecord(p.e0, c->stream);
		c->pending.push_back(p);
		return (int)c->pending.size() - 1;
	}
	void end(int id, uint64_t n_out, double bytes) override {
		if (id < 0) return;
		cudaEventRecord(c->pending[id].e1, c->stream);
		c->pending[id].closed = true;
		c->pending[id].rec.n_out = n_out;
		c->pending[id].rec.bytes = bytes;
	}
};

void resolve_profile(svb_ctx* c) {
	for (auto& p : c->pending) {
		if (p.closed) {   // scopes left by an exception (e.g. a batch that had to be split) never recorded e1
			float ms = 0;
			cudaEventSynchronize(p.e1);
			cudaEventElapsedTime(&ms, p.e0, p.e1);
			p.rec.ms = ms;
			c->prof.push_back(p.rec);
		}
		c->evPool.push_back(p.e0);
		c->evPool.push_back(p.e1);
	}
	c->pending.clear();
}

struct StageTimer {
	cudaEvent_t e0, e1;
	cudaStream_t s;
	StageTimer(cudaStream_t st) : s(st) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }
	double stop() {
		cudaEventRecord(e1, s);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		return ms;
	}
	~StageTimer() { cudaEventDestroy(e0); cudaEventDestroy(e1); }
};

// geom_octree.cpp:177-184: the bbox is narrowed to float, the root side is the max float side
void float_box(const double lo[3], const double hi[3], float bboxF[6], double& rootSide) {
	float s = 0.f;
	for (int k = 0; k < 3; ++k) {
		bboxF[k] = (float)lo[k];
		bboxF[3 + k] = (float)hi[k];
		float side = ((bboxF[3 + k] - bboxF[k]) * 0.5f) * 2.0f;
		if (k == 0 || side > s) s = side;
	}
	rootSide = (double)s;
}

int kind_of(uint32_t g, uint32_t L) { return g == L - 1 ? KIND_LEAF : (g == L - 2 ? KIND_K64 : KIND_INNER); }

struct TileHost {
	TileGeom g;
	uint32_t baseNode;   // index (Morton order) of the base leaf node
	int j;               // child slot inside it
	int ix, iy, iz;
};

template <class T>
std::vector<T> download(cudaStream_t s, const T* d, uint64_t n) {
	std::vector<T> h(n);
	if (n) SVB_CUDA(cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	return h;
}
template <class T>
void upload(cudaStream_t s, Pool& pool, DevBuf<T>& d, const std::vector<T>& h) {
	d.reset(pool, h.size() ? h.size() : 1);
	if (!h.empty()) SVB_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
	SVB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

// Everything a build keeps between its phases (local reduction -> [multi-GPU merge] -> finish).
struct svb_build_state {
	uint32_t L = 0, step = 0, s1 = 0;   // s1 = step + 1 = global level of the sub-octree roots (0: no sub-octrees)
	uint32_t rank = 0, world = 1;
	int tbits = 1, tileBits = 1;   // tbits: bits of a triangle rank inside the root tile (all triangles)
	int tbLocal = 1;               // bits of a triangle rank inside a sub-octree (max candidates per tile)
	bool allFlat = false;          // every triangle is flat (box mesh): slow-stream kernel variant without the general filter
	std::vector<LevelTable> tables;      // per global level
	std::vector<int> obits;
	DevBuf<uint64_t> dVoxels, dExact;
	DevBuf<uint32_t> rootKey;
	int rootChildMode = CH_UID_U32;
	uint64_t nNodesSVO = 0, nLastLevSVO = 0, pairs = 0, nBatches = 0;
	uint64_t baseNodesSVO = 0, baseVoxels = 0, basePairs = 0, baseExact = 0;   // base octree: built by every rank, counted once
	std::vector<uint64_t> svoCounts;
	double msVox = 0, msDedup = 0;
	// geometry
	float bboxF[6];
	double rootSide = 0;
	TileGeom rootG;
	TileGridHost grid1;
	DevBuf<int> dGrid1, dLocal1;         // single root tile: grid cell -> tile 0, tile 0 -> position 0
	DevBuf<int> dSelPos;                 // global tile_seq -> position in this rank's selection (-1: not ours); a batch is a range of positions
	DevBuf<uint32_t> dSeq1;
	// step mode
	std::vector<BatchLevel> base;        // base octree levels 0..step (reduced last)
	std::vector<uint8_t> baseLeafMask;
	std::vector<TileHost> tiles;         // all sub-octrees, in the reference's order
	TileGridHost grid;
	DevBuf<int> dGrid;
	DevBuf<uint32_t> tileRootRef;        // per tile: uid of its (reduced) root at level s1; UNSET if not built here
	std::vector<DevBuf<uint32_t>> l2g;   // multi-GPU merge: local uid -> global uid per level
	std::vector<bool> merged;
	RootPairsAll rootsAll;               // root pairs of all of this rank's sub-octrees, binned once (valid only during the local phase)
	DevBuf<uint32_t> mergeStatus;        // multi-GPU merge: 8 x u32 per level, written by merge_import, read back once by build_finish
	bool finished = false;
	uint64_t launches0 = 0;
	cudaEvent_t evStart = nullptr;
};

namespace {

typedef svb_build_state BuildState;

// Order keys (tile_seq | t* | path') must fit 63 bits at every level that is reduced through a 64-bit table.  The
// deepest level (3*(Lt-1) path bits) is the voxel-mask level, whose 256-entry table can hold a two-word key: if only
// that level overflows, it runs in wide mode (two passes, svb_dedup.cu::dedup_leaf); e.g. 64K^3 with step 7 and 12 M
// triangles needs 21 + 24 + 3*7 = 66 bits at the leaves but 63 above.
void set_order_key_width(BuildState& B, int Lt, int tb) {
	const int leafBits = B.tileBits + tb + 3 * (Lt - 1);
	const int innerBits = B.tileBits + tb + 3 * std::max(Lt - 2, 0);
	const char* force = getenv("SVB_WIDE_LEAF");
	bool wide = leafBits > 63 || (force && force[0] == '1');
	if ((wide ? innerBits : leafBits) > 63 || B.tileBits + tb > 63)
		throw Error(SVB_ERANGE, "order key exceeds 63 bits (" + std::to_string(B.tileBits) + " tile + " + std::to_string(tb) + " triangle-rank + " +
		                            std::to_string(3 * std::max(Lt - 2, 0)) + " path bits): increase step (fewer levels per sub-octree)");
	B.tables[B.L - 1].wide = wide;
}

// bottom-up reduction of one batch's levels [lo_l, Lt-1] into the tables
void dedup_batch(svb_ctx* c, BuildState& B, std::vector<BatchLevel>& lv, int Lt, uint32_t gbase, const uint32_t* d_tileSeq, const uint32_t* d_tileStart, int lo_l,
                 const std::vector<uint32_t>& hseq, int hi_l = -1, const LeafQuery* query = nullptr) {
	if (hi_l < 0) hi_l = Lt - 1;   // (Lt - 2: the leaf level has been reduced already, see run_tile_batch)
	const uint32_t seqLo = hseq.empty() ? 0 : *std::min_element(hseq.begin(), hseq.end()), seqHi = hseq.empty() ? 0 : *std::max_element(hseq.begin(), hseq.end());
	const bool seqMono = std::is_sorted(hseq.begin(), hseq.end());
	const int tb = gbase == 0 ? B.tbits : B.tbLocal;
	for (int l = hi_l; l >= lo_l; --l) {
		uint32_t g = gbase + l;
		BatchLevel& X = lv[l];
		DedupArgs a;
		a.N = X.n; a.code = X.code.p; a.tstar = X.tstar.p; a.mask = X.mask.p; a.childBase = X.childBase.p;
		a.l = l; a.tbits = tb; a.tileSeq = d_tileSeq; a.tileStart = d_tileStart;
		a.seqLo = seqLo; a.seqHi = seqHi; a.seqMonotone = seqMono;
		a.query = query;
		int ob = B.tileBits + tb + 3 * l;
		if (ob > B.obits[g]) B.obits[g] = ob;
		LevelTable& T = B.tables[g];
		if (T.kind == KIND_LEAF) {
			ProfScope ps(c, "dedup_leaf", g, X.n);
			dedup_leaf(c->stream, c->pool, T, a, B.dVoxels.p);
			ps.done(0, 37.0 * (double)X.n);
		} else {
			bool leafBelow = (kind_of(g + 1, B.L) == KIND_LEAF);
			a.childMode = leafBelow ? CH_MASK_U8 : CH_UID_U32;
			a.childRefs = leafBelow ? (const void*)lv[l + 1].mask.p : (const void*)lv[l + 1].ref.p;
			X.ref.reset(c->pool, X.n);
			a.ref = X.ref.p;
			uint64_t before = T.count;
			ProfScope ps(c, T.kind == KIND_K64 ? "dedup_k64" : "dedup_inner", g, X.n);
			dedup_level(c->stream, c->pool, T, a);
			ps.done(T.count, 37.0 * (double)X.n + 33.0 * (double)(T.count - before));
			lv[l + 1] = BatchLevel();   // the level below is no longer needed
		}
	}
}

// Voxelizes and reduces the sub-octrees sel[a..b) (indices into `tiles`, ascending).  Throws BatchTooBig
// (before anything was reduced) when the batch must be cut.
void run_tile_batch(svb_ctx* c, BuildState& B, const std::vector<TileHost>& tiles, const std::vector<uint32_t>& sel, uint32_t a, uint32_t b,
                    const TileGridHost& grid, const int* d_gridTile, const int* d_selPos, int Lt, uint32_t gbase, uint64_t budget, uint64_t nodeCap,
                    std::vector<BatchLevel>* keepLevels) {
	cudaStream_t s = c->stream;
	const uint32_t nt = b - a;
	std::vector<TileGeom> hg(nt);
	std::vector<uint32_t> hseq(nt);
	for (uint32_t i = 0; i < nt; ++i) { hg[i] = tiles[sel[a + i]].g; hseq[i] = sel[a + i]; }
	DevBuf<TileGeom> dTiles;
	DevBuf<uint32_t> dSeq;
	upload(s, c->pool, dTiles, hg);
	upload(s, c->pool, dSeq, hseq);

	std::vector<BatchLevel> lv;
	DevBuf<uint32_t> dTileStart;   // per tile of the batch: first root-pair index (needed again by the order keys)
	uint64_t pairs = 0;
	// First touches of the leaf level are only needed for voxel masks the leaf table has no entry for.  A batch that comes
	// after everything that table has seen is voxelized without them (no atomics / stores / 4 B per leaf node); if it does
	// meet an unknown mask (dedup_leaf_known fails, nothing reduced yet) it is voxelized again, this time with them.
	const uint32_t seqLo = *std::min_element(hseq.begin(), hseq.end());
	const bool leafLevelBatch = gbase != 0 && !keepLevels && Lt >= 2 && B.tables[gbase + Lt - 1].kind == KIND_LEAF;
	bool leafT = leafLevelBatch ? leaf_tstar_needed(B.tables[gbase + Lt - 1], seqLo) : true;
	// ... and the same one level up: on the 4^3 level only the nodes of entries that are NEW to its table need a first touch
	// (a few thousand per batch); they are queried directly (svb_dedup.cu::k_k64_query).  SVB_K64_NOTSTAR=0 keeps tracking it.
	auto k64_untracked_ok = [&] {
		if (leafT || Lt < 3) return false;
		const char* e = getenv("SVB_K64_NOTSTAR");
		if (e && e[0] == '0') return false;
		return k64_tstar_optional(B.tables[gbase + Lt - 2], seqLo);
	};
	bool k64U = k64_untracked_ok();
	bool leafDone = false;
	LeafQuery lqBatch;
	DevBuf<uint32_t> rootTriOwn;   // root pair -> triangle (outlives the voxelizer: the leaf query below reads it)
	const uint32_t* rootTri = nullptr;
	const bool rootsOnce = gbase != 0 && B.rootsAll.valid;
	uint64_t P = 0;
	DevBuf<uint64_t> exactBefore(c->pool, 1);
	SVB_CUDA(cudaMemcpyAsync(exactBefore.p, B.dExact.p, 8, cudaMemcpyDeviceToDevice, s));
	for (;;) {
	{
		StageTimer tm(s);
		ProfScope ps(c, "voxelize", gbase, nt);
		DevBuf<uint32_t> ptri, pnode;
		int cellLo[3] = {1 << 30, 1 << 30, 1 << 30}, cellHi[3] = {-1, -1, -1};
		for (uint32_t i = 0; i < nt; ++i) {
			const TileHost& th = tiles[sel[a + i]];
			cellLo[0] = std::min(cellLo[0], th.ix); cellLo[1] = std::min(cellLo[1], th.iy); cellLo[2] = std::min(cellLo[2], th.iz);
			cellHi[0] = std::max(cellHi[0], th.ix); cellHi[1] = std::max(cellHi[1], th.iy); cellHi[2] = std::max(cellHi[2], th.iz);
		}
		if (rootsOnce) batch_root_pairs(s, c->pool, B.rootsAll, a, nt, ptri, pnode, rootTri, dTileStart, P);
		else {
			make_root_pairs(s, c->pool, c->d_tris, c->T, grid, d_gridTile, d_selPos, a, nt, ptri, pnode, rootTriOwn, dTileStart, P, cellLo, cellHi);
			rootTri = rootTriOwn.p;
		}
		CtxProfHook hook(c);
		bool direct = true;
		for (uint32_t i = 0; i < nt && direct; ++i) direct = centre_chain_exact(hg[i], Lt);
		try {
			voxelize_batch(s, c->pool, c->d_tris, dTiles.p, nt, Lt, ptri, pnode, rootTri, dTileStart.p, P, budget, nodeCap, lv, pairs, B.dExact.p, direct, B.allFlat, leafT ? 0 : (k64U ? 2 : 1), c->profiling ? &hook : nullptr);
		} catch (const BatchTooBig&) {   // the range is cut and voxelized again: the aborted attempt's exact-test count must not stay in the counter
			SVB_CUDA(cudaMemcpyAsync(B.dExact.p, exactBefore.p, 8, cudaMemcpyDeviceToDevice, s));
			SVB_CUDA(cudaStreamSynchronize(s));
			throw;
		}
		ps.done(pairs, 36.0 * (double)c->T + 9.0 * (double)lv[Lt - 1].n);
		B.msVox += tm.stop();
	}
	if (leafT) break;
	{
		StageTimer tm(s);
		const uint32_t g = gbase + Lt - 1;
		BatchLevel& X = lv[Lt - 1];
		DedupArgs da;
		da.N = X.n; da.mask = X.mask.p; da.code = X.code.p; da.l = Lt - 1;
		da.tbits = B.tbLocal; da.tileSeq = dSeq.p; da.tileStart = dTileStart.p;
		LeafQuery lq;
		lq.tris = c->d_tris; lq.rootTri = rootTri; lq.P = P; lq.tiles = dTiles.p;
		lqBatch = lq;
		da.seqLo = seqLo; da.seqHi = *std::max_element(hseq.begin(), hseq.end()); da.seqMonotone = std::is_sorted(hseq.begin(), hseq.end());
		ProfScope ps(c, "dedup_leaf", g, X.n);
		leafDone = dedup_leaf_known(s, c->pool, B.tables[g], da, lq, B.dVoxels.p);
		ps.done(0, 37.0 * (double)X.n);
		B.msDedup += tm.stop();
		if (leafDone) {
			const int ob = B.tileBits + B.tbLocal + 3 * (Lt - 1);
			if (ob > B.obits[g]) B.obits[g] = ob;
			break;
		}
	}
	// an unknown voxel mask: once more, with first touches
	if (getenv("SVB_VX_STATS")) fprintf(stderr, "[vx-stats] batch [%u,%u): a voxel mask without an entry -- voxelizing again with leaf first touches\n", a, b);
	leafT = true;
	k64U = false;
	SVB_CUDA(cudaMemcpyAsync(B.dExact.p, exactBefore.p, 8, cudaMemcpyDeviceToDevice, s));
	lv.clear();
	}
	B.pairs += pairs;
	B.nBatches++;
	for (int l = 1; l < Lt; ++l) B.nNodesSVO += lv[l].n;
	B.nLastLevSVO += lv[Lt - 1].n;
	if (gbase == 0) { B.svoCounts.assign(Lt, 0); for (int l = 0; l < Lt; ++l) B.svoCounts[l] = lv[l].n; }
	if (keepLevels) {   // base octree of step mode: reduced last, on top of the sub-octree roots
		*keepLevels = std::move(lv);
		return;
	}
	StageTimer tm(s);
	if (gbase == 0) {
		dedup_batch(c, B, lv, Lt, 0, dSeq.p, dTileStart.p, 1, hseq);
		DedupArgs r;
		r.N = 1; r.code = lv[0].code.p; r.tstar = lv[0].tstar.p; r.mask = lv[0].mask.p; r.childBase = lv[0].childBase.p;
		bool leafBelow = (kind_of(1, B.L) == KIND_LEAF);
		r.childMode = leafBelow ? CH_MASK_U8 : CH_UID_U32;
		r.childRefs = leafBelow ? (const void*)lv[1].mask.p : (const void*)lv[1].ref.p;
		B.rootChildMode = r.childMode;
		root_key(s, r, B.rootKey.p);
	} else {
		dedup_batch(c, B, lv, Lt, gbase, dSeq.p, dTileStart.p, 0, hseq, leafDone ? Lt - 2 : Lt - 1, (leafDone && k64U) ? &lqBatch : nullptr);
		// remember what each sub-octree root was reduced to (uid, or the voxel mask for 1-level sub-octrees)
		if (B.tables[gbase].kind == KIND_LEAF) k_scatter_u8<<<blocks_for(nt, 256), 256, 0, s>>>(nt, dSeq.p, lv[0].mask.p, B.tileRootRef.p);
		else k_scatter_u32<<<blocks_for(nt, 256), 256, 0, s>>>(nt, dSeq.p, lv[0].ref.p, B.tileRootRef.p);
		SVB_KERNEL_CHECK();
	}
	B.msDedup += tm.stop();
}

// Batching: start with all of this rank's sub-octrees in one batch.  When the voxelizer predicts (or
// hits) an overflow it throws BatchTooBig before anything was reduced, carrying per-tile node counts
// of the level it reached; the range is then cut by cumulative weight into pieces that should fit.
// Batches are processed in ascending sub-octree order (order keys of later batches are larger, which
// the dedup tables rely on when they freeze an entry's representative).
void run_tiles_split(svb_ctx* c, BuildState& B, const std::vector<uint32_t>& sel, int Lt, uint32_t gbase, uint64_t budget, uint64_t nodeCap) {
	uint32_t a = 0;
	const uint32_t b = (uint32_t)sel.size();
	std::vector<uint32_t> plan(1, b);   // upcoming batch ends, ascending
	// The first batches of a share run the expensive variants (first touches tracked on every level, nothing frozen in the
	// tables yet); two small warm-up batches let the bulk of the share run as "later" batches.  Matters most when the share is
	// small, i.e. with many GPUs: a rank of an 8-GPU build has 3 - 4 batches in all.
	{
		const char* e = getenv("SVB_WARMUP_DIV");
		const uint32_t div = e ? (uint32_t)atoi(e) : 48u;
		if (div >= 2 && b >= 4 * div) { plan.insert(plan.begin(), b / div + b / (2 * div)); plan.insert(plan.begin(), b / (2 * div)); }
	}
	// A piece that comes out of a weight-based cut is not second-guessed by the voxelizer's own early prediction (nodeCap = 0
	// switches it off; the hard memory check stays): that prediction clamps the growth per level at 5, so one level further
	// down it over-estimates the leaf level by ~1.6x and used to cut EVERY piece of the 16K^3 city in two again -- nine aborted
	// attempts of five levels each (~30 ms of 570) and twice the batches (SVB_TRUST_CUTS=0: as before).
	const bool trustCuts = [] { const char* e = getenv("SVB_TRUST_CUTS"); return !(e && e[0] == '0'); }();
	std::vector<char> trusted(plan.size(), 0);
	while (a < b) {
		uint32_t e = plan.front();
		try {
			run_tile_batch(c, B, B.tiles, sel, a, e, B.grid, B.dGrid.p, B.dSelPos.p, Lt, gbase, budget, (trustCuts && trusted.front()) ? 0 : nodeCap, nullptr);
			a = e;
			plan.erase(plan.begin());
			trusted.erase(trusted.begin());
		} catch (const BatchTooBig& x) {
			if (getenv("SVB_VX_STATS")) fprintf(stderr, "[vx-stats] batch [%u,%u) cut at level %d (%d levels to go, growth %.2f)\n", a, e, x.level, x.remaining, x.growth);
			if (e - a <= 1) throw Error(SVB_ENOMEM, "a single sub-octree does not fit the batch budget; use a larger step");
			std::vector<uint32_t> cuts;
			if (x.weight.size() == e - a) {
				double scale = 1.0;
				for (int r = 0; r < x.remaining; ++r) scale *= x.growth;
				double total = 0;
				for (uint32_t w : x.weight) total += (double)w * scale;
				double cap = 0.7 * (double)nodeCap;
				uint32_t pieces = (uint32_t)std::max(2.0, std::ceil(total / cap));
				double per = total / pieces, acc = 0;
				for (uint32_t i = 0; i + 1 < e - a; ++i) {
					acc += (double)x.weight[i] * scale;
					if (acc >= per) { cuts.push_back(a + i + 1); acc = 0; }
				}
			}
			const bool byWeight = !cuts.empty();
			if (cuts.empty()) cuts.push_back(a + (e - a) / 2);
			plan.insert(plan.begin(), cuts.begin(), cuts.end());
			trusted.front() = byWeight ? 1 : 0;   // (the piece that ends at e)
			trusted.insert(trusted.begin(), cuts.size(), byWeight ? 1 : 0);
		}
	}
}

uint64_t batch_budget(svb_ctx* c) {
	// transient budget of one tile batch: a fraction of what the slab allocator can still hand out
	return c->batchBudget ? c->pool.live + c->batchBudget : c->pool.live + (uint64_t)(0.80 * (double)c->pool.headroom());
}

// ---- phase 1: voxelize + reduce this rank's share into rank-local tables
void build_local(svb_ctx* c, uint32_t L, uint32_t step, const double bmin[3], const double bmax[3], uint32_t rank, uint32_t world) {
	cudaStream_t s = c->stream;
	if (L < 2 || L > 21) throw Error(SVB_EINVAL, "levels must be in [2,21]");
	if (step > 0 && step + 1 >= L) throw Error(SVB_EINVAL, "step + 1 must be < levels");
	if (!c->d_tris && c->T) throw Error(SVB_EINVAL, "no triangles set");
	if (world == 0 || rank >= world) throw Error(SVB_EINVAL, "bad rank/world");
	if (world > 1 && step == 0) throw Error(SVB_EINVAL, "a sharded build needs step > 0 (sub-octrees are the unit of distribution)");
	c->out.clear();
	c->attr.reset();
	c->state = SVB_S_EMPTY;
	if (!c->profAccumulate) c->prof.clear();
	c->lastImageKind = -1;
	memset(&c->stats, 0, sizeof(c->stats));
	c->build.reset(new BuildState());
	BuildState& B = *c->build;
	B.launches0 = g_launches.load();
	cudaEventCreate(&B.evStart);
	cudaEventRecord(B.evStart, s);
	B.L = L; B.step = step; B.rank = rank; B.world = world;
	B.tbits = bits_for(c->T ? c->T - 1 : 0);
	B.tables.resize(L);
	B.obits.assign(L, 1);
	B.l2g.resize(L);
	B.merged.assign(L, false);
	if (world > 1) { B.mergeStatus.reset(c->pool, 8ull * L); B.mergeStatus.zero(); }
	for (uint32_t g = 1; g < L; ++g) table_init(s, c->pool, B.tables[g], kind_of(g, L), c->hashSeed);
	B.dVoxels.reset(c->pool, 1); B.dVoxels.zero();
	B.dExact.reset(c->pool, 1); B.dExact.zero();
	B.rootKey.reset(c->pool, 8); B.rootKey.fill_ff();

	B.allFlat = !getenv("SVB_NO_FLATONLY") && all_triangles_flat(s, c->pool, c->d_tris, c->T);
	float_box(bmin, bmax, B.bboxF, B.rootSide);
	const double rootSide = B.rootSide;
	B.rootG.cx = (bmin[0] + bmax[0]) * 0.5; B.rootG.cy = (bmin[1] + bmax[1]) * 0.5; B.rootG.cz = (bmin[2] + bmax[2]) * 0.5;   // bbox.center(), :214
	B.rootG.rootSide = rootSide;
	B.grid1.G = 1; B.grid1.cell = rootSide > 0 ? rootSide : 1.0;
	B.grid1.ox = B.rootG.cx - rootSide * 0.5; B.grid1.oy = B.rootG.cy - rootSide * 0.5; B.grid1.oz = B.rootG.cz - rootSide * 0.5;
	B.dGrid1.reset(c->pool, 1); B.dGrid1.zero();
	B.dLocal1.reset(c->pool, 1); B.dLocal1.zero();
	std::vector<TileHost> rootTile(1);
	rootTile[0].g = B.rootG; rootTile[0].baseNode = 0; rootTile[0].j = 0; rootTile[0].ix = rootTile[0].iy = rootTile[0].iz = 0;
	const std::vector<uint32_t> sel0(1, 0);
	const uint64_t budget = batch_budget(c);

	if (step == 0) {
		B.s1 = 0;
		B.tileBits = 1;
		set_order_key_width(B, (int)L, B.tbits);
		try {
			run_tile_batch(c, B, rootTile, sel0, 0, 1, B.grid1, B.dGrid1.p, B.dLocal1.p, (int)L, 0, budget, 0, nullptr);
		} catch (const BatchTooBig&) {
			throw Error(SVB_ENOMEM, "octree does not fit device memory in one piece; use step > 0");
		}
		return;
	}
	const uint32_t s1 = step + 1;
	B.s1 = s1;
	// ---- base octree (levels 0..step) over all triangles, exact hierarchical tests (every rank, identical)
	try {
		run_tile_batch(c, B, rootTile, sel0, 0, 1, B.grid1, B.dGrid1.p, B.dLocal1.p, (int)s1, 0, budget, 0, &B.base);
	} catch (const BatchTooBig&) {
		throw Error(SVB_ENOMEM, "base octree does not fit device memory");
	}
	B.baseNodesSVO = B.nNodesSVO;
	B.basePairs = B.pairs;
	B.baseExact = download(s, B.dExact.p, 1)[0];
	B.pairs = 0;
	B.dExact.zero();
	B.nNodesSVO = 0;
	B.nBatches = 0;
	B.nLastLevSVO = 0;            // geom_octree.cpp:318
	B.svoCounts.clear();
	// ---- enumerate the sub-octrees exactly like :335-344 (leaf node creation order i, child j = 7..0)
	const BatchLevel& BL = B.base[s1 - 1];
	std::vector<uint64_t> hcode = download(s, BL.code.p, BL.n);
	std::vector<uint32_t> htstar = download(s, BL.tstar.p, BL.n);
	B.baseLeafMask = download(s, BL.mask.p, BL.n);
	const std::vector<uint8_t>& hmask = B.baseLeafMask;
	const int lb = (int)s1 - 1;   // path digits of a base leaf node
	std::vector<uint32_t> order(BL.n);
	for (uint32_t i = 0; i < BL.n; ++i) order[i] = i;
	auto pathp = [&](uint32_t i) { uint64_t p = hcode[i]; return lb > 0 ? ((p & ~7ull) | (7ull - (p & 7ull))) : p; };
	std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
		if (htstar[x] != htstar[y]) return htstar[x] < htstar[y];
		return pathp(x) < pathp(y);
	});
	const int G = 1 << s1;
	std::vector<int> hgrid((size_t)G * G * G, -1);
	const double lhs = rootSide / (double)(1u << s1);   // getHalfSideD(stepLevels - 1), :306
	B.baseVoxels = 0;
	for (uint32_t oi = 0; oi < BL.n; ++oi) {
		uint32_t i = order[oi];
		B.baseVoxels += (uint64_t)__builtin_popcount(hmask[i]);
		if (!hmask[i]) continue;
		// centre of base leaf node i: the chain of :222-230
		double cx = B.rootG.cx, cy = B.rootG.cy, cz = B.rootG.cz, k = rootSide * 0.25;
		int ix = 0, iy = 0, iz = 0;
		for (int d = lb - 1; d >= 0; --d) {
			int dig = (int)((hcode[i] >> (3 * d)) & 7);
			cx = cx + ((dig & 4) ? k : -k); cy = cy + ((dig & 2) ? k : -k); cz = cz + ((dig & 1) ? k : -k);
			ix = (ix << 1) | ((dig >> 2) & 1); iy = (iy << 1) | ((dig >> 1) & 1); iz = (iz << 1) | (dig & 1);
			k *= 0.5;
		}
		for (int j = 7; j >= 0; --j) {
			if (!((hmask[i] >> j) & 1)) continue;
			TileHost t;
			double p2x = cx + ((j & 4) ? lhs : -lhs), p2y = cy + ((j & 2) ? lhs : -lhs), p2z = cz + ((j & 1) ? lhs : -lhs);
			double lo[3] = { std::min(cx, p2x), std::min(cy, p2y), std::min(cz, p2z) };
			double hi[3] = { std::max(cx, p2x), std::max(cy, p2y), std::max(cz, p2z) };
			float bf[6];
			float_box(lo, hi, bf, t.g.rootSide);   // :340-344 -> :177-184
			t.g.cx = (lo[0] + hi[0]) * 0.5; t.g.cy = (lo[1] + hi[1]) * 0.5; t.g.cz = (lo[2] + hi[2]) * 0.5;
			t.baseNode = i; t.j = j;
			t.ix = (ix << 1) | ((j >> 2) & 1); t.iy = (iy << 1) | ((j >> 1) & 1); t.iz = (iz << 1) | (j & 1);
			hgrid[((size_t)t.ix * G + t.iy) * G + t.iz] = (int)B.tiles.size();
			B.tiles.push_back(t);
		}
	}
	const uint64_t nTiles = B.tiles.size();
	B.tileBits = bits_for(nTiles ? nTiles - 1 : 0);
	const int Lt = (int)(L - s1);
	if (1 + B.tbits + 3 * lb > 63) throw Error(SVB_ERANGE, "order key of the base octree exceeds 63 bits: decrease step");
	B.grid.G = G; B.grid.cell = rootSide / (double)G;
	B.grid.ox = B.grid1.ox; B.grid.oy = B.grid1.oy; B.grid.oz = B.grid1.oz;
	upload(s, c->pool, B.dGrid, hgrid);
	// triangle ranks inside a sub-octree need as many bits as the busiest tile has candidate triangles (identical on
	// every rank: all ranks hold all triangles and the same tile grid)
	B.tbLocal = bits_for(std::max<uint32_t>(1, max_candidates_per_tile(s, c->pool, c->d_tris, c->T, B.grid, B.dGrid.p, nTiles)) - 1);
	set_order_key_width(B, Lt, B.tbLocal);
	// Order-key widths of the sub-octree levels are a function of the scene alone, never of what this rank happened to
	// reduce: a rank with an empty share (world > nTiles, an empty octant under SVB_SHARD=octant) must rank the merged
	// tables on the same bits as everyone else.
	for (uint32_t g = s1; g < L; ++g) B.obits[g] = std::max(B.obits[g], B.tileBits + B.tbLocal + 3 * (int)(g - s1));
	B.tileRootRef.reset(c->pool, nTiles ? nTiles : 1);
	B.tileRootRef.fill_ff();
	// ---- this rank's share: sub-octrees dealt round-robin in the reference's order (SVB_SHARD=octant: by
	//      top-level octant, the decomposition BASELINE.json names; unbalanced for flat scenes)
	std::vector<uint32_t> mine;
	const char* pol = getenv("SVB_SHARD");
	const bool byOctant = pol && pol[0] == 'o';
	for (uint32_t q = 0; q < nTiles; ++q) {
		uint32_t owner;
		if (byOctant) {
			const TileHost& t = B.tiles[q];
			uint32_t oct = (uint32_t)((((t.ix >> (s1 - 1)) & 1) << 2) | (((t.iy >> (s1 - 1)) & 1) << 1) | ((t.iz >> (s1 - 1)) & 1));
			owner = (uint32_t)(((uint64_t)oct * world) / 8);
		} else owner = q % world;
		if (owner == rank) mine.push_back(q);
	}
	{
		std::vector<int> hpos(nTiles ? nTiles : 1, -1);
		for (size_t i = 0; i < mine.size(); ++i) hpos[mine[i]] = (int)i;
		upload(s, c->pool, B.dSelPos, hpos);
	}
	// root pairs of all of this rank's sub-octrees, binned once (SVB_ROOTS_ONCE=0: once per tile batch, as before)
	if (!mine.empty() && !(getenv("SVB_ROOTS_ONCE") && getenv("SVB_ROOTS_ONCE")[0] == '0')) {
		int cellLo[3] = {1 << 30, 1 << 30, 1 << 30}, cellHi[3] = {-1, -1, -1};
		for (uint32_t q : mine) {
			const TileHost& th = B.tiles[q];
			cellLo[0] = std::min(cellLo[0], th.ix); cellLo[1] = std::min(cellLo[1], th.iy); cellLo[2] = std::min(cellLo[2], th.iz);
			cellHi[0] = std::max(cellHi[0], th.ix); cellHi[1] = std::max(cellHi[1], th.iy); cellHi[2] = std::max(cellHi[2], th.iz);
		}
		try {
			StageTimer tm(s);
			make_root_pairs_all(s, c->pool, c->d_tris, c->T, B.grid, B.dGrid.p, B.dSelPos.p, (uint32_t)mine.size(), B.rootsAll, cellLo, cellHi);
			B.msVox += tm.stop();
		} catch (const BatchTooBig&) {
			B.rootsAll = RootPairsAll();   // more than 2^32 root pairs in one piece: bin per batch
		}
	}
	// leaf-level nodes one batch may hold: 32-bit indices, and ~40 bytes of transient state per leaf node
	const uint64_t budgetTiles = batch_budget(c);
	uint64_t nodeCap = std::min<uint64_t>(3600000000ull, (budgetTiles - c->pool.live) / 40);
	if (!mine.empty()) run_tiles_split(c, B, mine, Lt, s1, budgetTiles, nodeCap);
	B.rootsAll = RootPairsAll();
}

// ---- phase 3: reduce the base octree on top of the (global) sub-octree roots, rank, materialise
void build_finish(svb_ctx* c, const uint64_t* totals) {
	if (!c->build) throw Error(SVB_EINVAL, "no build in progress");
	BuildState& B = *c->build;
	cudaStream_t s = c->stream;
	const uint32_t L = B.L, s1 = B.s1;
	if (B.world > 1) {   // the per-level imports ran stream-ordered; their unique counts and error flags are read back here, once
		std::vector<uint32_t> hs = download(s, B.mergeStatus.p, 8ull * L);
		for (uint32_t g = std::max(s1, 1u); g < L; ++g) {
			if (!B.merged[g]) throw Error(SVB_EINVAL, "svb_shard_finish: level " + std::to_string(g) + " has not been merged");
			merge_resolve(B.tables[g], &hs[8ull * g]);
		}
	}
	uint64_t leafVox = download(s, B.dVoxels.p, 1)[0];
	uint64_t nodesSVO = B.nNodesSVO, lastLev = B.nLastLevSVO, pairs = B.pairs, exact = download(s, B.dExact.p, 1)[0];
	if (totals) { leafVox = totals[0]; nodesSVO = totals[1]; lastLev = totals[2]; pairs = totals[3]; exact = totals[4]; }
	uint64_t nVoxels = leafVox, nTiles = 1;
	if (s1 > 0) {
		nTiles = B.tiles.size();
		nVoxels = B.baseVoxels + leafVox - nTiles;   // :352-354 "root doesn't count"
		nodesSVO += B.baseNodesSVO;
		pairs += B.basePairs;
		exact += B.baseExact;
		StageTimer tm(s);
		const BatchLevel& BL = B.base[s1 - 1];
		const std::vector<uint8_t>& hmask = B.baseLeafMask;
		// children of base leaf node n, in ascending child order, are sub-octree roots: gather their refs
		std::vector<uint32_t> cb(BL.n), seqOf;
		uint32_t acc = 0;
		for (uint32_t n = 0; n < BL.n; ++n) { cb[n] = acc; acc += (uint32_t)__builtin_popcount(hmask[n]); }
		seqOf.assign(acc ? acc : 1, 0);
		for (uint32_t q = 0; q < B.tiles.size(); ++q) {
			const TileHost& t = B.tiles[q];
			int r = __builtin_popcount(hmask[t.baseNode] & ((1u << t.j) - 1));
			seqOf[cb[t.baseNode] + r] = q;
		}
		// A sub-octree whose root came out EMPTY although the base octree set its bit (its cube is rebuilt from a
		// float-narrowed box, geom_octree.cpp:340-344 -> :177-184, so it can be a hair smaller than the base child
		// cube; a triangle that only touches the base cube's face then misses every child of the sub-octree):
		// the reference keeps the parent's bit, and its last toDAG pass skips the empty node, leaving
		// correspondences[] at 0 (:495, :519-526) -- the child pointer ends up at UNIQUE NODE 0 of level s1,
		// i.e. at the root of the first non-empty sub-octree in join order.  Reproduced here, bit for bit.
		{
			const bool maskMode = (kind_of(s1, L) == KIND_LEAF);
			std::vector<uint32_t> hroot = download(s, B.tileRootRef.p, B.tiles.size());
			auto empty = [&](uint32_t v) { return v == UNSET || (maskMode && v == 0); };
			uint32_t first = UNSET;
			for (uint32_t v : hroot) if (!empty(v)) { first = v; break; }
			bool patched = false;
			if (first != UNSET)
				for (uint32_t& v : hroot) if (empty(v)) { v = first; patched = true; }
			if (patched) SVB_CUDA(cudaMemcpyAsync(B.tileRootRef.p, hroot.data(), hroot.size() * 4ull, cudaMemcpyHostToDevice, s));
			SVB_CUDA(cudaStreamSynchronize(s));
		}
		DevBuf<uint32_t> dSeq, dCb, dZero;
		upload(s, c->pool, dSeq, seqOf);
		upload(s, c->pool, dCb, cb);
		upload(s, c->pool, dZero, std::vector<uint32_t>(1, 0));
		DevBuf<uint32_t> topRefs(c->pool, seqOf.size());
		if (acc) k_gather_u32<<<blocks_for(acc, 256), 256, 0, s>>>(acc, dSeq.p, B.tileRootRef.p, topRefs.p);
		SVB_KERNEL_CHECK();
		B.base[s1 - 1].childBase = std::move(dCb);
		const uint32_t* below = topRefs.p;
		int belowMode = (kind_of(s1, L) == KIND_LEAF) ? CH_MASK_U32 : CH_UID_U32;
		for (int l = (int)s1 - 1; l >= 0; --l) {
			BatchLevel& X = B.base[l];
			DedupArgs a;
			a.N = X.n; a.code = X.code.p; a.tstar = X.tstar.p; a.mask = X.mask.p; a.childBase = X.childBase.p;
			a.l = l; a.tbits = B.tbits; a.tileSeq = dZero.p; a.tileStart = dZero.p;
			a.childMode = belowMode; a.childRefs = below;
			if (l == 0) { B.rootChildMode = belowMode; root_key(s, a, B.rootKey.p); break; }
			int ob = 1 + B.tbits + 3 * l;
			if (ob > B.obits[l]) B.obits[l] = ob;
			X.ref.reset(c->pool, X.n);
			a.ref = X.ref.p;
			ProfScope ps(c, B.tables[l].kind == KIND_K64 ? "dedup_k64" : "dedup_inner", (uint32_t)l, X.n);
			dedup_level(s, c->pool, B.tables[l], a);
			ps.done(B.tables[l].count, 37.0 * (double)X.n);
			below = X.ref.p;
			belowMode = CH_UID_U32;
		}
		SVB_CUDA(cudaStreamSynchronize(s));
		B.msDedup += tm.stop();
	}
	double msFin;
	{
		StageTimer tm(s);
		finalize_levels(s, c->pool, B.tables, B.obits, B.rootKey.p, B.rootChildMode, c->out);
		msFin = tm.stop();
	}
	svb_stats& st = c->stats;
	st.nTotalVoxels = nVoxels;
	st.nNodesSVO = nodesSVO;
	st.nNodesLastLevSVO = lastLev;
	uint64_t nn = 1;
	for (uint32_t g = 1; g < L; ++g) nn += c->out[g].n;
	st.nNodesDAG = nn;              // geom_octree.cpp:471,516,544
	st.nNodes = nn;
	st.nNodesLastLevDAG = c->out[L - 1].n;
	st.nTiles = nTiles;
	st.nBatches = B.nBatches;
	st.nPairsTotal = pairs;
	st.nExactTests = exact;
	st.rootSide = B.rootSide;
	memcpy(st.bboxF, B.bboxF, sizeof(B.bboxF));
	st.msVoxelize = B.msVox;
	st.msDedup = B.msDedup;
	st.msFinalize = msFin;
	c->svoCounts = B.svoCounts;
	c->levels = L;
	c->state = SVB_S_DAG;
	{
		cudaEvent_t e1;
		cudaEventCreate(&e1);
		cudaEventRecord(e1, s);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, B.evStart, e1);
		st.msTotal = ms;
		cudaEventDestroy(e1);
		cudaEventDestroy(B.evStart);
	}
	st.nKernelLaunches = g_launches.load() - B.launches0;
	c->build.reset();
	resolve_profile(c);
}

template <class F>
int guarded(svb_ctx* c, F&& f) {
	if (!c) return SVB_EINVAL;
	try {
		if (cudaSetDevice(c->device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return SVB_ECUDA; }
		f();
		c->err.clear();
		return SVB_OK;
	} catch (const Error& e) {
		c->err = e.what();
		for (auto& p : c->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
		c->pending.clear();
		return e.code;
	} catch (const BatchTooBig&) {
		c->err = "batch too big";
		return SVB_ENOMEM;
	} catch (const std::exception& e) {
		c->err = e.what();
		return SVB_ECUDA;
	}
}

// A 64-bit tag collision found by an exact verify pass is not an error of the input: the stage runs again under the next
// seed (every tag changes), at most a few times.  f must leave no trace of a failed attempt behind (build_local starts from
// scratch; cross_merge_device replaces the octree only at its very end).
template <class F>
void with_hash_retries(svb_ctx* c, F&& f) {
	for (int attempt = 0;; ++attempt) {
		try {
			f();
			return;
		} catch (const Error& e) {
			if (e.code != SVB_ECOLLISION || attempt >= 4) throw;
			for (auto& p : c->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
			c->pending.clear();
			c->hashSeed++;
			c->nHashRetries++;
		}
	}
}

}  // namespace

// ===================================================================================== C ABI
extern "C" {

const char* svb_version(void) { return "svb 0.1 (sm_100a)"; }

svb_ctx* svb_create(int device) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;   // no device: no context, no fallback
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	svb_ctx* c = new svb_ctx();
	c->device = device;
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
	c->pool.stream = c->stream;
	return c;
}

svb_ctx* svb_create_on_stream(int device, void* cuda_stream) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	svb_ctx* c = new svb_ctx();
	c->device = device;
	c->stream = (cudaStream_t)cuda_stream;
	c->ownsStream = false;
	c->pool.stream = c->stream;
	return c;
}

void svb_destroy(svb_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	c->out.clear();
	c->attr.reset();
	c->build.reset();
	c->trisOwned.release();
	cudaStreamSynchronize(c->stream);
	c->image.release();
	c->staging.release();
	c->pool.release_all();
	for (auto& p : c->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
	for (cudaEvent_t e : c->evPool) cudaEventDestroy(e);
	if (c->ownsStream) cudaStreamDestroy(c->stream);
	delete c;
}

const char* svb_last_error(const svb_ctx* c) { return c ? c->err.c_str() : "null context (no CUDA device?)"; }
void* svb_stream(const svb_ctx* c) { return c ? (void*)c->stream : nullptr; }
int svb_synchronize(svb_ctx* c) {
	return guarded(c, [&] { SVB_CUDA(cudaStreamSynchronize(c->stream)); });
}

int svb_set_triangles(svb_ctx* c, const float* xyz9_host, uint64_t ntris) {
	return guarded(c, [&] {
		if (ntris && !xyz9_host) throw Error(SVB_EINVAL, "null triangle pointer");
		if (ntris >= (1ull << 32)) throw Error(SVB_ERANGE, "more than 2^32 triangles");
		c->trisOwned.reset(c->pool, ntris * 9 + 4);
		if (ntris) SVB_CUDA(cudaMemcpyAsync(c->trisOwned.p, xyz9_host, ntris * 36, cudaMemcpyHostToDevice, c->stream));
		SVB_CUDA(cudaStreamSynchronize(c->stream));
		c->d_tris = c->trisOwned.p;
		c->T = ntris;
	});
}

int svb_set_triangles_device(svb_ctx* c, const float* xyz9_dev, uint64_t ntris) {
	return guarded(c, [&] {
		if (ntris && !xyz9_dev) throw Error(SVB_EINVAL, "null triangle pointer");
		if (ntris >= (1ull << 32)) throw Error(SVB_ERANGE, "more than 2^32 triangles");
		c->trisOwned.release();
		c->d_tris = xyz9_dev;
		c->T = ntris;
	});
}

int svb_build(svb_ctx* c, uint32_t levels, uint32_t step, const double bmin[3], const double bmax[3], svb_stats* out) {
	int rc = guarded(c, [&] {
		if (!bmin || !bmax) throw Error(SVB_EINVAL, "null bbox");
		with_hash_retries(c, [&] {
			build_local(c, levels, step, bmin, bmax, 0, 1);
			build_finish(c, nullptr);
		});
		c->stats.nHashRetries = c->nHashRetries;
	});
	if (rc != SVB_OK && c) { c->out.clear(); c->state = SVB_S_EMPTY; c->build.reset(); }
	if (rc == SVB_OK && out) *out = c->stats;
	return rc;
}

// ---- multi-GPU: local phase, per-level exchange, finish
int svb_shard_build(svb_ctx* c, uint32_t levels, uint32_t step, const double bmin[3], const double bmax[3], uint32_t rank, uint32_t world) {
	int rc = guarded(c, [&] {
		if (!bmin || !bmax) throw Error(SVB_EINVAL, "null bbox");
		with_hash_retries(c, [&] { build_local(c, levels, step, bmin, bmax, rank, world); });   // (rank-local tables: the seed is a private matter)
		SVB_CUDA(cudaStreamSynchronize(c->stream));
	});
	if (rc != SVB_OK && c) { c->out.clear(); c->state = SVB_S_EMPTY; c->build.reset(); }
	return rc;
}

int svb_shard_info(const svb_ctx* c, uint32_t* firstLevel, uint32_t* lastLevel, uint64_t* nTiles, uint64_t counters[5]) {
	if (!c || !c->build) return SVB_EINVAL;
	const svb_build_state& B = *c->build;
	if (firstLevel) *firstLevel = B.s1;
	if (lastLevel) *lastLevel = B.L - 1;
	if (nTiles) *nTiles = B.tiles.size();
	if (counters) {
		uint64_t v[2] = {0, 0};
		cudaMemcpy(&v[0], B.dVoxels.p, 8, cudaMemcpyDeviceToHost);
		cudaMemcpy(&v[1], B.dExact.p, 8, cudaMemcpyDeviceToHost);
		counters[0] = v[0]; counters[1] = B.nNodesSVO; counters[2] = B.nLastLevSVO; counters[3] = B.pairs; counters[4] = v[1];
	}
	return SVB_OK;
}

int svb_shard_level_count(const svb_ctx* c, uint32_t level, uint64_t* n, uint32_t* recBytes) {
	if (!c || !c->build || level == 0 || level >= c->build->L) return SVB_EINVAL;
	const LevelTable& T = c->build->tables[level];
	if (n) *n = merge_count(T);
	if (recBytes) *recBytes = merge_rec_bytes(T.kind);
	return SVB_OK;
}

int svb_shard_export_level(svb_ctx* c, uint32_t level, void* d_out) {
	return guarded(c, [&] {
		if (!c->build || level < c->build->s1 || level >= c->build->L) throw Error(SVB_EINVAL, "bad level");
		svb_build_state& B = *c->build;
		LevelTable& T = B.tables[level];
		const uint32_t* l2gChild = nullptr;
		if (T.kind == KIND_INNER) {
			if (!B.merged[level + 1]) throw Error(SVB_EINVAL, "levels must be merged bottom-up");
			l2gChild = B.l2g[level + 1].p;
		}
		merge_export(c->stream, c->pool, T, l2gChild, d_out);   // stream-ordered on svb_stream(ctx)
	});
}

int svb_shard_import_level(svb_ctx* c, uint32_t level, const void* d_all, const uint64_t* counts, uint64_t strideBytes) {
	return guarded(c, [&] {
		if (!c->build || level < c->build->s1 || level >= c->build->L || !counts) throw Error(SVB_EINVAL, "bad level");
		svb_build_state& B = *c->build;
		if (B.world < 2) throw Error(SVB_EINVAL, "not a sharded build");
		merge_import(c->stream, c->pool, B.tables[level], d_all, counts, B.world, strideBytes, B.rank, B.l2g[level], B.mergeStatus.p + 8ull * level, c->mergeSeed);
		B.merged[level] = true;
		if (level == B.s1) {   // sub-octree roots now have global uids
			if (B.tables[level].kind != KIND_LEAF && !B.tiles.empty()) {
				k_remap_refs<<<blocks_for(B.tiles.size(), 256), 256, 0, c->stream>>>(B.tiles.size(), B.tileRootRef.p, B.l2g[level].p);
				SVB_KERNEL_CHECK();
			}
		}
	});
}

int svb_shard_export_roots(svb_ctx* c, void* d_out) {
	return guarded(c, [&] {
		if (!c->build) throw Error(SVB_EINVAL, "no build in progress");
		svb_build_state& B = *c->build;
		if (!B.merged[B.s1]) throw Error(SVB_EINVAL, "merge the sub-octree root level first");
		SVB_CUDA(cudaMemcpyAsync(d_out, B.tileRootRef.p, (B.tiles.size() ? B.tiles.size() : 1) * 4ull, cudaMemcpyDeviceToDevice, c->stream));
	});
}

int svb_shard_import_roots(svb_ctx* c, const void* d_all) {
	return guarded(c, [&] {
		if (!c->build) throw Error(SVB_EINVAL, "no build in progress");
		svb_build_state& B = *c->build;
		uint64_t n = B.tiles.size();
		if (n) {
			k_min_u32_rows<<<blocks_for(n, 256), 256, 0, c->stream>>>(n, B.world, (const uint32_t*)d_all, B.tileRootRef.p);
			SVB_KERNEL_CHECK();
		}
	});
}

int svb_shard_finish(svb_ctx* c, const uint64_t totals[5], svb_stats* out) {
	int rc = guarded(c, [&] {
		if (!c->build) throw Error(SVB_EINVAL, "no build in progress");
		build_finish(c, totals);
		c->stats.nHashRetries = c->nHashRetries;
	});
	if (rc != SVB_OK && c) { c->out.clear(); c->state = SVB_S_EMPTY; c->build.reset(); }
	if (rc == SVB_OK && out) *out = c->stats;
	return rc;
}

int svb_to_sdag(svb_ctx* c, svb_stats* out) {
	int rc = guarded(c, [&] {
		if (c->state != SVB_S_DAG) throw Error(SVB_EINVAL, "ERROR! This is not a DAG or SDAG!");   // geom_octree.cpp:560-563
		c->lastImageKind = -1;
		const uint64_t launches0 = g_launches.load();
		StageTimer tm(c->stream);
		uint64_t nn = to_sdag_device(c);
		c->stats.msSdag = tm.stop();
		c->stats.nKernelLaunches = g_launches.load() - launches0;
		c->stats.nHashRetries = c->nHashRetries;
		c->stats.nNodesSDAG = nn;   // geom_octree.cpp:578,664,683: root not counted
		c->stats.nNodes = nn;
		c->state = SVB_S_SDAG;
		resolve_profile(c);
	});
	if (rc == SVB_OK && out) *out = c->stats;
	return rc;
}

int svb_cross_merge(svb_ctx* c, svb_stats* out) {
	int rc = guarded(c, [&] {
		if (c->state != SVB_S_DAG) throw Error(SVB_EINVAL, "cross-level merge needs an octree in DAG state");
		c->lastImageKind = -1;
		const uint64_t launches0 = g_launches.load();
		StageTimer tm(c->stream);
		uint64_t nn = 0, removed = 0;
		with_hash_retries(c, [&] { removed = cross_merge_device(c, &nn); });
		c->stats.msCrossMerge = tm.stop();
		c->stats.nCrossLevelMerged = removed;   // ext.cpp:1517
		c->stats.nNodesDAG = nn;                // ext.cpp:1448
		c->stats.nNodes = nn;
		c->stats.nKernelLaunches = g_launches.load() - launches0;
		c->stats.nHashRetries = c->nHashRetries;
	});
	if (rc == SVB_OK && out) *out = c->stats;
	return rc;
}

int svb_set_merge_seed(svb_ctx* c, uint64_t seed) {
	if (!c) return SVB_EINVAL;
	c->mergeSeed = seed;
	return SVB_OK;
}

int svb_state(const svb_ctx* c) { return c ? c->state : SVB_S_EMPTY; }
uint32_t svb_levels(const svb_ctx* c) { return c ? c->levels : 0; }
int svb_get_stats(const svb_ctx* c, svb_stats* out) {
	if (!c || !out) return SVB_EINVAL;
	*out = c->stats;
	return SVB_OK;
}

int svb_level_count(const svb_ctx* c, uint32_t lev, uint64_t* n) {
	if (!c || !n || lev >= c->out.size()) return SVB_EINVAL;
	*n = c->out[lev].n;
	return SVB_OK;
}
int svb_level_count_svo(const svb_ctx* c, uint32_t lev, uint64_t* n) {
	if (!c || !n) return SVB_EINVAL;
	*n = lev < c->svoCounts.size() ? c->svoCounts[lev] : 0;
	return SVB_OK;
}

int svb_download_level(svb_ctx* c, uint32_t lev, uint8_t* mask, uint32_t* child8, uint8_t* mirror3, uint8_t* inv, uint32_t* childLevel8) {
	return guarded(c, [&] {
		if (lev >= c->out.size()) throw Error(SVB_EINVAL, "level out of range");
		OutLevel& o = c->out[lev];
		cudaStream_t s = c->stream;
		if (o.n) {
			if (mask) SVB_CUDA(cudaMemcpyAsync(mask, o.mask.p, o.n, cudaMemcpyDeviceToHost, s));
			if (child8) SVB_CUDA(cudaMemcpyAsync(child8, o.child.p, o.n * 32, cudaMemcpyDeviceToHost, s));
			if (mirror3) SVB_CUDA(cudaMemcpyAsync(mirror3, o.mirror.p, o.n * 3, cudaMemcpyDeviceToHost, s));
			if (inv) SVB_CUDA(cudaMemcpyAsync(inv, o.inv.p, o.n, cudaMemcpyDeviceToHost, s));
			if (childLevel8) {
				if (o.hasChildLevel) SVB_CUDA(cudaMemcpyAsync(childLevel8, o.childLevel.p, o.n * 32, cudaMemcpyDeviceToHost, s));
				else { SVB_CUDA(cudaStreamSynchronize(s)); for (uint64_t i = 0; i < o.n * 8; ++i) childLevel8[i] = lev + 1; }   // initChildLevels(), ext.cpp:21-30
			}
		}
		SVB_CUDA(cudaStreamSynchronize(s));
	});
}

// the previous round's path: copy the levels D2H and run the host encoders (csrc/host/encoders.cpp); SVB_ENCODE=host selects it
static uint64_t encode_host_path(svb_ctx* c, int kind) {
	svbhost::OctreeData o;
	o.levels.resize(c->levels);
	cudaStream_t s = c->stream;
	for (uint32_t l = 0; l < c->levels; ++l) {
		OutLevel& d = c->out[l];
		svbhost::LevelSoA& h = o.levels[l];
		h.n = d.n;
		h.mask.resize(d.n); h.child.resize(d.n * 8); h.mirror.resize(d.n * 3);
		if (d.n) {
			SVB_CUDA(cudaMemcpyAsync(h.mask.data(), d.mask.p, d.n, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaMemcpyAsync(h.child.data(), d.child.p, d.n * 32, cudaMemcpyDeviceToHost, s));
			SVB_CUDA(cudaMemcpyAsync(h.mirror.data(), d.mirror.p, d.n * 3, cudaMemcpyDeviceToHost, s));
			if (d.hasChildLevel) {
				h.childLevel.resize(d.n * 8);
				SVB_CUDA(cudaMemcpyAsync(h.childLevel.data(), d.childLevel.p, d.n * 32, cudaMemcpyDeviceToHost, s));
			}
		}
	}
	SVB_CUDA(cudaStreamSynchronize(s));
	memcpy(o.bboxF, c->stats.bboxF, 24);
	o.rootSide = c->stats.rootSide;
	o.nNodes = c->stats.nNodes;
	o.nVoxels = c->stats.nTotalVoxels;
	o.state = c->state;
	std::vector<uint8_t> img;
	std::string err;
	if (!svbhost::encode_file(o, kind, img, &err)) throw Error(SVB_EINVAL, err);
	memcpy(c->image.reserve(img.size()), img.data(), img.size());
	return img.size();
}

static void encode_cached(svb_ctx* c, int kind) {
	if (c->state != SVB_S_DAG && c->state != SVB_S_SDAG) throw Error(SVB_EINVAL, "nothing to encode");
	const char* e = getenv("SVB_ENCODE");
	const bool host = e && e[0] == 'h';
	const int key = kind + (host ? 16 : 0);
	if (c->lastImageKind == key && c->imageSize) return;
	c->lastImageKind = -1;
	c->imageSize = host ? encode_host_path(c, kind) : encode_device(c, kind);
	c->lastImageKind = key;
}

int64_t svb_encode(svb_ctx* c, int kind, uint8_t* buf, uint64_t cap) {
	int64_t size = -1;
	int rc = guarded(c, [&] {
		encode_cached(c, kind);
		size = (int64_t)c->imageSize;
		if (buf && cap >= c->imageSize) memcpy(buf, c->image.p, c->imageSize);
	});
	return rc == SVB_OK ? size : (int64_t)rc;
}

int svb_encode_view(svb_ctx* c, int kind, const uint8_t** image, uint64_t* size) {
	return guarded(c, [&] {
		if (!image || !size) throw Error(SVB_EINVAL, "null output pointer");
		encode_cached(c, kind);
		*image = c->image.p;
		*size = c->imageSize;
	});
}

int64_t svb_encode_levels(uint32_t levels, const uint64_t* counts, const uint8_t* mask, const uint32_t* child8,
                          const uint8_t* mirror3, const uint32_t* childLevel8, const float bboxF[6], double rootSide,
                          uint64_t nNodes, int state, int kind, uint8_t* buf, uint64_t cap) {
	if (!counts || !mask || !child8 || !bboxF || levels == 0) return SVB_EINVAL;
	try {
		svbhost::OctreeData o;
		o.levels.resize(levels);
		uint64_t off = 0;
		for (uint32_t l = 0; l < levels; ++l) {
			svbhost::LevelSoA& h = o.levels[l];
			h.n = counts[l];
			h.mask.assign(mask + off, mask + off + h.n);
			h.child.assign(child8 + off * 8, child8 + (off + h.n) * 8);
			if (mirror3) h.mirror.assign(mirror3 + off * 3, mirror3 + (off + h.n) * 3);
			else h.mirror.assign(h.n * 3, 0);
			if (childLevel8) h.childLevel.assign(childLevel8 + off * 8, childLevel8 + (off + h.n) * 8);
			off += h.n;
		}
		memcpy(o.bboxF, bboxF, 24);
		o.rootSide = rootSide;
		o.nNodes = nNodes;
		o.state = state;
		std::vector<uint8_t> img;
		if (!svbhost::encode_file(o, kind, img, nullptr)) return SVB_EINVAL;
		if (buf && cap >= img.size()) memcpy(buf, img.data(), img.size());
		return (int64_t)img.size();
	} catch (...) {
		return SVB_EINVAL;
	}
}

int svb_ssvdag_order_from_refs(const uint32_t* refs, const uint32_t* start, uint32_t nLevels, uint32_t* order) {
	if (!refs || !start || !order || nLevels == 0 || nLevels > 64) return SVB_EINVAL;
	try {
		svbhost::ssvdag_order_from_refs(refs, start, (int)nLevels, order);
		return SVB_OK;
	} catch (...) {
		return SVB_ENOMEM;
	}
}

int svb_upload_levels(svb_ctx* c, uint32_t levels, const uint64_t* counts, const uint8_t* mask, const uint32_t* child8,
                      const float bboxF[6], double rootSide, uint64_t nVoxels) {
	return guarded(c, [&] {
		if (!counts || !mask || !child8 || levels < 2 || levels > 32) throw Error(SVB_EINVAL, "bad arguments");
		// every later stage (toSDAG, cross merge, encoders) indexes the next level with these values unchecked
		{
			uint64_t off = 0;
			for (uint32_t l = 0; l < levels; ++l) {
				if (counts[l] >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "level too large");
				const uint64_t nNext = l + 1 < levels ? counts[l + 1] : 0;
				for (uint64_t i = off; i < off + counts[l]; ++i)
					for (int k = 0; k < 8; ++k) {
						const uint32_t ch = child8[i * 8 + k];
						const bool bit = (mask[i] >> k) & 1;
						if (l + 1 == levels) continue;                      // voxel masks: child slots are not read
						if (bit != (ch != NULLNODE) || (bit && ch >= nNext))
							throw Error(SVB_EINVAL, "svb_upload_levels: child pointer / mask mismatch or child index out of range at level " + std::to_string(l));
					}
				off += counts[l];
			}
		}
		c->lastImageKind = -1;
		cudaStream_t s = c->stream;
		c->out.clear();
		c->out.resize(levels);
		uint64_t off = 0, nn = 0;
		for (uint32_t l = 0; l < levels; ++l) {
			OutLevel& o = c->out[l];
			o.n = counts[l];
			o.mask.reset(c->pool, o.n); o.child.reset(c->pool, o.n * 8); o.mirror.reset(c->pool, o.n * 3); o.inv.reset(c->pool, o.n);
			o.mirror.zero(); o.inv.zero();
			if (o.n) {
				SVB_CUDA(cudaMemcpyAsync(o.mask.p, mask + off, o.n, cudaMemcpyHostToDevice, s));
				SVB_CUDA(cudaMemcpyAsync(o.child.p, child8 + off * 8, o.n * 32, cudaMemcpyHostToDevice, s));
			}
			off += o.n;
			nn += o.n;
		}
		SVB_CUDA(cudaStreamSynchronize(s));
		memset(&c->stats, 0, sizeof(c->stats));
		c->stats.nTotalVoxels = nVoxels;
		c->stats.nNodesDAG = nn;
		c->stats.nNodes = nn;
		c->stats.nNodesLastLevDAG = counts[levels - 1];
		c->stats.rootSide = rootSide;
		if (bboxF) memcpy(c->stats.bboxF, bboxF, 24);
		c->levels = levels;
		c->svoCounts.clear();
		c->state = SVB_S_DAG;
	});
}

int svb_decode_svdag(const uint8_t* file, uint64_t size, uint32_t* levels, uint64_t* counts, uint8_t* mask, uint32_t* child8,
                     float bboxF[6], double* rootSide, uint64_t* nNodes) {
	if (!file || !levels || !counts) return SVB_EINVAL;
	try {
		svbhost::OctreeData o;
		if (!svbhost::decode_svdag(file, size, o, nullptr)) return SVB_EINVAL;
		*levels = (uint32_t)o.levels.size();
		uint64_t off = 0;
		for (size_t l = 0; l < o.levels.size(); ++l) {
			counts[l] = o.levels[l].n;
			if (mask) memcpy(mask + off, o.levels[l].mask.data(), o.levels[l].n);
			if (child8) memcpy(child8 + off * 8, o.levels[l].child.data(), o.levels[l].n * 32);
			off += o.levels[l].n;
		}
		if (bboxF) memcpy(bboxF, o.bboxF, 24);
		if (rootSide) *rootSide = o.rootSide;
		if (nNodes) *nNodes = o.nNodes;
		return SVB_OK;
	} catch (...) {
		return SVB_EINVAL;
	}
}

int svb_load_svdag(svb_ctx* c, const uint8_t* file, uint64_t size, svb_stats* out) {
	if (!c || !file) return SVB_EINVAL;
	svbhost::OctreeData o;
	std::vector<uint64_t> counts;
	std::vector<uint8_t> mask;
	std::vector<uint32_t> child;
	try {   // nothing may escape across the C boundary (a hostile header can ask for absurd allocations)
		std::string err;
		if (!svbhost::decode_svdag(file, size, o, &err)) { c->err = err; return SVB_EINVAL; }
		for (auto& L : o.levels) {
			counts.push_back(L.n);
			mask.insert(mask.end(), L.mask.begin(), L.mask.end());
			child.insert(child.end(), L.child.begin(), L.child.end());
		}
	} catch (const std::exception& e) {
		c->err = std::string("svb_load_svdag: ") + e.what();
		return SVB_EINVAL;
	}
	int rc = svb_upload_levels(c, (uint32_t)o.levels.size(), counts.data(), mask.data(), child.data(), o.bboxF, o.rootSide, 0);
	if (rc == SVB_OK) {
		c->stats.nNodes = o.nNodes;       // Octree::_nNodes as stored in the file (GeomOctree(data, ..., stats), geom_octree.cpp:49-60)
		c->stats.nNodesDAG = o.nNodes;
		if (out) *out = c->stats;
	}
	return rc;
}

// ---- material-id leaves + Gray-coded attribute bit-trees (svb_attr.cu)
int svb_build_svo_materials(svb_ctx* c, uint32_t levels, const double bmin[3], const double bmax[3], const uint32_t* triMaterial, uint64_t* nLeafNodes) {
	int rc = guarded(c, [&] {
		if (!bmin || !bmax) throw Error(SVB_EINVAL, "null bbox");
		with_hash_retries(c, [&] { attr_build(c, levels, bmin, bmax, triMaterial); });
		if (nLeafNodes) *nLeafNodes = attr_leaf_count(c);
	});
	if (rc != SVB_OK && c) c->attr.reset();
	return rc;
}
int svb_download_leaf_materials(svb_ctx* c, uint8_t* mask, uint32_t* material8) {
	return guarded(c, [&] { attr_download(c, mask, material8); });
}
int svb_attribute_bit_trees(svb_ctx* c, uint32_t nbits, int gray, uint64_t* nodes, uint64_t* voxels) {
	return guarded(c, [&] {
		if (!nodes) throw Error(SVB_EINVAL, "null output");
		with_hash_retries(c, [&] { attr_bit_trees(c, nbits, gray, nodes, voxels); });
	});
}
uint32_t svb_gray_code(uint32_t a) { return a ^ (a >> 1); }

int svb_set_profiling(svb_ctx* c, int enabled) {
	if (!c) return SVB_EINVAL;
	c->profiling = enabled != 0;
	c->profAccumulate = enabled >= 2;   // 2 / 3: keep the records of successive builds (read and cleared by the caller: svb_profile_clear)
	c->profEmitOnly = enabled == 3;     // 3: only the "emit" launches are bracketed (bench.py: the roofline kernel inside the timed region)
	if (!enabled) c->prof.clear();
	if (enabled >= 2) {                 // the events of a timed region are created here, not between its launches
		cudaSetDevice(c->device);
		while (c->evPool.size() < 8192) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; c->evPool.push_back(e); }
	}
	return SVB_OK;
}
int svb_profile_clear(svb_ctx* c) {
	if (!c) return SVB_EINVAL;
	c->prof.clear();
	return SVB_OK;
}
int svb_profile_count(const svb_ctx* c) { return c ? (int)c->prof.size() : 0; }
int svb_profile_get(const svb_ctx* c, int i, svb_prof_rec* out) {
	if (!c || !out || i < 0 || i >= (int)c->prof.size()) return SVB_EINVAL;
	*out = c->prof[i];
	return SVB_OK;
}
int svb_set_batch_budget(svb_ctx* c, uint64_t bytes) {
	if (!c) return SVB_EINVAL;
	c->batchBudget = bytes;
	return SVB_OK;
}

}  // extern "C"
