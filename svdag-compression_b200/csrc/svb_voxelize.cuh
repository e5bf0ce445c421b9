// svb_voxelize.cuh -- host-side entry points of the voxelizer (svb_voxelize.cu).
#pragma once
#include "svb_classify.cuh"
#include "svb_internal.cuh"

namespace svb {

// Thrown (before anything is reduced) when a tile batch would outgrow its buffers.  Carries the
// per-tile node counts of the level it stopped at so the caller can cut the batch by weight.
struct BatchTooBig {
	int level = -1;                   // tile-local level the weights were taken at (-1: none)
	int remaining = 0;                // levels still to descend from there to the leaf level
	double growth = 4.0;              // observed node growth per level
	std::vector<uint32_t> weight;     // nodes per tile of the batch at `level`
};

struct TileGridHost {    // regular grid of tile cubes over the root cube (G = 2^(step+1) per axis; 1 for step 0)
	double ox = 0, oy = 0, oz = 0;   // min corner of the root cube
	double cell = 1;                 // tile side
	int G = 1;
};

// (triangle, tile) candidate pairs for the tiles of one batch, sorted by (tile, triangle).  On return pnode[q] = tile
// (index inside the batch), rootTri[q] = triangle, ptri[q] = q, tileStart[tile] = first q of the tile.
// d_gridTile: grid cell -> global tile_seq (-1 none); d_selPos: global tile_seq -> position in the caller's tile
// selection (-1 none); the batch is the positions [posFirst, posFirst + ntiles).
void make_root_pairs(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid,
                     const int* d_gridTile, const int* d_selPos, uint32_t posFirst, uint32_t ntiles, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode,
                     DevBuf<uint32_t>& rootTri, DevBuf<uint32_t>& tileStart, uint64_t& P, const int cellLo[3] = nullptr, const int cellHi[3] = nullptr);
// (cellLo / cellHi: grid-cell bounding box of the batch's tiles; triangles that cannot reach it are skipped early)

// The same for ALL sub-octrees of a caller's selection at once (positions [0, nSel)): the triangles are binned and the
// pairs sorted ONCE per build instead of once per tile batch (every batch used to walk all T triangles twice, work that
// does not shrink when the sub-octrees are spread over more GPUs).  A batch = positions [a, a + nt) is then a contiguous
// slice: batch_root_pairs() hands out the per-batch views (ptri[q] = q, pnode[q] = tile inside the batch, rootTri -> slice
// of R.rootTri, tileStart[tile] = first q of the tile).
struct RootPairsAll {
	DevBuf<uint32_t> rootTri, tileOf;     // per root pair, sorted by (selection position, triangle)
	DevBuf<uint32_t> tileBegin;           // nSel + 1
	std::vector<uint32_t> hTileBegin;     // host copy
	uint64_t P = 0;
	uint32_t nSel = 0;
	bool valid = false;
};
void make_root_pairs_all(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid, const int* d_gridTile, const int* d_selPos,
                         uint32_t nSel, RootPairsAll& R, const int cellLo[3] = nullptr, const int cellHi[3] = nullptr);
void batch_root_pairs(cudaStream_t s, Pool& pool, const RootPairsAll& R, uint32_t a, uint32_t nt, DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode,
                      const uint32_t*& rootTri, DevBuf<uint32_t>& tileStart, uint64_t& P);

// largest number of candidate triangles any tile of the grid has (bounds the tile-local triangle rank)
uint32_t max_candidates_per_tile(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T, const TileGridHost& grid, const int* d_gridTile, uint64_t nTiles);

// Level-synchronous SVO build of `ntiles` sub-octrees of Lt levels each.  Consumes the root pairs.  First-touch
// "triangles" (BatchLevel::tstar) are root pair indices q: monotone in the triangle id within a tile.
void voxelize_batch(cudaStream_t s, Pool& pool, const float* d_tris, const TileGeom* d_tiles, uint32_t ntiles, int Lt,
                    DevBuf<uint32_t>& ptri, DevBuf<uint32_t>& pnode, const uint32_t* rootTri, const uint32_t* tileStart, uint64_t P,
                    uint64_t budget_bytes, uint64_t nodeCap, std::vector<BatchLevel>& lv, uint64_t& pairsTotal, uint64_t* d_nExact, bool directCentre, bool allFlat,
                    int untracked = 0, ProfHook* prof = nullptr, DevBuf<uint32_t>* leafLastTri = nullptr);
// leafLastTri (attribute builds, svb_attr.cu): receives, per leaf node and voxel, 1 + the index of the LAST triangle in file
// order that touches the voxel (0: none) -- what GeomOctree::buildSVO(..., putMaterialIdInLeaves = true) keeps
// (geom_octree.cpp:210-211, :252).  Implies the exact classifier (every voxel decided by its own lane).
// prof: per-launch records "emit" (n_in = parent pairs, n_out = child pairs) and "children" (n_in = nodes, n_out = child nodes)
// with their algorithmic bytes.
// untracked = 1: the first touches of the deepest level are not tracked (lv[Lt-1].tstar stays empty); 2: nor those of the
// level above it (the 4^3 level).  Valid only when the caller can reduce those levels without them (dedup_leaf_known and
// DedupArgs::query, svb_dedup.cuh).

// true when every triangle is flat (box meshes): selects the slow-stream kernel without the general edge / plane filter
bool all_triangles_flat(cudaStream_t s, Pool& pool, const float* d_tris, uint64_t T);

// (centre_chain_exact(): svb_classify.cuh)

}  // namespace svb
