// svb_dedup.cuh -- host-side entry points of the per-level DAG reduction (svb_dedup.cu).
#pragma once
#include "svb_internal.cuh"

namespace svb {

// One level of nodes to be reduced.  Children of node n are refs[childBase[n] + r], r = rank of the
// child among the set bits of mask[n] (ascending child index).
struct LeafQuery;
struct DedupArgs {
	uint64_t N = 0;
	const uint64_t* code = nullptr;     // (tile_local << 3l) | path
	const uint32_t* tstar = nullptr;
	const uint8_t* mask = nullptr;      // structural child mask
	const uint32_t* childBase = nullptr;
	int childMode = CH_UID_U32;         // how to read the level below
	const void* childRefs = nullptr;    // u8 masks | u32 uids | u32 masks
	int l = 0;                          // octal digits of `path`
	int tbits = 0;                      // bits of a tile-local triangle rank
	const uint32_t* tileSeq = nullptr;  // device: tile_local -> global tile_seq (sub-octree sequence number)
	const uint32_t* tileStart = nullptr;// device: tile_local -> first root-pair index of the tile (tstar - tileStart = triangle rank)
	uint32_t* ref = nullptr;            // out: uid of each node's unique representative (NULLREF = empty node)
	// tile_seq range of the batch.  Batches reach a table in ascending sub-octree order (svb_api.cu::run_tiles_split), so
	// when seqLo exceeds everything the table has seen, no node of this batch can lower the order key of an entry that
	// already exists: such entries are frozen and their nodes need neither code nor tstar (svb_dedup.cu).
	uint32_t seqLo = 0, seqHi = 0;
	bool seqMonotone = false;           // tileSeq ascends with the tile-local index: order keys of the batch compare like (tstar, path')
	// tstar == nullptr on the 4^3 level of a later batch: its first touches were not tracked; the nodes of NEW entries get
	// theirs from a direct query of the triangles (svb_dedup.cu::k_k64_query), which needs:
	const LeafQuery* query = nullptr;
};

void table_init(cudaStream_t s, Pool& pool, LevelTable& T, int kind, uint64_t seed = 0);

// KIND_LEAF: nodes are bare 8-bit voxel masks. Accumulates popcounts into *d_voxels.
void dedup_leaf(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a, uint64_t* d_voxels);

// Must the voxelizer track the first touches (t*) of the leaf level for a batch whose smallest tile_seq is seqLo?  Not
// when the batch comes after everything the table has seen: then only voxel masks without an entry would need them.
bool leaf_tstar_needed(const LevelTable& T, uint32_t seqLo);
// Leaf level without t* (a.tstar may be null): true when every voxel mask of the batch already had an entry -- the level
// is then reduced (voxels counted, batch recorded).  false: nothing was changed; re-voxelize with t* and use dedup_leaf.
// The few leaf nodes whose voxel mask has no entry yet get their first touch from a direct query (LeafQuery: the first
// root pair of the node's tile whose triangle passes the reference-order predicate at every box on the node's path);
// a batch with more than a few 10^5 such nodes is refused instead.
struct LeafQuery {
	const float* tris = nullptr;        // 9 floats per triangle
	const uint32_t* rootTri = nullptr;  // root pair q -> triangle
	uint64_t P = 0;                     // root pairs of the batch
	const void* tiles = nullptr;        // TileGeom per tile of the batch
};
bool dedup_leaf_known(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a, const LeafQuery& lq, uint64_t* d_voxels);
// May the voxelizer skip the first touches of the 4^3 level for a batch whose smallest tile_seq is seqLo?  Yes when the
// batch comes after everything the level's table has seen (its existing entries are frozen) and the single-pass insert
// is in use: only nodes of NEW entries need a first touch then, and dedup_level queries those directly (DedupArgs::query).
bool k64_tstar_optional(const LevelTable& T, uint32_t seqLo);

// KIND_K64 / KIND_INNER.  Throws Error(SVB_ECOLLISION) if the exact verify pass finds two different
// keys behind one 64-bit tag.
void dedup_level(cudaStream_t s, Pool& pool, LevelTable& T, const DedupArgs& a);

// ---- multi-GPU merge of one level (see svb_dedup.cu): record size / count to exchange, export of this
// rank's unique nodes with global child uids, import of the all-gathered union into a fresh global table.
uint32_t merge_rec_bytes(int kind);
uint64_t merge_count(const LevelTable& T);
void merge_export(cudaStream_t s, Pool& pool, LevelTable& T, const uint32_t* l2gChild, void* d_out);
void merge_import(cudaStream_t s, Pool& pool, LevelTable& T, const void* d_all, const uint64_t* counts, uint32_t world, uint64_t strideBytes,
                  uint32_t myRank, DevBuf<uint32_t>& l2g, uint32_t* d_status, uint64_t mergeSeed = 0);
void merge_resolve(LevelTable& T, const uint32_t status[5]);

// Level 0 is never reduced (geom_octree.cpp:483): just resolve the root's children.
// rootKey: 8 x u32 (uids, or child masks when childMode is a MASK mode), NULLREF = none.
void root_key(cudaStream_t s, const DedupArgs& a, uint32_t* d_rootKey8);

// Rank the unique nodes of every level by order key and materialise the DAG levels.
// tables[g] for g in [1, L-1]; obits[g] = significant bits of the order keys of level g.
void finalize_levels(cudaStream_t s, Pool& pool, std::vector<LevelTable>& tables, const std::vector<int>& obits,
                     const uint32_t* d_rootKey8, int rootChildMode, std::vector<OutLevel>& out);

}  // namespace svb
