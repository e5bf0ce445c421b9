/* svb.h -- C ABI of the B200-native svbuilder hot path (libsvb.so).
 *
 * The reference (RvanderLaan/SVDAG-Compression, SymVox) has no plugin/FFI layer: the seam
 * is the C++ class GeomOctree (src/symvox/geom_octree.hpp:107-157) called by
 * src/svbuilder/main.cpp:147-271 and read by the encoders through getNodeData()
 * (src/symvox/encoded_octree.hpp:37).  Each entry point below names the reference
 * interface it stands in for.  Plain pointers and sizes only; every call returns
 * SVB_OK (0) or a negative SVB_E* code, and svb_last_error() gives the text.
 * One context per GPU; calls on one context must be serialised by the caller (the
 * reference's GeomOctree is not re-entrant either: member _clock, geom_octree.hpp:104).
 *
 * There is NO CPU fallback: every function that computes runs CUDA kernels for sm_100a
 * and fails with SVB_ECUDA when no device is usable.
 */
#ifndef SVB_H
#define SVB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_OK          0
#define SVB_EINVAL     -1   /* bad argument / wrong state (reference prints "ERROR! This is not a SVO!", geom_octree.cpp:466-469) */
#define SVB_ECUDA      -2   /* CUDA runtime error (no device, launch failure, ...) */
#define SVB_ENOMEM     -3   /* device memory exhausted even after splitting the tile batch */
#define SVB_ERANGE     -4   /* configuration exceeds an id/order-key bit budget (see DESIGN.md "order key") */
#define SVB_ECOLLISION -5   /* 64-bit node-key hash collisions detected by the exact verify pass under five seeds in a row (svb_build, svb_to_sdag
                               and svb_cross_merge re-run the stage with another seed by themselves), or one detected in the multi-GPU level
                               merge (svb_shard_finish): there every rank sees it, and the caller repeats the build after svb_set_merge_seed */
#define SVB_ENODEV     -6   /* no usable CUDA device (there is no CPU fallback) */

/* encoded file kinds (svb_encode / svb_raycast_depth): what EncodedSVDAG / EncodedUSSVDAG / EncodedSSVDAG save */
#define SVB_FILE_SVDAG   0   /* .svdag and -multi.svdag */
#define SVB_FILE_USSVDAG 1   /* .ussvdag */
#define SVB_FILE_SSVDAG  2   /* .ssvdag and .esvdag */

#define SVB_NULL_NODE 0xFFFFFFFEu   /* Octree::nullNode, src/symvox/octree.cpp:22 */

/* GeomOctree::State, geom_octree.hpp:35-40 */
enum svb_state { SVB_S_EMPTY = 0, SVB_S_SVO = 1, SVB_S_DAG = 2, SVB_S_SDAG = 3 };

/* GeomOctree::Stats (geom_octree.hpp:42-58) -- counts only, plus device-side timings. */
typedef struct svb_stats {
	uint64_t nTotalVoxels;
	uint64_t nNodesSVO, nNodesDAG, nNodesSDAG;
	uint64_t nNodesLastLevSVO, nNodesLastLevDAG;
	uint64_t nCrossLevelMerged;
	uint64_t nNodes;            /* Octree::_nNodes as the encoders would read it now (getNNodes()) */
	uint64_t nTiles;            /* sub-octrees built ("Building %zu subtress", geom_octree.cpp:326) */
	uint64_t nBatches;          /* tile batches the build was split into (device-memory bound) */
	uint64_t nPairsTotal;       /* (triangle,node) pairs classified == 1/8 of the SAT tests run */
	double   rootSide;          /* Octree::_rootSide (float value in a double, geom_octree.cpp:184) */
	float    bboxF[6];          /* Octree::_bbox (float-converted, geom_octree.cpp:177-180) */
	double   msVoxelize, msDedup, msFinalize, msSdag, msCrossMerge, msTotal; /* CUDA-event times of the last call */
	uint64_t nKernelLaunches;   /* CUDA kernels launched by the last svb_build / svb_to_sdag / svb_cross_merge call */
	uint64_t nExactTests;       /* (triangle, child) tests the interval filter left to the reference-order predicate */
	uint64_t nHashRetries;      /* stages of this context re-run under another hash seed after a detected 64-bit tag collision (normally 0) */
} svb_stats;

typedef struct svb_ctx svb_ctx;

/* Lifetime.  device = CUDA ordinal this context binds to. */
svb_ctx*    svb_create(int device);
/* The same on a CUDA stream (cudaStream_t) the CALLER owns and keeps alive until svb_destroy: everything the context does is
 * enqueued there.  What a host that already has a stream discipline uses (a torch stream: tensors handed to the
 * svb_shard_* calls, NCCL collectives and the context's kernels are then ordered by the stream itself). */
svb_ctx*    svb_create_on_stream(int device, void* cuda_stream);
void        svb_destroy(svb_ctx* ctx);
const char* svb_last_error(const svb_ctx* ctx);   /* never NULL */
const char* svb_version(void);
/* The CUDA stream (cudaStream_t) every call on this context enqueues its work on.  The svb_shard_export_* /
 * svb_shard_import_* calls are STREAM-ORDERED (they return without waiting for the device): a caller orders its
 * collective after an export and before the matching import either by issuing it on this stream (NCCL:
 * ncclAllGather(..., (cudaStream_t)svb_stream(ctx)); torch: torch.cuda.ExternalStream) or with events.  All other calls
 * return with their results complete. */
void* svb_stream(const svb_ctx* ctx);
int   svb_synchronize(svb_ctx* ctx);

/* Scene::getTrianglePtr()/getNRawTriangles() (src/symvox/scene.hpp:90-91): the flat float32
 * triangle soup, 9 floats per triangle, in file order.  The host variant copies H2D (the
 * pointer is only read during the call); the device variant borrows a device pointer that
 * must stay valid until the next svb_set_triangles* / svb_destroy. */
int svb_set_triangles(svb_ctx* ctx, const float* xyz9_host, uint64_t ntris);
int svb_set_triangles_device(svb_ctx* ctx, const float* xyz9_dev, uint64_t ntris);

/* GeomOctree::buildSVO + toDAG when step == 0 (main.cpp:158-170, geom_octree.cpp:171-280,
 * 462-548), GeomOctree::buildDAG when step > 0 (main.cpp:173-175, geom_octree.cpp:289-435),
 * followed by initChildLevels() (main.cpp:203).  bbox is the double-widened scene bbox
 * (main.cpp:150-155).  Leaves the context in state DAG with node order identical to the
 * reference's for the same (levels, step). */
int svb_build(svb_ctx* ctx, uint32_t levels, uint32_t step,
              const double bbox_min[3], const double bbox_max[3], svb_stats* out);

/* ---- multi-GPU (one context per GPU, one process per GPU) -------------------------------------
 * GeomOctree::buildDAG's decomposition (geom_octree.cpp:331-378: independent sub-octrees; :397-425:
 * join + global toDAG) spread over `world` devices.  Protocol, identical on every rank:
 *   svb_shard_build(rank, world)        builds the base octree (redundantly) and this rank's share of the
 *                                       sub-octrees, reduced into rank-local tables;
 *   svb_shard_info                      levels to merge [first,last] (= step+1 .. levels-1), tile count, and
 *                                       this rank's counters {leafVoxels, nodesSVO, lastLevSVO, pairs, exact};
 *   for level = last .. first:          svb_shard_level_count -> n records of recBytes each (known for ALL levels as
 *                                       soon as svb_shard_build returns: one count exchange serves the whole merge);
 *                                       svb_shard_export_level writes them to a DEVICE buffer; the caller
 *                                       all-gathers (NCCL) into world rows of strideBytes;
 *                                       svb_shard_import_level rebuilds the identical global table everywhere;
 *                                       (export / import are stream-ordered on svb_stream(ctx), no host sync: errors
 *                                       of an import -- SVB_ECOLLISION etc. -- surface in svb_shard_finish)
 *   svb_shard_export_roots / all-gather / svb_shard_import_roots   (nTiles x u32 per rank);
 *   svb_shard_finish(totals)            totals = the counters summed over ranks; reduces the base octree,
 *                                       ranks all levels, leaves every rank in state DAG with the same octree
 *                                       a single-GPU svb_build would have produced.
 * Device pointers are plain CUDA pointers (e.g. torch tensors' data_ptr()).  world == 1 callers use svb_build. */
int svb_shard_build(svb_ctx* ctx, uint32_t levels, uint32_t step, const double bbox_min[3], const double bbox_max[3],
                    uint32_t rank, uint32_t world);
int svb_shard_info(const svb_ctx* ctx, uint32_t* firstLevel, uint32_t* lastLevel, uint64_t* nTiles, uint64_t counters[5]);
int svb_shard_level_count(const svb_ctx* ctx, uint32_t level, uint64_t* n, uint32_t* recBytes);
int svb_shard_export_level(svb_ctx* ctx, uint32_t level, void* d_out);
int svb_shard_import_level(svb_ctx* ctx, uint32_t level, const void* d_all, const uint64_t* counts, uint64_t strideBytes);
int svb_shard_export_roots(svb_ctx* ctx, void* d_out);
int svb_shard_import_roots(svb_ctx* ctx, const void* d_all);
int svb_shard_finish(svb_ctx* ctx, const uint64_t totals[5], svb_stats* out);
/* Seed of the hashed keys of the level merge (default 0).  Must be the same on every rank.  After svb_shard_finish has
 * returned SVB_ECOLLISION -- on every rank alike, since all of them import identical records -- set seed + 1 everywhere
 * and run the protocol again (svdag-compression_b200/sharded.py and csrc/host/sharded_build.cpp do). */
int svb_set_merge_seed(svb_ctx* ctx, uint64_t seed);

/* GeomOctree::toSDAG(false,false) (geom_octree.cpp:551-697).  DAG -> SDAG. */
int svb_to_sdag(svb_ctx* ctx, svb_stats* out);

/* GeomOctree::mergeAcrossAllLevels() (geom_octree_extension.cpp:1192-1544).  DAG -> DAG with
 * cross-level pointers (childLevels meaningful). */
int svb_cross_merge(svb_ctx* ctx, svb_stats* out);

/* GeomOctree::getState()/getLevels()/getStats() */
int svb_state(const svb_ctx* ctx);
uint32_t svb_levels(const svb_ctx* ctx);
int svb_get_stats(const svb_ctx* ctx, svb_stats* out);

/* GeomOctree::getNodeData() (geom_octree.hpp:117): one level as SoA, copied D2H into
 * caller-owned buffers (any pointer may be NULL): mask[n], child8[n*8] (SVB_NULL_NODE = no
 * child), mirror3[n*3] (childrenMirroredBitmask x,y,z), inv[n] (invariantBitmask),
 * childLevel8[n*8] (Octree::Node::childLevels). */
int svb_level_count(const svb_ctx* ctx, uint32_t lev, uint64_t* n);
int svb_download_level(svb_ctx* ctx, uint32_t lev, uint8_t* mask, uint32_t* child8,
                       uint8_t* mirror3, uint8_t* inv, uint32_t* childLevel8);
/* "Reduced level %u from %lu to %lu nodes" (geom_octree.cpp:509): node count of a level
 * before the final dedup pass (the SVO level size for step 0; the concatenated sub-DAG level
 * size is NOT reproduced for step > 0, where 0 is returned). */
int svb_level_count_svo(const svb_ctx* ctx, uint32_t lev, uint64_t* n);

/* GeomOctree(const NodeData&, bbox, rootSide, levels, stats) (geom_octree.cpp:49-60), i.e. what
 * EncodedSVDAG::decode feeds back in (main.cpp:97-101): replace the context's octree by
 * host-provided DAG levels (state DAG) so toSDAG / cross-merge can run from a pre-built DAG.
 * counts[levels]; mask/child8 are the per-level arrays concatenated level by level. */
int svb_upload_levels(svb_ctx* ctx, uint32_t levels, const uint64_t* counts,
                      const uint8_t* mask, const uint32_t* child8,
                      const float bboxF[6], double rootSide, uint64_t nVoxels);

/* EncodedSVDAG::load + decode (encoded_svdag.cpp:43-74, :200-270; used by svbuilder when its input is a .svdag,
 * main.cpp:96-101): parse a .svdag image (host memory) into DAG levels and make them the context's octree
 * (state DAG), ready for svb_to_sdag / svb_cross_merge / svb_encode.  nTotalVoxels is 0 afterwards, as in the
 * reference.  Cross-level (-multi) files cannot be decoded (nor can the reference). */
int svb_load_svdag(svb_ctx* ctx, const uint8_t* file, uint64_t size, svb_stats* out);
/* The same decode without a GPU or a context: call with mask = child8 = NULL to get *levels and counts[0..*levels)
 * (counts must hold 32 entries), then again with buffers of sum(counts) and 8*sum(counts) elements. */
int svb_decode_svdag(const uint8_t* file, uint64_t size, uint32_t* levels, uint64_t* counts, uint8_t* mask, uint32_t* child8,
                     float bboxF[6], double* rootSide, uint64_t* nNodes);

/* EncodedSVDAG / EncodedUSSVDAG / EncodedSSVDAG ::encode(const GeomOctree&) + save()
 * (encoded_svdag.cpp:76-199, encoded_ussvdag.cpp:60-170, encoded_ssvdag.cpp:84-117,194-466):
 * copies the levels D2H and writes the exact file image the reference would save.
 * kind: 0 = .svdag (state DAG), 1 = .ussvdag (state SDAG), 2 = .ssvdag / .esvdag (DAG or SDAG).
 * Returns the image size in bytes (copied into buf when cap is large enough; call with
 * buf = NULL to size), or a negative SVB_E* code.  The formats are written ON THE GPU (csrc/svb_encode.cu): only the
 * finished image crosses PCIe, into a pinned buffer owned by the context.  (.ssvdag: the per-level reference counts
 * make one round trip to the host, where libstdc++'s own sort routines reproduce the tie order of the reference's
 * unstable std::sort, encoded_ssvdag.cpp:273-275.) */
int64_t svb_encode(svb_ctx* ctx, int kind, uint8_t* buf, uint64_t cap);
/* The same without the final copy: *image points at the context's pinned image (valid until the next call that
 * changes the octree or encodes another kind); what svbuilder hands to fwrite (main.cpp:220-271). */
int svb_encode_view(svb_ctx* ctx, int kind, const uint8_t** image, uint64_t* size);

/* Same encoders, fed from host arrays (no GPU, no context): lets a caller that already holds
 * GeomOctree::NodeData (e.g. decoded from a .svdag, main.cpp:97-101) write the file formats, and
 * lets the CPU test-suite pin the encoders without a device.  Arrays are concatenated level by
 * level; mirror3 / childLevel8 may be NULL (zeros / lev+1). nNodes is Octree::_nNodes as the
 * reference would hold it at that point; state is a svb_state. */
int64_t svb_encode_levels(uint32_t levels, const uint64_t* counts, const uint8_t* mask, const uint32_t* child8,
                          const uint8_t* mirror3, const uint32_t* childLevel8, const float bboxF[6], double rootSide,
                          uint64_t nNodes, int state, int kind, uint8_t* buf, uint64_t cap);

/* The host step of the .ssvdag encoder on its own (no GPU, no context): node order of every level from the per-level
 * reference counts, i.e. std::sort by count, descending, UNSTABLE, with exactly the tie order libstdc++'s std::sort
 * gives the reference (encoded_ssvdag.cpp:261-276) -- computed by the same library routines with the independent
 * halves of every partition running as parallel tasks.  refs / order: levels 0 .. nLevels-1 concatenated, level l at
 * [start[l], start[l+1]); level 0 (the root) is not sorted.  order[start[l] + r] = index of the node at rank r. */
int svb_ssvdag_order_from_refs(const uint32_t* refs, const uint32_t* start, uint32_t nLevels, uint32_t* order);

/* ---- material-id leaves + Gray-coded attribute bit-trees (BASELINE.json configs[3]) -------------------
 * svb_build_svo_materials = GeomOctree::buildSVO(levels, bbox, false, NULL, putMaterialIdInLeaves = true)
 * (geom_octree.cpp:171-280): while voxelizing triangle iTri the reference stores Scene::getTriangleMaterialId(iTri) in the
 * leaf node's child slot of every voxel the triangle touches (:210-211, :252), so a voxel keeps the material of the LAST
 * triangle in file order that touches it.  triMaterial: host array, one id per triangle.  *nLeafNodes = nNodesLastLevSVO.
 * svb_download_leaf_materials copies the leaf level out IN THE REFERENCE'S NODE ORDER: mask[n], material8[n * 8] with
 * SVB_NULL_NODE in the slots of unset voxels -- exactly `_data[levels-1][i].childrenBitmask / .children[0..7]` after the
 * call (pinned against the unmodified reference through oracle/ref_attr_driver.cpp).
 * svb_attribute_bit_trees is SELF-SPECIFIED (the reference names the idea in readme.md:8 and ships no code): the attribute
 * code of a voxel is its material id (gray = 0) or the reflected Gray code of it (gray = 1, svb_gray_code); bit-tree b is
 * the sparse voxel DAG of the voxels whose code has bit b set, reduced per level like the geometry; nodes[b] / voxels[b]
 * (either may be NULL except nodes) receive its node count (root included, 0 for an empty tree) and voxel count. */
int svb_build_svo_materials(svb_ctx* ctx, uint32_t levels, const double bbox_min[3], const double bbox_max[3],
                            const uint32_t* triMaterial, uint64_t* nLeafNodes);
int svb_download_leaf_materials(svb_ctx* ctx, uint8_t* mask, uint32_t* material8);
int svb_attribute_bit_trees(svb_ctx* ctx, uint32_t nbits, int gray, uint64_t* nodes, uint64_t* voxels);
uint32_t svb_gray_code(uint32_t a);

/* Per-kernel profile of the last svb_build/svb_to_sdag (enabled by svb_set_profiling(ctx,1)):
 * one record per dedup-family launch group, with the algorithmic byte count SURVEY.md §8(d)
 * assigns to it.  Used by bench.py for the roofline object. */
typedef struct svb_prof_rec {
	char     name[32];     /* kernel family, e.g. "dedup_leaf", "dedup_k64", "dedup_inner", "classify" */
	uint32_t level;        /* global octree level */
	uint64_t n_in;         /* units in (nodes or pairs) */
	uint64_t n_out;        /* unique nodes found so far at that level (dedup) / pairs out */
	double   ms;           /* CUDA-event duration on the context's stream */
	double   bytes;        /* algorithmic bytes of this launch group */
} svb_prof_rec;
int svb_set_profiling(svb_ctx* ctx, int enabled);   /* 0 off, 1 records of the last build, 2 records accumulate over builds until svb_profile_clear,
                                                        3 as 2 but only the launches of the "emit" family are bracketed (a timed region that wants one kernel's rate) */
int svb_profile_clear(svb_ctx* ctx);
int svb_profile_count(const svb_ctx* ctx);
int svb_profile_get(const svb_ctx* ctx, int i, svb_prof_rec* out);

/* Cap on device bytes one tile batch may use for its transient (pair/node) buffers;
 * 0 = automatic (a fraction of free memory).  Smaller values force more batches (tests). */
int svb_set_batch_budget(svb_ctx* ctx, uint64_t bytes);

/* ---- depth-image sanity check (BASELINE.json north_star) ----------------------------
 * CUDA DDA ray caster over an encoded file image, following the reference viewer's fragment shader in
 * DEPTH_MODE (shaders/octree_dda.frag.glsl:484-585, :846-858) with the uniforms
 * src/svviewer/octree_dda_renderer.cpp:195-211 derives from the file header.  `file`/`size`: the bytes of a
 * .svdag / -multi.svdag / .ussvdag / .ssvdag / .esvdag file (host memory); kind: SVB_FILE_*;
 * viewInv / projInv: the shader's viewMatInv / projMatInv (column-major float[16]); drawLevel 0 = all levels;
 * projectionFactor: the viewer's LOD factor (octree_dda_renderer.cpp:503-507; 1e30 disables LOD).
 * out_host receives width*height*3 floats, pixel (x, y) at 3*(y*width + x): (t, level, iterations) of the
 * hit, or zeros where the shader discards the fragment.  Images are deterministic (no FMA contraction, IEEE
 * division/sqrt), so two files of the same scene can be compared pixel-exactly.  Stateless; runs on `device`. */
int svb_raycast_depth(int device, const uint8_t* file, uint64_t size, int kind, const float viewInv[16], const float projInv[16],
                      uint32_t width, uint32_t height, uint32_t maxIters, uint32_t drawLevel, float projectionFactor, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SVB_H */
