// svb_encode.cu -- the reference's file formats written on the GPU.
//
//   EncodedSVDAG::encode    src/symvox/encoded_svdag.cpp:105-199     .svdag / -multi.svdag
//   EncodedUSSVDAG::encode  src/symvox/encoded_ussvdag.cpp:86-170    .ussvdag
//   EncodedSSVDAG::encode   src/symvox/encoded_ssvdag.cpp:194-466    .ssvdag / .esvdag
//
// The octree stays in HBM; what crosses PCIe is the finished file image (a 16K^3 city: 15 MB instead of 28 MB of node
// arrays), landing in a context-owned pinned buffer that svb_encode_view() hands out without another copy.
//
// .svdag / .ussvdag: one u32 stream, node = header word + one word per child for k = 7..0 holding the ABSOLUTE word
// offset of the child.  Sizes -> exclusive scan over all nodes of all levels -> every node writes its own words.
//
// .ssvdag: per level the nodes are ordered by reference count, descending, with the reference's UNSTABLE std::sort -- the
// tie order is part of the file.  The counts are histogrammed here (one atomicAdd per child pointer), travel to the host
// (4 B/node), are ordered by libstdc++'s own sort routines (csrc/host/encoders.cpp::ssvdag_order_from_refs: the very
// permutation the reference's call produces) and come back as a rank -> node table.  Then, bottom-up, pass A gives every
// node its address inside its encoded level (a node's size depends on its children's addresses: 16-bit pointer below
// 2^13, else 32-bit), and pass B writes headers, pointers and the 4^3 leaf bricks straight into the final image.
#include <cstring>

#include "svb_context.cuh"
#include "svb_encode.cuh"
#include "host/octree_data.hpp"

namespace svb {

namespace {

constexpr int ENC_THREADS = 256;
constexpr int ENC_MAX_LEVELS = 32;

struct EncLevels {   // kernel parameter (by value): the levels of the octree as device SoA pointers
	int L;
	uint32_t start[ENC_MAX_LEVELS + 1];   // index of a level's first node in the concatenation of all levels
	const uint8_t* mask[ENC_MAX_LEVELS];
	const uint32_t* child[ENC_MAX_LEVELS];
	const uint8_t* mirror[ENC_MAX_LEVELS];
	const uint32_t* childLevel[ENC_MAX_LEVELS];   // null: every child lives one level down
};

__device__ __forceinline__ int level_of(const EncLevels& E, uint32_t g) {
	int l = 0;
	while (l + 1 < E.L && g >= E.start[l + 1]) ++l;
	return l;
}

// ------------------------------------------------------------------ .svdag / .ussvdag
__global__ void __launch_bounds__(ENC_THREADS) k_ps_sizes(EncLevels E, uint32_t total, uint32_t* __restrict__ sz) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= total) return;
	const int l = level_of(E, g);
	sz[g] = (l + 1 < E.L) ? 1u + (uint32_t)__popc(E.mask[l][g - E.start[l]]) : 1u;   // leaf level: the mask word only
}
__global__ void __launch_bounds__(ENC_THREADS) k_ps_fill(EncLevels E, uint32_t total, const uint32_t* __restrict__ wordOf, int withMirror, uint32_t* __restrict__ data) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= total) return;
	const int l = level_of(E, g);
	const uint32_t i = g - E.start[l];
	uint32_t head = E.mask[l][i];
	if (withMirror) head |= ((uint32_t)E.mirror[l][i * 3ull + 2] << 24) | ((uint32_t)E.mirror[l][i * 3ull + 1] << 16) | ((uint32_t)E.mirror[l][i * 3ull] << 8);   // encoded_ussvdag.cpp:126-130
	uint32_t w = wordOf[g];
	data[w++] = head;
	if (l + 1 >= E.L) return;
	const uint32_t* ch = E.child[l] + i * 8ull;
	const uint32_t* cl = (!withMirror && E.childLevel[l]) ? E.childLevel[l] + i * 8ull : nullptr;   // USSVDAG ignores childLevels (encoded_ussvdag.cpp:137)
#pragma unroll
	for (int k = 7; k >= 0; --k) {
		const uint32_t c = ch[k];
		if (c == NULLNODE) continue;
		const uint32_t tl = cl ? cl[k] : (uint32_t)(l + 1);
		data[w++] = wordOf[E.start[tl] + c];   // encoded_svdag.cpp:155-170
	}
}

// ------------------------------------------------------------------ .ssvdag
// references from the level above (encoded_ssvdag.cpp:261-271); levels 1 .. L-2 are counted, the root keeps 0
__global__ void __launch_bounds__(ENC_THREADS) k_ss_refs(EncLevels E, uint32_t parents, uint32_t* __restrict__ refs) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;   // parents: nodes of levels 0 .. L-3
	if (g >= parents) return;
	const int l = level_of(E, g);
	const uint32_t* ch = E.child[l] + (uint64_t)(g - E.start[l]) * 8;
#pragma unroll
	for (int k = 0; k < 8; ++k) if (ch[k] != NULLNODE) atomicAdd(&refs[E.start[l + 1] + ch[k]], 1u);
}

// rank -> address for the brick level (one 8-byte brick per node, addressed by rank)
__global__ void __launch_bounds__(ENC_THREADS) k_ss_leaf_addr(uint32_t n, const uint32_t* __restrict__ order, uint32_t* __restrict__ addr) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n) addr[order[r]] = r;
}
// pass A: encoded size of the node at rank r: 1 header short + 1 or 2 shorts per child (encoded_ssvdag.cpp:352-430)
__global__ void __launch_bounds__(ENC_THREADS) k_ss_size(uint32_t n, const uint32_t* __restrict__ order, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ child,
                                                         const uint32_t* __restrict__ addrBelow, uint32_t* __restrict__ sz) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const uint32_t i = order[r];
	const unsigned m = mask[i];
	uint32_t s = 1;
#pragma unroll
	for (int c = 7; c >= 0; --c) {
		if (!((m >> c) & 1)) continue;
		const uint32_t a = addrBelow[child[i * 8ull + c]];
		s += (a < (1u << 13)) ? 1u : ((a < (1u << 30)) ? 2u : 0u);
	}
	sz[r] = s;
}
__global__ void __launch_bounds__(ENC_THREADS) k_ss_addr(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ off, uint32_t* __restrict__ addr) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n) addr[order[r]] = off[r];
}

// Layout of the file behind the 36-byte common header, in 16-bit units (everything in it is 2-byte aligned, not all of it
// 4-byte aligned):  u32 n16 | u16 inner[n16] | u32 nLeafBytes | u8 leaves[8 * nBricks] | u32 nOffsets | u32 levelOffsets[L-2]
struct SsLayout {
	uint32_t n16;                            // shorts of all inner levels
	uint32_t levelOffset[ENC_MAX_LEVELS];    // _levelOffsets (encoded_ssvdag.cpp:448-451)
	uint64_t bytes;                          // size of the region behind the common header
};
__global__ void k_ss_layout(int L, const uint64_t* __restrict__ totals, uint32_t nBricks, SsLayout* __restrict__ out) {
	if (threadIdx.x || blockIdx.x) return;
	uint64_t acc = 0;
	for (int i = 0; i < L - 2; ++i) { out->levelOffset[i] = (uint32_t)acc; acc += totals[i]; }
	out->n16 = (uint32_t)acc;
	out->bytes = 4 + 2 * acc + 4 + 8ull * nBricks + 4 + 4ull * (L - 2);
}
__device__ __forceinline__ void put_u32(uint16_t* p, uint32_t v) { p[0] = (uint16_t)(v & 0xFFFFu); p[1] = (uint16_t)(v >> 16); }   // little endian

// pass B, inner level: header + pointers of the node at rank r, at its final place in the image
__global__ void __launch_bounds__(ENC_THREADS) k_ss_fill(uint32_t n, int lev, const uint32_t* __restrict__ order, const uint32_t* __restrict__ off, const uint8_t* __restrict__ mask,
                                                         const uint32_t* __restrict__ child, const uint8_t* __restrict__ mirror, const uint32_t* __restrict__ addrBelow,
                                                         const SsLayout* __restrict__ lay, uint16_t* __restrict__ region) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const uint32_t i = order[r];
	uint16_t* enc = region + 2 + lay->levelOffset[lev];
	uint32_t w = off[r];
	const uint32_t headAt = w++;
	const unsigned m = mask[i];
	const unsigned mxm = mirror[i * 3ull], mym = mirror[i * 3ull + 1], mzm = mirror[i * 3ull + 2];
	uint32_t head = 0;
#pragma unroll
	for (int c = 7; c >= 0; --c) {
		if (!((m >> c) & 1)) continue;
		const uint32_t a = addrBelow[child[i * 8ull + c]];
		const uint32_t mx = (mxm >> c) & 1u, my = (mym >> c) & 1u, mz = (mzm >> c) & 1u;
		if (a < (1u << 13)) {
			head |= 1u << (2 * c);
			enc[w++] = (uint16_t)(a | (mx << 13) | (my << 14) | (mz << 15));
		} else if (a < (1u << 30)) {
			uint32_t pp = a;
			if (pp & (1u << 29)) { head |= 3u << (2 * c); pp &= ~(1u << 29); }   // bit 29 moves into the header (encoded_ssvdag.cpp:380-392)
			else head |= 2u << (2 * c);
			pp |= (mx << 29) | (my << 30) | (mz << 31);
			enc[w++] = (uint16_t)(pp >> 16);
			enc[w++] = (uint16_t)(pp & 0xFFFFu);
		}
	}
	enc[headAt] = (uint16_t)head;
}

// Node::mirror on a 2^3 voxel mask: slot i <- slot i ^ s (s = mx<<2 | my<<1 | mz, octree_node.cpp:88-195)
__device__ __forceinline__ unsigned mirror_mask8(unsigned m, unsigned s) {
	if (s & 1u) m = ((m & 0x55u) << 1) | ((m & 0xAAu) >> 1);
	if (s & 2u) m = ((m & 0x33u) << 2) | ((m & 0xCCu) >> 2);
	if (s & 4u) m = ((m & 0x0Fu) << 4) | ((m & 0xF0u) >> 4);
	return m;
}
// pass B, the two deepest levels fused into 4^3 bricks, bits re-ordered x-fastest (encoded_ssvdag.cpp:119-135, :280-351):
// brick bit x + 4y + 16z with (x, y, z) = 2 * (child slot bits 0, 1, 2) + (voxel bits 0, 1, 2)
__global__ void __launch_bounds__(ENC_THREADS) k_ss_bricks(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ child, const uint8_t* __restrict__ mirror,
                                                           const uint8_t* __restrict__ leafMask, const SsLayout* __restrict__ lay, uint16_t* __restrict__ region) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const uint32_t i = order[r];
	const unsigned mxm = mirror[i * 3ull], mym = mirror[i * 3ull + 1], mzm = mirror[i * 3ull + 2];
	uint64_t b = 0;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const uint32_t ch = child[i * 8ull + c];
		if (ch == NULLNODE) continue;
		const unsigned s = (((mxm >> c) & 1u) << 2) | (((mym >> c) & 1u) << 1) | ((mzm >> c) & 1u);
		const unsigned m = mirror_mask8(leafMask[ch], s);
		const uint64_t spread = (m & 0x03u) | ((m & 0x0Cu) << 2) | ((uint64_t)(m & 0x30u) << 12) | ((uint64_t)(m & 0xC0u) << 14);
		b |= spread << (2 * (c & 1) + 8 * ((c >> 1) & 1) + 32 * ((c >> 2) & 1));
	}
	uint16_t* dst = region + 2 + lay->n16 + 2 + 4ull * r;
	dst[0] = (uint16_t)b; dst[1] = (uint16_t)(b >> 16); dst[2] = (uint16_t)(b >> 32); dst[3] = (uint16_t)(b >> 48);
}
__global__ void k_ss_tail(int L, uint32_t nBricks, const SsLayout* __restrict__ lay, uint16_t* __restrict__ region) {
	if (threadIdx.x || blockIdx.x) return;
	put_u32(region, lay->n16);
	uint16_t* p = region + 2 + lay->n16;
	put_u32(p, nBricks * 8u);
	p += 2 + 4ull * nBricks;
	put_u32(p, (uint32_t)(L - 2));
	for (int i = 0; i < L - 2; ++i) put_u32(p + 2 + 2 * i, lay->levelOffset[i]);
}

EncLevels make_levels(svb_ctx* c) {
	EncLevels E;
	memset(&E, 0, sizeof(E));
	E.L = (int)c->levels;
	uint64_t acc = 0;
	for (int l = 0; l < E.L; ++l) {
		const OutLevel& o = c->out[l];
		E.start[l] = (uint32_t)acc;
		E.mask[l] = o.mask.p; E.child[l] = o.child.p; E.mirror[l] = o.mirror.p;
		E.childLevel[l] = o.hasChildLevel ? o.childLevel.p : nullptr;
		acc += o.n;
	}
	if (acc >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "octree too large for the 32-bit file formats");
	for (int l = E.L; l <= ENC_MAX_LEVELS; ++l) E.start[l] = (uint32_t)acc;
	return E;
}

void common_header(uint8_t* img, const svb_ctx* c) {   // EncodedOctree header: bbox, rootSide, levels, nNodes (low 4 bytes of the size_t)
	memcpy(img, c->stats.bboxF, 24);
	const float rs = (float)c->stats.rootSide;
	memcpy(img + 24, &rs, 4);
	const uint32_t L = c->levels, nn = (uint32_t)c->stats.nNodes;
	memcpy(img + 28, &L, 4);
	memcpy(img + 32, &nn, 4);
}

uint64_t encode_pointer_stream(svb_ctx* c, bool withMirror) {
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	const EncLevels E = make_levels(c);
	const uint32_t total = E.start[E.L];
	DevBuf<uint32_t> wordOf(pool, total ? total : 1);
	DevBuf<uint64_t> dWords(pool, 1);
	const unsigned nb = blocks_for(total, ENC_THREADS);
	k_ps_sizes<<<nb, ENC_THREADS, 0, s>>>(E, total, wordOf.p);
	SVB_KERNEL_CHECK();
	scan_u32(s, pool, wordOf.p, total, wordOf.p, dWords.p);
	uint64_t words = 0;
	SVB_CUDA(cudaMemcpyAsync(&words, dWords.p, 8, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	if (words >= 0xFFFFFFF0ull) throw Error(SVB_ERANGE, "more than 2^32 words in the pointer stream");
	DevBuf<uint32_t> data(pool, words ? words : 1);
	k_ps_fill<<<nb, ENC_THREADS, 0, s>>>(E, total, wordOf.p, withMirror ? 1 : 0, data.p);
	SVB_KERNEL_CHECK();
	const uint64_t size = 44 + 4 * words;
	uint8_t* img = c->image.reserve(size);
	SVB_CUDA(cudaMemcpyAsync(img + 44, data.p, 4 * words, cudaMemcpyDeviceToHost, s));
	common_header(img, c);
	const uint32_t w32 = (uint32_t)words;
	memcpy(img + 36, &w32, 4);   // _firstLeafPtr: computed after the leaf level too, so it equals the word count (encoded_svdag.cpp:129-137)
	memcpy(img + 40, &w32, 4);
	SVB_CUDA(cudaStreamSynchronize(s));
	return size;
}

uint64_t encode_ssvdag(svb_ctx* c) {
	cudaStream_t s = c->stream;
	Pool& pool = c->pool;
	const EncLevels E = make_levels(c);
	const int L = E.L;
	if (L < 3) throw Error(SVB_EINVAL, "SSVDAG needs at least 3 levels");
	for (int l = 0; l <= L - 2; ++l)
		if (c->out[l].n > (1ull << 30)) throw Error(SVB_EINVAL, "level too big for 30-bit pointers");   // encoded_ssvdag.cpp:251-254
	const uint32_t nOrdered = E.start[L - 1];   // nodes of levels 0 .. L-2
	const uint32_t nParents = E.start[L - 2];   // nodes of levels 0 .. L-3
	// ---- reference counts -> host -> node order -> device
	DevBuf<uint32_t> refs(pool, nOrdered ? nOrdered : 1), order(pool, nOrdered ? nOrdered : 1);
	refs.zero();
	if (nParents) {
		k_ss_refs<<<blocks_for(nParents, ENC_THREADS), ENC_THREADS, 0, s>>>(E, nParents, refs.p);
		SVB_KERNEL_CHECK();
	}
	uint32_t* hRefs = (uint32_t*)c->staging.reserve(8ull * nOrdered + 16);
	uint32_t* hOrder = hRefs + nOrdered;
	SVB_CUDA(cudaMemcpyAsync(hRefs, refs.p, 4ull * nOrdered, cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	svbhost::ssvdag_order_from_refs(hRefs, E.start, L - 1, hOrder);
	SVB_CUDA(cudaMemcpyAsync(order.p, hOrder, 4ull * nOrdered, cudaMemcpyHostToDevice, s));
	// ---- pass A, bottom-up: addresses
	DevBuf<uint32_t> addr(pool, nOrdered ? nOrdered : 1), off(pool, nOrdered ? nOrdered : 1);
	DevBuf<uint64_t> totals(pool, ENC_MAX_LEVELS);
	DevBuf<SsLayout> lay(pool, 1);
	totals.zero();
	const uint32_t nBricks = (uint32_t)c->out[L - 2].n;
	if (nBricks) {
		k_ss_leaf_addr<<<blocks_for(nBricks, ENC_THREADS), ENC_THREADS, 0, s>>>(nBricks, order.p + E.start[L - 2], addr.p + E.start[L - 2]);
		SVB_KERNEL_CHECK();
	}
	for (int lev = L - 3; lev >= 0; --lev) {
		const uint32_t n = (uint32_t)c->out[lev].n;
		if (!n) continue;
		const unsigned nb = blocks_for(n, ENC_THREADS);
		k_ss_size<<<nb, ENC_THREADS, 0, s>>>(n, order.p + E.start[lev], E.mask[lev], E.child[lev], addr.p + E.start[lev + 1], off.p + E.start[lev]);
		SVB_KERNEL_CHECK();
		scan_u32(s, pool, off.p + E.start[lev], n, off.p + E.start[lev], totals.p + lev);
		k_ss_addr<<<nb, ENC_THREADS, 0, s>>>(n, order.p + E.start[lev], off.p + E.start[lev], addr.p + E.start[lev]);
		SVB_KERNEL_CHECK();
	}
	k_ss_layout<<<1, 32, 0, s>>>(L, totals.p, nBricks, lay.p);
	SVB_KERNEL_CHECK();
	SsLayout hl;
	SVB_CUDA(cudaMemcpyAsync(&hl, lay.p, sizeof(SsLayout), cudaMemcpyDeviceToHost, s));
	SVB_CUDA(cudaStreamSynchronize(s));
	// ---- pass B: the image
	DevBuf<uint16_t> region(pool, hl.bytes / 2 + 8);
	for (int lev = L - 3; lev >= 0; --lev) {
		const uint32_t n = (uint32_t)c->out[lev].n;
		if (!n) continue;
		k_ss_fill<<<blocks_for(n, ENC_THREADS), ENC_THREADS, 0, s>>>(n, lev, order.p + E.start[lev], off.p + E.start[lev], E.mask[lev], E.child[lev], E.mirror[lev],
		                                                              addr.p + E.start[lev + 1], lay.p, region.p);
		SVB_KERNEL_CHECK();
	}
	if (nBricks) {
		k_ss_bricks<<<blocks_for(nBricks, ENC_THREADS), ENC_THREADS, 0, s>>>(nBricks, order.p + E.start[L - 2], E.child[L - 2], E.mirror[L - 2], E.mask[L - 1], lay.p, region.p);
		SVB_KERNEL_CHECK();
	}
	k_ss_tail<<<1, 32, 0, s>>>(L, nBricks, lay.p, region.p);
	SVB_KERNEL_CHECK();
	const uint64_t size = 36 + hl.bytes;
	uint8_t* img = c->image.reserve(size);
	SVB_CUDA(cudaMemcpyAsync(img + 36, region.p, hl.bytes, cudaMemcpyDeviceToHost, s));
	common_header(img, c);
	SVB_CUDA(cudaStreamSynchronize(s));
	return size;
}

}  // namespace

// kind: SVB_FILE_*; the image lands in c->image (pinned).  State checks as in the reference's encode() methods.
uint64_t encode_device(svb_ctx* c, int kind) {
	if (c->levels == 0 || c->levels > ENC_MAX_LEVELS) throw Error(SVB_EINVAL, "nothing to encode");
	if (kind == SVB_FILE_SVDAG) {
		if (c->state != SVB_S_DAG) throw Error(SVB_EINVAL, "FAILED! Octree is not in DAG state");          // encoded_svdag.cpp:109-112
		return encode_pointer_stream(c, false);
	}
	if (kind == SVB_FILE_USSVDAG) {
		if (c->state != SVB_S_SDAG) throw Error(SVB_EINVAL, "FAILED! Octree is not in SDAG state");        // encoded_ussvdag.cpp:90-93
		return encode_pointer_stream(c, true);
	}
	if (kind == SVB_FILE_SSVDAG) {
		if (c->state != SVB_S_DAG && c->state != SVB_S_SDAG) throw Error(SVB_EINVAL, "FAILED! Octree is not in SDAG state");   // encoded_ssvdag.cpp:218-221
		return encode_ssvdag(c);
	}
	throw Error(SVB_EINVAL, "unknown encoding");
}

}  // namespace svb
