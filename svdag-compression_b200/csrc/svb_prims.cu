// svb_prims.cu -- scan and radix-sort primitives (hand-written; no CUB/Thrust on the hot path).
//
//  * exclusive scans: reduce-then-scan over 2048-item tiles (block sums -> single-block
//    scan of the sums in 64 bit -> re-scan with the tile offset).  Inputs are read twice
//    (1 B or 4 B per item), outputs written once: HBM-bound, 2 passes.
//  * radix sort: stable LSD, 8-bit digits, (u64 key, u32 payload); used to rank the unique
//    nodes of a level by their order key (U_l elements, not N_l).
#include "svb_classify.cuh"
#include <cstdlib>

#include "svb_internal.cuh"

namespace svb {

std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct Popc8In {
	const uint8_t* p;
	__device__ void load(uint64_t base, uint64_t n, uint32_t v[SCAN_ITEMS]) const {
		if (base + SCAN_ITEMS <= n) {
			uint2 w = *reinterpret_cast<const uint2*>(p + base);   // base is a multiple of 8, p 16B aligned
			uint32_t a = w.x, b = w.y;
#pragma unroll
			for (int i = 0; i < 4; ++i) { v[i] = __popc((a >> (8 * i)) & 0xFF); v[4 + i] = __popc((b >> (8 * i)) & 0xFF); }
		} else {
#pragma unroll
			for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? __popc((uint32_t)p[base + i]) : 0;
		}
	}
};
struct U32In {
	const uint32_t* p;
	__device__ void load(uint64_t base, uint64_t n, uint32_t v[SCAN_ITEMS]) const {
#pragma unroll
		for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? p[base + i] : 0;
	}
};

__device__ inline uint32_t warp_incl_scan(uint32_t x, int lane) {
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
		if (lane >= d) x += y;
	}
	return x;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
__device__ inline uint32_t block_excl_scan(uint32_t x, uint32_t* total) {
	__shared__ uint32_t wsum[SCAN_THREADS / 32];
	__shared__ uint32_t wtot;
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = warp_incl_scan(x, lane);
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t s = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0;
		uint32_t si = warp_incl_scan(s, lane);
		if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
		if (lane == SCAN_THREADS / 32 - 1) wtot = si;
	}
	__syncthreads();
	uint32_t r = inc - x + wsum[w];
	*total = wtot;
	__syncthreads();
	return r;
}

// rel (optional): the exclusive prefix of every thread's 8 items inside its tile, i.e. tile-relative offsets at a granularity of
// 8 items -- what lets the consumer of the offsets (k_emit_warp, svb_voxelize.cu) work warp by warp without a CTA-wide scan
template <class In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(In in, uint64_t n, uint32_t* blockSums, uint32_t* __restrict__ rel = nullptr) {
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	in.load(base, n, v);
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
	uint32_t tot;
	const uint32_t ex = block_excl_scan(s, &tot);
	if (rel) rel[(uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x] = ex;
	if (threadIdx.x == 0) blockSums[blockIdx.x] = tot;
}

// single block: exclusive scan of nb u32 sums into u64 offsets, grand total to *total (SCAN_TILE sums per trip)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const uint32_t* __restrict__ sums, uint64_t nb, uint64_t* __restrict__ offs, uint64_t* __restrict__ total) {
	__shared__ uint64_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint64_t base = 0; base < nb; base += SCAN_TILE) {
		const uint64_t i0 = base + (uint64_t)threadIdx.x * SCAN_ITEMS;
		uint32_t v[SCAN_ITEMS];
		uint32_t x = 0;
#pragma unroll
		for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = (i0 + j < nb) ? sums[i0 + j] : 0; x += v[j]; }
		uint32_t tot;
		uint32_t ex = block_excl_scan(x, &tot);
		uint64_t o = carry + ex;
#pragma unroll
		for (int j = 0; j < SCAN_ITEMS; ++j) { if (i0 + j < nb) offs[i0 + j] = o; o += v[j]; }
		__syncthreads();
		if (threadIdx.x == 0) carry += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}

// The same with 1024 threads x 16 sums per trip and 64-bit arithmetic throughout: the single-CTA pass over the tile sums
// is a serial chain of trips (two barriers each), 4 of them per level -- 8x fewer trips (default; SVB_SCAN_WIDE=0: the
// kernel above).
constexpr int SW_THREADS = 1024, SW_ITEMS = 16;
__device__ __forceinline__ void scan_sums_wide_body(const uint32_t* __restrict__ sums, uint64_t nb, uint64_t* __restrict__ offs, uint64_t* __restrict__ total) {
	__shared__ unsigned long long wsum[SW_THREADS / 32];
	__shared__ unsigned long long carry;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint64_t base = 0; base < nb; base += (uint64_t)SW_THREADS * SW_ITEMS) {
		const uint64_t i0 = base + (uint64_t)threadIdx.x * SW_ITEMS;
		uint32_t v[SW_ITEMS];
		if (i0 + SW_ITEMS <= nb) {   // i0 is a multiple of 16: 64-byte aligned
#pragma unroll
			for (int j = 0; j < SW_ITEMS / 4; ++j) {
				const uint4 q = reinterpret_cast<const uint4*>(sums + i0)[j];
				v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
			}
		} else {
#pragma unroll
			for (int j = 0; j < SW_ITEMS; ++j) v[j] = (i0 + j < nb) ? sums[i0 + j] : 0;
		}
		unsigned long long x = 0;
#pragma unroll
		for (int j = 0; j < SW_ITEMS; ++j) x += v[j];
		unsigned long long inc = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
			if (lane >= d) inc += y;
		}
		const unsigned long long c0 = carry;   // (written after the second barrier of the previous trip)
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		if (w == 0) {
			const unsigned long long sv = wsum[lane];
			unsigned long long si = sv;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, si, d);
				if (lane >= d) si += y;
			}
			wsum[lane] = si - sv;
			if (lane == 31) carry = c0 + si;
		}
		__syncthreads();
		unsigned long long o = c0 + wsum[w] + (inc - x);
		if (i0 + SW_ITEMS <= nb) {
#pragma unroll
			for (int j = 0; j < SW_ITEMS; j += 2) {
				ulonglong2 pr;
				pr.x = o; o += v[j];
				pr.y = o; o += v[j + 1];
				reinterpret_cast<ulonglong2*>(offs + i0)[j >> 1] = pr;
			}
		} else {
#pragma unroll
			for (int j = 0; j < SW_ITEMS; ++j) { if (i0 + j < nb) offs[i0 + j] = o; o += v[j]; }
		}
		__syncthreads();   // wsum / carry are rewritten by the next trip
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SW_THREADS) k_scan_sums_wide(const uint32_t* __restrict__ sums, uint64_t nb, uint64_t* __restrict__ offs, uint64_t* __restrict__ total) {
	scan_sums_wide_body(sums, nb, offs, total);
}
// up to four independent scans in one launch, one CTA each: the four tile-sum scans of a voxelizer level used to be four
// launches of a single CTA in a row on the stream (~10 us each, ~800 per build, the GPU idle behind each of them)
struct ScanJobs4 {
	const uint32_t* sums[4];
	uint64_t nb[4];
	uint64_t* offs[4];
	uint64_t* total[4];
};
__global__ void __launch_bounds__(SW_THREADS) k_scan_sums_wide_multi(ScanJobs4 J) {
	const int j = blockIdx.x;
	if (J.nb[j] == 0) { if (threadIdx.x == 0) *J.total[j] = 0; return; }
	scan_sums_wide_body(J.sums[j], J.nb[j], J.offs[j], J.total[j]);
}
static void launch_scan_sums(cudaStream_t s, const uint32_t* sums, uint64_t nb, uint64_t* offs, uint64_t* total) {
	const char* e = getenv("SVB_SCAN_WIDE");
	if (e && e[0] == '0') k_scan_sums<<<1, SCAN_THREADS, 0, s>>>(sums, nb, offs, total);
	else k_scan_sums_wide<<<1, SW_THREADS, 0, s>>>(sums, nb, offs, total);
}

// pair streams: per tile of SCAN_TILE pairs, the number of child pairs (popcount of the hit masks) and the number of
// those whose flags put them into the flat stream (pair_is_fast, svb_classify.cuh), in one pass
// (flatTri != 0: the second count is over the pairs of FLAT triangles instead -- the ones k_slow_leaves decides in place)
__global__ void __launch_bounds__(SCAN_THREADS) k_pair_reduce(const uint8_t* __restrict__ hit, const uint16_t* __restrict__ fl, uint64_t n,
                                                              uint32_t* __restrict__ sumsA, uint32_t* __restrict__ sumsB, uint32_t* __restrict__ rel, int flatTri) {
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t a = 0, b = 0;
	if (base + SCAN_ITEMS <= n) {
		uint2 w = *reinterpret_cast<const uint2*>(hit + base);     // base is a multiple of 8, hit and fl 16B aligned
		if (w.x | w.y) {
			uint4 f = *reinterpret_cast<const uint4*>(fl + base);
			const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
			for (int i = 0; i < SCAN_ITEMS; ++i) {
				uint32_t c = __popc(((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xFF);
				uint32_t g = (fw[i >> 1] >> (16 * (i & 1))) & 0xFFFF;
				a += c;
				if (flatTri ? (g & (7u << FL_FLAT)) != 0 : pair_is_fast(g)) b += c;
			}
		}
	} else {
		for (int i = 0; i < SCAN_ITEMS; ++i)
			if (base + i < n) {
				uint32_t c = __popc((uint32_t)hit[base + i]);
				a += c;
				if (c && (flatTri ? (fl[base + i] & (7u << FL_FLAT)) != 0 : pair_is_fast(fl[base + i]))) b += c;
			}
	}
	uint32_t packed = a | (b << 16);   // a, b <= 8 * SCAN_ITEMS per thread, <= 8 * SCAN_TILE = 16384 per tile: fits 16 bits each
	uint32_t tot;
	const uint32_t ex = block_excl_scan(packed, &tot);   // (an exclusive prefix stays below 16384 in either half: no carry between them)
	if (rel) rel[(uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x] = ex;
	if (threadIdx.x == 0) { sumsA[blockIdx.x] = tot & 0xFFFF; sumsB[blockIdx.x] = tot >> 16; }
}

template <class In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(In in, uint64_t n, const uint64_t* offs, uint32_t* out) {
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	in.load(base, n, v);
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
	uint32_t tot;
	uint32_t ex = block_excl_scan(s, &tot);
	uint64_t o = offs[blockIdx.x] + ex;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		if (base + i < n) out[base + i] = (uint32_t)o;
		o += v[i];
	}
}

template <class In>
void scan_impl(cudaStream_t s, Pool& pool, In in, uint64_t n, uint32_t* out, uint64_t* d_total) {
	if (n == 0) { SVB_CUDA(cudaMemsetAsync(d_total, 0, 8, s)); return; }
	uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
	DevBuf<uint32_t> sums(pool, nb);
	DevBuf<uint64_t> offs(pool, nb);
	k_scan_reduce<In><<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, sums.p);
	SVB_KERNEL_CHECK();
	launch_scan_sums(s, sums.p, nb, offs.p, d_total);
	SVB_KERNEL_CHECK();
	k_scan_apply<In><<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, offs.p, out);
	SVB_KERNEL_CHECK();
}

// ---------------------------------------------------------------- radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_CHUNK = RS_THREADS * RS_ROUNDS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* keys, uint64_t n, int shift, uint32_t* hist, unsigned nblocks) {
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	uint64_t base = (uint64_t)blockIdx.x * RS_CHUNK;
	for (int r = 0; r < RS_ROUNDS; ++r) {
		uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1u);
	}
	__syncthreads();
	hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* keys, const uint32_t* vals, uint64_t n, int shift,
                                                            const uint32_t* histScan, unsigned nblocks, uint64_t* okeys, uint32_t* ovals) {
	__shared__ uint32_t off[256];
	__shared__ uint32_t wcnt[RS_WARPS][256];
	off[threadIdx.x] = histScan[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint64_t base = (uint64_t)blockIdx.x * RS_CHUNK;
	for (int r = 0; r < RS_ROUNDS; ++r) {
		for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_THREADS) (&wcnt[0][0])[j] = 0;
		__syncthreads();
		uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		bool valid = i < n;
		uint64_t k = valid ? keys[i] : 0;
		uint32_t v = valid ? vals[i] : 0;
		unsigned d = valid ? (unsigned)((k >> shift) & 0xFF) : 256u + lane;   // invalid lanes never match
		unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
		unsigned below = __popc(peers & ((1u << lane) - 1));
		if (valid && below == 0) wcnt[w][d] = __popc(peers);
		__syncthreads();
		if (valid) {
			uint32_t pos = off[d] + below;
			for (int ww = 0; ww < w; ++ww) pos += wcnt[ww][d];
			okeys[pos] = k;
			ovals[pos] = v;
		}
		__syncthreads();
		{
			uint32_t add = 0;
#pragma unroll
			for (int ww = 0; ww < RS_WARPS; ++ww) add += wcnt[ww][threadIdx.x];
			off[threadIdx.x] += add;
		}
		__syncthreads();
	}
}

}  // namespace

void scan_popc8(cudaStream_t s, Pool& pool, const uint8_t* bytes, uint64_t n, uint32_t* out, uint64_t* d_total) {
	scan_impl(s, pool, Popc8In{bytes}, n, out, d_total);
}
uint64_t scan_tile_items() { return SCAN_TILE; }

void scan_tiles_popc8(cudaStream_t s, Pool& pool, const uint8_t* bytes, uint64_t n, DevBuf<uint64_t>& tileOffs, uint64_t* d_total, DevBuf<uint32_t>* rel) {
	uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
	tileOffs.reset(pool, nb ? nb : 1);
	if (rel) rel->reset(pool, nb ? nb * SCAN_THREADS : 1);
	if (n == 0) { SVB_CUDA(cudaMemsetAsync(d_total, 0, 8, s)); return; }
	DevBuf<uint32_t> sums(pool, nb);
	k_scan_reduce<Popc8In><<<(unsigned)nb, SCAN_THREADS, 0, s>>>(Popc8In{bytes}, n, sums.p, rel ? rel->p : nullptr);
	SVB_KERNEL_CHECK();
	launch_scan_sums(s, sums.p, nb, tileOffs.p, d_total);
	SVB_KERNEL_CHECK();
}

void scan_tiles_pairs(cudaStream_t s, Pool& pool, const uint8_t* hit, const uint16_t* flags, uint64_t n,
                      DevBuf<uint64_t>& tileOffsA, DevBuf<uint64_t>& tileOffsB, uint64_t* d_totalA, uint64_t* d_totalB, DevBuf<uint32_t>* rel, int flatTri) {
	uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
	tileOffsA.reset(pool, nb ? nb : 1);
	tileOffsB.reset(pool, nb ? nb : 1);
	if (rel) rel->reset(pool, nb ? nb * SCAN_THREADS : 1);
	if (n == 0) { SVB_CUDA(cudaMemsetAsync(d_totalA, 0, 8, s)); SVB_CUDA(cudaMemsetAsync(d_totalB, 0, 8, s)); return; }
	DevBuf<uint32_t> sumsA(pool, nb), sumsB(pool, nb);
	k_pair_reduce<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(hit, flags, n, sumsA.p, sumsB.p, rel ? rel->p : nullptr, flatTri);
	SVB_KERNEL_CHECK();
	launch_scan_sums(s, sumsA.p, nb, tileOffsA.p, d_totalA);
	SVB_KERNEL_CHECK();
	launch_scan_sums(s, sumsB.p, nb, tileOffsB.p, d_totalB);
	SVB_KERNEL_CHECK();
}

void scan_level_tiles(cudaStream_t s, Pool& pool, const uint8_t* nodeMask, uint64_t nNodes, const uint8_t* hitF, uint64_t nF,
                      const uint8_t* hitS, const uint16_t* flagsS, uint64_t nS,
                      DevBuf<uint64_t>& nodeOffs, DevBuf<uint64_t>& offF, DevBuf<uint64_t>& offS, DevBuf<uint64_t>& offSF,
                      DevBuf<uint32_t>* relF, DevBuf<uint32_t>* relS, uint64_t* d_tot4, int flatTri) {
	const uint64_t nbN = (nNodes + SCAN_TILE - 1) / SCAN_TILE, nbF = (nF + SCAN_TILE - 1) / SCAN_TILE, nbS = (nS + SCAN_TILE - 1) / SCAN_TILE;
	nodeOffs.reset(pool, nbN ? nbN : 1);
	offF.reset(pool, nbF ? nbF : 1);
	offS.reset(pool, nbS ? nbS : 1);
	offSF.reset(pool, nbS ? nbS : 1);
	if (relF) relF->reset(pool, nbF ? nbF * SCAN_THREADS : 1);
	if (relS) relS->reset(pool, nbS ? nbS * SCAN_THREADS : 1);
	DevBuf<uint32_t> sumsN(pool, nbN ? nbN : 1), sumsF(pool, nbF ? nbF : 1), sumsA(pool, nbS ? nbS : 1), sumsB(pool, nbS ? nbS : 1);
	if (nbN) { k_scan_reduce<Popc8In><<<(unsigned)nbN, SCAN_THREADS, 0, s>>>(Popc8In{nodeMask}, nNodes, sumsN.p, nullptr); SVB_KERNEL_CHECK(); }
	if (nbF) { k_scan_reduce<Popc8In><<<(unsigned)nbF, SCAN_THREADS, 0, s>>>(Popc8In{hitF}, nF, sumsF.p, relF ? relF->p : nullptr); SVB_KERNEL_CHECK(); }
	if (nbS) { k_pair_reduce<<<(unsigned)nbS, SCAN_THREADS, 0, s>>>(hitS, flagsS, nS, sumsA.p, sumsB.p, relS ? relS->p : nullptr, flatTri); SVB_KERNEL_CHECK(); }
	ScanJobs4 J;
	J.sums[0] = sumsN.p; J.nb[0] = nbN; J.offs[0] = nodeOffs.p; J.total[0] = d_tot4 + 0;
	J.sums[1] = sumsF.p; J.nb[1] = nbF; J.offs[1] = offF.p; J.total[1] = d_tot4 + 1;
	J.sums[2] = sumsA.p; J.nb[2] = nbS; J.offs[2] = offS.p; J.total[2] = d_tot4 + 2;
	J.sums[3] = sumsB.p; J.nb[3] = nbS; J.offs[3] = offSF.p; J.total[3] = d_tot4 + 3;
	k_scan_sums_wide_multi<<<4, SW_THREADS, 0, s>>>(J);
	SVB_KERNEL_CHECK();
}

void scan_u32(cudaStream_t s, Pool& pool, const uint32_t* in, uint64_t n, uint32_t* out, uint64_t* d_total) {
	scan_impl(s, pool, U32In{in}, n, out, d_total);
}

void radix_sort_pairs(cudaStream_t s, Pool& pool, uint64_t* keys, uint32_t* vals, uint64_t n, int bits) {
	if (n <= 1) return;
	if (bits > 64) bits = 64;
	int passes = (bits + 7) / 8;
	if (passes < 1) passes = 1;
	unsigned nb = (unsigned)((n + RS_CHUNK - 1) / RS_CHUNK);
	DevBuf<uint64_t> k2(pool, n);
	DevBuf<uint32_t> v2(pool, n);
	DevBuf<uint32_t> hist(pool, (uint64_t)256 * nb);
	DevBuf<uint64_t> tot(pool, 1);
	uint64_t* ka = keys; uint32_t* va = vals;
	uint64_t* kb = k2.p; uint32_t* vb = v2.p;
	for (int p = 0; p < passes; ++p) {
		k_rs_hist<<<nb, RS_THREADS, 0, s>>>(ka, n, 8 * p, hist.p, nb);
		SVB_KERNEL_CHECK();
		scan_u32(s, pool, hist.p, (uint64_t)256 * nb, hist.p, tot.p);
		k_rs_scatter<<<nb, RS_THREADS, 0, s>>>(ka, va, n, 8 * p, hist.p, nb, kb, vb);
		SVB_KERNEL_CHECK();
		uint64_t* tk = ka; ka = kb; kb = tk;
		uint32_t* tv = va; va = vb; vb = tv;
	}
	if (ka != keys) {
		SVB_CUDA(cudaMemcpyAsync(keys, ka, n * 8, cudaMemcpyDeviceToDevice, s));
		SVB_CUDA(cudaMemcpyAsync(vals, va, n * 4, cudaMemcpyDeviceToDevice, s));
	}
}

}  // namespace svb
