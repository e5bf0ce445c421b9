// svb_raycast.cu -- CUDA DDA ray caster over encoded SVDAG / USSVDAG / SSVDAG files: the depth-image sanity check
// BASELINE.json's north_star asks for ("a CUDA DDA ray-cast depth image of both outputs compared pixel-exact").
//
// Follows the reference viewer's fragment shader in DEPTH_MODE (shaders/octree_dda.frag.glsl:484-585 trace_ray,
// :376-481 DDA primitives, :158-185 / :190-216 / :221-345 node fetches, :592-608 camera ray, :846-858 main) with the
// uniforms src/svviewer/octree_dda_renderer.cpp:195-211 derives from the file header.  One thread per pixel, 8x8
// pixel tiles per CTA so that neighbouring rays walk the same nodes (the DAG words come through L1/L2).
//
// This translation unit is compiled with --fmad=false: every float operation is rounded separately, in the order the
// shader writes it, so the image can be compared bit for bit with the CPU restatement (oracle/dda_oracle.c) and images
// of different files (reference-built vs GPU-built, .svdag vs -multi.svdag vs .ussvdag) with each other.
// Conventions where GLSL is silent: mat4*vec4 and dot() sum left to right, normalize(v) = v / sqrt(dot(v,v)).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>

#include "../../include/svb.h"

namespace {

struct DagView {
	int kind;                      // SVB_FILE_SVDAG / USSVDAG / SSVDAG
	uint32_t levels, innerLevels;  // INNER_LEVELS = levels-1 (SVDAG, USSVDAG) or levels-2 (SSVDAG)
	float bbmin[3], bbmax[3], rootSide;
	const uint32_t* nodes;         // SVDAG / USSVDAG
	const uint16_t* inner;         // SSVDAG
	const uint2* leaves;           // SSVDAG 4^3 leaves
	uint32_t levelOffsets[32];
};

struct Cam { float vi[16], pi[16]; };

constexpr int RC_STACK = 34;

struct Trav {
	float t, cell;
	float tnext[3], inv[3];
	int idx[3], loc[3], dlt[3], mir[3];
	int node, size;
	uint32_t hdr, level, child;
	uint2 leaf;
};

__device__ __forceinline__ uint32_t linear_child(const Trav& s) {   // voxel_to_linear_idx, :149-152
	const int n1 = s.size - 1;
	const int x = s.mir[0] ? n1 - s.loc[0] : s.loc[0];
	const int y = s.mir[1] ? n1 - s.loc[1] : s.loc[1];
	const int z = s.mir[2] ? n1 - s.loc[2] : s.loc[2];
	return (uint32_t)(z + s.size * (y + s.size * x));
}

__device__ __forceinline__ int child_mask2(uint32_t hdr, uint32_t c) { return (int)((hdr >> (2 * c)) & 3u); }

template <int KIND>
__device__ __forceinline__ bool voxel_bit(const DagView& d, const Trav& s) {
	if (KIND != SVB_FILE_SSVDAG) return (s.hdr >> s.child) & 1u;                       // :172-174
	if (s.level < d.innerLevels) return child_mask2(s.hdr, s.child) != 0;               // :295-301
	const uint32_t w = (s.child & 32u) ? s.leaf.y : s.leaf.x;                           // :288-293
	return (w >> (s.child & 31u)) & 1u;
}

template <int KIND>
__device__ __forceinline__ void fetch(const DagView& d, Trav& s) {
	if (KIND != SVB_FILE_SSVDAG) s.hdr = d.nodes[s.node];                               // :176-178
	else if (s.level < d.innerLevels) s.hdr = d.inner[(uint32_t)s.node + d.levelOffsets[s.level]];   // :303-305
	else s.leaf = d.leaves[s.node];                                                      // :306-307
}

template <int KIND>
__device__ __forceinline__ void descend_pointer(const DagView& d, Trav& s) {
	if (KIND == SVB_FILE_SVDAG) {                                                        // :180-184
		s.node = (int)d.nodes[s.node + __popc((s.hdr & 0xFFu) >> s.child)];
	} else if (KIND == SVB_FILE_USSVDAG) {                                               // :206-215
		const uint32_t h = s.hdr;
		s.node = (int)d.nodes[s.node + __popc((h & 0xFFu) >> s.child)];
		s.mir[0] ^= (h >> (s.child + 8)) & 1u;
		s.mir[1] ^= (h >> (s.child + 16)) & 1u;
		s.mir[2] ^= (h >> (s.child + 24)) & 1u;
	} else {                                                                             // :310-345
		int off = 1 + (int)d.levelOffsets[s.level] + s.node;
		for (uint32_t i = 7; i > s.child; --i) off += min(child_mask2(s.hdr, i), 2);   // == childIndir table (renderer.cpp:533-545)
		int p = (int)d.inner[off];
		const int cm = child_mask2(s.hdr, s.child);
		s.mir[0] ^= (p >> 13) & 1; s.mir[1] ^= (p >> 14) & 1; s.mir[2] ^= (p >> 15) & 1;
		p &= ~(7 << 13);
		if (cm > 1) p = (int)(((uint32_t)p << 16) | (uint32_t)d.inner[off + 1]) | ((cm & 1) << 29);
		s.node = p;
	}
}

__device__ __forceinline__ void crossings(Trav& s, const float ro[3]) {                  // :386-389 / :427-431 / :447-451
#pragma unroll
	for (int a = 0; a < 3; ++a) s.tnext[a] = ((float)(s.idx[a] + max(s.dlt[a], 0)) * s.cell - ro[a]) * s.inv[a];
}

__device__ __forceinline__ void step_axis(const Trav& s, int m[3]) {                     // :398-402
	m[0] = (s.tnext[0] < s.tnext[1]) && (s.tnext[0] <= s.tnext[2]);
	m[1] = (s.tnext[1] < s.tnext[2]) && (s.tnext[1] <= s.tnext[0]);
	m[2] = (s.tnext[2] < s.tnext[0]) && (s.tnext[2] <= s.tnext[1]);
}

__device__ __forceinline__ void one_level_down(Trav& s, const float ro[3], const float rd[3]) {   // :439-455
	s.level++;
	s.cell *= 0.5f;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const float pa = ro[a] + s.t * rd[a];
		const float pc = (float)(s.idx[a] * 2 + 1) * s.cell;
		s.idx[a] = s.idx[a] * 2 + (pc < pa ? 1 : 0);
	}
	crossings(s, ro);
#pragma unroll
	for (int a = 0; a < 3; ++a) s.loc[a] = s.idx[a] & (s.size - 1);
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float minf2(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float maxf2(float a, float b) { return a < b ? b : a; }

__device__ __forceinline__ void mat_vec(const float* m, float x, float y, float z, float w, float o[4]) {
#pragma unroll
	for (int i = 0; i < 4; ++i) o[i] = ((m[i] * x + m[4 + i] * y) + m[8 + i] * z) + m[12 + i] * w;
}
__device__ __forceinline__ void normalize(float v[3]) {
	const float l = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
	v[0] = v[0] / l; v[1] = v[1] / l; v[2] = v[2] / l;
}

template <int KIND>
__global__ void __launch_bounds__(64) k_raycast(DagView d, Cam cam, uint32_t W, uint32_t H, uint32_t maxIters, uint32_t drawLevel, float projFactor, float* __restrict__ out) {
	const uint32_t tilesX = (W + 7) / 8;
	const uint32_t px = (blockIdx.x % tilesX) * 8 + (threadIdx.x & 7), py = (blockIdx.x / tilesX) * 8 + (threadIdx.x >> 3);
	if (px >= W || py >= H) return;
	float* o = out + 3ull * ((uint64_t)py * W + px);
	o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;   // "discard"
	// ---- camera ray (:592-608)
	const float sx = (((float)px + 0.5f) / (float)W) * 2.0f - 1.0f, sy = (((float)py + 0.5f) / (float)H) * 2.0f - 1.0f;
	float w0[4], w1[4], ro[3], rd[3];
	mat_vec(cam.pi, sx, sy, 0.f, 1.f, w0);
	mat_vec(cam.pi, sx, sy, 1.f, 1.f, w1);
#pragma unroll
	for (int a = 0; a < 3; ++a) rd[a] = w1[a] / w1[3] - w0[a] / w0[3];
	normalize(rd);
	mat_vec(cam.vi, 0.f, 0.f, 0.f, 1.f, w0);
	mat_vec(cam.vi, rd[0], rd[1], rd[2], 1.f, w1);
#pragma unroll
	for (int a = 0; a < 3; ++a) { ro[a] = w0[a]; rd[a] = w1[a] - w0[a]; }
	normalize(rd);
	// ---- transform_ray (:484-512)
	const float half = d.rootSide / 2.0f;
	const float scale = 1.0f / (2.0f * half);
	float tmin = 0.f * scale, tmax = 1e30f * scale;
	float tlo = 0.0f, thi = 0.f;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const float centre = (d.bbmin[a] + d.bbmax[a]) * 0.5f;
		const float omin = centre - half;
		const float sg = sgnf(rd[a]);
		ro[a] = (ro[a] - omin) * scale;
		if (rd[a] * sg < 1e-4f) rd[a] = sg * 1e-4f;
		const float t1 = ((d.bbmin[a] - omin) * scale - ro[a]) / rd[a], t2 = ((d.bbmax[a] - omin) * scale - ro[a]) / rd[a];
		const float lo = minf2(t1, t2), hi = maxf2(t1, t2);
		// intersectAABB :357-366: tMinMax = (max(max(lo.x, 0), max(lo.y, lo.z)), min(hi.x, min(hi.y, hi.z)))
		if (a == 0) { tlo = maxf2(lo, 0.0f); thi = hi; }
		else if (a == 1) { w0[0] = lo; w0[1] = hi; }
		else { tlo = maxf2(tlo, maxf2(w0[0], lo)); thi = minf2(thi, minf2(w0[1], hi)); }
	}
	tmin = maxf2(tlo, tmin + 1e-10f);
	tmax = minf2(thi, tmax);
	if (!(tlo < thi)) return;   // -4: out of the scene bbox
	// ---- init (:515-531) + dda_init (:376-390)
	Trav s;
	int stNode[RC_STACK], stMask[RC_STACK];
	uint32_t stHdr[RC_STACK];
	int sp = 0;
	s.t = tmin;
	s.level = 0; s.cell = 0.5f; s.size = 2; s.node = 0;
	s.hdr = 0; s.leaf = make_uint2(0, 0);
	{
		const float tt = s.t + 1.0f / (256.f * 1024.f);
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			s.inv[a] = 1.0f / rd[a];
			s.dlt[a] = (int)sgnf(rd[a]);
			s.mir[a] = 0;
			s.idx[a] = (int)((ro[a] + tt * rd[a]) / s.cell);
		}
		crossings(s, ro);
#pragma unroll
		for (int a = 0; a < 3; ++a) s.loc[a] = s.idx[a] % 2;
	}
	fetch<KIND>(d, s);
	s.child = linear_child(s);
	const uint32_t leafSize = (KIND == SVB_FILE_SSVDAG) ? 4 : 2;
	const uint32_t maxLevel = min(d.innerLevels, drawLevel - 1);
	const float oscale = 2.0f * half;
	uint32_t it = 0;
	// ---- trace_ray loop (:556-581)
	do {
		if (!voxel_bit<KIND>(d, s)) {
			int m[3];
			step_axis(s, m);                                                          // dda_next :397-414
#pragma unroll
			for (int a = 0; a < 3; ++a) { s.idx[a] += m[a] * s.dlt[a]; s.loc[a] += m[a] * s.dlt[a]; }
			s.t = ((float)m[0] * s.tnext[0] + (float)m[1] * s.tnext[1]) + (float)m[2] * s.tnext[2];
#pragma unroll
			for (int a = 0; a < 3; ++a) s.tnext[a] += (float)m[a] * s.cell * fabsf(s.inv[a]);
			const bool inside = s.loc[0] >= 0 && s.loc[1] >= 0 && s.loc[2] >= 0 && s.loc[0] < s.size && s.loc[1] < s.size && s.loc[2] < s.size;
			if (!inside) {
				if (sp == 0) return;                                                  // -1: left the scene without a hit
				--sp;                                                                 // up_in :423-437
				const uint32_t lev = ((uint32_t)stMask[sp] >> 3) & 255u;
				const uint32_t dl = s.level - lev;
				s.node = stNode[sp]; s.hdr = stHdr[sp]; s.level = lev;
				s.mir[0] = stMask[sp] & 1; s.mir[1] = (stMask[sp] >> 1) & 1; s.mir[2] = (stMask[sp] >> 2) & 1;
				s.cell *= (float)(1 << dl);
				s.size = lev < d.innerLevels ? 2 : (int)leafSize;
#pragma unroll
				for (int a = 0; a < 3; ++a) { s.idx[a] >>= dl; s.loc[a] = s.idx[a] & 1; }
				crossings(s, ro);
			}
		} else {
			if (s.level >= maxLevel || (s.cell * projFactor) < s.t) {                   // hit (resolution_ok :368-370)
				const float tv = s.t * oscale;
				if (tv > 0.f) { o[0] = tv; o[1] = (float)s.level; o[2] = (float)it; }   // DEPTH_MODE main :846-858
				return;
			}
			int m[3];                                                                  // down_in :457-481
			step_axis(s, m);
			const int nx = s.loc[0] + m[0] * s.dlt[0], ny = s.loc[1] + m[1] * s.dlt[1], nz = s.loc[2] + m[2] * s.dlt[2];
			if (nx >= 0 && ny >= 0 && nz >= 0 && nx < 2 && ny < 2 && nz < 2 && sp < RC_STACK) {
				stNode[sp] = s.node; stHdr[sp] = s.hdr;
				stMask[sp] = s.mir[0] | (s.mir[1] << 1) | (s.mir[2] << 2) | (int)(s.level << 3);
				++sp;
			}
			descend_pointer<KIND>(d, s);
			one_level_down(s, ro, rd);
			if (s.level == d.innerLevels) {
				s.size = (int)leafSize;
				for (uint32_t vc = leafSize / 2; vc > 1; vc >>= 1) one_level_down(s, ro, rd);
			}
			fetch<KIND>(d, s);
		}
		s.child = linear_child(s);
		++it;
	} while (s.t < tmax && it < maxIters);
	// -3 (too many iterations) / -2 (out of t bounds): no hit
}

struct DevMem {
	void* p = nullptr;
	~DevMem() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 16); }
};

}  // namespace

extern "C" int svb_raycast_depth(int device, const uint8_t* file, uint64_t size, int kind, const float viewInv[16], const float projInv[16],
                                 uint32_t width, uint32_t height, uint32_t maxIters, uint32_t drawLevel, float projectionFactor, float* out_host) {
	if (!file || !viewInv || !projInv || !out_host || width == 0 || height == 0) return SVB_EINVAL;
	if (kind != SVB_FILE_SVDAG && kind != SVB_FILE_USSVDAG && kind != SVB_FILE_SSVDAG) return SVB_EINVAL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SVB_ENODEV;   // no device: no image, no fallback
	if (cudaSetDevice(device) != cudaSuccess) return SVB_ECUDA;
	DagView d;
	memset(&d, 0, sizeof(d));
	d.kind = kind;
	if (size < 48) return SVB_EINVAL;
	memcpy(d.bbmin, file, 12); memcpy(d.bbmax, file + 12, 12);
	memcpy(&d.rootSide, file + 24, 4); memcpy(&d.levels, file + 28, 4);
	if (d.levels < 2 || d.levels > 30) return SVB_EINVAL;
	DevMem a, b, img;
	if (kind != SVB_FILE_SSVDAG) {   // encoded_svdag.cpp:76-103 / encoded_ussvdag.cpp:60-84: header 44 B, then `count` words
		uint32_t count; memcpy(&count, file + 40, 4);
		if (size < 44 + 4ull * count) return SVB_EINVAL;
		if (a.alloc(4ull * count) != cudaSuccess) return SVB_ENOMEM;
		if (cudaMemcpy(a.p, file + 44, 4ull * count, cudaMemcpyHostToDevice) != cudaSuccess) return SVB_ECUDA;
		d.nodes = (const uint32_t*)a.p;
		d.innerLevels = d.levels - 1;
	} else {                         // encoded_ssvdag.cpp:84-117: header 36 B, u16 inner[], u8 leaves[], u32 levelOffsets[]
		uint64_t o = 36;
		uint32_t nInner, nLeafBytes, nOff;
		memcpy(&nInner, file + o, 4); o += 4;
		const uint8_t* inner = file + o; o += 2ull * nInner;
		if (o + 4 > size) return SVB_EINVAL;
		memcpy(&nLeafBytes, file + o, 4); o += 4;
		const uint8_t* leaves = file + o; o += nLeafBytes;
		if (o + 4 > size) return SVB_EINVAL;
		memcpy(&nOff, file + o, 4); o += 4;
		if (o + 4ull * nOff > size || nOff > 32) return SVB_EINVAL;
		memcpy(d.levelOffsets, file + o, 4ull * nOff);
		if (a.alloc(2ull * nInner + 8) != cudaSuccess || b.alloc((uint64_t)nLeafBytes + 8) != cudaSuccess) return SVB_ENOMEM;
		if (cudaMemset(a.p, 0, 2ull * nInner + 8) != cudaSuccess) return SVB_ECUDA;
		if (cudaMemcpy(a.p, inner, 2ull * nInner, cudaMemcpyHostToDevice) != cudaSuccess) return SVB_ECUDA;
		if (cudaMemcpy(b.p, leaves, nLeafBytes, cudaMemcpyHostToDevice) != cudaSuccess) return SVB_ECUDA;
		d.inner = (const uint16_t*)a.p;
		d.leaves = (const uint2*)b.p;
		d.innerLevels = d.levels - 2;
	}
	if (drawLevel == 0) drawLevel = d.levels;   // _drawLevel = getNLevels() (octree_dda_renderer.cpp:222)
	Cam cam;
	memcpy(cam.vi, viewInv, 64); memcpy(cam.pi, projInv, 64);
	const size_t bytes = 12ull * width * height;
	if (img.alloc(bytes) != cudaSuccess) return SVB_ENOMEM;
	const unsigned grid = ((width + 7) / 8) * ((height + 7) / 8);
	float* dout = (float*)img.p;
	if (kind == SVB_FILE_SVDAG) k_raycast<SVB_FILE_SVDAG><<<grid, 64>>>(d, cam, width, height, maxIters, drawLevel, projectionFactor, dout);
	else if (kind == SVB_FILE_USSVDAG) k_raycast<SVB_FILE_USSVDAG><<<grid, 64>>>(d, cam, width, height, maxIters, drawLevel, projectionFactor, dout);
	else k_raycast<SVB_FILE_SSVDAG><<<grid, 64>>>(d, cam, width, height, maxIters, drawLevel, projectionFactor, dout);
	if (cudaGetLastError() != cudaSuccess) return SVB_ECUDA;
	if (cudaMemcpy(out_host, dout, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return SVB_ECUDA;
	return SVB_OK;
}
