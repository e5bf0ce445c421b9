// svb_sat.cuh -- triangle/box overlap predicate, bit-exact with the reference's
// testTriBox / planeBoxOverlap (src/symvox/test_triangle_box.cpp:105-184, :38-56).
//
// Every product/sum is an explicit round-to-nearest intrinsic (__dmul_rn/__dadd_rn/__dsub_rn)
// so the compiler can never contract a multiply-add into an FMA: the reference is built for
// baseline x86-64 (no -march in its CMakeLists.txt), i.e. with separately rounded operations.
// The predicate is a pure conjunction of 13 axis tests on finite doubles, so the ORDER of the
// tests is free; the cheap box-axis tests run first (most children are rejected there).
// Vector conventions (SpaceLand's are not in the reference tree -> oracle/sl_shim):
// dot = a0*b0 + a1*b1 + a2*b2 left to right, cross = textbook.
#pragma once
#include <cmath>

// Host/device portability of the predicate sources: nvcc compiles them for the kernels; tests/classify_harness.cpp
// compiles the SAME text with g++ -ffp-contract=off (separately rounded operations, std::fma correctly rounded)
// to check the filter against the predicate on the CPU.  libsvb.so never runs the host variant.
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define SVB_HD __host__ __device__ __forceinline__
#define SVB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define SVB_HD inline
#define SVB_HD_NOINLINE inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif
#endif
#ifdef __CUDA_ARCH__
#define SVB_DADD(a, b) __dadd_rn((a), (b))
#define SVB_DSUB(a, b) __dsub_rn((a), (b))
#define SVB_DMUL(a, b) __dmul_rn((a), (b))
#define SVB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SVB_FFS(x) __ffs(x)
#define SVB_FFSLL(x) __ffsll((long long)(x))
#define SVB_POPC(x) __popc(x)
#elif defined(SVB_HOST_STRICT_OPS)
// host pass of the SECOND test harness, built with -ffp-contract=fast -mfma so that every plain a * b + c of the interval
// filter is contracted into an FMA, as nvcc contracts it on the device (85-108 DFMA per k_classify_filtered instantiation):
// the explicitly rounded operations go through out-of-line functions that hold a single operation each -- nothing to
// contract with -- and stay separately rounded, like __dadd_rn / __dmul_rn on the device.
namespace svb { namespace hostops {
__attribute__((noinline)) inline double add(double a, double b) { return a + b; }
__attribute__((noinline)) inline double sub(double a, double b) { return a - b; }
__attribute__((noinline)) inline double mul(double a, double b) { return a * b; }
} }
#define SVB_DADD(a, b) ::svb::hostops::add((double)(a), (double)(b))
#define SVB_DSUB(a, b) ::svb::hostops::sub((double)(a), (double)(b))
#define SVB_DMUL(a, b) ::svb::hostops::mul((double)(a), (double)(b))
#define SVB_FMA(a, b, c) std::fma((double)(a), (double)(b), (double)(c))
#define SVB_FFS(x) __builtin_ffs((int)(x))
#define SVB_FFSLL(x) __builtin_ffsll((long long)(x))
#define SVB_POPC(x) __builtin_popcount((unsigned)(x))
#else   // host pass: plain IEEE operations; the translation unit must be built with -ffp-contract=off
#define SVB_DADD(a, b) ((double)(a) + (double)(b))
#define SVB_DSUB(a, b) ((double)(a) - (double)(b))
#define SVB_DMUL(a, b) ((double)(a) * (double)(b))
#define SVB_FMA(a, b, c) std::fma((double)(a), (double)(b), (double)(c))
#define SVB_FFS(x) __builtin_ffs((int)(x))
#define SVB_FFSLL(x) __builtin_ffsll((long long)(x))
#define SVB_POPC(x) __builtin_popcount((unsigned)(x))
#endif

namespace svb {

SVB_HD bool sat_axis_reject(double pa, double pb, double rad) {
	double mn = (pa < pb) ? pa : pb;
	double mx = (pa < pb) ? pb : pa;
	return (mn > rad) || (mx < -rad);
}

// a*u - b*w
SVB_HD double sat_ms(double a, double u, double b, double w) { return SVB_DSUB(SVB_DMUL(a, u), SVB_DMUL(b, w)); }
// -a*u + b*w
SVB_HD double sat_nma(double a, double u, double b, double w) { return SVB_DADD(SVB_DMUL(-a, u), SVB_DMUL(b, w)); }

// c: box centre, h: half side, t: 9 floats (v0 v1 v2)
SVB_HD bool tri_box_overlap(double cx, double cy, double cz, double h, const float* __restrict__ t) {
	const double v0x = SVB_DSUB((double)t[0], cx), v0y = SVB_DSUB((double)t[1], cy), v0z = SVB_DSUB((double)t[2], cz);
	const double v1x = SVB_DSUB((double)t[3], cx), v1y = SVB_DSUB((double)t[4], cy), v1z = SVB_DSUB((double)t[5], cz);
	const double v2x = SVB_DSUB((double)t[6], cx), v2y = SVB_DSUB((double)t[7], cy), v2z = SVB_DSUB((double)t[8], cz);
	const double nh = -h;
	// box axes (test_triangle_box.cpp:165-174)
	{
		double mn = fmin(fmin(v0x, v1x), v2x), mx = fmax(fmax(v0x, v1x), v2x);
		if (mn > h || mx < nh) return false;
		mn = fmin(fmin(v0y, v1y), v2y); mx = fmax(fmax(v0y, v1y), v2y);
		if (mn > h || mx < nh) return false;
		mn = fmin(fmin(v0z, v1z), v2z); mx = fmax(fmax(v0z, v1z), v2z);
		if (mn > h || mx < nh) return false;
	}
	const double e0x = SVB_DSUB(v1x, v0x), e0y = SVB_DSUB(v1y, v0y), e0z = SVB_DSUB(v1z, v0z);
	const double e1x = SVB_DSUB(v2x, v1x), e1y = SVB_DSUB(v2y, v1y), e1z = SVB_DSUB(v2z, v1z);
	const double e2x = SVB_DSUB(v0x, v2x), e2y = SVB_DSUB(v0y, v2y), e2z = SVB_DSUB(v0z, v2z);
	double fex, fey, fez;
	// edge 0: X01(v0,v2) Y02(v0,v2) Z12(v1,v2)   (:137-142)
	fex = fabs(e0x); fey = fabs(e0y); fez = fabs(e0z);
	if (sat_axis_reject(sat_ms(e0z, v0y, e0y, v0z), sat_ms(e0z, v2y, e0y, v2z), SVB_DMUL(SVB_DADD(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e0z, v0x, e0x, v0z), sat_nma(e0z, v2x, e0x, v2z), SVB_DMUL(SVB_DADD(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e0y, v1x, e0x, v1y), sat_ms(e0y, v2x, e0x, v2y), SVB_DMUL(SVB_DADD(fey, fex), h))) return false;
	// edge 1: X01(v0,v2) Y02(v0,v2) Z0(v0,v1)    (:144-149)
	fex = fabs(e1x); fey = fabs(e1y); fez = fabs(e1z);
	if (sat_axis_reject(sat_ms(e1z, v0y, e1y, v0z), sat_ms(e1z, v2y, e1y, v2z), SVB_DMUL(SVB_DADD(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e1z, v0x, e1x, v0z), sat_nma(e1z, v2x, e1x, v2z), SVB_DMUL(SVB_DADD(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e1y, v0x, e1x, v0y), sat_ms(e1y, v1x, e1x, v1y), SVB_DMUL(SVB_DADD(fey, fex), h))) return false;
	// edge 2: X2(v0,v1) Y1(v0,v1) Z12(v1,v2)     (:151-156)
	fex = fabs(e2x); fey = fabs(e2y); fez = fabs(e2z);
	if (sat_axis_reject(sat_ms(e2z, v0y, e2y, v0z), sat_ms(e2z, v1y, e2y, v1z), SVB_DMUL(SVB_DADD(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e2z, v0x, e2x, v0z), sat_nma(e2z, v1x, e2x, v1z), SVB_DMUL(SVB_DADD(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e2y, v1x, e2x, v1y), sat_ms(e2y, v2x, e2x, v2y), SVB_DMUL(SVB_DADD(fey, fex), h))) return false;
	// plane (:179-180, :38-56)
	const double nx = sat_ms(e0y, e1z, e0z, e1y);
	const double ny = sat_ms(e0z, e1x, e0x, e1z);
	const double nz = sat_ms(e0x, e1y, e0y, e1x);
	double vminx, vmaxx, vminy, vmaxy, vminz, vmaxz;
	if (nx > 0.0) { vminx = SVB_DSUB(nh, v0x); vmaxx = SVB_DSUB(h, v0x); } else { vminx = SVB_DSUB(h, v0x); vmaxx = SVB_DSUB(nh, v0x); }
	if (ny > 0.0) { vminy = SVB_DSUB(nh, v0y); vmaxy = SVB_DSUB(h, v0y); } else { vminy = SVB_DSUB(h, v0y); vmaxy = SVB_DSUB(nh, v0y); }
	if (nz > 0.0) { vminz = SVB_DSUB(nh, v0z); vmaxz = SVB_DSUB(h, v0z); } else { vminz = SVB_DSUB(h, v0z); vmaxz = SVB_DSUB(nh, v0z); }
	double d0 = SVB_DADD(SVB_DADD(SVB_DMUL(nx, vminx), SVB_DMUL(ny, vminy)), SVB_DMUL(nz, vminz));
	if (d0 > 0.0) return false;
	double d1 = SVB_DADD(SVB_DADD(SVB_DMUL(nx, vmaxx), SVB_DMUL(ny, vmaxy)), SVB_DMUL(nz, vmaxz));
	return d1 >= 0.0;
}

}  // namespace svb
