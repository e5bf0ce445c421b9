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
#include <cuda_runtime.h>

namespace svb {

__device__ __forceinline__ bool sat_axis_reject(double pa, double pb, double rad) {
	double mn = (pa < pb) ? pa : pb;
	double mx = (pa < pb) ? pb : pa;
	return (mn > rad) || (mx < -rad);
}

// a*u - b*w
__device__ __forceinline__ double sat_ms(double a, double u, double b, double w) { return __dsub_rn(__dmul_rn(a, u), __dmul_rn(b, w)); }
// -a*u + b*w
__device__ __forceinline__ double sat_nma(double a, double u, double b, double w) { return __dadd_rn(__dmul_rn(-a, u), __dmul_rn(b, w)); }

// c: box centre, h: half side, t: 9 floats (v0 v1 v2)
__device__ __forceinline__ bool tri_box_overlap(double cx, double cy, double cz, double h, const float* __restrict__ t) {
	const double v0x = __dsub_rn((double)t[0], cx), v0y = __dsub_rn((double)t[1], cy), v0z = __dsub_rn((double)t[2], cz);
	const double v1x = __dsub_rn((double)t[3], cx), v1y = __dsub_rn((double)t[4], cy), v1z = __dsub_rn((double)t[5], cz);
	const double v2x = __dsub_rn((double)t[6], cx), v2y = __dsub_rn((double)t[7], cy), v2z = __dsub_rn((double)t[8], cz);
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
	const double e0x = __dsub_rn(v1x, v0x), e0y = __dsub_rn(v1y, v0y), e0z = __dsub_rn(v1z, v0z);
	const double e1x = __dsub_rn(v2x, v1x), e1y = __dsub_rn(v2y, v1y), e1z = __dsub_rn(v2z, v1z);
	const double e2x = __dsub_rn(v0x, v2x), e2y = __dsub_rn(v0y, v2y), e2z = __dsub_rn(v0z, v2z);
	double fex, fey, fez;
	// edge 0: X01(v0,v2) Y02(v0,v2) Z12(v1,v2)   (:137-142)
	fex = fabs(e0x); fey = fabs(e0y); fez = fabs(e0z);
	if (sat_axis_reject(sat_ms(e0z, v0y, e0y, v0z), sat_ms(e0z, v2y, e0y, v2z), __dmul_rn(__dadd_rn(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e0z, v0x, e0x, v0z), sat_nma(e0z, v2x, e0x, v2z), __dmul_rn(__dadd_rn(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e0y, v1x, e0x, v1y), sat_ms(e0y, v2x, e0x, v2y), __dmul_rn(__dadd_rn(fey, fex), h))) return false;
	// edge 1: X01(v0,v2) Y02(v0,v2) Z0(v0,v1)    (:144-149)
	fex = fabs(e1x); fey = fabs(e1y); fez = fabs(e1z);
	if (sat_axis_reject(sat_ms(e1z, v0y, e1y, v0z), sat_ms(e1z, v2y, e1y, v2z), __dmul_rn(__dadd_rn(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e1z, v0x, e1x, v0z), sat_nma(e1z, v2x, e1x, v2z), __dmul_rn(__dadd_rn(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e1y, v0x, e1x, v0y), sat_ms(e1y, v1x, e1x, v1y), __dmul_rn(__dadd_rn(fey, fex), h))) return false;
	// edge 2: X2(v0,v1) Y1(v0,v1) Z12(v1,v2)     (:151-156)
	fex = fabs(e2x); fey = fabs(e2y); fez = fabs(e2z);
	if (sat_axis_reject(sat_ms(e2z, v0y, e2y, v0z), sat_ms(e2z, v1y, e2y, v1z), __dmul_rn(__dadd_rn(fez, fey), h))) return false;
	if (sat_axis_reject(sat_nma(e2z, v0x, e2x, v0z), sat_nma(e2z, v1x, e2x, v1z), __dmul_rn(__dadd_rn(fez, fex), h))) return false;
	if (sat_axis_reject(sat_ms(e2y, v1x, e2x, v1y), sat_ms(e2y, v2x, e2x, v2y), __dmul_rn(__dadd_rn(fey, fex), h))) return false;
	// plane (:179-180, :38-56)
	const double nx = sat_ms(e0y, e1z, e0z, e1y);
	const double ny = sat_ms(e0z, e1x, e0x, e1z);
	const double nz = sat_ms(e0x, e1y, e0y, e1x);
	double vminx, vmaxx, vminy, vmaxy, vminz, vmaxz;
	if (nx > 0.0) { vminx = __dsub_rn(nh, v0x); vmaxx = __dsub_rn(h, v0x); } else { vminx = __dsub_rn(h, v0x); vmaxx = __dsub_rn(nh, v0x); }
	if (ny > 0.0) { vminy = __dsub_rn(nh, v0y); vmaxy = __dsub_rn(h, v0y); } else { vminy = __dsub_rn(h, v0y); vmaxy = __dsub_rn(nh, v0y); }
	if (nz > 0.0) { vminz = __dsub_rn(nh, v0z); vmaxz = __dsub_rn(h, v0z); } else { vminz = __dsub_rn(h, v0z); vmaxz = __dsub_rn(nh, v0z); }
	double d0 = __dadd_rn(__dadd_rn(__dmul_rn(nx, vminx), __dmul_rn(ny, vminy)), __dmul_rn(nz, vminz));
	if (d0 > 0.0) return false;
	double d1 = __dadd_rn(__dadd_rn(__dmul_rn(nx, vmaxx), __dmul_rn(ny, vmaxy)), __dmul_rn(nz, vmaxz));
	return d1 >= 0.0;
}

}  // namespace svb
