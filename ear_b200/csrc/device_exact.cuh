// Device-side float32 arithmetic in the reference's operation order.
//
// Everything on the geometry path (hit point, normals, reflection, sampling blend, lengths) must
// round exactly like the x86-64 SSE build of the reference: one IEEE rounding per operation, no
// FMA contraction.  This translation unit is compiled with -fmad=false AND every operation here
// goes through an explicit round-to-nearest intrinsic, so neither flag changes nor inlining can
// fuse them.  Vector helpers follow GMTL's order as fixed in oracle/shim/gmtl/gmtl.h:
//   dot = (a0*b0 + a1*b1) + a2*b2,  cross = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0).
#pragma once
#ifndef EARB_HOST_EMULATION
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace earb {

struct V3 { float x, y, z; };

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ V3 vscale(V3 a, float s) { return mk(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
__device__ __forceinline__ V3 vneg(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float vdot(V3 a, V3 b) {
	return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
	return mk(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
	          fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
__device__ __forceinline__ float vlength(V3 a) { return fsqrt(vdot(a, a)); }
// gmtl::normalize: divide each component by the length (no reciprocal), untouched if zero
__device__ __forceinline__ V3 vnormalized(V3 a) {
	const float len = vlength(a);
	if (len != 0.0f) { a.x = fdiv(a.x, len); a.y = fdiv(a.y, len); a.z = fdiv(a.z, len); }
	return a;
}
// gmtl::reflect: v - 2 (v.n) n   (src/Scene.cpp:71-72, 221)
__device__ __forceinline__ V3 vreflect(V3 v, V3 n) {
	const float d = vdot(v, n);
	return mk(fsub(v.x, fmul(2.0f, fmul(d, n.x))), fsub(v.y, fmul(2.0f, fmul(d, n.y))),
	          fsub(v.z, fmul(2.0f, fmul(d, n.z))));
}
// std::fpclassify(x) != FP_NORMAL  (src/Scene.cpp:35-41): zero, subnormal, inf, nan are invalid
__device__ __forceinline__ bool invalid_float(float x) {
	const uint32_t e = (__float_as_uint(x) >> 23) & 0xffu;
	return e == 0u || e == 0xffu;
}

// powf as the reference's libm computes it (glibc powf is exp2(y*log2 x) in double, rounded once:
// <= 0.52 ulp, subnormal results kept).  CUDA's own powf flushes subnormal results to zero, which
// changes the FP_NORMAL classification of tiny contributions (src/Scene.cpp:254), so the same
// double-precision route is taken here; the result equals glibc's except when the exact value sits
// within ~1e-13 (relative) of a float32 rounding boundary.
#ifndef EARB_HOST_EMULATION
__device__ __forceinline__ float pow_ref(float x, float y) {
	if (y == 0.0f || x == 1.0f) return 1.0f;
	// results below 2^-150 round to +0: decide that from a single-precision estimate and skip the double path
	// (x^1000 lobes underflow for x < 0.9, i.e. for ~90 % of the specular-lobe evaluations)
	if (x > 0.0f && x < 1.0f && y > 0.0f && y * __log2f(x) < -156.0f) return 0.0f;
	return __double2float_rn(exp2((double)y * log2((double)x)));
}
// pow(af, l) with log2(af) hoisted: the air-absorption base is fixed per context
__device__ __forceinline__ double log2_ref(float x) { return log2((double)x); }
__device__ __forceinline__ float pow_ref_hoisted(float x, double log2_x, float y) {
	if (y == 0.0f || x == 1.0f) return 1.0f;
	return __double2float_rn(exp2((double)y * log2_x));
}
#endif

// gmtl::intersectDoubleSided (Moeller-Trumbore, non-culling, EPSILON 1e-5; src/Mesh.cpp:40,65)
// with e1 = v1 - v0 and e2 = v2 - v0 precomputed in float32 (same values the reference forms).
__device__ __forceinline__ bool moeller_trumbore(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float& t) {
	const float EPSILON = 0.00001f;
	const V3 p = vcross(d, e2);
	const float det = vdot(e1, p);
	if (det > -EPSILON && det < EPSILON) return false;
	const float inv = fdiv(1.0f, det);
	const V3 tv = vsub(o, v0);
	const float u = fmul(vdot(tv, p), inv);
	if (u < 0.0f || u > 1.0f) return false;
	const V3 q = vcross(tv, e1);
	const float v = fmul(vdot(d, q), inv);
	if (v < 0.0f || fadd(u, v) > 1.0f) return false;
	t = fmul(vdot(e2, q), inv);
	return t >= 0.0f;
}

// ---------------- Philox4x32-10 ----------------
// key = seed, counter = (ray id lo, ray id hi, draw index, context).  Every draw is one block: a
// Sample_Sphere try takes its three uniforms from words 0..2, Material::Bounce takes word 0.  The only
// per-ray state is the 32-bit draw index, so a ray can be parked in memory between kernels.
struct Rng {
	uint32_t k0, k1, ray_lo, ray_hi, ctx, block;
	__device__ __forceinline__ void start(uint64_t seed, uint32_t context, uint64_t ray, uint32_t first_block = 0) {
		k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
		ray_lo = (uint32_t)ray; ray_hi = (uint32_t)(ray >> 32);
		ctx = context; block = first_block;
	}
	__device__ __forceinline__ void draw(uint32_t& w0, uint32_t& w1, uint32_t& w2) {
		uint32_t c0 = ray_lo, c1 = ray_hi, c2 = block++, c3 = ctx, ka = k0, kb = k1;
#pragma unroll
		for (int r = 0; r < 10; ++r) {
			const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
			const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
			const uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
			c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
			ka += 0x9E3779B9u; kb += 0xBB67AE85u;
		}
		w0 = c0; w1 = c1; w2 = c2;
	}
	// 24-bit uniforms in [0,1); stand in for gmtl::Math::unitRandom = rand()/RAND_MAX
	static __device__ __forceinline__ float to_unit(uint32_t x) { return fmul((float)(x >> 8), 1.0f / 16777216.0f); }
	__device__ __forceinline__ void unit3(float& a, float& b, float& c) {
		uint32_t w0, w1, w2;
		draw(w0, w1, w2);
		a = to_unit(w0); b = to_unit(w1); c = to_unit(w2);
	}
	__device__ __forceinline__ float unit1() {
		uint32_t w0, w1, w2;
		draw(w0, w1, w2);
		return to_unit(w0);
	}
};

// Sample_Sphere, src/Distributions.h:48-58 (cube rejection, 0.001 <= |v|^2 <= 1)
__device__ __forceinline__ V3 sample_sphere(Rng& rng) {
	for (;;) {
		float u1, u2, u3;
		rng.unit3(u1, u2, u3);
		const float f1 = fsub(fmul(u1, 2.0f), 1.0f);
		const float f2 = fsub(fmul(u2, 2.0f), 1.0f);
		const float f3 = fsub(fmul(u3, 2.0f), 1.0f);
		V3 v = mk(f1, f2, f3);
		const float l = vdot(v, v);
		if (l < 0.001f || l > 1.0f) continue;
		const float s = fsqrt(l);
		return mk(fdiv(v.x, s), fdiv(v.y, s), fdiv(v.z, s));
	}
}
// Sample_Hemi(v, n), src/Distributions.h:62-67 (uniform hemisphere by rejection on n.v < 0)
__device__ __forceinline__ V3 sample_hemi(Rng& rng, V3 n) {
	for (;;) {
		const V3 v = sample_sphere(rng);
		if (vdot(n, v) < 0.0f) continue;
		return v;
	}
}
// Sample_Hemi(v, n, reflection, factor), src/Distributions.h:71-75
__device__ __forceinline__ V3 sample_hemi_blend(Rng& rng, V3 n, V3 refl, float factor) {
	const V3 h = sample_hemi(rng, n);
	const V3 v = vadd(vscale(h, fsub(1.0f, factor)), vscale(refl, factor));
	return vnormalized(v);
}

}  // namespace earb
