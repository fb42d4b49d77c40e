// BVH traversal: closest hit (K2, replaces Mesh::RayIntersection, src/Mesh.cpp:33-56) and any hit
// (K4, replaces Scene::Connect / Mesh::LineIntersection, src/Scene.cpp:84-96, src/Mesh.cpp:58-71).
//
// One ray per thread, while-while loop, short stack in registers/local memory.  Nodes are 64 B
// (both children's boxes, four 16-byte read-only loads), triangle records 48 B (three loads); the
// whole structure of the 1M-triangle hall (~80 MB) is L2-resident on B200 (126 MB).
// Box tests use FMA and are conservative (padded boxes + per-child interval slack, see
// bvh_build.cpp); triangle tests use the reference's exact float32 expression.
#pragma once
#include "device_exact.cuh"

namespace earb {

struct SceneDev {
	const float4* nodes;      // 4 float4 per node
	const float4* tris;       // 3 float4 per triangle record (leaf order)
	const float4* materials;  // [M][B] {refl, refr, kept, spec}
	int32_t n_tris, n_materials, n_bands;
};

constexpr int32_t kEmptyChildDev = 0x7fffffff;
constexpr int kStackSize = 64;   // builder bounds depth by 30 + log2(T) (bvh_build.cpp kSahDepth)

struct RaySetup {
	float idx, idy, idz, oox, ooy, ooz;
};
__device__ __forceinline__ RaySetup make_setup(V3 o, V3 d) {
	const float tiny = 1e-20f;
	const float dx = fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x);
	const float dy = fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y);
	const float dz = fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z);
	RaySetup s;
	s.idx = 1.0f / dx; s.idy = 1.0f / dy; s.idz = 1.0f / dz;
	s.oox = o.x * s.idx; s.ooy = o.y * s.idy; s.ooz = o.z * s.idz;
	return s;
}

// slab test of one child box against the ray interval [lo_t, hi_t]
__device__ __forceinline__ bool slab(const RaySetup& s, float lx, float hx, float ly, float hy, float lz, float hz,
                                     float lo_t, float hi_t, float& tnear) {
	const float x0 = fmaf(lx, s.idx, -s.oox), x1 = fmaf(hx, s.idx, -s.oox);
	const float y0 = fmaf(ly, s.idy, -s.ooy), y1 = fmaf(hy, s.idy, -s.ooy);
	const float z0 = fmaf(lz, s.idz, -s.ooz), z1 = fmaf(hz, s.idz, -s.ooz);
	const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), lo_t));
	const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), hi_t));
	tnear = tn;
	return tn <= tf;
}

// ANY_HIT = false: argmin over (t, original index) of triangles with MT hit and t > 0.001 (t < 1e6);
//                  returns the original triangle index or -1, `best_t`, and the record slot.
// ANY_HIT = true : d is the UNNORMALISED segment x - p; returns 1 as soon as a triangle has 1e-5 < t < 1.
template <bool ANY_HIT>
__device__ __forceinline__ int32_t traverse(const SceneDev& sc, V3 o, V3 d, float& best_t, int32_t& best_slot) {
	const RaySetup rs = make_setup(o, d);
	int32_t stack[kStackSize];
	int sp = 0;
	int32_t node = 0;
	int32_t best_idx = -1;
	best_t = ANY_HIT ? 1.0f : 1000000.0f;
	best_slot = -1;
	for (;;) {
		// ---- inner nodes ----
		while (node >= 0 && node != kEmptyChildDev) {
			const float4 a = __ldg(sc.nodes + 4 * (size_t)node);
			const float4 b = __ldg(sc.nodes + 4 * (size_t)node + 1);
			const float4 c = __ldg(sc.nodes + 4 * (size_t)node + 2);
			const float4 dd = __ldg(sc.nodes + 4 * (size_t)node + 3);
			const int32_t c0 = __float_as_int(dd.x), c1 = __float_as_int(dd.y);
			const float s0 = dd.z, s1 = dd.w;
			float t0, t1;
			const bool h0 = c0 != kEmptyChildDev && slab(rs, a.x, a.y, a.z, a.w, c.x, c.y, -s0, best_t + s0, t0);
			const bool h1 = c1 != kEmptyChildDev && slab(rs, b.x, b.y, b.z, b.w, c.z, c.w, -s1, best_t + s1, t1);
			if (h0 && h1) {
				const bool swap = t1 < t0;
				node = swap ? c1 : c0;
				if (sp < kStackSize) stack[sp++] = swap ? c0 : c1;
			} else if (h0) node = c0;
			else if (h1) node = c1;
			else if (sp > 0) node = stack[--sp];
			else node = kEmptyChildDev;
		}
		if (node == kEmptyChildDev) break;
		// ---- leaf ----
		const int32_t code = ~node;
		const int32_t first = code >> 3, count = (code & 7) + 1;
		for (int32_t i = 0; i < count; ++i) {
			const float4 r0 = __ldg(sc.tris + 3 * (size_t)(first + i));
			const float4 r1 = __ldg(sc.tris + 3 * (size_t)(first + i) + 1);
			const float4 r2 = __ldg(sc.tris + 3 * (size_t)(first + i) + 2);
			float t;
			if (moeller_trumbore(mk(r0.x, r0.y, r0.z), mk(r1.x, r1.y, r1.z), mk(r2.x, r2.y, r2.z), o, d, t)) {
				if (ANY_HIT) {
					if (t > 1e-5f && t < 1.0f) return 1;
				} else {
					const int32_t idx = __float_as_int(r0.w);
					// strict t < d in file order == lowest original index among equal t (src/Mesh.cpp:40)
					if (t > 0.001f && (t < best_t || (t == best_t && idx < best_idx))) {
						best_t = t; best_idx = idx; best_slot = first + i;
					}
				}
			}
		}
		if (sp > 0) node = stack[--sp]; else break;
	}
	return ANY_HIT ? 0 : best_idx;
}

}  // namespace earb
