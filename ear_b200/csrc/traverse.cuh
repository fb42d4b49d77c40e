// BVH traversal: closest hit (K2, replaces Mesh::RayIntersection, src/Mesh.cpp:33-56) and any hit
// (K4, replaces Scene::Connect / Mesh::LineIntersection, src/Scene.cpp:84-96, src/Mesh.cpp:58-71).
//
// One ray per lane, but the loop is WARP-SYNCHRONOUS: every iteration the warp votes and performs
// either one node step (lanes sitting on an inner node) or one leaf step (lanes sitting on a leaf);
// lanes whose turn it is not simply wait.  Leaves are postponed until enough lanes have one
// (kLeafVote) so both kinds of step run at high lane occupancy, and there is no divergent
// control flow for the compiler to fail to reconverge (the first version of this kernel averaged
// 2.8 active lanes per instruction -- profiles/r1_ncu_render_v0.txt).
//   * nodes are 64 B = two 256-bit loads (LDG.E.ENL2.256), triangle records 64 B of which the test
//     needs 48 (one 256-bit + one 128-bit load); the 4th quarter (unit normal) is fetched for the
//     winner only.  The whole structure of the 1M-triangle hall (~96 MB) is L2-resident on B200.
//   * the traversal stack lives in shared memory, [entry][thread] so a warp's access is
//     conflict-free; entries carry the child's entry distance so stale subtrees are culled on pop.
//   * box tests use FMA and are conservative; triangle tests use the reference's exact float32
//     expression (device_exact.cuh).  A child is culled only if the ray enters its padded box
//     later than best_t * (1 + kKappa) + s0 [default], or later than best_t + slack(child)
//     [EXACT: rigorous per-child bound from bvh_build.cpp, ~2x the node visits].
#pragma once
#include "device_exact.cuh"

namespace earb {

struct SceneDev {
	const float4* nodes;      // 4 float4 per node
	const float4* tris;       // 4 float4 per triangle record (leaf order)
	const float4* materials;  // [M][B] {refl, refr, kept, spec}
	int32_t n_tris, n_materials, n_bands;
	float s0;                 // absolute interval margin (see bvh_build.cpp)
	int32_t exact;            // 1: rigorous per-child slack
	int32_t leaf_vote;        // lanes that must wait on a leaf before the warp runs a leaf step
	int32_t fetch_vote;       // idle lanes that trigger a fetch of new work in the persistent kernels
	int32_t vis_cap;          // visibility-map texel lists longer than this are traced through the BVH instead
	int32_t vis_prefix;       // ... after their first vis_prefix (nearest) entries failed to block the query
	int2* spill;              // [spill_rows][spill_threads] overflow of the shared-memory traversal stacks
	int32_t spill_threads;    // columns of `spill`; kernels that walk the BVH launch at most this many threads
	int32_t spill_rows;       // sized at scene creation from the depth of the tree: 3 pushes per level at most
	const float4* emitters;   // emitter triangles of mesh sources: (v0, area) (v1, 0) (v2, 0) (normal, 0), or null
};

constexpr int32_t kEmptyChildDev = 0x7fffffff;
#ifndef EARB_STACK_ENTRIES
#define EARB_STACK_ENTRIES 20   // 20 x 128 x 8 B = 20 KB per block: ten blocks fit an SM's shared memory
#endif
constexpr int kStackEntries = EARB_STACK_ENTRIES;       // shared-memory entries per lane; deeper pushes go to SceneDev::spill
constexpr int kStackSpill = 160;        // host emulation only: rows of its local overflow array
constexpr int kLeafVote = 12;           // do a leaf step once this many lanes wait on a leaf
constexpr float kKappa = 1.0f / 1024.0f;

// L2 residency (EARB_L2_HINTS, default on): the scene image (nodes + triangle records, 85 MB for the 1M-triangle
// hall) is re-read by every ray and fits the 126 MB L2, while the ray pool and the work lists (GBs per launch) stream
// through once.  Image loads ask L2 to keep their lines (evict_last), stream loads / stores to drop theirs first
// (evict_first); without the hints the streams evicted the tree (L2 hit rate 62 % in the closest-hit kernel,
// profiles/r1_ncu_final_wf_traverse_kernel_16mi.txt).  256-bit accesses carry the priority in the instruction,
// narrower ones take it from a createpolicy descriptor.
#ifndef EARB_L2_HINTS
#define EARB_L2_HINTS 1
#endif
struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p) {
	F8 r;
#ifdef EARB_HOST_EMULATION
	r.lo = p[0]; r.hi = p[1];
	return r;
#else
#if EARB_L2_HINTS
	asm volatile("ld.global.nc.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
	             : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
	             : "l"(p));
	return r;
#endif
}
#ifndef EARB_HOST_EMULATION
__device__ __forceinline__ uint64_t l2_policy_keep() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
#endif
// 128-bit load of scene-image data (third quarter of a triangle record, normals, materials)
__device__ __forceinline__ float4 ldg_keep(const float4* p) {
#if defined(EARB_HOST_EMULATION) || !EARB_L2_HINTS
	return __ldg(p);
#else
	float4 r;
	asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
	             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(l2_policy_keep()));
	return r;
#endif
}
#ifndef EARB_HOST_EMULATION
// streamed pool / list accesses: 16-, 8- and 4-byte forms
template <class T> __device__ __forceinline__ T ld_stream(const T* p);
template <class T> __device__ __forceinline__ void st_stream(T* p, T v);
#if EARB_L2_HINTS
template <> __device__ __forceinline__ float4 ld_stream<float4>(const float4* p) {
	float4 r;
	asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(l2_policy_stream()));
	return r;
}
template <> __device__ __forceinline__ uint4 ld_stream<uint4>(const uint4* p) {
	uint4 r;
	asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(l2_policy_stream()));
	return r;
}
template <> __device__ __forceinline__ uint2 ld_stream<uint2>(const uint2* p) {
	uint2 r;
	asm volatile("ld.global.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(l2_policy_stream()));
	return r;
}
template <> __device__ __forceinline__ int2 ld_stream<int2>(const int2* p) {
	int2 r;
	asm volatile("ld.global.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(l2_policy_stream()));
	return r;
}
template <> __device__ __forceinline__ int ld_stream<int>(const int* p) {
	int r;
	asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(l2_policy_stream()));
	return r;
}
template <> __device__ __forceinline__ void st_stream<float4>(float4* p, float4 v) {
	asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(l2_policy_stream()) : "memory");
}
template <> __device__ __forceinline__ void st_stream<uint4>(uint4* p, uint4 v) {
	asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(l2_policy_stream()) : "memory");
}
template <> __device__ __forceinline__ void st_stream<uint2>(uint2* p, uint2 v) {
	asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(l2_policy_stream()) : "memory");
}
template <> __device__ __forceinline__ void st_stream<int2>(int2* p, int2 v) {
	asm volatile("st.global.L2::cache_hint.v2.s32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(l2_policy_stream()) : "memory");
}
template <> __device__ __forceinline__ void st_stream<int>(int* p, int v) {
	asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_stream()) : "memory");
}
#else
template <class T> __device__ __forceinline__ T ld_stream(const T* p) { return *p; }
template <class T> __device__ __forceinline__ void st_stream(T* p, T v) { *p = v; }
#endif
#endif

struct RaySetup { float idx, idy, idz, oox, ooy, ooz; };
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

// Per-lane traversal stack: kStackEntries entries in shared memory (column `threadIdx.x` of an [entry][thread]
// array: conflict-free), addressed through explicit shared-space instructions.  Deeper pushes (never seen with SAH
// trees of the benchmark scenes; a 4-wide node pushes at most 3 entries per level, and scene creation checks
// 3 * depth against the total capacity) overflow into a GLOBAL side buffer, one column per thread.  The first
// version spilled into a local array: indexing it dynamically forced the whole traversal state into a 400-byte
// local frame (33.8 M local stores per launch, profiles/r1_ncu_final_wf_traverse_kernel_16mi.txt).
struct LaneStack {
#ifdef EARB_HOST_EMULATION
	int2* smem;       // this lane's column: entry k at smem[k * stride]
	int stride;
	int2 spill[kStackSpill];
	__device__ __forceinline__ void bind(int2* column, int stride_entries) { smem = column; stride = stride_entries; }
	__device__ __forceinline__ void put(int k, int2 e) { smem[k * stride] = e; }
	__device__ __forceinline__ int2 get(int k) const { return smem[k * stride]; }
	__device__ __forceinline__ void put_deep(const SceneDev&, int k, int2 e) { spill[k] = e; }
	__device__ __forceinline__ int2 get_deep(const SceneDev&, int k) const { return spill[k]; }
#else
	uint32_t base;    // shared-space byte address of this lane's column
	uint32_t pitch;   // bytes between consecutive entries of a column
	__device__ __forceinline__ void bind(int2* column, int stride_entries) {
		base = (uint32_t)__cvta_generic_to_shared(column);
		pitch = (uint32_t)stride_entries * (uint32_t)sizeof(int2);
	}
	__device__ __forceinline__ void put(int k, int2 e) {
		asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + (uint32_t)k * pitch), "r"(e.x), "r"(e.y));
	}
	__device__ __forceinline__ int2 get(int k) const {
		int2 e;
		asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(base + (uint32_t)k * pitch));
		return e;
	}
	static __device__ __forceinline__ size_t deep_at(const SceneDev& sc, int k) {
		return (size_t)k * (size_t)sc.spill_threads + (size_t)(blockIdx.x * blockDim.x + threadIdx.x);
	}
	__device__ __forceinline__ void put_deep(const SceneDev& sc, int k, int2 e) { sc.spill[deep_at(sc, k)] = e; }
	__device__ __forceinline__ int2 get_deep(const SceneDev& sc, int k) const { return sc.spill[deep_at(sc, k)]; }
#endif
	int sp;
	__device__ __forceinline__ void push(const SceneDev& sc, int32_t node, float key) {
		const int2 e = make_int2(node, __float_as_int(key));
		if (sp < kStackEntries) put(sp, e);
		else if (sp < kStackEntries + sc.spill_rows) put_deep(sc, sp - kStackEntries, e);
		++sp;
	}
	__device__ __forceinline__ int2 pop(const SceneDev& sc) {
		--sp;
		return sp < kStackEntries ? get(sp) : get_deep(sc, min(sp - kStackEntries, sc.spill_rows - 1));
	}
};

// Per-lane traversal state shared by the warp-synchronous drivers (traverse_warp below and the
// persistent wavefront kernels in wavefront.cuh).
struct TravState {
	int32_t node;        // >= 0 inner node, < 0 leaf, kEmptyChildDev: nothing (left) to do
	float best_t;
	int32_t best_idx;    // closest: original triangle index or -1; any-hit: 1 once occluded
	int32_t best_slot;   // closest: triangle record slot of the winner
	V3 o, d;
	RaySetup rs;
	LaneStack st;
	template <bool ANY_HIT>
	__device__ __forceinline__ void begin(V3 origin, V3 dir) {
		o = origin; d = dir; rs = make_setup(origin, dir);
		node = 0; st.sp = 0;
		best_idx = ANY_HIT ? 0 : -1;
		best_t = ANY_HIT ? 1.0f : 1000000.0f;
		best_slot = -1;
	}
};

template <bool EXACT>
__device__ __forceinline__ void pop_next(const SceneDev& sc, TravState& ts) {
	const float limit = EXACT ? ts.best_t : fmaf(ts.best_t, kKappa, ts.best_t) + sc.s0;
	int32_t node = kEmptyChildDev;
	while (ts.st.sp > 0 && node == kEmptyChildDev) {   // one exit test, no break: the lanes reconverge every round (-1 %)
		const int2 e = ts.st.pop(sc);
		if (__int_as_float(e.y) <= limit) node = e.x;
	}
	ts.node = node;
}

__device__ __forceinline__ float half_bits_to_float(uint32_t h) {   // fp16 bits -> float (no inf/nan in the table)
	const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
	if (e == 0u) return (float)m * 5.9604645e-8f;                      // subnormal: m * 2^-24
	return __int_as_float((int)(((e + 112u) << 23) | (m << 13)));
}
__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) { const uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
// c[k] for k in 0..3 out of registers.  Written as a ternary chain this compiled into a BRANCH tree (BSSY / BRA / BSYNC,
// ~12 instructions and a divergent reconvergence per pick, four picks per node step: profiles/r2_ncu_wf_traverse_kernel.txt);
// three selects on the two bits of k instead.
__device__ __forceinline__ int32_t pick4(int32_t c0, int32_t c1, int32_t c2, int32_t c3, uint32_t k) {
#ifdef EARB_HOST_EMULATION
	return k == 0u ? c0 : k == 1u ? c1 : k == 2u ? c2 : c3;
#else
	int32_t r;
	asm("{\n\t.reg .pred p, q;\n\t.reg .b32 a, b;\n\t"
	    "and.b32 a, %5, 1;\n\tsetp.ne.b32 p, a, 0;\n\t"
	    "and.b32 b, %5, 2;\n\tsetp.ne.b32 q, b, 0;\n\t"
	    "selp.b32 a, %2, %1, p;\n\tselp.b32 b, %4, %3, p;\n\tselp.b32 %0, b, a, q;\n\t}"
	    : "=r"(r) : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(k));
	return r;
#endif
}

#ifndef EARB_DECODE_PRMT
#define EARB_DECODE_PRMT 0   // measured on B200 (profiles/r2_ab_closest.txt): the PRMT form is 4 % SLOWER than I2F.U8 -- the kernel
                             // is bound by issue slots and divergence, not by the XU pipe, and the form costs 6 more FMAs per node
#endif
#ifdef EARB_HOST_EMULATION
static float g_emul_decode_bias = 0.00390625f;
#endif
// float 2^15 + byte k of w (see node_step)
__device__ __forceinline__ float plane_byte(uint32_t w, int k) {
#ifdef EARB_HOST_EMULATION
	return __uint_as_float(0x47000000u | (((w >> (8 * k)) & 0xffu) << 8));
#else
	// result bytes (low to high): 0x00, byte k of w, 0x00, 0x47  <-  selector nibbles 4, k, 5, 7 over {w, 0x47000000}
	return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504u | ((uint32_t)k << 4)));
#endif
}

// one inner-node step for a lane sitting on an inner (4-wide, 8-bit quantised) node
template <bool EXACT>
__device__ __forceinline__ void node_step(const SceneDev& sc, TravState& ts) {
	const F8 n0 = ldg256(sc.nodes + 4 * (size_t)ts.node);
	const F8 n1 = ldg256(sc.nodes + 4 * (size_t)ts.node + 2);
	const RaySetup& rs = ts.rs;
	const uint32_t ex = __float_as_uint(n0.lo.w);
	// plane(q) = lo + q * 2^e  ->  t(q) = q * (2^e * idir) + (lo - o) * idir
	const float ax = __int_as_float((int)((ex & 0xffu) << 23)) * rs.idx;
	const float ay = __int_as_float((int)(((ex >> 8) & 0xffu) << 23)) * rs.idy;
	const float az = __int_as_float((int)(((ex >> 16) & 0xffu) << 23)) * rs.idz;
	const float bx = fmaf(n0.lo.x, rs.idx, -rs.oox), by = fmaf(n0.lo.y, rs.idy, -rs.ooy), bz = fmaf(n0.lo.z, rs.idz, -rs.ooz);
	// entry / exit planes by direction sign (one select per axis per node instead of min/max per child)
	const uint32_t qlx = __float_as_uint(n1.lo.x), qly = __float_as_uint(n1.lo.y), qlz = __float_as_uint(n1.lo.z);
	const uint32_t qhx = __float_as_uint(n1.lo.w), qhy = __float_as_uint(n1.hi.x), qhz = __float_as_uint(n1.hi.y);
	const bool ngx = rs.idx < 0.0f, ngy = rs.idy < 0.0f, ngz = rs.idz < 0.0f;
	const uint32_t nx = ngx ? qhx : qlx, fx = ngx ? qlx : qhx;
	const uint32_t ny = ngy ? qhy : qly, fy = ngy ? qly : qhy;
	const uint32_t nz = ngz ? qhz : qlz, fz = ngz ? qlz : qhz;
	const float hi_rel = fmaf(ts.best_t, kKappa, ts.best_t) + sc.s0;
	const uint32_t sl01 = __float_as_uint(n1.hi.z), sl23 = __float_as_uint(n1.hi.w);
#if EARB_DECODE_PRMT
	// Plane bytes become floats without the conversion unit: one byte permute builds the bit pattern of 2^15 + q
	// (exponent 0x47, the byte in mantissa bits 8..15, where one ulp of 2^15 is 2^-8 ... so q lands on weight 1) and the
	// offset 2^15 * a is folded into the per-axis constant.  Folding rounds that constant at magnitude 2^15 |a|
	// (error <= 2^-9 grid steps in t); the near constant is lowered and the far one raised by 2^-8 grid steps --
	// -(2^15 +- 2^-8) are exact floats -- so the decoded slab still contains the quantised box.  (I2F.U8 runs on
	// the quarter-rate XU pipe: 24 per node step kept it 45 % busy.)
#ifdef EARB_HOST_EMULATION
	const float bias = g_emul_decode_bias;   // test knob: 0 shows that the margin is load-bearing (tests/test_host_logic.py)
#else
	const float bias = 0.00390625f;          // 2^-8 grid steps
#endif
	const float m_lo = -(32768.0f + bias), m_hi = -(32768.0f - bias);
	const float m_near_x = ngx ? m_hi : m_lo, m_far_x = ngx ? m_lo : m_hi;
	const float m_near_y = ngy ? m_hi : m_lo, m_far_y = ngy ? m_lo : m_hi;
	const float m_near_z = ngz ? m_hi : m_lo, m_far_z = ngz ? m_lo : m_hi;
	const float bnx = fmaf(ax, m_near_x, bx), bfx = fmaf(ax, m_far_x, bx);
	const float bny = fmaf(ay, m_near_y, by), bfy = fmaf(ay, m_far_y, by);
	const float bnz = fmaf(az, m_near_z, bz), bfz = fmaf(az, m_far_z, bz);
#endif
	uint32_t key[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
#if EARB_DECODE_PRMT
		const float tnx = fmaf(plane_byte(nx, k), ax, bnx), tfx = fmaf(plane_byte(fx, k), ax, bfx);
		const float tny = fmaf(plane_byte(ny, k), ay, bny), tfy = fmaf(plane_byte(fy, k), ay, bfy);
		const float tnz = fmaf(plane_byte(nz, k), az, bnz), tfz = fmaf(plane_byte(fz, k), az, bfz);
#else
		const int sh = 8 * k;
		const float tnx = fmaf((float)((nx >> sh) & 0xffu), ax, bx), tfx = fmaf((float)((fx >> sh) & 0xffu), ax, bx);
		const float tny = fmaf((float)((ny >> sh) & 0xffu), ay, by), tfy = fmaf((float)((fy >> sh) & 0xffu), ay, by);
		const float tnz = fmaf((float)((nz >> sh) & 0xffu), az, bz), tfz = fmaf((float)((fz >> sh) & 0xffu), az, bz);
#endif
		float slack = 0.0f;
		if (EXACT) slack = half_bits_to_float(((k < 2 ? sl01 : sl23) >> (16 * (k & 1))) & 0xffffu);
		const float lo_t = EXACT ? -slack : -sc.s0;
		const float hi_t = EXACT ? ts.best_t + slack : hi_rel;
		const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, lo_t));
		const float tf = fminf(fminf(tfx, tfy), fminf(tfz, hi_t));
		// sort key: entry distance (minus slack in EXACT mode), clamped at 0, low two mantissa bits = child slot;
		// it only ever under-estimates the entry distance, so culling by it on pop stays conservative
		const float kd = fmaxf(EXACT ? tn - slack : tn, 0.0f);
		key[k] = (tn <= tf) ? ((__float_as_uint(kd) & ~3u) | (uint32_t)k) : 0xffffffffu;
	}
	cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
	const int32_t c0 = __float_as_int(n0.hi.x), c1 = __float_as_int(n0.hi.y), c2 = __float_as_int(n0.hi.z), c3 = __float_as_int(n0.hi.w);
	if (key[0] == 0xffffffffu) { pop_next<EXACT>(sc, ts); return; }
	// farther children go on the stack, farthest first.  The keys are sorted, so the valid ones are a prefix; with
	// room for all three in the shared-memory part (nearly always) the pushes are three predicated stores -- the
	// per-push branches of the general path ran with 2-5 lanes active and made up ~20 % of the kernel's instructions
	const bool h1 = key[1] != 0xffffffffu, h2 = key[2] != 0xffffffffu, h3 = key[3] != 0xffffffffu;
	const int2 e1 = make_int2(pick4(c0, c1, c2, c3, key[1] & 3u), (int)(key[1] & ~3u));
	const int2 e2 = make_int2(pick4(c0, c1, c2, c3, key[2] & 3u), (int)(key[2] & ~3u));
	const int2 e3 = make_int2(pick4(c0, c1, c2, c3, key[3] & 3u), (int)(key[3] & ~3u));
	if (ts.st.sp + 3 <= kStackEntries) {
		int sp = ts.st.sp;
		if (h3) ts.st.put(sp, e3);
		sp += h3 ? 1 : 0;
		if (h2) ts.st.put(sp, e2);
		sp += h2 ? 1 : 0;
		if (h1) ts.st.put(sp, e1);
		sp += h1 ? 1 : 0;
		ts.st.sp = sp;
	} else {
		if (h3) ts.st.push(sc, e3.x, __int_as_float(e3.y));
		if (h2) ts.st.push(sc, e2.x, __int_as_float(e2.y));
		if (h1) ts.st.push(sc, e1.x, __int_as_float(e1.y));
	}
	ts.node = pick4(c0, c1, c2, c3, key[0] & 3u);
}

// one leaf step for a lane sitting on a leaf: tests ONE triangle and stays on the leaf while it has more, so a
// warp's leaf round costs one triangle test whatever the leaf sizes are (a per-leaf loop idles the lanes with
// short leaves while the longest one finishes)
template <bool ANY_HIT, bool EXACT>
__device__ __forceinline__ void leaf_step(const SceneDev& sc, TravState& ts) {
	const int32_t code = ~ts.node;
	const int32_t first = code >> 3, rest = code & 7;
	const float4* rec = sc.tris + 4 * (size_t)first;
	const F8 r01 = ldg256(rec);
	const float4 r2 = ldg_keep(rec + 2);
	float t;
	if (moeller_trumbore(mk(r01.lo.x, r01.lo.y, r01.lo.z), mk(r01.hi.x, r01.hi.y, r01.hi.z), mk(r2.x, r2.y, r2.z), ts.o, ts.d, t)) {
		if (ANY_HIT) {
			if (t > 1e-5f && t < 1.0f) ts.best_idx = 1;
		} else {
			const int32_t idx = __float_as_int(r01.lo.w);
			// strict t < d in file order == lowest original index among equal t (src/Mesh.cpp:40)
			if (t > 0.001f && (t < ts.best_t || (t == ts.best_t && idx < ts.best_idx))) {
				ts.best_t = t; ts.best_idx = idx; ts.best_slot = first;
			}
		}
	}
	if (ANY_HIT && ts.best_idx) ts.node = kEmptyChildDev;
	else if (rest > 0) ts.node = ~(((first + 1) << 3) | (rest - 1));
	else pop_next<EXACT>(sc, ts);
}

// Speculative variant used by the persistent kernels: the leaf is PARKED in `pend` (one triangle is tested per
// call) while the lane keeps walking inner nodes from its stack; best_t only shrinks later, which is conservative.
template <bool ANY_HIT>
__device__ __forceinline__ void pend_step(const SceneDev& sc, TravState& ts, int32_t& pend) {
	const int32_t code = ~pend;
	const int32_t first = code >> 3, rest = code & 7;
	const float4* rec = sc.tris + 4 * (size_t)first;
	const F8 r01 = ldg256(rec);
	const float4 r2 = ldg_keep(rec + 2);
	float t;
	if (moeller_trumbore(mk(r01.lo.x, r01.lo.y, r01.lo.z), mk(r01.hi.x, r01.hi.y, r01.hi.z), mk(r2.x, r2.y, r2.z), ts.o, ts.d, t)) {
		if (ANY_HIT) {
			if (t > 1e-5f && t < 1.0f) ts.best_idx = 1;
		} else {
			const int32_t idx = __float_as_int(r01.lo.w);
			if (t > 0.001f && (t < ts.best_t || (t == ts.best_t && idx < ts.best_idx))) {   // src/Mesh.cpp:40
				ts.best_t = t; ts.best_idx = idx; ts.best_slot = first;
			}
		}
	}
	if (ANY_HIT && ts.best_idx) { pend = 0; ts.node = kEmptyChildDev; ts.st.sp = 0; }   // occluded: the query is over
	else pend = rest > 0 ? ~(((first + 1) << 3) | (rest - 1)) : 0;
}

// Warp-synchronous traversal of one ray per lane; must be called by all 32 lanes of the warp
// (inactive lanes pass active = false).
// ANY_HIT = false: argmin over (t, original index) of triangles with MT hit and t > 0.001 (t < 1e6):
//                  best_idx = original triangle index or -1, best_t, best_slot = record slot.
// ANY_HIT = true : d is the UNNORMALISED segment x - p; best_idx = 1 as soon as a triangle has 1e-5 < t < 1.
template <bool ANY_HIT, bool EXACT>
__device__ __forceinline__ void traverse_warp(const SceneDev& sc, int2* stack_smem, int stack_stride, bool active, V3 o,
                                              V3 d, float& best_t, int32_t& best_idx, int32_t& best_slot) {
	TravState ts;
	ts.st.bind(stack_smem, stack_stride);
	ts.begin<ANY_HIT>(o, d);
	if (!active) ts.node = kEmptyChildDev;
	for (;;) {
		const bool inner = ts.node >= 0 && ts.node != kEmptyChildDev;
		const bool leaf = ts.node < 0;
		const unsigned m_inner = __ballot_sync(0xffffffffu, inner);
		const unsigned m_leaf = __ballot_sync(0xffffffffu, leaf);
		if ((m_inner | m_leaf) == 0u) break;
		if (m_inner != 0u && __popc(m_leaf) < sc.leaf_vote) {
			if (inner) node_step<EXACT>(sc, ts);
		} else {
			if (leaf) leaf_step<ANY_HIT, EXACT>(sc, ts);
		}
	}
	best_t = ts.best_t; best_idx = ts.best_idx; best_slot = ts.best_slot;
}

}  // namespace earb
