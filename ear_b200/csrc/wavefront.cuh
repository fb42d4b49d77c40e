// Wavefront engine for Scene::Render's bounce loop (src/Scene.cpp:124-284).
//
// Rays live in a pool of N slots in global memory (SoA, 16-byte vectors).  One ITERATION advances every
// live ray by one bounce through four kernels; all hand-offs are device-side lists built with warp
// ballots + one atomic per warp, so nothing returns to the host between bounces:
//
//   wf_shade_kernel     K3 + K6 + K1: consume last iteration's hit (material, reflect / transmit, resample,
//                       air + surface absorption, cut-offs), write the shading record, enqueue one
//                       occlusion query per facing recorder, REFILL dead slots from the global ray queue
//                       (emission), and compact the slots that still need a closest-hit query into trav_list
//   wf_traverse<false>  K2: persistent closest-hit traversal over trav_list  (Mesh::RayIntersection)
//   wf_traverse<true>   K4: persistent any-hit traversal over q_list        (Scene::Connect); visible
//                       queries are compacted into vis_list
//   wf_splat_kernel     K5: contribution weight + Recorder::Record for every visible query
//
// The traversal kernels carry ONLY traversal state (about 56 registers), so many warps are
// resident to hide the L2 latency of the node fetches, and lanes that finish early fetch the next ray from
// the list instead of waiting for the slowest lane of their warp (vote-driven: node step / leaf step /
// fetch, whichever the warp needs).  The fused single-kernel version this replaces held the whole ray
// state in registers (120 regs, 16 warps/SM) and averaged 11 of 32 lanes per instruction.
#pragma once
#include "traverse.cuh"

namespace earb {

struct WfPool {
	float4* ro;         // [N] origin.xyz, intensity
	float4* rd;         // [N] direction.xyz, path length
	uint4* rm;          // [N] ray id lo, ray id hi, context | bounce << 16 (bounce 0 = empty slot), Philox draw index
	int2* hit;          // [N] (t bits, triangle record slot or -1), written by the closest-hit kernel
	float4* sh0;        // [N] hit point.xyz, intensity after the surface            } shading record of the bounce
	float4* sh1;        // [N] facing normal.xyz, path length up to the hit point    } whose occlusion queries are
	float4* sh2;        // [N] incoming unit direction.xyz, specularity (sign bit set: transmitted)  } in flight
	int* trav_list;     // [N] slots that need a closest-hit query, binned by (direction octant, origin cell)
	uint2* trav_tmp;    // [N] (slot, bin) as appended by the shade kernel, before binning
	uint2* q_tmp;       // [N*R] queries as appended by the shade kernel (bin in bits 16..30 of y), before binning
	int* trav_rank;     // [N] position of the appended ray inside its bin (the value the shade kernel's counting atomic returned)
	int* q_rank;        // [N*R] same for the queries
	int* bins;          // [kRayBins + kSortBins (+ block sums)] histogram -> offsets of the two counting sorts (closest rays, queries)
	float cell_origin[3], cell_scale[3];   // world -> [0,16) cell coordinates of the scene bounds
	int slot_bits;      // split of the query word between slot index and recorder index (see kMaxSlotBits)
	int sort_queries;   // 0: occlusion queries keep their slot order (EAR_B200_SORT_QUERIES)
	int ray_key;        // how closest-hit rays are binned (EAR_B200_RAY_KEY): 0 octant+cell12, 1 octant+axis order+cell12, 2 octant+cell15
	double* ctx_log2af; // [n_ctx] log2(absorption_factor), hoisted out of pow(af, length)
	uint2* q_list;      // [N*R] occlusion queries: x = slot | recorder << 24, y = context | (bounce & 1) << 31
	uint2* vis_list;    // [N*R] the unoccluded ones
	int* counts;        // 0 trav_count, 1 q_count, 2 vis_count, 3 trav_cursor, 4 q_cursor
	uint2* vis_sorted;  // [N*R] vis_list counting-sorted by (context, recorder) for the windowed splat (aliases q_tmp, dead by then)
	int* pair_count;    // [pairs] entries per (context, recorder) / scatter cursor
	int* pair_base;     // [pairs + 1] exclusive scan of the counts
	const float4* qx;   // harness only: explicit segment end point per query (else the recorder position)
	int q_count_idx, q_cursor_idx;   // which counters the any-hit kernel uses (1, 4 normally; 5, 6 for the map fallback list)
	int n_slots;
};

constexpr int kSortBins = 32768;   // queries: 15-bit keys = 3 recorder bits + 12 Morton cell bits
constexpr int kRayBins = 262144;   // closest-hit rays: up to 18-bit keys (3 octant bits + 15 more, see ray_bin)
constexpr int kRayScanBlocks = kRayBins / 1024;
constexpr int kQueryScanBlocks = kSortBins / 1024;
constexpr int kScanBlocks = kRayScanBlocks + kQueryScanBlocks;
// bins layout: [0, kRayBins) rays | [kRayBins, +kSortBins) queries | [.., +kScanBlocks) prefixes of the 1024-bin blocks
constexpr int kBinsTotal = kRayBins + kSortBins + kScanBlocks;
constexpr int32_t kWaitLeaf = 0x7ffffffe;   // closest-hit lane waiting for its parked leaf (kEmptyChildDev - 1)
constexpr int kMaxSlotBits = 28;   // a query word is slot | recorder << slot_bits; slot_bits = min(28, 32 - bits(n_rec)) per call

// ---------------------------------------------------------------------------------------------------
// K2 / K4: persistent traversal with dynamic fetch
// ---------------------------------------------------------------------------------------------------
// resident blocks per SM of the closest-hit kernel: 10 x 128 threads at 48 registers and 20 shared-memory stack entries.
// Measured per 4e7 rays (profiles/r2_ab_closest.txt): 6 blocks (72 regs) 757 ms, 8 (62) 685, 9 (56) 640, 10 (48, 20-entry
// stack) 629, 12 (40 regs, spills, 16-entry stack) 718 -- the kernel is latency-bound, warps in flight pay until spills start
#ifndef EARB_TRAV_MIN_BLOCKS
#define EARB_TRAV_MIN_BLOCKS 10
#endif
#ifndef EARB_ANYHIT_MIN_BLOCKS
#define EARB_ANYHIT_MIN_BLOCKS 9   // measured at C5 (profiles/r2_ab_closest.txt): 6 blocks 814 ms, 8 758, 9 706, 10 729
#endif
template <bool ANY_HIT, bool EXACT>
__global__ void __launch_bounds__(kBlock, EXACT ? 6 : (ANY_HIT ? EARB_ANYHIT_MIN_BLOCKS : EARB_TRAV_MIN_BLOCKS)) wf_traverse_kernel(SceneDev sc, WfPool pool, RenderParams p) {
	extern __shared__ int2 stack_smem[];
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const int total = ANY_HIT ? pool.counts[pool.q_count_idx] : pool.counts[0];
	int* cursor = pool.counts + (ANY_HIT ? pool.q_cursor_idx : 3);
	TravState ts;
	ts.st.bind(stack_smem + threadIdx.x, blockDim.x); ts.st.sp = 0;
	ts.node = kEmptyChildDev;
	ts.best_t = 0.0f; ts.best_idx = 0; ts.best_slot = -1;
	ts.o = mk(0, 0, 0); ts.d = mk(0, 0, 0); ts.rs = make_setup(ts.o, ts.d);
	bool has_job = false, exhausted = total <= 0;
	uint2 job = make_uint2(0u, 0u);   // closest: x = slot; any-hit: the query
	int32_t pend = 0;                 // parked leaf (0 = none): tested one triangle per leaf round
	for (;;) {
		// Any-hit: a lane that reaches a leaf parks it and keeps walking inner nodes (it only blocks on a second
		// leaf): order does not matter for a yes/no answer and the leaf rounds fill up (-2 % time).  Closest hit:
		// the lane waits for its leaf, because the hit it may find prunes the rest of its walk (parking measured +10 %).
		if (ts.node < 0 && pend == 0) {
			pend = ts.node;
			if (ANY_HIT) pop_next<EXACT>(sc, ts); else ts.node = kWaitLeaf;
		}
		const bool inner = ts.node >= 0 && ts.node < kWaitLeaf;
		const bool parked = pend != 0;
		const unsigned m_inner = __ballot_sync(0xffffffffu, inner);
		const unsigned m_pend = __ballot_sync(0xffffffffu, parked);
		const unsigned m_idle = ~(m_inner | m_pend);
		const bool busy = (m_inner | m_pend) != 0u;
		if (!exhausted && (__popc(m_idle) >= sc.fetch_vote || !busy)) {
			// ---------------- retire finished lanes, fetch new work ----------------
			const bool idle = !inner && !parked;
			if (ANY_HIT) {
				const bool visible = idle && has_job && ts.best_idx == 0;
				const unsigned m_vis = __ballot_sync(0xffffffffu, visible);
				if (m_vis) {
					int base = 0;
					if (lane == 0) base = atomicAdd(pool.counts + 2, __popc(m_vis));
					base = __shfl_sync(0xffffffffu, base, 0);
					if (visible) st_stream(pool.vis_list + base + __popc(m_vis & lt_mask), job);
				}
			} else if (idle && has_job) {
				st_stream(pool.hit + job.x, make_int2(__float_as_int(ts.best_t), ts.best_slot));
			}
			int base = 0;
			const int want = __popc(m_idle);
			if (lane == 0) base = atomicAdd(cursor, want);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base + want >= total) exhausted = true;
			if (idle) {
				const int my = base + __popc(m_idle & lt_mask);
				has_job = my < total;
				if (has_job) {
					if (ANY_HIT) {
						job = ld_stream(pool.q_list + my);
						const uint32_t slot = job.x & ((1u << pool.slot_bits) - 1u), r = job.x >> pool.slot_bits, c = job.y & 0xffffu;
						const float4 s0 = ld_stream(pool.sh0 + slot);
						V3 x;
						if (pool.qx) { const float4 e = pool.qx[my]; x = mk(e.x, e.y, e.z); }
						else { const float* rp = p.rec[(size_t)c * p.n_rec + r].position; x = mk(rp[0], rp[1], rp[2]); }
						const V3 pnt = mk(s0.x, s0.y, s0.z);
						ts.begin<true>(pnt, vsub(x, pnt));   // LineSeg(p, x) = Ray(p, x - p)
					} else {
						job.x = (uint32_t)ld_stream(pool.trav_list + my);
						const float4 o4 = ld_stream(pool.ro + job.x), d4 = ld_stream(pool.rd + job.x);
						ts.begin<false>(mk(o4.x, o4.y, o4.z), mk(d4.x, d4.y, d4.z));
					}
				}
			}
			continue;
		}
		if (!busy) {
			// exhausted and nothing in flight: retire what is left and leave
			if (ANY_HIT) {
				const bool visible = has_job && ts.best_idx == 0;
				const unsigned m_vis = __ballot_sync(0xffffffffu, visible);
				if (m_vis) {
					int base = 0;
					if (lane == 0) base = atomicAdd(pool.counts + 2, __popc(m_vis));
					base = __shfl_sync(0xffffffffu, base, 0);
					if (visible) st_stream(pool.vis_list + base + __popc(m_vis & lt_mask), job);
				}
			} else if (has_job) {
				st_stream(pool.hit + job.x, make_int2(__float_as_int(ts.best_t), ts.best_slot));
			}
			break;
		}
		if (m_inner != 0u && __popc(m_pend) < sc.leaf_vote) {
			if (inner) node_step<EXACT>(sc, ts);
		} else {
			if (parked) {
				pend_step<ANY_HIT>(sc, ts, pend);
				if (!ANY_HIT && pend == 0) pop_next<EXACT>(sc, ts);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// K6 compaction with binning: the ray / query lists are counting-sorted by a 15-bit key so that lanes of a
// warp (which fetch consecutive list entries) start in the same ~4 m cell and walk the same top of the BVH:
// identical node addresses coalesce into one L1 wavefront, and all resident warps work on a narrow spatial
// window of the scene at any time.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread4(uint32_t v) {   // 4 bits -> every third bit
	return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}
__device__ __forceinline__ uint32_t cell_key(const WfPool& pool, float x, float y, float z) {
	const int cx = min(15, max(0, (int)((x - pool.cell_origin[0]) * pool.cell_scale[0])));
	const int cy = min(15, max(0, (int)((y - pool.cell_origin[1]) * pool.cell_scale[1])));
	const int cz = min(15, max(0, (int)((z - pool.cell_origin[2]) * pool.cell_scale[2])));
	return spread4((uint32_t)cx) | (spread4((uint32_t)cy) << 1) | (spread4((uint32_t)cz) << 2);   // 12-bit Morton code
}
__device__ __forceinline__ uint32_t spread5(uint32_t v) { return spread4(v & 15u) | ((v & 16u) << 8); }
// bin of a closest-hit ray: warps fetch consecutive list entries, so a finer key means more alike rays per warp
__device__ __forceinline__ uint32_t ray_bin(const WfPool& pool, float ox, float oy, float oz, float dx, float dy, float dz) {
	const uint32_t oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
	if (pool.ray_key == 2 || pool.ray_key == 3) {
		const int cx = min(31, max(0, (int)((ox - pool.cell_origin[0]) * pool.cell_scale[0] * 2.0f)));
		const int cy = min(31, max(0, (int)((oy - pool.cell_origin[1]) * pool.cell_scale[1] * 2.0f)));
		const int cz = min(31, max(0, (int)((oz - pool.cell_origin[2]) * pool.cell_scale[2] * 2.0f)));
		const uint32_t cell = spread5((uint32_t)cx) | (spread5((uint32_t)cy) << 1) | (spread5((uint32_t)cz) << 2);
		return pool.ray_key == 2 ? (oct << 15) | cell : (cell << 3) | oct;
	}
	const uint32_t cell = cell_key(pool, ox, oy, oz);
	if (pool.ray_key == 1) {
		const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
		const uint32_t order = (ax > ay ? 1u : 0u) | (ay > az ? 2u : 0u) | (ax > az ? 4u : 0u);
		return (((oct << 3) | order) << 12) | cell;
	}
	if (pool.ray_key == 4) return (cell << 3) | oct;
	return (oct << 12) | cell;
}
// exclusive scan of the two histograms, in place, two levels: every block scans 1024 bins and leaves its total in the
// block-prefix area; wf_scan_top_kernel turns the totals into prefixes (the scatter adds them).
__device__ __forceinline__ int block_exclusive_scan(int sum, int* warp_tot, int& total) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int inc = sum;
	for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) warp_tot[wid] = inc;
	__syncthreads();
	if (wid == 0) {
		int w = warp_tot[lane];
		for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
		warp_tot[lane] = w;
	}
	__syncthreads();
	total = warp_tot[31];
	return inc - sum + (wid ? warp_tot[wid - 1] : 0);
}
__global__ void __launch_bounds__(1024) wf_scan_kernel(WfPool pool) {
	__shared__ int warp_tot[32];
	int total;
	int* h = pool.bins + blockIdx.x * 1024;     // ray bins and query bins are contiguous: block b owns bins [1024 b, 1024 b + 1024)
	const int v = h[threadIdx.x];
	h[threadIdx.x] = block_exclusive_scan(v, warp_tot, total);
	if (threadIdx.x == 0) pool.bins[kRayBins + kSortBins + blockIdx.x] = total;
}
// block totals -> exclusive prefixes, separately for the ray blocks and the query blocks
__global__ void __launch_bounds__(kScanBlocks) wf_scan_top_kernel(WfPool pool) {
	__shared__ int warp_tot[32];
	int* h = pool.bins + kRayBins + kSortBins;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int v = h[threadIdx.x];
	int inc = v;
	for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) warp_tot[wid] = inc;
	__syncthreads();
	const int first_warp = threadIdx.x < kRayScanBlocks ? 0 : kRayScanBlocks / 32;   // both segments are whole warps
	int before = 0;
	for (int w = first_warp; w < wid; ++w) before += warp_tot[w];
	h[threadIdx.x] = before + inc - v;
}
// scatter the appended entries to their bins.  The position inside the bin is the value the shade kernel's counting
// atomic returned (kept in trav_rank / q_rank), so this pass is a plain gather / scatter: the first version counted
// with fire-and-forget atomics and took a second, returning atomic per entry here -- whose round trip was its whole cost.
__global__ void __launch_bounds__(256) wf_scatter_kernel(WfPool pool) {
	const int n_trav = pool.counts[0], n_q = pool.sort_queries ? pool.counts[1] : 0;
	const int stride = gridDim.x * blockDim.x;
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	for (int i = tid; i < n_trav; i += stride) {
		const uint2 e = ld_stream(pool.trav_tmp + i);
		const int pos = pool.bins[e.y] + pool.bins[kRayBins + kSortBins + (e.y >> 10)] + ld_stream(pool.trav_rank + i);
		st_stream(pool.trav_list + pos, (int)e.x);
	}
	for (int i = tid; i < n_q; i += stride) {
		const uint2 e = ld_stream(pool.q_tmp + i);
		const uint32_t bin = (e.y >> 16) & 0x7fffu;
		const int pos = pool.bins[kRayBins + bin] + pool.bins[kRayBins + kSortBins + kRayScanBlocks + (bin >> 10)] + ld_stream(pool.q_rank + i);
		st_stream(pool.q_list + pos, e);
	}
}

// ---------------------------------------------------------------------------------------------------
// K3 + K6 + K1: shade, refill, enqueue
// ---------------------------------------------------------------------------------------------------
#ifndef EARB_SHADE_MIN_BLOCKS
#define EARB_SHADE_MIN_BLOCKS 4
#endif
#ifndef EARB_COOP_REJECTION
#define EARB_COOP_REJECTION 1
#endif
// One try of Sample_Sphere / Sample_Hemi (src/Distributions.h:48-67) from the three words of a Philox block.
__device__ __forceinline__ V3 rejection_candidate(uint32_t w0, uint32_t w1, uint32_t w2) {
	return mk(fsub(fmul(Rng::to_unit(w0), 2.0f), 1.0f), fsub(fmul(Rng::to_unit(w1), 2.0f), 1.0f), fsub(fmul(Rng::to_unit(w2), 2.0f), 1.0f));
}
// sphere: 0.001 <= |cand|^2 <= 1; hemisphere: n . (cand / |cand|) >= 0 as well.  |cand| <= 1 and the float evaluation of
// n.v is off by ~4e-7 at most, so n.cand < -1e-5 means the exact test rejects and n.cand > 1e-5 means it accepts; only in
// between (probability ~1e-5 per try) is the reference's expression evaluated.
__device__ __forceinline__ bool rejection_try(uint32_t w0, uint32_t w1, uint32_t w2, V3 n, bool hemi) {
	const V3 cand = rejection_candidate(w0, w1, w2);
	const float l = vdot(cand, cand);
	const bool in_sphere = !(l < 0.001f || l > 1.0f);
	const float dc = fmaf(n.x, cand.x, fmaf(n.y, cand.y, n.z * cand.z));
	bool accept = in_sphere && (!hemi || dc > 1e-5f);
	if (in_sphere && hemi && fabsf(dc) <= 1e-5f) {   // too close to call on the unnormalised candidate
		const float s = fsqrt(l);
		accept = !(vdot(n, mk(fdiv(cand.x, s), fdiv(cand.y, s), fdiv(cand.z, s))) < 0.0f);
	}
	return accept;
}
#ifndef EARB_DEFER_NORMALISE
#define EARB_DEFER_NORMALISE 1
#endif
__global__ void __launch_bounds__(256, EARB_SHADE_MIN_BLOCKS) wf_shade_kernel(SceneDev sc, WfPool pool, RenderParams p) {
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;   // n_slots is a multiple of the block size
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	LocalCounters lc = {0, 0, 0, 0, 0, 0};
	uint4 m = ld_stream(pool.rm + slot);
	int bounce = (int)(m.z >> 16);
	int c = (int)(m.z & 0xffffu);
	bool alive = bounce != 0;
	const bool had_ray = alive;
	float4 ro = make_float4(0, 0, 0, 0), rd = make_float4(0, 0, 0, 0);
	// what this lane needs from the shared sampling loop below
	enum { kNone = 0, kBounce = 1, kEmit = 2 };
	int mode = kNone;
	bool mesh_emit = false;   // this slot starts a ray of a mesh source: hemisphere about the emitter's normal, bounce 0 is recorded
	V3 n = mk(0, 0, 0), pnt = mk(0, 0, 0), blend = mk(0, 0, 0), prev_dir = mk(0, 0, 0), o = mk(0, 0, 0);
	float intensity = 0.0f, path = 0.0f, spec = 0.0f, kept = 0.0f, t = 0.0f;
	bool refract = false;
	unsigned long long ray = ((unsigned long long)m.y << 32) | m.x;
	Rng rng;
	rng.start(p.seed, stream_key(p, c), ray, m.w);

	// ---- K6 + K1 (first half): slots that are empty on entry take the next ray id of the shard.  (A ray that ends in
	// this launch frees its slot for the NEXT launch: one idle iteration per ~50, and emission shares the sampling
	// loop with the bounces instead of running on one lane of the warp.) ----
	const unsigned dead = __ballot_sync(0xffffffffu, !alive);
	if (dead) {
		unsigned long long base = 0;
		const int want = __popc(dead);
		if (lane == 0) {
			base = *(volatile unsigned long long*)p.next_work;
			if ((long long)base < p.total_work) base = atomicAdd(p.next_work, (unsigned long long)want);
		}
		base = __shfl_sync(0xffffffffu, base, 0);
		const long long w = (long long)base + __popc(dead & lt_mask);
		if (!alive && w < p.total_work) {
			c = 0;
			while (c + 1 < p.n_ctx && w >= p.work_prefix[c + 1]) ++c;
			ray = (unsigned long long)(p.first_ray + (w - p.work_prefix[c]));
			rng.start(p.seed, stream_key(p, c), ray, 0);
			m.x = (uint32_t)ray; m.y = (uint32_t)(ray >> 32);
			++lc.rays;
			mode = kEmit;
			if (p.ctx[c].source_kind == EAR_B200_MESH_SOURCE) {
				// AbstractSoundFile::SoundRay of a mesh source (src/SoundFile.cpp:216-221): Mesh::SamplePoint picks a triangle
				// by area (src/Mesh.cpp:143-154: x = rangeRandom(0, total_area) = r * total_area + 0; x -= area until x < 0),
				// Triangle::SamplePoint (src/Triangle.cpp:44-53) a point in it; the direction is Sample_Hemi about the
				// triangle's own normal (the shared loop below).  If x never drops below 0 the reference keeps its
				// default-constructed zeros for point and normal.
				mesh_emit = true;
				const float total = p.ctx_emit_area[c];
				float x = fadd(fmul(rng.unit1(), total), 0.0f);
				const float4* et = sc.emitters + 4 * (size_t)p.ctx[c].emitter_first;
				const int n_e = p.ctx[c].emitter_count;
				for (int i = 0; i < n_e; ++i) {
					const float4 e0 = __ldg(et + 4 * i);
					x = fsub(x, e0.w);
					if (x < 0.0f) {
						const float4 e1 = __ldg(et + 4 * i + 1), e2 = __ldg(et + 4 * i + 2), e3 = __ldg(et + 4 * i + 3);
						float r1, r2, unused;
						rng.unit3(r1, r2, unused);
						const float sr1 = fsqrt(r1);
						pnt = vadd(vadd(vscale(mk(e0.x, e0.y, e0.z), fsub(1.0f, sr1)), vscale(mk(e1.x, e1.y, e1.z), fmul(sr1, fsub(1.0f, r2)))),
						           vscale(mk(e2.x, e2.y, e2.z), fmul(sr1, r2)));
						n = mk(e3.x, e3.y, e3.z);
						break;
					}
				}
			}
		}
	}
	const long long out_row = (long long)(ray - (unsigned long long)p.first_ray);   // parity harness: output row

	// ---- K3 (first half): consume the hit of the previous launch (Scene::Bounce, src/Scene.cpp:49-63) ----
	if (alive) {
		ro = ld_stream(pool.ro + slot); rd = ld_stream(pool.rd + slot);
		const int2 h = ld_stream(pool.hit + slot);
		++lc.segments;                                                            // one Scene::Bounce call
		o = mk(ro.x, ro.y, ro.z);
		const V3 d = mk(rd.x, rd.y, rd.z);
		intensity = ro.w; path = rd.w;
		if (p.hits) p.hits[out_row * p.max_bounces + bounce] = h.y >= 0 ? __float_as_int(__ldg(sc.tris + 4 * (size_t)h.y).w) : -1;
		if (h.y < 0) alive = false;                                               // escaped (src/Scene.cpp:166)
		else {
			t = __int_as_float(h.x);
			const float4 r1 = ldg_keep(sc.tris + 4 * (size_t)h.y + 1);
			const float4 r3 = ldg_keep(sc.tris + 4 * (size_t)h.y + 3);
			prev_dir = vnormalized(d);                                            // prev_ray_dir (:277) of this bounce
			const V3 tri_n = mk(r3.x, r3.y, r3.z);
			pnt = vadd(o, vscale(d, t));                                          // src/Mesh.cpp:48
			n = (vdot(tri_n, d) > 0.0f) ? vscale(tri_n, -1.0f) : tri_n;           // :49-53
			const float4 mat = __ldg(sc.materials + (size_t)__float_as_int(r1.w) * sc.n_bands + p.ctx[c].band);
			// Material::Bounce (src/Material.cpp:76-83); its comparisons against 0.0001 are in double
			if ((double)mat.x < 0.0001 && (double)mat.y < 0.0001) refract = false;
			else refract = !(rng.unit1() <= fdiv(mat.x, fadd(mat.x, mat.y)));
			spec = mat.w; kept = mat.z;
			if (refract) { n = vneg(n); blend = d; }                              // src/Scene.cpp:65-69
			else blend = vreflect(d, n);                                          // :71-73
			mode = kBounce;
		}
	}

	// ---- Sample_Sphere / Sample_Hemi (src/Distributions.h:48-67) for every lane that needs a direction: ONE flat
	// rejection loop (a try = one Philox block; sphere: 0.001 <= |v|^2 <= 1; hemisphere: n.v >= 0), so that a warp
	// runs max-over-lanes tries once instead of nesting the two rejections ----
#if EARB_DEFER_NORMALISE
	V3 v = mk(0, 0, 0);
	{
		// The hemisphere test n.v >= 0 is on the NORMALISED candidate v = cand / |cand| in the reference.  Its outcome is
		// decided on the unnormalised candidate whenever that is clear-cut (rejection_try).  The sqrt and the three
		// divisions then run ONCE per lane after the loop, with the warp converged, instead of once per sphere-accepted
		// try with ~5 lanes active (ncu: that block was 24 % of this kernel's instructions).
		//
		// A hemisphere try succeeds with probability pi/12 = 0.26, so a warp that lets every lane run its own tries needs
		// the MAXIMUM over 32 lanes -- 13.9 rounds measured, 8.7 lanes active on average (ncu: half of this kernel's
		// instructions).  Tries are independent Philox blocks (ray id, block index), so once half of the lanes are done
		// the idle ones evaluate LATER tries of the pending rays: T = 2, 4 .. 32 consecutive blocks per pending ray and
		// round; the first accepted one in block order wins, which is exactly the try the serial loop stops at.
		const bool hemi = mode == kBounce || mesh_emit;   // Sample_Hemi; point sources emit over the sphere
		bool pending = mode != kNone;
		uint32_t a0 = 0x80000000u, a1 = 0x80000000u, a2 = 0xffffff00u;   // words of the accepted try (default: (0, 0, 1))
#if EARB_COOP_REJECTION
		__shared__ uint4 coop_a[8][32];
		__shared__ float4 coop_b[8][32];
		const int wid = threadIdx.x >> 5;
#endif
		unsigned m_pending = __ballot_sync(0xffffffffu, pending);
		while (m_pending) {   // warp-uniform: one back edge
#if EARB_COOP_REJECTION
			const int n_pend = __popc(m_pending);
			if (n_pend <= 16) {
				const int log_t = 31 - __clz(32 / n_pend);          // T = 2^log_t tries per pending ray this round
				const int rank = __popc(m_pending & lt_mask);
				if (pending) {
					coop_a[wid][rank] = make_uint4(rng.ray_lo, rng.ray_hi, rng.ctx, rng.block);
					coop_b[wid][rank] = make_float4(n.x, n.y, n.z, hemi ? 1.0f : 0.0f);
				}
				__syncwarp();
				const int j = lane >> log_t, k = lane & ((1 << log_t) - 1);
				bool ok = false;
				uint32_t w0 = 0, w1 = 0, w2 = 0;
				if (j < n_pend) {
					const uint4 ra = coop_a[wid][j];
					const float4 rb = coop_b[wid][j];
					Rng other = rng;                                   // same seed words for the whole call
					other.ray_lo = ra.x; other.ray_hi = ra.y; other.ctx = ra.z; other.block = ra.w + (uint32_t)k;
					other.draw(w0, w1, w2);
					ok = rejection_try(w0, w1, w2, mk(rb.x, rb.y, rb.z), rb.w != 0.0f);
				}
				const unsigned m_ok = __ballot_sync(0xffffffffu, ok);
				__syncwarp();                                          // the slots are rewritten next round
				const unsigned span = log_t == 5 ? 0xffffffffu : ((1u << (1 << log_t)) - 1u);
				const unsigned mine = pending ? ((m_ok >> (rank << log_t)) & span) : 0u;
				const int first = __ffs(mine) - 1;                     // earliest accepted try of this lane's ray, or -1
				const int src = first >= 0 ? (rank << log_t) + first : lane;
				const uint32_t g0 = __shfl_sync(0xffffffffu, w0, src), g1 = __shfl_sync(0xffffffffu, w1, src), g2 = __shfl_sync(0xffffffffu, w2, src);
				if (pending) {
					if (first >= 0) { a0 = g0; a1 = g1; a2 = g2; rng.block += (uint32_t)first + 1u; pending = false; }
					else rng.block += 1u << log_t;
				}
			} else
#endif
			if (pending) {
				uint32_t w0, w1, w2;
				rng.draw(w0, w1, w2);
				if (rejection_try(w0, w1, w2, n, hemi)) { a0 = w0; a1 = w1; a2 = w2; pending = false; }
			}
			m_pending = __ballot_sync(0xffffffffu, pending);
		}
		if (mode != kNone) {
			const V3 cand = rejection_candidate(a0, a1, a2);
			const float s = fsqrt(vdot(cand, cand));
			v = mk(fdiv(cand.x, s), fdiv(cand.y, s), fdiv(cand.z, s));
		}
	}
#else
	V3 v = mk(0, 0, 0);
	{
		bool pending = mode != kNone;
		while (pending) {
			float u1, u2, u3;
			rng.unit3(u1, u2, u3);
			const V3 cand = mk(fsub(fmul(u1, 2.0f), 1.0f), fsub(fmul(u2, 2.0f), 1.0f), fsub(fmul(u3, 2.0f), 1.0f));
			const float l = vdot(cand, cand);
			if (!(l < 0.001f || l > 1.0f)) {
				const float s = fsqrt(l);
				v = mk(fdiv(cand.x, s), fdiv(cand.y, s), fdiv(cand.z, s));
				pending = (mode == kBounce || mesh_emit) && (vdot(n, v) < 0.0f);
			}
		}
	}
#endif
	__syncwarp();

	bool shaded = false;
	if (mode == kBounce) {
		// Sample_Hemi(v, n, reflection, factor), src/Distributions.h:71-75, then src/Scene.cpp:76-78, 154-175
		v = vnormalized(vadd(vscale(v, fsub(1.0f, spec)), vscale(blend, spec)));
		const float seg = vlength(vsub(pnt, o));
		intensity = fmul(intensity, pow_ref_hoisted(p.ctx[c].absorption_factor, pool.ctx_log2af[c], seg));   // :154
		path = fadd(path, seg);
		intensity = fmul(intensity, kept);                                        // :169-171
		ro = make_float4(pnt.x, pnt.y, pnt.z, intensity);
		rd = make_float4(v.x, v.y, v.z, path);
		if (invalid_float(intensity)) alive = false;                              // :175
		else {
			shaded = p.n_rec > 0;
			if (shaded) {
				st_stream(pool.sh0 + slot, ro);
				st_stream(pool.sh1 + slot, make_float4(n.x, n.y, n.z, path));
				st_stream(pool.sh2 + slot, make_float4(prev_dir.x, prev_dir.y, prev_dir.z,
				                             refract ? __int_as_float(__float_as_int(spec) | (int)0x80000000) : spec));
			}
			if ((double)intensity < 0.00000001) alive = false;                    // :275
			else if (bounce + 1 >= p.max_bounces) alive = false;                  // loop bound (:143)
		}
	} else if (mode == kEmit) {
		// AbstractSoundFile::SoundRay, point source (src/SoundFile.cpp:223-226).  Bounce 0 of the reference loop
		// records nothing for point sources (src/Scene.cpp:185) and leaves intensity 1, path 0.
		const float* sp = p.ctx[c].source_position;
		ro = mesh_emit ? make_float4(pnt.x, pnt.y, pnt.z, 1.0f) : make_float4(sp[0], sp[1], sp[2], 1.0f);
		rd = make_float4(v.x, v.y, v.z, 0.0f);
		bounce = 0;
		alive = 1 < p.max_bounces;
		if (mesh_emit && p.n_rec > 0) {
			// "In case the sound source emits from a mesh, the direct sound is sampled regardless" (src/Scene.cpp:185):
			// bounce 0 connects the emission point to every recorder with dot := 1 and no surface term (:205-216);
			// specularity 2 marks the record for the splat kernel
			shaded = true;
			st_stream(pool.sh0 + slot, ro);
			st_stream(pool.sh1 + slot, make_float4(n.x, n.y, n.z, 0.0f));
			st_stream(pool.sh2 + slot, make_float4(0.0f, 0.0f, 0.0f, 2.0f));
		}
	}
	m.w = rng.block;
	if (!alive && (had_ray || mode == kEmit) && p.final_state) {   // parity harness: state the ray ended with
		float* fs = p.final_state + 8 * out_row;
		fs[0] = ro.x; fs[1] = ro.y; fs[2] = ro.z; fs[3] = rd.x; fs[4] = rd.y; fs[5] = rd.z; fs[6] = ro.w; fs[7] = rd.w;
	}

	// ---- K4 enqueue: Scene::Connect is called for every recorder (src/Scene.cpp:188-195); its answer is only
	// used when dot(lsdir, n) > 0 (:209), so only those queries are traced.
	// Two passes over the recorders.  Pass 1 decides `facing` per (lane, recorder) and keeps the warp's ballot per recorder
	// in shared memory; ONE counter atomic then reserves the warp's range of the query list.  Pass 2 writes the entries,
	// four recorders at a time, so that the four rank atomics of a group are in flight together.  (The first version
	// took a counter atomic AND a rank atomic per recorder and waited for both before moving on: with 64 recorders 85 %
	// of this kernel's stall samples sat on those two round trips, profiles/r2_ncu_c5_wf_shade_kernel.txt.) ----
	// (One position atomic per BLOCK instead of per warp -- warp totals meeting in shared memory behind two barriers -- was
	// tried and lost: shade 181 -> 192 ms per 4e7 rays at C4, 36 -> 40 ms per 3e6 rays at C5.)
	// K6's atomics (list position of the warp's live rays, rank of each ray inside its bin) are issued early, between the
	// atomics of the query enqueue, so that the round trips overlap; their results are consumed at the end of the kernel
	const unsigned live = __ballot_sync(0xffffffffu, alive);
	const uint32_t my_ray_bin = alive ? ray_bin(pool, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z) : 0u;
	int live_base = 0, live_rank = 0;
	bool ray_atomics_issued = false;
	auto issue_ray_atomics = [&]() {
		if (ray_atomics_issued) return;
		ray_atomics_issued = true;
		if (live) {
			if (lane == 0) live_base = atomicAdd(pool.counts + 0, __popc(live));
			if (alive) live_rank = atomicAdd(pool.bins + my_ray_bin, 1);
		}
	};
	if (p.n_rec == 1) {
		// one recorder: nothing to batch -- the single-pass form (two passes cost 4 % of this kernel here)
		bool facing = false;
		if (shaded) {
			++lc.occlusion;
			const float* x = p.rec[c].position;
			const V3 seg = vsub(mk(x[0], x[1], x[2]), pnt);
			const float du = fmaf(seg.x, n.x, fmaf(seg.y, n.y, seg.z * n.z));
			const float mag = fabsf(seg.x) + fabsf(seg.y) + fabsf(seg.z);
			if (mesh_emit) facing = true;
			else if (fabsf(du) > 1e-4f * mag) facing = du > 0.0f;   // see the general form below
			else facing = vdot(vnormalized(seg), n) > 0.0f;
		}
		const unsigned mq = __ballot_sync(0xffffffffu, facing);
		if (mq) {
			int base = 0;
			if (lane == 0) base = atomicAdd(pool.counts + 1, __popc(mq));
			base = __shfl_sync(0xffffffffu, base, 0);
			if (facing) {
				const int at = base + __popc(mq & lt_mask);
				const uint32_t word_y = (uint32_t)c | ((uint32_t)(bounce & 1) << 31);
				if (pool.sort_queries) {
					const uint32_t bin = cell_key(pool, pnt.x, pnt.y, pnt.z);
					st_stream(pool.q_rank + at, atomicAdd(pool.bins + kRayBins + bin, 1));
					st_stream(pool.q_tmp + at, make_uint2((uint32_t)slot, word_y | (bin << 16)));
				} else {
					st_stream(pool.q_list + at, make_uint2((uint32_t)slot, word_y));
				}
			}
		}
	} else if (p.n_rec > 1) {
		__shared__ unsigned q_mask[8][256];   // [warp of the block][recorder] (n_rec <= 255)
		const int wq = threadIdx.x >> 5;
		int warp_total = 0;
		if (shaded) lc.occlusion += (unsigned long long)p.n_rec;
		for (int r = 0; r < p.n_rec; ++r) {
			bool facing = false;
			if (shaded) {
				const float* x = p.rec[(size_t)c * p.n_rec + r].position;
				// dot(lsdir, n) > 0 with lsdir = normalize(x - p) (src/Scene.cpp:202-209).  The sign is that of dot(x - p, n);
				// normalising (three divisions by the rounded length) and the three roundings of the dot move the value by a
				// few ulp of |x - p| at most, so the unnormalised dot decides whenever it is clearly away from zero and the
				// exact expression is evaluated only in between.
				const V3 seg = vsub(mk(x[0], x[1], x[2]), pnt);
				const float du = fmaf(seg.x, n.x, fmaf(seg.y, n.y, seg.z * n.z));
				const float mag = fabsf(seg.x) + fabsf(seg.y) + fabsf(seg.z);
				if (mesh_emit) facing = true;
				else if (fabsf(du) > 1e-4f * mag) facing = du > 0.0f;
				else facing = vdot(vnormalized(seg), n) > 0.0f;
			}
			const unsigned mq = __ballot_sync(0xffffffffu, facing);
			if (lane == 0) q_mask[wq][r] = mq;
			warp_total += __popc(mq);
		}
		__syncwarp();
		if (warp_total) {   // warp-uniform
			int run = 0;
			if (lane == 0) run = atomicAdd(pool.counts + 1, warp_total);
			issue_ray_atomics();   // K6's two atomics go out before anything waits for the one above
			run = __shfl_sync(0xffffffffu, run, 0);
			const uint32_t cell = pool.sort_queries ? cell_key(pool, pnt.x, pnt.y, pnt.z) : 0u;
			// The rays of a warp were fetched in bin order, so most of its hit points share a cell and their rank atomics
			// hit the SAME counter (32-way same-address returning atomics: at C5 they were 85 % of this kernel, 184 of 218 ms
			// per 3e6 rays).  Lanes with equal cells are grouped once; per recorder the first facing lane of a group takes
			// one atomic for all of them.
			const unsigned m_shaded = __ballot_sync(0xffffffffu, shaded);
			const unsigned peers = (shaded && pool.sort_queries) ? __match_any_sync(m_shaded, cell) : 0u;
			const uint32_t word_y = (uint32_t)c | ((uint32_t)(bounce & 1) << 31);
			constexpr int kGroup = 4;
			for (int r0 = 0; r0 < p.n_rec; r0 += kGroup) {   // warp-uniform
				int at[kGroup], rank[kGroup], who[kGroup];     // who = leader lane | position inside the group << 8
				bool mine[kGroup];
#pragma unroll
				for (int g = 0; g < kGroup; ++g) {
					const int r = r0 + g;
					const unsigned mq = r < p.n_rec ? q_mask[wq][r] : 0u;
					mine[g] = (mq >> lane) & 1u;
					at[g] = run + __popc(mq & lt_mask);
					run += __popc(mq);
					rank[g] = 0; who[g] = lane;
					if (mine[g] && pool.sort_queries) {
						const unsigned group = peers & mq;               // facing lanes of this recorder in my cell
						const int leader = __ffs(group) - 1;
						who[g] = leader | (__popc(group & lt_mask) << 8);
						if (lane == leader) rank[g] = atomicAdd(pool.bins + kRayBins + ((((uint32_t)r & 7u) << 12) | cell), __popc(group));
					}
				}
#pragma unroll
				for (int g = 0; g < kGroup; ++g) {
					const int first = __shfl_sync(0xffffffffu, rank[g], who[g] & 31);   // every lane takes part
					if (!mine[g]) continue;
					const uint32_t r = (uint32_t)(r0 + g);
					const uint32_t word_x = (uint32_t)slot | (r << pool.slot_bits);
					if (pool.sort_queries) {
						const uint32_t bin = ((r & 7u) << 12) | cell;
						st_stream(pool.q_rank + at[g], first + (who[g] >> 8));
						st_stream(pool.q_tmp + at[g], make_uint2(word_x, word_y | (bin << 16)));
					} else {
						st_stream(pool.q_list + at[g], make_uint2(word_x, word_y));
					}
				}
			}
		}
	}
	bounce = alive ? bounce + 1 : 0;
	m.z = (uint32_t)c | ((uint32_t)bounce << 16);
	st_stream(pool.rm + slot, m);
	if (alive) { st_stream(pool.ro + slot, ro); st_stream(pool.rd + slot, rd); }
	// ---- K6 compaction: slots that need a closest-hit query next, binned by (direction octant, origin cell) ----
	issue_ray_atomics();
	if (live) {
		const int base = __shfl_sync(0xffffffffu, live_base, 0);
		if (alive) {
			const int at = base + __popc(live & lt_mask);
			st_stream(pool.trav_rank + at, live_rank);
			st_stream(pool.trav_tmp + at, make_uint2((uint32_t)slot, my_ray_bin));
		}
	}
	// the launch loop stops when no slot holds a ray and the shard's queue is dry
	// Counters: warp sums -> block sums in shared memory -> ONE atomic per counter and block, spread over kCounterParts
	// partial rows (folded into p.counters when the wavefront loop ends).  The first version added every warp's sums
	// straight to p.counters: 1.5 M same-address 64-bit atomics per launch of 16 Mi slots, and how fast the L2 retires
	// those depended on where the caller's counter block happened to live (shade 400 ms per 1e8 rays through
	// ear_b200_trace_device with a torch tensor, 530 ms through ear_b200_render: profiles/r2_ab_closest.txt).
	__shared__ unsigned long long w_cnt[8][3];
	const unsigned long long v0 = warp_sum(lc.rays), v1 = warp_sum(lc.segments), v2 = warp_sum(lc.occlusion);
	if (lane == 0) { w_cnt[threadIdx.x >> 5][0] = v0; w_cnt[threadIdx.x >> 5][1] = v1; w_cnt[threadIdx.x >> 5][2] = v2; }
	__syncthreads();
	if (threadIdx.x < 3) {
		unsigned long long v = 0;
#pragma unroll
		for (int w = 0; w < 8; ++w) v += w_cnt[w][threadIdx.x];
		if (v) atomicAdd(p.counter_parts + (size_t)(blockIdx.x % kCounterParts) * 4 + threadIdx.x, v);
	}
}

// adds the partial rows of the shade kernel's counters (rays, segments, occlusion queries) to the call's counters and
// clears them for the next call
__global__ void __launch_bounds__(kCounterParts) wf_fold_counters_kernel(RenderParams p) {
	__shared__ unsigned long long part[kCounterParts / 32][3];
	unsigned long long v[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		v[k] = p.counter_parts[(size_t)threadIdx.x * 4 + k];
		p.counter_parts[(size_t)threadIdx.x * 4 + k] = 0;
		v[k] = warp_sum(v[k]);
	}
	if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) part[threadIdx.x >> 5][k] = v[k];
	__syncthreads();
	if (threadIdx.x < 3) {
		unsigned long long t = 0;
		for (int w = 0; w < kCounterParts / 32; ++w) t += part[w][threadIdx.x];
		if (t) atomicAdd(p.counters + threadIdx.x, t);
	}
}

// ---------------------------------------------------------------------------------------------------
// K5: contribution weight (src/Scene.cpp:197-263) + Recorder::Record for every visible query
// ---------------------------------------------------------------------------------------------------
// weight of one visible query; false when the reference rejects it (INVALID_FLOAT, src/Scene.cpp:254)
__device__ __forceinline__ bool splat_weight(const WfPool& pool, const RenderParams& p, uint2 q, const ear_b200_recorder& rec, uint32_t slot,
                                             uint32_t c, V3& lsdir, float& contrib, float& dist) {
	const float4 s0 = ld_stream(pool.sh0 + slot), s1 = ld_stream(pool.sh1 + slot), s2 = ld_stream(pool.sh2 + slot);
	const V3 pnt = mk(s0.x, s0.y, s0.z), n = mk(s1.x, s1.y, s1.z), prev_dir = mk(s2.x, s2.y, s2.z);
	const bool refract = (__float_as_uint(s2.w) >> 31) != 0u;   // sign bit carries the bounce type
	const float spec = fabsf(s2.w);
	const float intensity = s0.w, path = s1.w;
	const float af = p.ctx[c].absorption_factor;
	const V3 segv = vsub(mk(rec.position[0], rec.position[1], rec.position[2]), pnt);
	lsdir = vnormalized(segv);
	float factor;
	if (spec > 1.5f) factor = 1.0f;                                              // bounce 0 of a mesh source: no surface term (:216)
	else if (!refract) {                                                         // :219-235
		const V3 rv = vreflect(prev_dir, n);
		const float diff = -vdot(n, prev_dir);
		const float dsp = vdot(rv, lsdir);
		const float specf = (0.0f < dsp) ? dsp : 0.0f;
		factor = fadd(fmul(fmul(spec, 1001.0f), pow_ref(specf, 1000.0f)), fmul(fsub(1.0f, spec), diff));
	} else {                                                                     // :236-247
		const float diff = vdot(n, prev_dir);
		const float dsp = vdot(prev_dir, lsdir);
		const float specf = (0.0f < dsp) ? dsp : 0.0f;
		factor = fadd(fmul(fmul(spec, 1001.0f), pow_ref(specf, 1000.0f)), fmul(fsub(1.0f, spec), diff));
	}
	contrib = spec > 1.5f ? intensity : fmul(intensity, factor);
	const float l = vlength(segv);                                               // :250
	contrib = fmul(contrib, pow_ref_hoisted(af, pool.ctx_log2af[c], l));
	contrib = fmul(contrib, fdiv(2.0f, fmul(fmul(fmul(4.0f, PI_F), l), l)));     // INV_HEMI_2, :252
	if (invalid_float(contrib)) return false;
	if (q.y >> 31) contrib = fmul(contrib, -1.0f);                               // odd bounce (:257)
	dist = fadd(path, l);
	return true;
}

#ifndef EARB_SPLAT_MIN_BLOCKS
#define EARB_SPLAT_MIN_BLOCKS 5    // resident 256-thread blocks per SM asked of the compiler (5 <-> 48 registers)
#endif
__global__ void __launch_bounds__(256, EARB_SPLAT_MIN_BLOCKS) wf_splat_kernel(WfPool pool, RenderParams p) {
	const int total = pool.counts[2];
	LocalCounters lc = {0, 0, 0, 0, 0, 0};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
		const uint2 q = ld_stream(pool.vis_list + i);
		const uint32_t slot = q.x & ((1u << pool.slot_bits) - 1u), r = q.x >> pool.slot_bits, c = q.y & 0xffffu;
		const ear_b200_recorder& rec = p.rec[(size_t)c * p.n_rec + r];
		V3 lsdir; float contrib, dist;
		if (!splat_weight(pool, p, q, rec, slot, c, lsdir, contrib, dist)) continue;
		const size_t track = ((size_t)c * p.n_rec + r) * p.tpr;
		GlobalSink sink = {p.hist + track * p.n_bins, p.range + track * 2, p.n_bins};
		record(rec, sink, p.n_bins, lsdir, contrib, fdiv(dist, 343.0f), dist, p.ctx[c].band, lc);
	}
	const unsigned long long v3 = warp_sum(lc.contributions), v4 = warp_sum(lc.bin_updates), v5 = warp_sum(lc.dropped);
	if ((threadIdx.x & 31) == 0 && (v3 | v4 | v5)) {
		atomicAdd(p.counters + 3, v3); atomicAdd(p.counters + 4, v4);
		if (v5) atomicAdd(p.counters + 5, v5);
	}
}

// ---------------------------------------------------------------------------------------------------
// K5, privatised form (EAR_B200_SPLAT=window): the visible queries of an iteration are counting-sorted by
// (context, recorder); a block then owns a run of one recorder's contributions, accumulates them into a
// shared-memory copy of a TIME WINDOW of that recorder's track(s) and flushes the window once, lane i of a warp
// adding bin i -- one coalesced RED per 32 touched bins instead of one RED per ramp sample.  It works because a pool
// that runs in lockstep (bounce cap << pool generations) delivers, in one iteration, rays of nearly the same bounce
// number: their arrival times cluster in a window a few thousand bins wide.  Samples outside the window go
// straight to the global histogram, so the result is the same sum in a different order.
// ---------------------------------------------------------------------------------------------------
constexpr int kPrivWindow = 16384;    // floats of shared memory per block (one mono track, or 2 x 8192 for stereo)
constexpr int kPrivChunk = 32768;     // visible queries per block pass
constexpr int kPrivMaxPairs = 65536;  // (context, recorder) pairs the sort supports; beyond that the direct form runs

__global__ void __launch_bounds__(256) wf_vis_count_kernel(WfPool pool, RenderParams p) {
	const int total = pool.counts[2];
	const int stride = gridDim.x * blockDim.x;
	for (int base = blockIdx.x * blockDim.x; base < total; base += stride) {   // warp-uniform trip count
		const int i = base + threadIdx.x;
		const bool on = i < total;
		int pair = -1;
		if (on) {
			const uint2 q = ld_stream(pool.vis_list + i);
			pair = (int)(q.y & 0xffffu) * p.n_rec + (int)(q.x >> pool.slot_bits);
		}
		// one atomic per distinct pair in the warp
		const unsigned peers = __match_any_sync(0xffffffffu, pair);
		if (on && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(pool.pair_count + pair, __popc(peers));
	}
}
// exclusive scan of the pair counts (one block; n_pairs <= kPrivMaxPairs), cursors reset
__global__ void __launch_bounds__(1024) wf_vis_scan_kernel(WfPool pool, int n_pairs) {
	__shared__ int warp_tot[32];
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n_pairs; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = i < n_pairs ? pool.pair_count[i] : 0;
		int total;
		const int ex = block_exclusive_scan(v, warp_tot, total);
		if (i < n_pairs) { pool.pair_base[i] = carry + ex; pool.pair_count[i] = 0; }
		__syncthreads();
		if (threadIdx.x == 0) carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) pool.pair_base[n_pairs] = carry;
}
__global__ void __launch_bounds__(256) wf_vis_scatter_kernel(WfPool pool, RenderParams p) {
	const int total = pool.counts[2];
	const int stride = gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	for (int base = blockIdx.x * blockDim.x; base < total; base += stride) {
		const int i = base + threadIdx.x;
		const bool on = i < total;
		int pair = -1;
		uint2 q = make_uint2(0u, 0u);
		if (on) {
			q = ld_stream(pool.vis_list + i);
			pair = (int)(q.y & 0xffffu) * p.n_rec + (int)(q.x >> pool.slot_bits);
		}
		const unsigned peers = __match_any_sync(0xffffffffu, pair);
		const int leader = __ffs(peers) - 1;
		int at = 0;
		if (on && lane == leader) at = atomicAdd(pool.pair_count + pair, __popc(peers));   // pair_count doubles as the cursor
		at = __shfl_sync(0xffffffffu, at, leader);
		if (on) st_stream(pool.vis_sorted + pool.pair_base[pair] + at + __popc(peers & ((1u << lane) - 1u)), q);
	}
}

__device__ __forceinline__ int block_sum_i(int v, int* scratch) {
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	v = (threadIdx.x & 31) < (blockDim.x >> 5) ? scratch[threadIdx.x & 31] : 0;
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	return v;
}
__device__ __forceinline__ int block_min_i32(int v, int* scratch) {
	for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	v = (threadIdx.x & 31) < (blockDim.x >> 5) ? scratch[threadIdx.x & 31] : 0x7fffffff;
	for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	__syncthreads();
	return v;
}

__global__ void __launch_bounds__(256) wf_splat_window_kernel(WfPool pool, RenderParams p) {
	extern __shared__ float win[];            // kPrivWindow floats
	__shared__ int scratch[32];
	const int total = pool.counts[2];
	const int n_chunks = (total + kPrivChunk - 1) / kPrivChunk;
	LocalCounters lc = {0, 0, 0, 0, 0, 0};
	for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
		const int end = min(total, (chunk + 1) * kPrivChunk);
		int pos = chunk * kPrivChunk;
		while (pos < end) {   // block-uniform: one run of a single (context, recorder) pair per round
			const uint2 q0 = pool.vis_sorted[pos];
			const uint32_t r0 = q0.x >> pool.slot_bits, c0 = q0.y & 0xffffu;
			const int pair = (int)c0 * p.n_rec + (int)r0;
			const int run_end = min(end, pool.pair_base[pair + 1]);
			const ear_b200_recorder& rec = p.rec[pair];
			const int n_tr = rec.kind == EAR_B200_STEREO ? 2 : 1;
			const int width = kPrivWindow / n_tr;
			// window placement: mean first bin of a sample of the run (one entry per thread)
			const int len = run_end - pos;
			int my_s = 0, my_n = 0;
			if ((int)threadIdx.x < len) {
				const uint2 q = pool.vis_sorted[pos + (int)(((long long)threadIdx.x * len) / blockDim.x)];
				const uint32_t slot = q.x & ((1u << pool.slot_bits) - 1u);
				const float4 s0 = pool.sh0[slot];
				const float path = pool.sh1[slot].w;
				const V3 segv = vsub(mk(rec.position[0], rec.position[1], rec.position[2]), mk(s0.x, s0.y, s0.z));
				my_s = min(p.n_bins >> 8, __double2int_rz((double)fdiv(fadd(path, vlength(segv)), 343.0f) * 44100.0) >> 8);   // in units of 256 bins: the sum stays in range
				my_n = 1;
			}
			const int sum = block_sum_i(my_s, scratch), cnt = block_sum_i(my_n, scratch);
			int lo = (int)(((long long)sum << 8) / max(cnt, 1)) - width / 2;
			lo = max(0, min(lo, p.n_bins - width)) & ~31;
			for (int j = threadIdx.x; j < kPrivWindow; j += blockDim.x) win[j] = 0.0f;
			__syncthreads();
			const size_t track = (size_t)pair * p.tpr;
			WindowSink sink = {p.hist + track * p.n_bins, p.n_bins, win, lo, width, {0x7fffffff, 0x7fffffff}, {-1, -1}};
			for (int i = pos + threadIdx.x; i < run_end; i += blockDim.x) {
				const uint2 q = ld_stream(pool.vis_sorted + i);
				const uint32_t slot = q.x & ((1u << pool.slot_bits) - 1u);
				V3 lsdir; float contrib, dist;
				if (!splat_weight(pool, p, q, rec, slot, c0, lsdir, contrib, dist)) continue;
				record(rec, sink, p.n_bins, lsdir, contrib, fdiv(dist, 343.0f), dist, p.ctx[c0].band, lc);
			}
			__syncthreads();
			// flush: lane i of a warp adds bin i -> one coalesced RED per 32 touched bins
			for (int k = 0; k < n_tr; ++k) {
				float* tr = p.hist + (track + k) * p.n_bins + lo;
				const int room = min(width, p.n_bins - lo);
				for (int j = threadIdx.x; j < room; j += blockDim.x) {
					const float v = win[k * width + j];
					if (v != 0.0f) atomicAdd(tr + j, v);
				}
				const int t_lo = block_min_i32(sink.t_lo[k], scratch);
				const int t_hi = -block_min_i32(-sink.t_hi[k], scratch);
				if (threadIdx.x == 0 && t_hi >= 0) touch_range(p.range + (track + k) * 2, t_lo, t_hi);
			}
			__syncthreads();
			pos = run_end;
		}
	}
	const unsigned long long v3 = warp_sum(lc.contributions), v4 = warp_sum(lc.bin_updates), v5 = warp_sum(lc.dropped);
	if ((threadIdx.x & 31) == 0 && (v3 | v4 | v5)) {
		atomicAdd(p.counters + 3, v3); atomicAdd(p.counters + 4, v4);
		if (v5) atomicAdd(p.counters + 5, v5);
	}
}

// ---------------------------------------------------------------------------------------------------
// H1 harness glue: explicit rays / segments through the same persistent traversal kernels
// ---------------------------------------------------------------------------------------------------
__global__ void wf_load_rays_kernel(WfPool pool, const float* origins, const float* dirs, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	pool.ro[i] = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], 1.0f);
	pool.rd[i] = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], 0.0f);
	pool.trav_list[i] = i;
	if (i == 0) { pool.counts[0] = n; pool.counts[3] = 0; }
}
__global__ void wf_store_hits_kernel(SceneDev sc, WfPool pool, int n, int32_t* tri_index, float* t_out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int2 h = pool.hit[i];
	tri_index[i] = h.y >= 0 ? __float_as_int(__ldg(sc.tris + 4 * (size_t)h.y).w) : -1;
	t_out[i] = __int_as_float(h.x);
}
__global__ void wf_load_segments_kernel(WfPool pool, float4* qx, const float* pp, const float* xx, int n, uint8_t* out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	pool.sh0[i] = make_float4(pp[3 * i], pp[3 * i + 1], pp[3 * i + 2], 0.0f);
	qx[i] = make_float4(xx[3 * i], xx[3 * i + 1], xx[3 * i + 2], 0.0f);
	pool.q_list[i] = make_uint2((uint32_t)i, 0u);
	out[i] = 1;   // occluded unless the traversal reports the query visible
	if (i == 0) { pool.counts[1] = n; pool.counts[2] = 0; pool.counts[4] = 0; pool.counts[5] = 0; pool.counts[6] = 0; }
}
__global__ void wf_mark_visible_kernel(WfPool pool, uint8_t* out) {
	const int total = pool.counts[2];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) out[pool.vis_list[i].x & ((1u << pool.slot_bits) - 1u)] = 0;
}
__global__ void wf_ctx_table_kernel(WfPool pool, RenderParams p) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < p.n_ctx) pool.ctx_log2af[c] = log2_ref(p.ctx[c].absorption_factor);
}
// Start of an iteration: the eight list counters and the bins of the two counting sorts back to zero.  A kernel, not
// cudaMemsetAsync: on a stream of its own the driver ran the two memsets of an iteration on a copy engine, and the
// hand-over from the copy engine to the next kernel cost ~0.2 ms per iteration -- 32 ms per 4e7 rays, 95 ms per 1e8, all of it
// showing up in front of the shade kernel (profiles/r2_ab_closest.txt, "diag2"); on the legacy default stream it did not.
__global__ void __launch_bounds__(256) wf_clear_kernel(WfPool pool) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < 8) pool.counts[i] = 0;
	if (i < kBinsTotal) pool.bins[i] = 0;
}
__global__ void wf_fill_int_kernel(int32_t* dst, long long n, int32_t v) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}

}  // namespace earb
