// Recorder visibility maps: K4 (Scene::Connect / Mesh::LineIntersection, src/Scene.cpp:84-96, src/Mesh.cpp:58-71)
// for the render loop, where every occlusion query ends at one of a handful of fixed points (the recorders).
//
// For a recorder position X the directions around X are cut into a cube map (6 faces x res x res texels); each
// texel holds the list of triangles whose projection from X, dilated by the lateral slop of the float32
// Moeller-Trumbore test, touches the texel.  A query P -> X then looks up the texel of (P - X) and runs the
// reference's exact float test on that short list -- no tree walk, no stack, no divergence beyond the list
// length.  The list is a SUPERSET of the triangles the float test can accept for any segment through X in that
// texel (the line must pass within h of the triangle; h as in bvh_build.cpp), so the yes/no answer is the same as
// the O(T) loop's.  Texels with long lists (grazing views along a wall) fall back to the BVH any-hit kernel.
#pragma once
#include "wavefront.cuh"
#include "vismap_geom.cuh"

namespace earb {

#ifndef EARB_VIS_BATCH
#define EARB_VIS_BATCH 2
#endif
constexpr int kVisBatch = EARB_VIS_BATCH;   // candidates whose loads are issued together in the lookup loop
constexpr int kVisMaxList = 256;  // longer texel lists are not stored: those queries walk the BVH instead.  Per 2e7 rays, with
                                  // conservative rasterisation and two-candidate rounds: 64 -> 125 ms, 96 -> 113.5, 128 -> 108,
                                  // 192 -> 102, 255 -> 90 (the hall's longest list is 239: no fallback at all).  Before those two
                                  // changes the optimum was 96.

// pass 0: count list lengths; pass 1: fill the lists (cursor = running fill position per texel).
// One (triangle, face) footprint per thread; footprints of more than a few texels are spread over the warp
// (a near triangle covers thousands of texels: left to one thread its atomics serialise into hundreds of ms).
// Texels whose list is longer than the cap are never looked up (the query goes to the BVH), so pass 1 skips them:
// their offset carries kVisOverlong and they own no items.
constexpr int kVisOverlong = (int)0x80000000u;

template <int PASS>
__device__ __forceinline__ void vis_emit(int texel, int t, int* counts_or_cursor, const int* offsets, int* items) {
	if (PASS == 0) atomicAdd(counts_or_cursor + texel, 1);
	else {
		const int o = offsets[texel];
		if (o >= 0) items[o + atomicAdd(counts_or_cursor + texel, 1)] = t;
	}
}

template <int PASS>
__global__ void __launch_bounds__(128) vis_build_kernel(SceneDev sc, double x0, double x1, double x2, int res, double reach, double maxabs,
                                                        int* counts_or_cursor, const int* offsets, int* items, int id_bits) {
	// neighbouring triangles cover the same texels: taking them in bit-reversed order keeps the atomics of the
	// threads in flight on different addresses
	const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long id = (long long)(__brevll(tid) >> (64 - id_bits));
	const int lane = threadIdx.x & 31;
	const double X[3] = {x0, x1, x2};
	int i0 = 0, i1 = -1, j0 = 0, j1 = -1, t = 0, face = 0;
	float edge[9] = {0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 1.0f};   // accepts every texel
	bool valid = id < 6LL * sc.n_tris;
	if (valid) {
		bool has_edges;
		t = (int)(id / 6); face = (int)(id % 6);
		valid = vis_footprint(sc, t, X, res, face, reach, maxabs, i0, i1, j0, j1, edge, has_edges);
		if (!has_edges) { for (int k = 0; k < 3; ++k) { edge[3 * k] = 0.0f; edge[3 * k + 1] = 0.0f; edge[3 * k + 2] = 1.0f; } }
	}
	const int w = valid ? i1 - i0 + 1 : 0;
	const int area = valid ? w * (j1 - j0 + 1) : 0;
	constexpr int kOwn = 4;
	if (area > 0 && area <= kOwn)
		for (int k = 0; k < area; ++k) {
			const int i = i0 + k % w, j = j0 + k / w;
			if (vis_covers(edge, res, i, j)) vis_emit<PASS>((face * res + j) * res + i, t, counts_or_cursor, offsets, items);
		}
	unsigned big = __ballot_sync(0xffffffffu, area > kOwn);
	while (big) {
		const int src = __ffs(big) - 1;
		big &= big - 1;
		const int b_i0 = __shfl_sync(0xffffffffu, i0, src), b_j0 = __shfl_sync(0xffffffffu, j0, src), b_w = __shfl_sync(0xffffffffu, w, src);
		const int b_area = __shfl_sync(0xffffffffu, area, src), b_t = __shfl_sync(0xffffffffu, t, src), b_face = __shfl_sync(0xffffffffu, face, src);
		float e[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) e[k] = __shfl_sync(0xffffffffu, edge[k], src);
		for (int k = lane; k < b_area; k += 32) {
			const int i = b_i0 + k % b_w, j = b_j0 + k / b_w;
			if (vis_covers(e, res, i, j)) vis_emit<PASS>((b_face * res + j) * res + i, b_t, counts_or_cursor, offsets, items);
		}
	}
}

// Sort-based build (default).  The fill pass above takes one RETURNING atomic per (triangle, texel) entry -- 132 M of
// them for the 1M-triangle hall, 34 ms -- and leaves every list in arbitrary order.  Instead, with NO atomic at all: pass 0
// counts the entries each footprint emits; one scan gives every footprint its place in a flat array; pass 1 writes
// (key, triangle) pairs there, key = texel << d | distance of the triangle from the recorder in d bits; ONE radix sort
// (cub::DeviceRadixSort) then yields all lists at once, each ordered nearest-first -- the order the lookups want (a
// blocked query stops at the first triangle its segment crosses) -- and the list offsets are read off the sorted keys
// (vis_offsets_from_keys_kernel).  (The first version of pass 0 also counted list lengths with one RED per entry: 10 ms
// of its 14.6.)
template <int PASS>
__global__ void __launch_bounds__(128) vis_emit_kernel(SceneDev sc, double x0, double x1, double x2, int res, double reach, double maxabs,
                                                       int id_bits, int* pair_count, const int* pair_base,
                                                       uint32_t* keys, int* vals, int dist_bits, float dist_scale) {
	const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long id = (long long)(__brevll(tid) >> (64 - id_bits));
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const double X[3] = {x0, x1, x2};
	int i0 = 0, i1 = -1, j0 = 0, j1 = -1, t = 0, face = 0;
	float edge[9] = {0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 1.0f};
	bool valid = id < 6LL * sc.n_tris;
	uint32_t dq = 0;
	if (valid) {
		bool has_edges;
		t = (int)(id / 6); face = (int)(id % 6);
		valid = vis_footprint(sc, t, X, res, face, reach, maxabs, i0, i1, j0, j1, edge, has_edges);
		if (!has_edges) { for (int k = 0; k < 3; ++k) { edge[3 * k] = 0.0f; edge[3 * k + 1] = 0.0f; edge[3 * k + 2] = 1.0f; } }
		if (PASS == 1 && valid) {
			const float4 r0 = __ldg(sc.tris + 4 * (size_t)t), r1 = __ldg(sc.tris + 4 * (size_t)t + 1), r2 = __ldg(sc.tris + 4 * (size_t)t + 2);
			const float cx = r0.x + (r1.x + r2.x) * (1.0f / 3.0f) - (float)x0, cy = r0.y + (r1.y + r2.y) * (1.0f / 3.0f) - (float)x1,
			            cz = r0.z + (r1.z + r2.z) * (1.0f / 3.0f) - (float)x2;
			dq = (uint32_t)min((float)((1u << dist_bits) - 1u), sqrtf(cx * cx + cy * cy + cz * cz) * dist_scale);
		}
	}
	const int w = valid ? i1 - i0 + 1 : 0;
	const int area = valid ? w * (j1 - j0 + 1) : 0;
	const int base = PASS == 1 ? pair_base[tid] : 0;
	int mine = 0;
	constexpr int kOwn = 4;
	if (area > 0 && area <= kOwn)
		for (int k = 0; k < area; ++k) {
			const int i = i0 + k % w, j = j0 + k / w;
			if (vis_covers(edge, res, i, j)) {
				const int texel = (face * res + j) * res + i;
				if (PASS == 1) { keys[base + mine] = ((uint32_t)texel << dist_bits) | dq; vals[base + mine] = t; }
				++mine;
			}
		}
	unsigned big = __ballot_sync(0xffffffffu, area > kOwn);
	while (big) {
		const int src = __ffs(big) - 1;
		big &= big - 1;
		const int b_i0 = __shfl_sync(0xffffffffu, i0, src), b_j0 = __shfl_sync(0xffffffffu, j0, src), b_w = __shfl_sync(0xffffffffu, w, src);
		const int b_area = __shfl_sync(0xffffffffu, area, src), b_t = __shfl_sync(0xffffffffu, t, src), b_face = __shfl_sync(0xffffffffu, face, src);
		const int b_base = __shfl_sync(0xffffffffu, base, src);
		const uint32_t b_dq = __shfl_sync(0xffffffffu, dq, src);
		float e[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) e[k] = __shfl_sync(0xffffffffu, edge[k], src);
		int run = 0;
		for (int k0 = 0; k0 < b_area; k0 += 32) {   // warp-uniform trip count: every lane votes
			const int k = k0 + lane;
			const int i = b_i0 + k % b_w, j = b_j0 + k / b_w;
			const bool covered = k < b_area && vis_covers(e, res, i, j);
			const unsigned m = __ballot_sync(0xffffffffu, covered);
			if (PASS == 1 && covered) {
				const int texel = (b_face * res + j) * res + i;
				const int at = b_base + run + __popc(m & lt_mask);
				keys[at] = ((uint32_t)texel << dist_bits) | b_dq; vals[at] = b_t;
			}
			run += __popc(m);
		}
		if (lane == src) mine = run;
	}
	if (PASS == 0) pair_count[tid] = mine;
}

// 64-bit total of the per-footprint entry counts (the scan that places them works in 32 bits: a map with 2^31 entries or
// more is not built, its recorder's queries walk the BVH)
__global__ void __launch_bounds__(256) vis_pair_total_kernel(const int* pair_count, size_t n, unsigned long long* total) {
	__shared__ unsigned long long part[8];
	unsigned long long s = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += (unsigned long long)pair_count[i];
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int k = 1; k < 8; ++k) s += part[k];
		atomicAdd(total, s);
	}
}

// offsets[t] = first position in the sorted pair array whose texel is >= t (offsets[n_tex] = total): entry i closes the
// texels after its predecessor's up to its own.  Long runs of empty texels (a cube face that sees nothing) are filled by
// the whole warp.
__global__ void __launch_bounds__(256) vis_offsets_from_keys_kernel(const uint32_t* keys, int total, int dist_bits, int n_tex, int* offsets) {
	const int lane = threadIdx.x & 31;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long base = (long long)blockIdx.x * blockDim.x; base <= total; base += stride) {   // warp-uniform trip count
		const long long i = base + threadIdx.x;
		int first = 0, last = -1;   // texels [first, last] get offset i
		if (i <= total) {
			last = i < total ? (int)(keys[i] >> dist_bits) : n_tex;
			first = i > 0 ? (int)(keys[i - 1] >> dist_bits) + 1 : 0;
		}
		constexpr int kInline = 4;
		if (last - first < kInline) for (int t = first; t <= last; ++t) offsets[t] = (int)i;
		unsigned big = __ballot_sync(0xffffffffu, last - first >= kInline);
		while (big) {
			const int src = __ffs(big) - 1;
			big &= big - 1;
			const int b_first = __shfl_sync(0xffffffffu, first, src), b_last = __shfl_sync(0xffffffffu, last, src);
			const int b_i = (int)__shfl_sync(0xffffffffu, (int)i, src);
			for (int t = b_first + lane; t <= b_last; t += 32) offsets[t] = b_i;
		}
	}
}

// lists longer than `cap` carry kVisOverlong in their offset (readers mask the bit off their neighbour's word)
__global__ void __launch_bounds__(256) vis_flag_overlong_kernel(int* offsets, int n_tex, int cap) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_tex) return;
	const int beg = offsets[t] & ~kVisOverlong, end = offsets[t + 1] & ~kVisOverlong;
	if (end - beg > cap) offsets[t] = beg | kVisOverlong;
}

// Exclusive scan of the list lengths in three launches (block sums, scan of the sums, offsets).  Lengths above
// `cap` count as zero and mark their offset with kVisOverlong.  out[n] = total (negative on overflow).
constexpr int kVisScanBlocks = 1024;

// cap < 0: every list is stored (the sort-based build keeps the long ones too, flagged; they are never looked up)
__device__ __forceinline__ int vis_len(int v, int cap) { return (cap >= 0 && v > cap) ? 0 : v; }

__global__ void __launch_bounds__(1024) vis_scan_sums_kernel(const int* in, long long* sums, int n, int per, int cap) {
	__shared__ long long part[32];
	const int beg = min(n, (int)blockIdx.x * per), end = min(n, beg + per);
	long long s = 0;
	for (int i = beg + threadIdx.x; i < end; i += 1024) s += vis_len(in[i], cap);
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x < 32) {
		s = part[threadIdx.x];
		for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (threadIdx.x == 0) sums[blockIdx.x] = s;
	}
}

__global__ void __launch_bounds__(1024) vis_scan_top_kernel(long long* sums, int* out, int n) {
	__shared__ long long part[kVisScanBlocks];
	part[threadIdx.x] = sums[threadIdx.x];
	__syncthreads();
	if (threadIdx.x == 0) {
		long long run = 0;
		for (int i = 0; i < kVisScanBlocks; ++i) { const long long v = part[i]; part[i] = run; run += v; }
		out[n] = run < 0x7fffffffLL ? (int)run : -1;
	}
	__syncthreads();
	sums[threadIdx.x] = part[threadIdx.x];
}

__global__ void __launch_bounds__(1024) vis_scan_offsets_kernel(const int* in, const long long* sums, int* out, int n, int per, int cap, int flag_cap) {
	__shared__ int warp_tot[32];
	__shared__ int tile_tot;
	const int beg = min(n, (int)blockIdx.x * per), end = min(n, beg + per);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	long long run = sums[blockIdx.x];
	for (int base = beg; base < end; base += 1024) {   // block-uniform trip count
		const int i = base + threadIdx.x;
		const int raw = i < end ? in[i] : 0;
		const int v = vis_len(raw, cap);
		int inc = v;
		for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
		if (lane == 31) warp_tot[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			int wv = warp_tot[lane], winc = wv;
			for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += u; }
			warp_tot[lane] = winc - wv;
			if (lane == 31) tile_tot = winc;
		}
		__syncthreads();
		if (i < end) out[i] = (int)(run + warp_tot[wid] + inc - v) | (raw > flag_cap ? kVisOverlong : 0);
		run += tile_tot;
		__syncthreads();
	}
}

// Texel lists ordered by distance from the recorder, nearest triangle first (one warp per texel).  A query that is
// blocked stops at the first triangle its segment crosses; in a venue of several rooms the blocker of nearly every
// query from another room is a wall of the RECORDER's room, i.e. among the nearest entries of the texel -- with the
// fill order (arbitrary: atomics) a blocked query tests a third of its list on average, with this order one or two
// entries.  Visible queries test the whole list either way, so the order is only established when the render loop
// sees that most queries are blocked (launch_wavefront).  Key: squared distance from X to the triangle's centroid.
constexpr int kVisSortMax = 256;   // longest list that is stored (kVisMaxList) -- lists are sorted in shared memory
__global__ void __launch_bounds__(256) vis_sort_kernel(SceneDev sc, float x0, float x1, float x2, const int* offsets, int* items, int n_tex) {
	__shared__ float s_key[8][kVisSortMax];
	__shared__ int s_item[8][kVisSortMax];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tex; t += warps) {
		const int beg = offsets[t];
		if (beg < 0) continue;                                   // over the cap: not stored
		const int len = (offsets[t + 1] & ~kVisOverlong) - beg;
		if (len < 2 || len > kVisSortMax) continue;
		int n2 = 2;
		while (n2 < len) n2 <<= 1;
		for (int i = lane; i < n2; i += 32) {
			float key = INFINITY; int item = -1;
			if (i < len) {
				item = items[beg + i];
				const float4 r0 = __ldg(sc.tris + 4 * (size_t)item), r1 = __ldg(sc.tris + 4 * (size_t)item + 1), r2 = __ldg(sc.tris + 4 * (size_t)item + 2);
				const float cx = r0.x + (r1.x + r2.x) * (1.0f / 3.0f) - x0, cy = r0.y + (r1.y + r2.y) * (1.0f / 3.0f) - x1, cz = r0.z + (r1.z + r2.z) * (1.0f / 3.0f) - x2;
				key = cx * cx + cy * cy + cz * cz;
			}
			s_key[wid][i] = key; s_item[wid][i] = item;
		}
		__syncwarp();
		for (int k = 2; k <= n2; k <<= 1)
			for (int j = k >> 1; j > 0; j >>= 1) {
				for (int i = lane; i < n2; i += 32) {
					const int l = i ^ j;
					if (l > i) {
						const bool up = (i & k) == 0;
						const float a = s_key[wid][i], b = s_key[wid][l];
						// ties keep the lower item first so that the order is deterministic
						const bool swap = up ? (a > b || (a == b && s_item[wid][i] > s_item[wid][l])) : (a < b || (a == b && s_item[wid][i] < s_item[wid][l]));
						if (swap) {
							s_key[wid][i] = b; s_key[wid][l] = a;
							const int ti = s_item[wid][i]; s_item[wid][i] = s_item[wid][l]; s_item[wid][l] = ti;
						}
					}
				}
				__syncwarp();
			}
		for (int i = lane; i < len; i += 32) items[beg + i] = s_item[wid][i];
		__syncwarp();
	}
}

// K4 through the maps: one query per thread.  Visible queries go to vis_list; queries whose texel list is too
// long (or whose recorder has no map) go to `q_bvh` for the BVH any-hit kernel.
#ifndef EARB_VISMAP_MIN_BLOCKS
#define EARB_VISMAP_MIN_BLOCKS 5   // resident 256-thread blocks per SM asked of the compiler (5 <-> 48 registers)
#endif
__global__ void __launch_bounds__(256, EARB_VISMAP_MIN_BLOCKS) wf_vismap_kernel(SceneDev sc, WfPool pool, RenderParams p, const VisMapDev* maps,
                                                        const int* map_of, uint2* q_bvh) {
	const int total = pool.counts[1];
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	const int stride = gridDim.x * blockDim.x;
	for (int base_i = blockIdx.x * blockDim.x; base_i < total; base_i += stride) {   // warp-uniform trip count
		const int i = base_i + threadIdx.x;
		bool visible = false, fallback = false;
		uint2 q = make_uint2(0u, 0u);
		if (i < total) {
			q = ld_stream(pool.q_list + i);
			const uint32_t slot = q.x & ((1u << pool.slot_bits) - 1u), r = q.x >> pool.slot_bits, c = q.y & 0xffffu;
			const int mi = map_of[c * p.n_rec + r];
			if (mi < 0) fallback = true;
			else {
				const VisMapDev mp = maps[mi];
				const float4 s0 = ld_stream(pool.sh0 + slot);
				const V3 pnt = mk(s0.x, s0.y, s0.z), x = mk(mp.x[0], mp.x[1], mp.x[2]);
				const int texel = vis_texel(mp, pnt.x - x.x, pnt.y - x.y, pnt.z - x.z);
				const int raw = mp.offsets[texel];
				const bool overlong = raw < 0;   // list longer than the cap
				const int beg = raw & ~kVisOverlong;
				int end = mp.offsets[texel + 1] & ~kVisOverlong;
				// An over-long list is never walked to its end (that is what the BVH is for), but "blocked" needs only ONE
				// crossing triangle, whichever list it came from: the sort-based build stores these lists too, nearest first,
				// and in a venue of several rooms the blocker of most queries is a wall next to the recorder -- so the first
				// vis_prefix entries settle most of them; the rest go to the any-hit kernel as before.  (The atomic build
				// does not store over-long lists: beg == end there.)
				if (overlong) end = min(end, beg + sc.vis_prefix);
				{
					const V3 d = vsub(x, pnt);   // LineSeg(p, x) = Ray(p, x - p)
					visible = true;
					// The loop is a chain of dependent loads (item -> triangle record -> test; ncu: long-scoreboard stalls
					// 15 per issue).  Two candidates per round, their loads in flight before the first test (measured: 1 -> 123.4,
					// 2 -> 112.5, 4 -> 127.1 ms per 2e7 rays).  The
					// tail of a list repeats its last entry (testing a triangle twice cannot change a yes/no answer).
					for (int k = beg; k < end && visible; k += kVisBatch) {
						int item[kVisBatch];
#pragma unroll
						for (int j = 0; j < kVisBatch; ++j) item[j] = __ldg(mp.items + min(k + j, end - 1));
						F8 r01[kVisBatch];
						float4 r2[kVisBatch];
#pragma unroll
						for (int j = 0; j < kVisBatch; ++j) {
							const float4* rec = sc.tris + 4 * (size_t)item[j];
							r01[j] = ldg256(rec);
							r2[j] = ldg_keep(rec + 2);
						}
#pragma unroll
						for (int j = 0; j < kVisBatch; ++j) {
							float t;
							if (moeller_trumbore(mk(r01[j].lo.x, r01[j].lo.y, r01[j].lo.z), mk(r01[j].hi.x, r01[j].hi.y, r01[j].hi.z),
							                     mk(r2[j].x, r2[j].y, r2[j].z), pnt, d, t) &&
							    t > 1e-5f && t < 1.0f) visible = false;
						}
					}
					if (overlong && visible) { visible = false; fallback = true; }   // not blocked by the prefix: the BVH decides
				}
			}
		}
		const unsigned m_vis = __ballot_sync(0xffffffffu, visible), m_fb = __ballot_sync(0xffffffffu, fallback);
		if (m_vis) {
			int b = 0;
			if (lane == 0) b = atomicAdd(pool.counts + 2, __popc(m_vis));
			b = __shfl_sync(0xffffffffu, b, 0);
			if (visible) st_stream(pool.vis_list + b + __popc(m_vis & lt_mask), q);
		}
		if (m_fb) {
			int b = 0;
			if (lane == 0) b = atomicAdd(pool.counts + 5, __popc(m_fb));
			b = __shfl_sync(0xffffffffu, b, 0);
			if (fallback) st_stream(q_bvh + b + __popc(m_fb & lt_mask), q);
		}
	}
}

}  // namespace earb
