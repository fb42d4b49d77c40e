// K0 on the device: binned-SAH BVH build over the Mesh/Triangle soup, collapse to 4-wide quantised nodes and the
// triangle records, all in HBM -- the host only sequences launches and reads one counter per level.
//
// Same structure as the host builder (bvh_build.cpp): per-triangle padded boxes + interval slack, binary SAH tree with
// 16 bins per axis (three axes per pass, strict '<' over axis-then-bin order, leaf-vs-split test for <= 4 triangles),
// collapse by "expand the inner child with the largest area until there are four", child boxes quantised OUTWARD
// to 8 bits per plane.  The tree is not byte-identical to the host's (float reductions run in another order, equal-cost
// ties may fall differently), which is immaterial: the traversal returns the reference's winner for ANY tree whose
// boxes contain the padded triangle boxes.  What differs is how the work is laid out:
//
//   top phase     level-synchronous over all nodes with more than kSmallNode triangles: one pass over the primitive
//                 array per step (bounds, bins, flags, scatter), every block accumulating into shared memory for the
//                 at most two nodes its 256 primitives can belong to and issuing one set of global atomics;
//                 a warp per node evaluates the 45 split candidates; the partition is a stable scatter driven by
//                 ONE exclusive scan of the "goes left" flags of the whole array, ping-ponging two primitive arrays.
//   bottom phase  every subtree of <= kSmallNode triangles is built by one block entirely in shared memory
//                 (warp per node: shuffle reductions, shared-memory bins, ballot-driven stable partition).
//   collapse      breadth-first: a thread per wide node expands its binary node, a scan numbers the inner children,
//                 a second kernel quantises and writes the 64-byte node.  Siblings are adjacent.
//   records       triangle records in leaf order, normals in the reference's float order (src/Triangle.cpp:41).
//
// Order of primitives and shape of the tree are deterministic (scans, not atomics, decide positions; atomics only
// allocate scratch node ids, which the breadth-first renumbering forgets).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "bvh_build.h"
#include "device_exact.cuh"

// device allocations go through the library's cache (ear_b200.cu)
template <class T> static cudaError_t dev_alloc(T** out, size_t bytes);
static void dev_free(void* p);

namespace earb {
namespace dbvh {

constexpr int kBins = 16;
constexpr int kSmallNode = 512;      // subtrees up to this size are built by one block in shared memory
constexpr int kTopBlock = 256;       // < kSmallNode / 2 + 1: a block's primitives belong to at most two large nodes
constexpr int kSahDepth = 30;        // past this many top levels: split by position (bounded depth for adversarial input)
constexpr int kBinWords = 3 * kBins * 7;
constexpr int kBoundWords = 13;      // lo[3] hi[3] clo[3] chi[3] slack

// order-preserving float <-> int (atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
#define DBVH_INF_LO 0x7f800000           /* f2o(+inf) */
#define DBVH_INF_HI ((int)0x807fffff)    /* f2o(-inf) = 0xff800000 ^ 0x7fffffff */

// scratch binary node
struct BNode {
	float lo[3]; int32_t left;     // left < 0: leaf
	float hi[3]; int32_t right;
	int32_t first, count;
	float slack; int32_t pad;
};
static_assert(sizeof(BNode) == 48, "BNode");

struct Globals {          // device-resident scalars
	int lo[3], hi[3];     // scene bounds, ordered ints
	float s0, diagonal, maxabs, reach;
	int node_count;       // scratch binary nodes allocated
	int next_active;      // large nodes of the next level
	int small_count;      // subtrees handed to the bottom phase
	int wide_next;        // collapse: inner children of the current level
	int depth;
};

struct Level {            // per active (large) node of the current level
	int* node;            // [cap] BNode id
	int* acc;             // [cap][kBoundWords] ordered-int bounds accumulators
	int* bins;            // [cap][kBinWords]  per (axis, bin): lo[3] hi[3] (ordered ints) count
	int* axis;            // [cap] split axis, -1: split by position
	int* bin;             // [cap] last bin that goes left
	float* cmin;          // [cap][3]
	float* scale;         // [cap][3]
	int* n_left;          // [cap]
	int* child_seg;       // [cap][2] index of the child among the next level's large nodes, or -1
};

// ---------------------------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------------------------
__global__ void init_globals_kernel(Globals* g) {
	for (int k = 0; k < 3; ++k) { g->lo[k] = DBVH_INF_LO; g->hi[k] = DBVH_INF_HI; }
	g->node_count = 0; g->next_active = 0; g->small_count = 0; g->wide_next = 0; g->depth = 0;
}

__global__ void __launch_bounds__(256) scene_bounds_kernel(const float* __restrict__ verts, int n, Globals* g) {
	__shared__ int s_lo[3], s_hi[3];
	if (threadIdx.x < 3) { s_lo[threadIdx.x] = DBVH_INF_LO; s_hi[threadIdx.x] = DBVH_INF_HI; }
	__syncthreads();
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * 3; i += (long long)gridDim.x * blockDim.x) {
		const float* p = verts + 3 * i;
		for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], p[a]); hi[a] = fmaxf(hi[a], p[a]); }
	}
	for (int a = 0; a < 3; ++a) {
		for (int o = 16; o; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
		if ((threadIdx.x & 31) == 0) { atomicMin(&s_lo[a], f2o(lo[a])); atomicMax(&s_hi[a], f2o(hi[a])); }
	}
	__syncthreads();
	if (threadIdx.x < 3) { atomicMin(&g->lo[threadIdx.x], s_lo[threadIdx.x]); atomicMax(&g->hi[threadIdx.x], s_hi[threadIdx.x]); }
}

__global__ void scene_constants_kernel(Globals* g, int n, float margin_scale) {
	float diag2 = 0.0f, maxabs = 0.0f;
	for (int a = 0; a < 3; ++a) {
		float lo = o2f(g->lo[a]), hi = o2f(g->hi[a]);
		if (n == 0) { lo = 0.0f; hi = 0.0f; g->lo[a] = f2o(0.0f); g->hi[a] = f2o(0.0f); }
		const float d = hi - lo;
		diag2 += d * d;
		maxabs = fmaxf(maxabs, fmaxf(fabsf(lo), fabsf(hi)));
	}
	const float eps = 5.9604645e-8f;
	g->diagonal = sqrtf(diag2);
	g->maxabs = maxabs;
	g->reach = 2.0f * g->diagonal + 1.0f;
	g->s0 = margin_scale * 64.0f * eps * (maxabs + g->reach);
}

// per-triangle padded box and interval slack (bvh_build.cpp, "exactness contract")
__global__ void __launch_bounds__(256) prims_kernel(const float* __restrict__ verts, int n, const Globals* g, float margin_scale,
                                                    float4* plo, float4* phi) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* p = verts + 9 * (size_t)i;
	const float eps = 5.9604645e-8f;
	const float reach = g->reach, maxabs = g->maxabs;
	float e1[3], e2[3], cr[3];
	for (int k = 0; k < 3; ++k) { e1[k] = p[3 + k] - p[k]; e2[k] = p[6 + k] - p[k]; }
	cr[0] = e1[1] * e2[2] - e1[2] * e2[1]; cr[1] = e1[2] * e2[0] - e1[0] * e2[2]; cr[2] = e1[0] * e2[1] - e1[1] * e2[0];
	const float l1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
	const float l2 = sqrtf(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
	const float cl = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
	float sinphi = (l1 > 0 && l2 > 0) ? cl / (l1 * l2) : 1.0f;
	sinphi = fmaxf(sinphi, 1e-3f);
	const float pad = margin_scale * (16.0f * eps * reach / sinphi + 8.0f * eps * (maxabs + reach));
	float lo[3], hi[3];
	for (int k = 0; k < 3; ++k) {
		lo[k] = fminf(p[k], fminf(p[3 + k], p[6 + k])) - pad;
		hi[k] = fmaxf(p[k], fmaxf(p[3 + k], p[6 + k])) + pad;
	}
	const float slack = fminf(margin_scale * 16.0f * eps * reach * l1 * l2 / 1e-5f + pad, 4.0f * reach);
	plo[i] = make_float4(lo[0], lo[1], lo[2], __int_as_float(i));
	phi[i] = make_float4(hi[0], hi[1], hi[2], slack);
}

__device__ __forceinline__ float half_area(const float lo[3], const float hi[3]) {
	const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
	return dx * dy + dy * dz + dz * dx;
}

// ---------------------------------------------------------------------------------------------------------------
// top phase
// ---------------------------------------------------------------------------------------------------------------
__global__ void fill_int_kernel(int* dst, long long n, int v) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void level_init_kernel(Level lv, int n_active) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_active * kBoundWords) {
		const int w = i % kBoundWords;
		lv.acc[i] = (w < 3 || (w >= 6 && w < 9)) ? DBVH_INF_LO : (w == 12 ? 0 : DBVH_INF_HI);
	}
	if (i < n_active * kBinWords) {
		const int w = i % 7;
		lv.bins[i] = w < 3 ? DBVH_INF_LO : (w < 6 ? DBVH_INF_HI : 0);
	}
}

// the (at most two) large nodes the primitives of this block belong to
__device__ __forceinline__ void block_segments(int seg, int* s_seg, int* s_pos) {
	if (threadIdx.x == 0) { s_pos[0] = 0x7fffffff; s_pos[1] = -1; s_seg[0] = -1; s_seg[1] = -1; }
	__syncthreads();
	if (seg >= 0) { atomicMin(&s_pos[0], (int)threadIdx.x); atomicMax(&s_pos[1], (int)threadIdx.x); }
	__syncthreads();
	if (seg >= 0 && (int)threadIdx.x == s_pos[0]) s_seg[0] = seg;
	if (seg >= 0 && (int)threadIdx.x == s_pos[1]) s_seg[1] = seg;
	__syncthreads();
}

__global__ void __launch_bounds__(kTopBlock) top_bounds_kernel(const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                               const int* __restrict__ seg_of, int n, Level lv) {
	__shared__ int s_seg[2], s_pos[2];
	__shared__ int s_acc[2][kBoundWords];
	const int i = blockIdx.x * kTopBlock + threadIdx.x;
	const int seg = i < n ? seg_of[i] : -1;
	block_segments(seg, s_seg, s_pos);
	if (s_seg[0] < 0) return;   // no large node in this block (block-uniform)
	if (threadIdx.x < 2 * kBoundWords) {
		const int w = threadIdx.x % kBoundWords;
		s_acc[threadIdx.x / kBoundWords][w] = (w < 3 || (w >= 6 && w < 9)) ? DBVH_INF_LO : (w == 12 ? 0 : DBVH_INF_HI);
	}
	__syncthreads();
	if (seg >= 0) {
		const int slot = seg == s_seg[0] ? 0 : 1;
		const float4 lo = plo[i], hi = phi[i];
		const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
		const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
		for (int a = 0; a < 3; ++a) {
			atomicMin(&s_acc[slot][a], f2o(l[a])); atomicMax(&s_acc[slot][3 + a], f2o(h[a]));
			atomicMin(&s_acc[slot][6 + a], f2o(c[a])); atomicMax(&s_acc[slot][9 + a], f2o(c[a]));
		}
		atomicMax(&s_acc[slot][12], __float_as_int(hi.w));   // slack >= 0: plain int order
	}
	__syncthreads();
	if (threadIdx.x < 2 * kBoundWords) {
		const int slot = threadIdx.x / kBoundWords, w = threadIdx.x % kBoundWords;
		const int sg = s_seg[slot];
		if (sg >= 0 && !(slot == 1 && s_seg[1] == s_seg[0])) {
			int* dst = lv.acc + (size_t)sg * kBoundWords + w;
			if (w < 3 || (w >= 6 && w < 9)) atomicMin(dst, s_acc[slot][w]); else atomicMax(dst, s_acc[slot][w]);
		}
	}
}

// bin ranges of every active node from its centroid bounds (one thread per node)
__global__ void level_ranges_kernel(Level lv, int n_active) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_active) return;
	const int* acc = lv.acc + (size_t)k * kBoundWords;
	for (int a = 0; a < 3; ++a) {
		const float clo = o2f(acc[6 + a]), chi = o2f(acc[9 + a]);
		lv.cmin[3 * k + a] = clo;
		lv.scale[3 * k + a] = chi > clo ? (float)kBins / (chi - clo) : 0.0f;
	}
}

__device__ __forceinline__ int bin_of(float lo, float hi, float cmin, float scale) {
	return min(max((int)((0.5f * (lo + hi) - cmin) * scale), 0), kBins - 1);
}

__global__ void __launch_bounds__(kTopBlock) top_bins_kernel(const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                             const int* __restrict__ seg_of, int n, Level lv) {
	__shared__ int s_seg[2], s_pos[2];
	__shared__ int s_bins[2][kBinWords];
	const int i = blockIdx.x * kTopBlock + threadIdx.x;
	const int seg = i < n ? seg_of[i] : -1;
	block_segments(seg, s_seg, s_pos);
	if (s_seg[0] < 0) return;
	for (int j = threadIdx.x; j < 2 * kBinWords; j += kTopBlock) {
		const int w = j % 7;
		s_bins[j / kBinWords][j % kBinWords] = w < 3 ? DBVH_INF_LO : (w < 6 ? DBVH_INF_HI : 0);
	}
	__syncthreads();
	if (seg >= 0) {
		const int slot = seg == s_seg[0] ? 0 : 1;
		const float4 lo = plo[i], hi = phi[i];
		const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
		const int ol[3] = {f2o(lo.x), f2o(lo.y), f2o(lo.z)}, oh[3] = {f2o(hi.x), f2o(hi.y), f2o(hi.z)};
		for (int a = 0; a < 3; ++a) {
			const int b = bin_of(l[a], h[a], lv.cmin[3 * seg + a], lv.scale[3 * seg + a]);
			int* bw = &s_bins[slot][(a * kBins + b) * 7];
			atomicMin(bw + 0, ol[0]); atomicMin(bw + 1, ol[1]); atomicMin(bw + 2, ol[2]);
			atomicMax(bw + 3, oh[0]); atomicMax(bw + 4, oh[1]); atomicMax(bw + 5, oh[2]);
			atomicAdd(bw + 6, 1);
		}
	}
	__syncthreads();
	for (int j = threadIdx.x; j < 2 * kBinWords; j += kTopBlock) {
		const int slot = j / kBinWords, w = j % kBinWords;
		const int sg = s_seg[slot];
		if (sg < 0 || (slot == 1 && s_seg[1] == s_seg[0])) continue;
		if (s_bins[slot][w - w % 7 + 6] == 0) continue;   // empty bin
		int* dst = lv.bins + (size_t)sg * kBinWords + w;
		const int v = s_bins[slot][w];
		if (w % 7 < 3) atomicMin(dst, v); else if (w % 7 < 6) atomicMax(dst, v); else atomicAdd(dst, v);
	}
}

// SAH over the 3 x 15 candidate planes of one node's bins; lanes take candidates.  The winner is the candidate with
// the least cost, first in (axis, bin) order among equals -- the order of the host builder's sweep.
// bins: ordered-int words [axis][bin][7].  Returns cost (INFINITY: no usable candidate); axis / bin by reference.
__device__ __forceinline__ float warp_best_split(const int* bins, const float* scale, int& best_axis, int& best_bin, int& n_left) {
	const int lane = threadIdx.x & 31;
	float best = INFINITY; int code = 0x7fffffff, nl_best = 0;
	for (int cand = lane; cand < 3 * (kBins - 1); cand += 32) {
		const int axis = cand / (kBins - 1), b = cand % (kBins - 1);
		if (!(scale[axis] > 0.0f)) continue;
		float llo[3] = {INFINITY, INFINITY, INFINITY}, lhi[3] = {-INFINITY, -INFINITY, -INFINITY};
		float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
		int cl = 0, cr = 0;
		for (int k = 0; k < kBins; ++k) {
			const int* bw = bins + (axis * kBins + k) * 7;
			const int c = bw[6];
			if (c == 0) continue;
			if (k <= b) { cl += c; for (int a = 0; a < 3; ++a) { llo[a] = fminf(llo[a], o2f(bw[a])); lhi[a] = fmaxf(lhi[a], o2f(bw[3 + a])); } }
			else { cr += c; for (int a = 0; a < 3; ++a) { rlo[a] = fminf(rlo[a], o2f(bw[a])); rhi[a] = fmaxf(rhi[a], o2f(bw[3 + a])); } }
		}
		if (cl == 0 || cr == 0) continue;
		const float cost = half_area(llo, lhi) * (float)cl + half_area(rlo, rhi) * (float)cr;
		if (cost < best || (cost == best && cand < code)) { best = cost; code = cand; nl_best = cl; }
	}
	for (int o = 16; o; o >>= 1) {
		const float ob = __shfl_xor_sync(0xffffffffu, best, o);
		const int oc = __shfl_xor_sync(0xffffffffu, code, o), on = __shfl_xor_sync(0xffffffffu, nl_best, o);
		if (ob < best || (ob == best && oc < code)) { best = ob; code = oc; nl_best = on; }
	}
	best_axis = code == 0x7fffffff ? -1 : code / (kBins - 1);
	best_bin = code == 0x7fffffff ? -1 : code % (kBins - 1);
	n_left = nl_best;
	return best;
}

// one warp per active node: choose the split, create the two children, route them to the next level or to the
// bottom phase
__global__ void __launch_bounds__(128) top_split_kernel(Level lv, int n_active, BNode* nodes, Globals* g, int level, int* next_node,
                                                        int* small_roots) {
	const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (k >= n_active) return;
	const int id = lv.node[k];
	const int* acc = lv.acc + (size_t)k * kBoundWords;
	const int first = nodes[id].first, count = nodes[id].count;
	int axis, bin, nl;
	const float scale[3] = {lv.scale[3 * k], lv.scale[3 * k + 1], lv.scale[3 * k + 2]};
	warp_best_split(lv.bins + (size_t)k * kBinWords, scale, axis, bin, nl);
	if (level >= kSahDepth || axis < 0 || nl == 0 || nl == count) { axis = -1; nl = count / 2; }   // split by position
	if (lane == 0) {
		BNode& nd = nodes[id];
		for (int a = 0; a < 3; ++a) { nd.lo[a] = o2f(acc[a]); nd.hi[a] = o2f(acc[3 + a]); }
		nd.slack = __int_as_float(acc[12]);
		const int l = atomicAdd(&g->node_count, 2), r = l + 1;
		nd.left = l; nd.right = r;
		nodes[l].first = first; nodes[l].count = nl; nodes[l].left = -1; nodes[l].right = -1;
		nodes[r].first = first + nl; nodes[r].count = count - nl; nodes[r].left = -1; nodes[r].right = -1;
		lv.axis[k] = axis; lv.bin[k] = bin; lv.n_left[k] = nl;
		for (int c = 0; c < 2; ++c) {
			const int cid = c ? r : l, cc = c ? count - nl : nl;
			int sg = -1;
			if (cc > kSmallNode) { sg = atomicAdd(&g->next_active, 1); next_node[sg] = cid; }
			else small_roots[atomicAdd(&g->small_count, 1)] = cid;
			lv.child_seg[2 * k + c] = sg;
		}
	}
}

// "goes left" flag of every primitive of a large node (0 elsewhere); scanned over the whole array
__global__ void __launch_bounds__(256) top_flags_kernel(const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                        const int* __restrict__ seg_of, int n, Level lv, const BNode* nodes, int* flags) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int seg = seg_of[i];
	int f = 0;
	if (seg >= 0) {
		const int axis = lv.axis[seg];
		if (axis < 0) f = (i - nodes[lv.node[seg]].first) < lv.n_left[seg] ? 1 : 0;
		else {
			const float4 lo = plo[i], hi = phi[i];
			const float l = axis == 0 ? lo.x : axis == 1 ? lo.y : lo.z, h = axis == 0 ? hi.x : axis == 1 ? hi.y : hi.z;
			f = bin_of(l, h, lv.cmin[3 * seg + axis], lv.scale[3 * seg + axis]) <= lv.bin[seg] ? 1 : 0;
		}
	}
	flags[i] = f;
}

// exclusive scan of an int array in three launches (per-block sums, scan of the sums, offsets)
constexpr int kScanBlock = 1024;
constexpr int kScanPer = 8;     // elements per thread
__device__ __forceinline__ int block_scan_excl(int v, int* warp_tot, int& total) {
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int inc = v;
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
	const int r = inc - v + (wid ? warp_tot[wid - 1] : 0);
	__syncthreads();
	return r;
}
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(const int* in, int n, int* sums) {
	__shared__ int warp_tot[32];
	const int base = blockIdx.x * kScanBlock * kScanPer + threadIdx.x * kScanPer;
	int s = 0;
	for (int j = 0; j < kScanPer; ++j) if (base + j < n) s += in[base + j];
	int total;
	block_scan_excl(s, warp_tot, total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanBlock) scan_top_kernel(int* sums, int n_blocks) {
	__shared__ int warp_tot[32];
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n_blocks; base += kScanBlock) {
		const int i = base + threadIdx.x;
		const int v = i < n_blocks ? sums[i] : 0;
		int total;
		const int ex = block_scan_excl(v, warp_tot, total);
		if (i < n_blocks) sums[i] = carry + ex;
		__syncthreads();
		if (threadIdx.x == 0) carry += total;
		__syncthreads();
	}
}
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const int* in, int n, const int* sums, int* out) {
	__shared__ int warp_tot[32];
	const int base = blockIdx.x * kScanBlock * kScanPer + threadIdx.x * kScanPer;
	int v[kScanPer], s = 0;
	for (int j = 0; j < kScanPer; ++j) { v[j] = base + j < n ? in[base + j] : 0; s += v[j]; }
	int total;
	int run = sums[blockIdx.x] + block_scan_excl(s, warp_tot, total);
	for (int j = 0; j < kScanPer; ++j) { if (base + j < n) out[base + j] = run; run += v[j]; }
}

// stable scatter of the primitives of every large node to its children's ranges; everything else is copied through
__global__ void __launch_bounds__(256) top_scatter_kernel(const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                          const int* __restrict__ seg_of, const int* __restrict__ flags,
                                                          const int* __restrict__ scan, int n, Level lv, const BNode* nodes,
                                                          float4* plo_out, float4* phi_out, int* seg_out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int seg = seg_of[i];
	int dst = i, nseg = -1;
	if (seg >= 0) {
		const int first = nodes[lv.node[seg]].first;
		const int rank_left = scan[i] - scan[first];          // lefts of this node before i
		if (flags[i]) { dst = first + rank_left; nseg = lv.child_seg[2 * seg]; }
		else { dst = first + lv.n_left[seg] + (i - first - rank_left); nseg = lv.child_seg[2 * seg + 1]; }
	}
	plo_out[dst] = plo[i]; phi_out[dst] = phi[i]; seg_out[dst] = nseg;
}

// ---------------------------------------------------------------------------------------------------------------
// bottom phase: one block per subtree of <= kSmallNode primitives, entirely in shared memory
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBotWarps = 4;
// work item of the bottom phase: scratch node id; first (10 bits) | count (10 bits, <= kSmallNode) | depth (12 bits)
struct BotItem { int node; unsigned packed; };
__device__ __forceinline__ BotItem bot_item(int node, int first, int count, int depth) {
	BotItem it; it.node = node; it.packed = (unsigned)first | ((unsigned)count << 10) | ((unsigned)min(depth, 4095) << 20); return it;
}
static_assert(kSmallNode <= 1023, "BotItem packs first and count into 10 bits each");

__global__ void __launch_bounds__(32 * kBotWarps) bottom_kernel(const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                                float4* plo_out, float4* phi_out, BNode* nodes, Globals* g,
                                                                const int* __restrict__ small_roots, int max_leaf, float sah_ct) {
	__shared__ float4 s_lo[2][kSmallNode], s_hi[2][kSmallNode];
	__shared__ BotItem s_list[2][kSmallNode + 1];   // a level holds at most one item per primitive
	__shared__ int s_n[2];
	__shared__ int s_bins[kBotWarps][kBinWords];
	__shared__ int s_node_base;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	const int root = small_roots[blockIdx.x];
	const int g_first = nodes[root].first, total = nodes[root].count;
	for (int i = threadIdx.x; i < total; i += blockDim.x) { s_lo[0][i] = plo[g_first + i]; s_hi[0][i] = phi[g_first + i]; }
	if (threadIdx.x == 0) {
		s_list[0][0] = bot_item(root, 0, total, 0); s_n[0] = 1; s_n[1] = 0;
		s_node_base = atomicAdd(&g->node_count, 2 * total);      // a subtree of c primitives has at most 2c - 1 nodes
	}
	__syncthreads();
	int cur = 0;   // index of the current list; the primitives of every listed node are in buffer `cur`
	__shared__ int s_alloc;
	if (threadIdx.x == 0) s_alloc = 0;
	__syncthreads();
	while (s_n[cur] > 0) {
		const int n_items = s_n[cur];
		for (int it = wid; it < n_items; it += kBotWarps) {
			const BotItem packed_item = s_list[cur][it];
			struct { int node, first, count, depth; } item = {packed_item.node, (int)(packed_item.packed & 1023u),
			                                                  (int)((packed_item.packed >> 10) & 1023u), (int)(packed_item.packed >> 20)};
			const float4* L = s_lo[cur] + item.first;
			const float4* H = s_hi[cur] + item.first;
			// ---- bounds ----
			float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
			float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
			float slack = 0.0f;
			for (int i = lane; i < item.count; i += 32) {
				const float4 l = L[i], h = H[i];
				const float pl[3] = {l.x, l.y, l.z}, ph[3] = {h.x, h.y, h.z};
				for (int a = 0; a < 3; ++a) {
					const float c = 0.5f * (pl[a] + ph[a]);
					lo[a] = fminf(lo[a], pl[a]); hi[a] = fmaxf(hi[a], ph[a]);
					clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
				}
				slack = fmaxf(slack, h.w);
			}
			for (int o = 16; o; o >>= 1) {
				for (int a = 0; a < 3; ++a) {
					lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
					clo[a] = fminf(clo[a], __shfl_xor_sync(0xffffffffu, clo[a], o)); chi[a] = fmaxf(chi[a], __shfl_xor_sync(0xffffffffu, chi[a], o));
				}
				slack = fmaxf(slack, __shfl_xor_sync(0xffffffffu, slack, o));
			}
			// ---- split decision ----
			int axis = -1, bin = -1, nl = 0;
			bool leaf = item.count <= 1;
			float scale[3] = {0.0f, 0.0f, 0.0f};
			if (!leaf) {
				for (int a = 0; a < 3; ++a) scale[a] = chi[a] > clo[a] ? (float)kBins / (chi[a] - clo[a]) : 0.0f;
				int* bins = s_bins[wid];
				for (int j = lane; j < kBinWords; j += 32) { const int w = j % 7; bins[j] = w < 3 ? DBVH_INF_LO : (w < 6 ? DBVH_INF_HI : 0); }
				__syncwarp();
				for (int i = lane; i < item.count; i += 32) {
					const float4 l = L[i], h = H[i];
					const float pl[3] = {l.x, l.y, l.z}, ph[3] = {h.x, h.y, h.z};
					const int ol[3] = {f2o(l.x), f2o(l.y), f2o(l.z)}, oh[3] = {f2o(h.x), f2o(h.y), f2o(h.z)};
					for (int a = 0; a < 3; ++a) {
						int* bw = bins + (a * kBins + bin_of(pl[a], ph[a], clo[a], scale[a])) * 7;
						atomicMin(bw + 0, ol[0]); atomicMin(bw + 1, ol[1]); atomicMin(bw + 2, ol[2]);
						atomicMax(bw + 3, oh[0]); atomicMax(bw + 4, oh[1]); atomicMax(bw + 5, oh[2]);
						atomicAdd(bw + 6, 1);
					}
				}
				__syncwarp();
				const float best = warp_best_split(bins, scale, axis, bin, nl);
				__syncwarp();
				if (item.count <= max_leaf) {
					const float split_cost = axis < 0 ? INFINITY : sah_ct + best / fmaxf(half_area(lo, hi), 1e-30f);
					if (!(split_cost < (float)item.count)) leaf = true;
				}
				if (!leaf && (axis < 0 || nl == 0 || nl == item.count || item.depth >= 2 * kSahDepth)) { axis = -1; nl = item.count / 2; }
			}
			// ---- write the node; leaves send their primitives to the output array ----
			int l_id = -1, r_id = -1;
			if (!leaf && lane == 0) { l_id = s_node_base + atomicAdd(&s_alloc, 2); r_id = l_id + 1; }
			l_id = __shfl_sync(0xffffffffu, l_id, 0); r_id = __shfl_sync(0xffffffffu, r_id, 0);
			if (lane == 0) {
				BNode& nd = nodes[item.node];
				for (int a = 0; a < 3; ++a) { nd.lo[a] = lo[a]; nd.hi[a] = hi[a]; }
				nd.slack = slack;
				nd.first = g_first + item.first; nd.count = item.count;
				nd.left = leaf ? -1 : l_id; nd.right = leaf ? -1 : r_id;
			}
			if (leaf) {
				for (int i = lane; i < item.count; i += 32) { plo_out[g_first + item.first + i] = L[i]; phi_out[g_first + item.first + i] = H[i]; }
				continue;
			}
			// ---- stable partition into the other buffer ----
			float4* OL = s_lo[cur ^ 1] + item.first;
			float4* OH = s_hi[cur ^ 1] + item.first;
			int done_l = 0, done_r = 0;
			const float cm = axis == 0 ? clo[0] : axis == 1 ? clo[1] : clo[2];
			const float sc = axis == 0 ? scale[0] : axis == 1 ? scale[1] : scale[2];
			for (int base = 0; base < item.count; base += 32) {
				const int i = base + lane;
				const bool on = i < item.count;
				float4 l = make_float4(0, 0, 0, 0), h = make_float4(0, 0, 0, 0);
				bool left = false;
				if (on) {
					l = L[i]; h = H[i];
					if (axis < 0) left = i < nl;
					else {
						const float a_lo = axis == 0 ? l.x : axis == 1 ? l.y : l.z, a_hi = axis == 0 ? h.x : axis == 1 ? h.y : h.z;
						left = bin_of(a_lo, a_hi, cm, sc) <= bin;
					}
				}
				const unsigned m_on = __ballot_sync(0xffffffffu, on), m_l = __ballot_sync(0xffffffffu, on && left);
				const unsigned m_r = m_on & ~m_l;
				if (on) {
					const int dst = left ? done_l + __popc(m_l & lt_mask) : nl + done_r + __popc(m_r & lt_mask);
					OL[dst] = l; OH[dst] = h;
				}
				done_l += __popc(m_l); done_r += __popc(m_r);
			}
			if (lane == 0) {
				const int at = atomicAdd(&s_n[cur ^ 1], 2);
				s_list[cur ^ 1][at] = bot_item(l_id, item.first, nl, item.depth + 1);
				s_list[cur ^ 1][at + 1] = bot_item(r_id, item.first + nl, item.count - nl, item.depth + 1);
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) s_n[cur] = 0;
		cur ^= 1;
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// collapse to 4-wide nodes, breadth-first
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int expand4(const BNode* nodes, int id, int kid[4]) {
	kid[0] = nodes[id].left; kid[1] = nodes[id].right; kid[2] = kid[3] = -1;
	int n_kids = 2;
	while (n_kids < 4) {
		int best = -1; float best_area = -1.0f;
		for (int k = 0; k < n_kids; ++k) {
			const BNode& c = nodes[kid[k]];
			if (c.left < 0) continue;
			const float area = half_area(c.lo, c.hi);
			if (area > best_area) { best_area = area; best = k; }
		}
		if (best < 0) break;
		const int b = kid[best];
		kid[best] = nodes[b].left;
		kid[n_kids++] = nodes[b].right;
	}
	return n_kids;
}
__global__ void __launch_bounds__(256) collapse_count_kernel(const BNode* nodes, const int* queue, int n_queue, int4* kids_out, int* inner) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_queue) return;
	int kid[4];
	const int n_kids = expand4(nodes, queue[q], kid);
	int m = 0;
	for (int k = 0; k < n_kids; ++k) if (nodes[kid[k]].left >= 0) ++m;
	kids_out[q] = make_int4(kid[0], kid[1], kid[2], kid[3]);
	inner[q] = m;
}
__device__ __forceinline__ unsigned short half_up_dev(float f) {   // smallest fp16 >= f (f >= 0)
	if (!(f > 0.0f)) return 0;
	if (f >= 65504.0f) return 0x7bff;
	const unsigned bits = __float_as_uint(f);
	const int e = (int)((bits >> 23) & 0xff) - 127;
	if (e < -14) { const unsigned q = (unsigned)ceilf(f * 16777216.0f); return (unsigned short)min(q, 0x400u); }
	const unsigned man = bits & 0x7fffffu;
	unsigned h = (unsigned)((e + 15) << 10) | (man >> 13);
	if (man & 0x1fffu) ++h;
	return (unsigned short)min(h, 0x7bffu);
}
__global__ void __launch_bounds__(256) collapse_emit_kernel(const BNode* nodes, const int* queue, int n_queue, const int4* kids_in,
                                                            const int* inner_scan, int out_base, int next_base, Node* out, int* next_queue) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_queue) return;
	const BNode& bn = nodes[queue[q]];
	const int4 k4 = kids_in[q];
	const int kid[4] = {k4.x, k4.y, k4.z, k4.w};
	Node nd;
	double step[3];
	for (int a = 0; a < 3; ++a) {
		nd.lo[a] = bn.lo[a];
		const double ext = (double)bn.hi[a] - (double)bn.lo[a];
		int e = -126;
		if (ext > 0.0) {
			int ex2;
			const double m = frexp(ext / 255.0, &ex2);
			e = m == 0.5 ? ex2 - 1 : ex2;
		}
		e = max(-126, min(127, e));
		while (e < 127 && ldexp(255.0, e) < ext) ++e;
		nd.ex[a] = (uint8_t)(e + 127);
		step[a] = ldexp(1.0, e);
	}
	nd.pad = 0;
	int m = 0;
	const int at = inner_scan[q];
	for (int k = 0; k < 4; ++k) {
		if (kid[k] < 0) {
			nd.child[k] = kEmptyChild;
			for (int a = 0; a < 3; ++a) { nd.q[a][k] = 255; nd.q[3 + a][k] = 0; }
			nd.slack[k] = 0;
			continue;
		}
		const BNode& c = nodes[kid[k]];
		if (c.left < 0) nd.child[k] = c.count ? ~((c.first << 3) | (c.count - 1)) : kEmptyChild;
		else { nd.child[k] = next_base + at + m; next_queue[at + m] = kid[k]; ++m; }
		nd.slack[k] = half_up_dev(c.slack);
		for (int a = 0; a < 3; ++a) {
			const double l = ((double)c.lo[a] - (double)nd.lo[a]) / step[a];
			const double h = ((double)c.hi[a] - (double)nd.lo[a]) / step[a];
			nd.q[a][k] = (uint8_t)fmax(0.0, fmin(255.0, floor(l)));
			nd.q[3 + a][k] = (uint8_t)fmax(0.0, fmin(255.0, ceil(h)));
		}
	}
	out[out_base + q] = nd;
}
// a scene of one leaf (or none): one node with one child, so traversal can always start with a node fetch
__global__ void single_leaf_kernel(const BNode* nodes, int n, Node* out) {
	Node nd;
	const BNode& bn = nodes[0];
	for (int a = 0; a < 3; ++a) { nd.lo[a] = n ? bn.lo[a] : 0.0f; nd.ex[a] = 127; }
	nd.pad = 0;
	for (int k = 0; k < 4; ++k) {
		nd.child[k] = kEmptyChild; nd.slack[k] = 0;
		for (int a = 0; a < 3; ++a) { nd.q[a][k] = 255; nd.q[3 + a][k] = 0; }
	}
	if (n) {
		nd.child[0] = ~((0 << 3) | (n - 1));
		nd.slack[0] = half_up_dev(bn.slack);
		for (int a = 0; a < 3; ++a) {
			const double ext = (double)bn.hi[a] - (double)bn.lo[a];
			int e = -126;
			if (ext > 0.0) { int ex2; const double m = frexp(ext / 255.0, &ex2); e = m == 0.5 ? ex2 - 1 : ex2; }
			e = max(-126, min(127, e));
			while (e < 127 && ldexp(255.0, e) < ext) ++e;
			nd.ex[a] = (uint8_t)(e + 127);
			nd.q[a][0] = 0;
			nd.q[3 + a][0] = (uint8_t)fmax(0.0, fmin(255.0, ceil(ext / ldexp(1.0, e))));
		}
	}
	out[0] = nd;
}

// triangle records in leaf order (bvh_build.h: TriRecord); the normal in the reference's float order
__global__ void __launch_bounds__(256) records_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tri_material,
                                                      const float4* __restrict__ plo, int n, TriRecord* out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int t = __float_as_int(plo[i].w);
	const float* p = verts + 9 * (size_t)t;
	TriRecord r;
	for (int k = 0; k < 3; ++k) { r.v0[k] = p[k]; r.e1[k] = fsub(p[3 + k], p[k]); r.e2[k] = fsub(p[6 + k], p[k]); }
	r.index = t; r.material = tri_material ? tri_material[t] : 0; r.pad0 = 0; r.pad1 = 0;
	const float cx = fsub(fmul(r.e1[1], r.e2[2]), fmul(r.e1[2], r.e2[1]));
	const float cy = fsub(fmul(r.e1[2], r.e2[0]), fmul(r.e1[0], r.e2[2]));
	const float cz = fsub(fmul(r.e1[0], r.e2[1]), fmul(r.e1[1], r.e2[0]));
	const float len = fsqrt(fadd(fadd(fmul(cx, cx), fmul(cy, cy)), fmul(cz, cz)));
	r.normal[0] = cx; r.normal[1] = cy; r.normal[2] = cz;
	if (len != 0.0f) { r.normal[0] = fdiv(cx, len); r.normal[1] = fdiv(cy, len); r.normal[2] = fdiv(cz, len); }
	out[i] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
struct Result {
	Node* d_nodes = nullptr;        // [n_nodes]    (caller frees)
	TriRecord* d_tris = nullptr;    // [n]          (caller frees)
	int n_nodes = 0, depth = 0;
	float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, diagonal = 0, s0 = 0;
};

struct Scratch {   // freed on every return path
	std::vector<void*> ptrs;
	~Scratch() { for (void* p : ptrs) dev_free(p); }
	template <class T> cudaError_t get(T** out, size_t count) {
		void* p = nullptr;
		const cudaError_t e = dev_alloc(&p, std::max<size_t>(count, 1) * sizeof(T));
		if (e == cudaSuccess) { ptrs.push_back(p); *out = (T*)p; }
		return e;
	}
};

#define DBVH_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e__); return false; } } while (0)

static void exclusive_scan(const int* in, int n, int* out, int* sums, cudaStream_t st) {
	const int blocks = (n + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer);
	scan_sums_kernel<<<blocks, kScanBlock, 0, st>>>(in, n, sums);
	scan_top_kernel<<<1, kScanBlock, 0, st>>>(sums, blocks);
	scan_apply_kernel<<<blocks, kScanBlock, 0, st>>>(in, n, sums, out);
}

// d_verts [n][3][3], d_mat [n] on the current device.  On success `out` owns two device allocations.
static bool build(const float* d_verts, const int32_t* d_mat, int n, cudaStream_t st, Result& out, std::string& err) {
	Scratch sc;
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!dbg) return;
		cudaStreamSynchronize(st);
		const auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[ear_b200] device bvh: %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	const char* knob = std::getenv("EAR_B200_BVH_MARGIN_SCALE");
	const float margin_scale = knob ? (float)std::atof(knob) : 1.0f;
	int max_leaf = kMaxLeaf; float sah_ct = 1.0f;
	if (const char* e = std::getenv("EAR_B200_MAX_LEAF")) max_leaf = std::max(1, std::min(8, std::atoi(e)));
	if (const char* e = std::getenv("EAR_B200_SAH_CT")) sah_ct = (float)std::atof(e);
	Globals* g = nullptr;
	DBVH_TRY(sc.get(&g, 1));
	init_globals_kernel<<<1, 1, 0, st>>>(g);
	if (n > 0) scene_bounds_kernel<<<std::min(1024, (n * 3 + 255) / 256), 256, 0, st>>>(d_verts, n, g);
	scene_constants_kernel<<<1, 1, 0, st>>>(g, n, margin_scale);
	const int np = std::max(n, 1);
	float4 *plo[2], *phi[2];
	int* seg[2];
	int *flags, *scan, *sums;
	BNode* nodes;
	for (int b = 0; b < 2; ++b) { DBVH_TRY(sc.get(&plo[b], np)); DBVH_TRY(sc.get(&phi[b], np)); DBVH_TRY(sc.get(&seg[b], np)); }
	DBVH_TRY(sc.get(&flags, np)); DBVH_TRY(sc.get(&scan, np + 1));
	DBVH_TRY(sc.get(&sums, (np + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer) + 1));
	// scratch binary nodes: the top phase makes < 2 n / kSmallNode + 2, every bottom subtree reserves 2 c slots
	const size_t node_cap = 2 * (size_t)np + 4 * ((size_t)np / kSmallNode + 2) + 8;
	DBVH_TRY(sc.get(&nodes, node_cap));
	const int cap = 2 * (np / kSmallNode + 2);   // most large nodes one level can hold
	Level lv;
	int* next_node;
	int* small_roots;
	DBVH_TRY(sc.get(&lv.node, cap)); DBVH_TRY(sc.get(&next_node, cap));
	DBVH_TRY(sc.get(&lv.acc, (size_t)cap * kBoundWords)); DBVH_TRY(sc.get(&lv.bins, (size_t)cap * kBinWords));
	DBVH_TRY(sc.get(&lv.axis, cap)); DBVH_TRY(sc.get(&lv.bin, cap)); DBVH_TRY(sc.get(&lv.n_left, cap));
	DBVH_TRY(sc.get(&lv.cmin, 3 * (size_t)cap)); DBVH_TRY(sc.get(&lv.scale, 3 * (size_t)cap)); DBVH_TRY(sc.get(&lv.child_seg, 2 * (size_t)cap));
	DBVH_TRY(sc.get(&small_roots, (size_t)np / 1 + 2));   // every small root holds >= 1 primitive
	const int grid_n = (np + 255) / 256;
	lap("scratch allocation");
	if (n > 0) prims_kernel<<<grid_n, 256, 0, st>>>(d_verts, n, g, margin_scale, plo[0], phi[0]);
	lap("bounds + primitives");
	// root
	BNode root{};
	root.left = root.right = -1; root.first = 0; root.count = n;
	DBVH_TRY(cudaMemcpyAsync(nodes, &root, sizeof(root), cudaMemcpyHostToDevice, st));
	Globals hg{};
	int cur = 0;
	int n_active = 0;
	{
		// node 0 is allocated; it is either the first large node or the only small root
		const int one = 1, zero = 0;
		DBVH_TRY(cudaMemcpyAsync(&g->node_count, &one, sizeof(int), cudaMemcpyHostToDevice, st));
		if (n > kSmallNode) {
			n_active = 1;
			DBVH_TRY(cudaMemcpyAsync(lv.node, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
			fill_int_kernel<<<std::min(1024, grid_n), 256, 0, st>>>(seg[0], n, 0);
		} else if (n > 0) {
			DBVH_TRY(cudaMemcpyAsync(small_roots, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
			DBVH_TRY(cudaMemcpyAsync(&g->small_count, &one, sizeof(int), cudaMemcpyHostToDevice, st));
		}
	}
	// ---- top phase ----
	for (int level = 0; n_active > 0; ++level) {
		const int init_grid = (std::max(n_active * kBinWords, n_active * kBoundWords) + 255) / 256;
		level_init_kernel<<<init_grid, 256, 0, st>>>(lv, n_active);
		const int zero = 0;
		DBVH_TRY(cudaMemcpyAsync(&g->next_active, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
		const int top_grid = (n + kTopBlock - 1) / kTopBlock;
		top_bounds_kernel<<<top_grid, kTopBlock, 0, st>>>(plo[cur], phi[cur], seg[cur], n, lv);
		level_ranges_kernel<<<(n_active + 127) / 128, 128, 0, st>>>(lv, n_active);
		top_bins_kernel<<<top_grid, kTopBlock, 0, st>>>(plo[cur], phi[cur], seg[cur], n, lv);
		top_split_kernel<<<(n_active * 32 + 127) / 128, 128, 0, st>>>(lv, n_active, nodes, g, level, next_node, small_roots);
		top_flags_kernel<<<grid_n, 256, 0, st>>>(plo[cur], phi[cur], seg[cur], n, lv, nodes, flags);
		exclusive_scan(flags, n, scan, sums, st);
		top_scatter_kernel<<<grid_n, 256, 0, st>>>(plo[cur], phi[cur], seg[cur], flags, scan, n, lv, nodes, plo[cur ^ 1], phi[cur ^ 1], seg[cur ^ 1]);
		DBVH_TRY(cudaGetLastError());
		DBVH_TRY(cudaMemcpyAsync(&hg, g, sizeof(hg), cudaMemcpyDeviceToHost, st));
		DBVH_TRY(cudaStreamSynchronize(st));
		n_active = hg.next_active;
		if (n_active > cap) { err = "device BVH build: level overflow (internal error)"; return false; }
		std::swap(lv.node, next_node);
		cur ^= 1;
	}
	lap("top phase");
	// ---- bottom phase ----
	DBVH_TRY(cudaMemcpyAsync(&hg, g, sizeof(hg), cudaMemcpyDeviceToHost, st));
	DBVH_TRY(cudaStreamSynchronize(st));
	if (hg.small_count > 0) {
		// leaves write their primitives into the other buffer: it then holds the final (leaf) order
		bottom_kernel<<<hg.small_count, 32 * kBotWarps, 0, st>>>(plo[cur], phi[cur], plo[cur ^ 1], phi[cur ^ 1], nodes, g, small_roots, max_leaf, sah_ct);
		DBVH_TRY(cudaGetLastError());
	}
	lap("bottom phase");
	const float4* final_lo = plo[cur ^ 1];
	// ---- records ----
	DBVH_TRY(dev_alloc(&out.d_tris, (size_t)np * sizeof(TriRecord)));
	if (n > 0) records_kernel<<<grid_n, 256, 0, st>>>(d_verts, d_mat, final_lo, n, out.d_tris);
	lap("triangle records");
	// ---- collapse ----
	DBVH_TRY(cudaMemcpyAsync(&hg, g, sizeof(hg), cudaMemcpyDeviceToHost, st));
	DBVH_TRY(cudaStreamSynchronize(st));
	if ((size_t)hg.node_count > node_cap) { err = "device BVH build: node scratch overflow (internal error)"; return false; }
	BNode hroot;
	DBVH_TRY(cudaMemcpyAsync(&hroot, nodes, sizeof(hroot), cudaMemcpyDeviceToHost, st));
	DBVH_TRY(cudaStreamSynchronize(st));
	const size_t wide_cap = (size_t)hg.node_count / 2 + 2;     // inner binary nodes bound the wide nodes
	DBVH_TRY(dev_alloc(&out.d_nodes, wide_cap * sizeof(Node)));
	if (hroot.left < 0) {
		single_leaf_kernel<<<1, 1, 0, st>>>(nodes, n, out.d_nodes);
		out.n_nodes = 1; out.depth = 1;
	} else {
		int* queue[2];
		int4* kids;
		int *inner, *inner_scan, *qsums;
		DBVH_TRY(sc.get(&queue[0], wide_cap)); DBVH_TRY(sc.get(&queue[1], wide_cap));
		DBVH_TRY(sc.get(&kids, wide_cap)); DBVH_TRY(sc.get(&inner, wide_cap)); DBVH_TRY(sc.get(&inner_scan, wide_cap + 1));
		DBVH_TRY(sc.get(&qsums, wide_cap / (kScanBlock * kScanPer) + 2));
		const int zero = 0;
		DBVH_TRY(cudaMemcpyAsync(queue[0], &zero, sizeof(int), cudaMemcpyHostToDevice, st));
		int n_queue = 1, base = 0, q = 0, depth = 0;
		while (n_queue > 0) {
			++depth;
			const int grid_q = (n_queue + 255) / 256;
			collapse_count_kernel<<<grid_q, 256, 0, st>>>(nodes, queue[q], n_queue, kids, inner);
			exclusive_scan(inner, n_queue, inner_scan, qsums, st);
			int last_scan = 0, last_inner = 0;
			DBVH_TRY(cudaMemcpyAsync(&last_scan, inner_scan + n_queue - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
			DBVH_TRY(cudaMemcpyAsync(&last_inner, inner + n_queue - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
			DBVH_TRY(cudaStreamSynchronize(st));
			const int n_next = last_scan + last_inner;
			if ((size_t)base + n_queue + n_next > wide_cap) { err = "device BVH build: wide node overflow (internal error)"; return false; }
			collapse_emit_kernel<<<grid_q, 256, 0, st>>>(nodes, queue[q], n_queue, kids, inner_scan, base, base + n_queue, out.d_nodes, queue[q ^ 1]);
			DBVH_TRY(cudaGetLastError());
			base += n_queue;
			n_queue = n_next;
			q ^= 1;
		}
		out.n_nodes = base; out.depth = depth;
	}
	DBVH_TRY(cudaStreamSynchronize(st));
	lap("collapse");
	if (dbg) std::fprintf(stderr, "[ear_b200] device bvh: %d triangles, %d scratch nodes, %d wide nodes, depth %d, %d bottom subtrees\n", n, hg.node_count, out.n_nodes, out.depth, hg.small_count);
	for (int a = 0; a < 3; ++a) {
		int lo = hg.lo[a], hi = hg.hi[a];
		lo = lo >= 0 ? lo : lo ^ 0x7fffffff; hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
		std::memcpy(&out.lo[a], &lo, 4); std::memcpy(&out.hi[a], &hi, 4);
	}
	out.diagonal = hg.diagonal; out.s0 = hg.s0;
	return true;
}

}  // namespace dbvh
}  // namespace earb
