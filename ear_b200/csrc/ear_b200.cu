// libear_b200.so -- CUDA kernels (sm_100a) + the C ABI declared in include/ear_b200.h.
//
// Replaces the reference's Scene::Render bounce loop (src/Scene.cpp:111-318) and what it calls.
// Kernels (names follow SURVEY.md section 2.3):
//   K1 emission + K3 material/resample + K6 refill/compaction (`wf_shade_kernel`), K2 closest hit and K4 any hit
//   (`wf_traverse_kernel`, `wf_vismap_kernel`), K5 weight/splat (`wf_splat_kernel`) form the wavefront engine
//   (wavefront.cuh); K0 builds the BVH on the device (bvh_device.cuh); K7 (`scale_kernel`, `direct_kernel`)
//   finalises tracks.  The harness entry points (first_hit, occluded, trace_paths) run the production kernels
//   over explicit rays / segments / ray ids.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/ear_b200.h"
#include "bvh_build.h"
#include "traverse.cuh"

using namespace earb;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int32_t fail(const std::string& msg) { g_last_error = msg; return 1; }

// Device allocations go through a small per-process cache: cudaMalloc / cudaFree of the multi-GB ray pool, the BVH
// scratch and the visibility maps cost tens of milliseconds per scene (more on a busy host), which is most of what a
// short render spends outside its kernels.  A released block is kept (per device, by size) and handed to the next
// request of about that size; the cache is trimmed when it holds more than EAR_B200_CACHE_MB (default 32768) and when
// an allocation fails.  Contents are never assumed: every user initialises what it reads.
namespace devcache {
struct Block { void* p; size_t bytes; int device; };
static std::mutex g_lock;
static std::unordered_map<void*, Block> g_live;
static std::multimap<std::pair<int, size_t>, void*> g_free;   // (device, bytes) -> block
static size_t g_cached = 0;
static size_t limit() {
	static size_t v = 0;
	if (!v) { const char* e = std::getenv("EAR_B200_CACHE_MB"); v = (size_t)(e ? std::max(0, std::atoi(e)) : 32768) << 20; if (!v) v = 1; }
	return v;
}
static void trim_locked(size_t keep) {
	while (g_cached > keep && !g_free.empty()) {
		auto it = std::prev(g_free.end());
		int cur = 0;
		cudaGetDevice(&cur);
		cudaSetDevice(it->first.first);
		cudaFree(it->second);
		cudaSetDevice(cur);
		g_cached -= it->first.second;
		g_free.erase(it);
	}
}
// blocks below this size are not cached (EAR_B200_CACHE_MIN_KB): they hold the hot words of the engine (list counters, queue
// head, context tables), cost microseconds to allocate, and the shade kernel ran up to 30 % slower when they sat in recycled
// blocks (profiles/r2_ab_closest.txt, diag3-diag6)
static size_t min_cached() {
	static size_t v = ~(size_t)0;
	if (v == ~(size_t)0) { const char* e = std::getenv("EAR_B200_CACHE_MIN_KB"); v = (size_t)(e ? std::max(0, std::atoi(e)) : 1024) << 10; }
	return v;
}
static cudaError_t alloc(void** out, size_t bytes) {
	bytes = std::max<size_t>((bytes + 255) / 256 * 256, 256);
	if (bytes < min_cached()) return cudaMalloc(out, bytes);   // release() finds it unknown and hands it to cudaFree
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> g(g_lock);
	auto it = g_free.lower_bound(std::make_pair(dev, bytes));
	if (it != g_free.end() && it->first.first == dev && it->first.second <= bytes + bytes / 8 + (1u << 20)) {
		*out = it->second;
		g_live[*out] = Block{*out, it->first.second, dev};
		g_cached -= it->first.second;
		g_free.erase(it);
		return cudaSuccess;
	}
	cudaError_t e = cudaMalloc(out, bytes);
	if (e != cudaSuccess) { cudaGetLastError(); trim_locked(0); e = cudaMalloc(out, bytes); }
	if (e == cudaSuccess) g_live[*out] = Block{*out, bytes, dev};
	return e;
}
static void release(void* p) {
	if (!p) return;
	std::lock_guard<std::mutex> g(g_lock);
	auto it = g_live.find(p);
	if (it == g_live.end()) { cudaFree(p); return; }
	g_free.emplace(std::make_pair(it->second.device, it->second.bytes), p);
	g_cached += it->second.bytes;
	g_live.erase(it);
	trim_locked(limit());
}
static void trim_all() { std::lock_guard<std::mutex> g(g_lock); trim_locked(0); }
}  // namespace devcache
template <class T> static cudaError_t dev_alloc(T** out, size_t bytes) { return devcache::alloc((void**)out, bytes); }
static void dev_free(void* p) { devcache::release(p); }

// Page-locked host blocks for the tracks a render hands out (one block per result).  cudaHostAlloc of tens of MB costs
// milliseconds, and a pageable destination makes the download a staged copy through page faults; a released block is kept
// for the next result of about that size (bounded: at most kHostCacheBlocks blocks).
namespace hostcache {
struct Block { void* p; size_t bytes; };
static std::mutex g_lock;
static std::vector<Block> g_free;
constexpr size_t kHostCacheBlocks = 4;
static void* take(size_t bytes, size_t* got) {
	bytes = std::max<size_t>((bytes + 4095) / 4096 * 4096, 4096);
	{
		std::lock_guard<std::mutex> g(g_lock);
		size_t best = g_free.size();
		for (size_t i = 0; i < g_free.size(); ++i)
			if (g_free[i].bytes >= bytes && g_free[i].bytes <= bytes + bytes / 4 + (1u << 20) && (best == g_free.size() || g_free[i].bytes < g_free[best].bytes)) best = i;
		if (best != g_free.size()) {
			const Block b = g_free[best];
			g_free.erase(g_free.begin() + (long)best);
			*got = b.bytes;
			return b.p;
		}
	}
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	*got = bytes;
	return p;
}
static size_t byte_limit() {   // page-locked memory is a scarce host resource: EAR_B200_HOST_CACHE_MB (default 8192)
	static size_t v = 0;
	if (!v) { const char* e = std::getenv("EAR_B200_HOST_CACHE_MB"); v = ((size_t)(e ? std::max(0, std::atoi(e)) : 8192) << 20) + 1; }
	return v;
}
static void give(void* p, size_t bytes) {
	if (!p) return;
	std::lock_guard<std::mutex> g(g_lock);
	g_free.push_back(Block{p, bytes});
	size_t held = 0;
	for (const auto& b : g_free) held += b.bytes;
	while (!g_free.empty() && (g_free.size() > kHostCacheBlocks || held > byte_limit())) {   // drop the oldest
		held -= g_free.front().bytes;
		cudaFreeHost(g_free.front().p);
		g_free.erase(g_free.begin());
	}
}
static void trim_all() {
	std::lock_guard<std::mutex> g(g_lock);
	for (auto& b : g_free) cudaFreeHost(b.p);
	g_free.clear();
}
}  // namespace hostcache

// call-scoped device buffer: freed on every return path (the ABI functions leave through CUDA_TRY on errors)
template <class T>
struct DevBuf {
	T* p = nullptr;
	DevBuf() = default;
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	~DevBuf() { if (p) dev_free(p); }
	cudaError_t alloc(size_t count) { return dev_alloc(&p, std::max<size_t>(count, 1) * sizeof(T)); }
	operator T*() const { return p; }
};
#define CUDA_TRY(expr)                                                                          \
	do {                                                                                        \
		cudaError_t err__ = (expr);                                                             \
		if (err__ != cudaSuccess)                                                               \
			return fail(std::string(#expr) + ": " + cudaGetErrorString(err__));                 \
	} while (0)

// ------------------------------------------------------------------------------------------
// device-side parameter blocks
// ------------------------------------------------------------------------------------------
constexpr int kCounterParts = 256;   // partial rows of the shade kernel's counters (wavefront.cuh)
struct RenderParams {
	const ear_b200_context* ctx;     // [n_ctx]
	const ear_b200_recorder* rec;    // [n_ctx][n_rec]
	const long long* work_prefix;    // [n_ctx + 1] rays of this shard, prefix-summed over contexts
	const float* ctx_emit_area;      // [n_ctx] Mesh::total_area of a mesh source's emitter (float sum in file order), else 0
	int32_t n_ctx, n_rec, max_bounces, n_bins;
	int32_t tpr;                     // tracks per recorder in hist / range: 2 when the call has a stereo recorder, else 1
	unsigned long long seed;
	long long first_ray;
	long long total_work;
	float* hist;                     // [n_ctx][n_rec][tpr][n_bins]
	uint32_t* range;                 // [n_ctx][n_rec][tpr][2] {first_sample, real_length}
	unsigned long long* counters;    // [8]
	unsigned long long* counter_parts;   // [kCounterParts][4] partial rows of counters 0..2 (shade kernel), folded at the end of a trace
	unsigned long long* next_work;   // work-queue head
	int32_t* hits;                   // parity harness: [rays][max_bounces] triangle hit per bounce (or null)
	float* final_state;              // parity harness: [rays][8] (or null)
};

struct LocalCounters {
	unsigned long long rays, segments, occlusion, contributions, bin_updates, dropped;
};

// Philox "context" word of a ray's streams: the context's position in the call unless the caller pinned it
__device__ __forceinline__ uint32_t stream_key(const RenderParams& p, int c) {
	const int32_t id = p.ctx[c].stream_id;
	return id > 0 ? (uint32_t)(id - 1) : (uint32_t)c;
}

#define PI_F 3.14159265f  /* src/Distributions.h:37 */

// FloatBuffer::operator[] bookkeeping (src/Recorder.cpp:52-59): first_sample = min touched index,
// real_length = max touched index.  Read first, touch the atomic only when it would move.
__device__ __forceinline__ void touch_range(uint32_t* range, int lo, int hi) {
	const uint32_t cur_first = __ldcg(range), cur_real = __ldcg(range + 1);
	if ((uint32_t)lo < cur_first) atomicMin(range, (uint32_t)lo);
	if ((uint32_t)hi > cur_real) atomicMax(range + 1, (uint32_t)hi);
}

// Where Record() adds its samples.  GlobalSink: straight into the HBM/L2-resident histogram (one RED.ADD.F32 per bin).
// WindowSink (wf_splat_window_kernel): into a per-block shared-memory copy of a time window of the recorder's tracks,
// flushed once per block; samples outside the window fall through to the global histogram.
struct GlobalSink {
	float* tracks;      // [tpr][n_bins] of this (context, recorder)
	uint32_t* range;    // [tpr][2]
	int n_bins;
	__device__ __forceinline__ void add(int k, int idx, float v) { atomicAdd(tracks + (size_t)k * n_bins + idx, v); }
	__device__ __forceinline__ void touched(int k, int lo, int hi) { touch_range(range + 2 * k, lo, hi); }
};
struct WindowSink {
	float* tracks;
	int n_bins;
	float* win;         // shared memory: [tracks of the recorder][width]
	int lo, width;      // the window covers bins [lo, lo + width) of each track
	int t_lo[2], t_hi[2];   // bins this thread touched, per track (merged per block at the flush)
	__device__ __forceinline__ void add(int k, int idx, float v) {
		const unsigned rel = (unsigned)(idx - lo);
		if (rel < (unsigned)width) atomicAdd(win + k * width + (int)rel, v);
		else atomicAdd(tracks + (size_t)k * n_bins + idx, v);
	}
	__device__ __forceinline__ void touched(int k, int lo_bin, int hi_bin) { t_lo[k] = min(t_lo[k], lo_bin); t_hi[k] = max(t_hi[k], hi_bin); }
};

// One linearly decaying ramp of `w` bins starting at bin `s` (the USE_FILTER splat shared by
// MonoRecorder::Record, src/MonoRecorder.cpp:83-97, and StereoRecorder::Record, :120-129).
template <class Sink>
__device__ __forceinline__ void splat_ramp(Sink& sink, int k, int n_bins, int s, int w, float ampl, float step, LocalCounters& lc) {
	int lo = 0x7fffffff, hi = -1;
	for (int i = 0; i < w; ++i) {
		const int idx = s + i;
		if (idx >= 0) {                         // StereoRecorder::_Sample drops negative bins (:93)
			if (idx < n_bins) {
				sink.add(k, idx, ampl);
				lo = min(lo, idx); hi = max(hi, idx);
				++lc.bin_updates;
			} else ++lc.dropped;
		}
		ampl = fsub(ampl, step);
	}
	if (hi >= 0) sink.touched(k, lo, hi);
}

template <class Sink>
__device__ __forceinline__ void record(const ear_b200_recorder& rec, Sink& sink, int n_bins, V3 dir, float a, float t, float dist,
                                       int band, LocalCounters& lc) {
	++lc.contributions;
	const float width = fsqrt(dist);
	const float ampl = fdiv(fmul(2.0f, a), width);
	const int w = (int)ceilf(width);
	if (rec.kind == EAR_B200_STEREO) {
		const float dt = vdot(dir, mk(rec.right_ear[0], rec.right_ear[1], rec.right_ear[2]));
		const float time_difference = fdiv(rec.head_size, 343.0f);
		const int s_right = __double2int_rz((double)fsub(t, fmul(dt, time_difference)) * 44100.0);
		const int s_left = __double2int_rz((double)fadd(t, fmul(dt, time_difference)) * 44100.0);
		float ampl_left = ampl, ampl_right = ampl;
		const float factor = pow_ref(rec.head_absorption[band], fmul(fabsf(dt), rec.head_size));
		if (dt < 0) ampl_right = fmul(ampl_right, fmul(factor, factor));
		else ampl_left = fmul(ampl_left, fmul(factor, factor));
		splat_ramp(sink, 0, n_bins, s_left, w, ampl_left, fdiv(ampl_left, (float)w), lc);
		splat_ramp(sink, 1, n_bins, s_right, w, ampl_right, fdiv(ampl_right, (float)w), lc);
	} else {
		const int s = __double2int_rz((double)t * 44100.0);
		splat_ramp(sink, 0, n_bins, s, w, ampl, fdiv(ampl, (float)w), lc);
	}
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

constexpr int kBlock = 128;

#include "wavefront.cuh"
#include "vismap.cuh"
#include "post.cuh"
#include "bvh_device.cuh"

// ---- K7: Scene::Render's tail, src/Scene.cpp:286-316 ----
// FloatBuffer::Multiply over [first_sample, real_length) -- the last touched bin is NOT scaled
// (src/Recorder.cpp:85-89).  mode 0: x 1/amount, mode 1: x gain^2.
__global__ void scale_kernel(RenderParams p, int mode) {
	const int n_tracks = p.n_ctx * p.n_rec * p.tpr;
	for (int track = blockIdx.y; track < n_tracks; track += gridDim.y) {   // (ctx * n_rec + rec) * tpr + k
		const int c = track / (p.tpr * p.n_rec);
		const int r = (track / p.tpr) % p.n_rec;
		if ((track % p.tpr) && p.rec[(size_t)c * p.n_rec + r].kind != EAR_B200_STEREO) continue;
		const uint32_t first = p.range[2 * track], real = p.range[2 * track + 1];
		float f;
		if (mode == 0) {
			// `amount` is a float bumped by 1.0f per ray (src/Scene.cpp:122,127): it saturates at 2^24
			const long long ns = p.ctx[c].num_samples;
			f = fdiv(1.0f, (float)(ns < 16777216 ? ns : 16777216));
		} else f = fmul(p.ctx[c].gain, p.ctx[c].gain);
		float* tr = p.hist + (size_t)track * p.n_bins;
		for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < real; i += gridDim.x * blockDim.x)
			tr[i] = fmul(tr[i], f);
	}
}
// Direct sound (src/Scene.cpp:299-311): one lane per (context, recorder).
template <bool EXACT>
__global__ void __launch_bounds__(kBlock) direct_kernel(SceneDev sc, RenderParams p) {
	extern __shared__ int2 stack_smem[];
	const int n_pairs = p.n_ctx * p.n_rec;
	for (int base = blockIdx.x * blockDim.x; base < n_pairs; base += gridDim.x * blockDim.x) {   // block-uniform trip count
		const int i = base + threadIdx.x;
		const bool active = i < n_pairs;
		const int j = active ? i : 0;
		const int c = j / p.n_rec;
		const ear_b200_context cx = p.ctx[c];
		const ear_b200_recorder& rec = p.rec[j];
		const V3 listener = mk(rec.position[0], rec.position[1], rec.position[2]);
		const V3 source = mk(cx.source_position[0], cx.source_position[1], cx.source_position[2]);
		float t; int32_t occ, slot;
		traverse_warp<true, EXACT>(sc, stack_smem + threadIdx.x, blockDim.x, active && cx.source_kind == EAR_B200_POINT_SOURCE,
		                           listener, vsub(source, listener), t, occ, slot);
		// "not a mesh source" (src/Scene.cpp:299): mesh emitters have no direct lobe
		if (!active || occ || cx.source_kind != EAR_B200_POINT_SOURCE) continue;
		const V3 dist = vsub(listener, source);
		const float len = vlength(dist);
		const V3 dir = vnormalized(dist);
		const float a = fmul(fmul(fdiv(1.0f, fmul(fmul(fmul(4.0f, PI_F), len), len)), pow_ref(cx.absorption_factor, len)),
		                     cx.dry_level);
		LocalCounters lc = {0, 0, 0, 0, 0, 0};
		GlobalSink sink = {p.hist + (size_t)j * p.tpr * p.n_bins, p.range + (size_t)j * p.tpr * 2, p.n_bins};
		record(rec, sink, p.n_bins, dir, a, fdiv(len, 343.0f), len, cx.band, lc);
		if (p.counters) {   // the direct lobe is a Record() call like any other (src/Scene.cpp:308)
			atomicAdd(p.counters + 3, lc.contributions); atomicAdd(p.counters + 4, lc.bin_updates);
			atomicAdd(p.counters + 5, lc.dropped);
		}
	}
}
__global__ void init_range_kernel(uint32_t* range, int n_tracks) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_tracks) { range[2 * i] = 3 * EAR_B200_SAMPLE_RATE - 1; range[2 * i + 1] = 0; }
}

constexpr size_t kStackBytes = (size_t)kStackEntries * kBlock * sizeof(int2);

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct ear_b200_scene {
	int device = 0;
	SceneDev dev{};
	unsigned char* d_image = nullptr;   // [header | nodes | tris | materials], one allocation
	size_t image_bytes = 0;
	int2* d_spill = nullptr;            // traversal-stack overflow (SceneDev::spill)
	float4* d_emitters = nullptr;       // emitter triangles of mesh sources, 4 float4 each (ear_b200_scene_set_emitters)
	int32_t n_emitters = 0;
	std::vector<float> emitter_area;    // host copy of the areas: Mesh::total_area per context is summed on the host
	std::vector<float> emitter_verts;   // host copy of the table (replicas on other GPUs are given the same emitters)
	float4* d_nodes = nullptr;          // views into d_image
	float4* d_tris = nullptr;
	float4* d_materials = nullptr;
	std::vector<float> materials;
	int32_t n_tris = 0, n_materials = 0, n_bands = 0, n_nodes = 0, depth = 0;
	float diagonal = 0.0f;
	float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
	double bvh_build_ms = 0.0;
	cudaStream_t stream = nullptr;
	int sm_count = 148;
	int min_blocks = 4;
	int max_slots = 1 << 24;        // most rays in flight in the wavefront pool (EAR_B200_SLOTS); per 8e7 rays: 8 Mi 1.50e9,
	                                // 16 Mi 1.55e9, 32 Mi 1.56e9, 64 Mi 1.55e9 seg/s (bigger launches amortise the persistent
	                                // kernels' tails and sort better, but the shade kernel walks every slot)
	bool slots_forced = false;
	int pool_generations = 0;       // 0: choose from the bounce cap; k: pool = work / k slots (EAR_B200_GENERATIONS)
	int check_every = 8;            // iterations between host checks for completion
	int sort_queries = -1;          // counting-sort the occlusion queries by (recorder, cell): -1 = for up to 4 recorders (EAR_B200_SORT_QUERIES=0/1)
	int splat_mode = 0;             // 0: one RED per ramp sample; 1: shared-memory time-window privatisation (EAR_B200_SPLAT=window)
	int grid_wave = 1;              // lookups / splat launch one wave of resident blocks (EAR_B200_GRID_WAVE)
	int ray_key = 2;                // binning of closest-hit rays (EAR_B200_RAY_KEY, see ray_bin; 2 measured best)
	WfPool pool{};
	size_t pool_slots = 0, pool_queries = 0, log2af_cap = 0;
	int* h_counts = nullptr;        // pinned
	unsigned long long* d_scratch_counters = nullptr;
	ear_b200_stats stats{};
	// recorder visibility maps (vismap.cuh), cached per recorder position
	struct VisMapHost { float x[3]; int res; int* d_offsets; int* d_items; size_t n_items; bool sorted; };
	int vismap_build = 0;           // 0: sort-based build (no returning atomics, lists nearest-first); 1: count + atomic fill (EAR_B200_VISMAP_BUILD=atomic)
	int vismap_sort = -1;           // -1: order the texel lists by distance once most queries turn out blocked; 0 never; 1 at build time (EAR_B200_VISMAP_SORT)
	std::vector<VisMapHost> vismaps;
	int vismap_res = -1;            // -1: choose from the triangle count; 0: disabled (EAR_B200_VISMAP_RES)
	VisMapDev* d_maps = nullptr; int* d_map_of = nullptr; size_t map_of_cap = 0;
	uint2* d_q_bvh = nullptr; size_t q_bvh_cap = 0;
	float* d_post = nullptr; size_t post_cap = 0;   // post-chain scratch: [n_tracks] float + [n_tracks] uint32
	int* d_vis_counts = nullptr; long long* d_vis_sums = nullptr; size_t vis_scratch_cap = 0;
	size_t vis_budget = 0, vis_bytes = 0; bool vis_budget_spent = false;   // memory guard for the maps
	std::vector<ear_b200_recorder> h_rec;   // host copy of the recorders of the current call
	float maxabs = 0.0f;
	// event pairs recorded around every engine launch, harvested at the engine's sync points
	struct Timed { cudaEvent_t a, b; int cls; };
	std::vector<Timed> ev_pool;
	size_t ev_used = 0;
	cudaStream_t last_stream = nullptr;
	// scratch reused across calls
	ear_b200_context* d_ctx = nullptr; ear_b200_recorder* d_rec = nullptr; long long* d_prefix = nullptr; float* d_ctx_area = nullptr;
	unsigned long long* d_queue = nullptr;
	unsigned long long* d_counter_parts = nullptr;   // [kCounterParts][4], all zero between calls
	size_t ctx_cap = 0, rec_cap = 0;
};

extern "C" const char* ear_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t ear_b200_abi_version(void) { return EAR_B200_ABI_VERSION; }
extern "C" void ear_b200_release_cached_memory(void) { devcache::trim_all(); hostcache::trim_all(); }
extern "C" int32_t ear_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" void ear_b200_scene_destroy(ear_b200_scene* s);
extern "C" void ear_b200_group_destroy(ear_b200_group* g);

// Device image of a scene: [header | nodes | triangle records | materials] in ONE allocation, so a scene built
// on one GPU can be replicated to the others with a single NVLink broadcast instead of N host BVH builds.
struct ImageHeader {
	uint32_t magic, version;
	int32_t n_tris, n_materials, n_bands, n_nodes, depth, reserved;
	float diagonal, s0, maxabs, reserved_f;
	float lo[3], hi[3];
	uint64_t off_nodes, off_tris, off_materials, bytes;
};
constexpr uint32_t kImageMagic = 0x42524145u;   // "EARB"
constexpr size_t kImageAlign = 256;
static_assert(sizeof(ImageHeader) <= kImageAlign, "header must fit its slot");

static size_t image_round(size_t x) { return (x + kImageAlign - 1) / kImageAlign * kImageAlign; }

static void image_layout(ImageHeader& h) {
	h.off_nodes = kImageAlign;
	h.off_tris = h.off_nodes + image_round(std::max<size_t>((size_t)h.n_nodes, 1) * sizeof(Node));
	h.off_materials = h.off_tris + image_round(std::max<size_t>((size_t)h.n_tris, 1) * sizeof(TriRecord));
	h.bytes = h.off_materials + image_round((size_t)h.n_materials * h.n_bands * 4 * sizeof(float));
}

static void read_slot_knob(ear_b200_scene* s) {
	const char* sl = std::getenv("EAR_B200_SLOTS");
	if (sl) { s->max_slots = std::max(256, std::min(1 << kMaxSlotBits, std::atoi(sl))); s->slots_forced = true; }
}

static int32_t ensure_pool(ear_b200_scene* s, size_t slots, size_t queries);

// everything of scene creation that does not depend on where the image came from
static int32_t scene_finish(ear_b200_scene* s, const ImageHeader& h) {
	s->n_tris = h.n_tris; s->n_materials = h.n_materials; s->n_bands = h.n_bands;
	s->n_nodes = h.n_nodes; s->depth = h.depth; s->diagonal = h.diagonal; s->maxabs = h.maxabs;
	for (int k = 0; k < 3; ++k) { s->lo[k] = h.lo[k]; s->hi[k] = h.hi[k]; }
	s->image_bytes = (size_t)h.bytes;
	s->d_nodes = (float4*)(s->d_image + h.off_nodes);
	s->d_tris = (float4*)(s->d_image + h.off_tris);
	s->d_materials = (float4*)(s->d_image + h.off_materials);
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
	s->sm_count = prop.multiProcessorCount;
	{   // diagnosis knob: EAR_B200_STREAM=null runs the scene's own work on the legacy default stream, =blocking on a blocking stream
		const char* sk = std::getenv("EAR_B200_STREAM");
		if (sk && std::strcmp(sk, "null") == 0) s->stream = nullptr;
		else if (sk && std::strcmp(sk, "blocking") == 0) CUDA_TRY(cudaStreamCreate(&s->stream));
		else CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	}
	CUDA_TRY(dev_alloc(&s->d_queue, sizeof(unsigned long long)));
	CUDA_TRY(dev_alloc(&s->d_counter_parts, kCounterParts * 4 * sizeof(unsigned long long)));
	CUDA_TRY(cudaMemset(s->d_counter_parts, 0, kCounterParts * 4 * sizeof(unsigned long long)));
	s->dev.nodes = s->d_nodes; s->dev.tris = s->d_tris; s->dev.materials = s->d_materials;
	s->dev.n_tris = h.n_tris; s->dev.n_materials = h.n_materials; s->dev.n_bands = h.n_bands;
	s->dev.s0 = h.s0;
	// overflow rows of the traversal stacks: a walk holds at most 3 entries per level of the 4-wide tree (+ slack)
	s->dev.spill_threads = s->sm_count * 16 * kBlock;   // 16 blocks of kBlock threads is the most an SM can hold
	s->dev.spill_rows = std::max(1, 3 * h.depth + 4 - kStackEntries);
	CUDA_TRY(dev_alloc(&s->d_spill, (size_t)s->dev.spill_rows * s->dev.spill_threads * sizeof(int2)));
	s->dev.spill = s->d_spill;
	// EAR_B200_EXACT_SLACK=1 selects the rigorous per-child interval bound (about 2x the node visits)
	const char* ex = std::getenv("EAR_B200_EXACT_SLACK");
	s->dev.exact = (ex && std::atoi(ex) != 0) ? 1 : 0;
	// tuning knobs (defaults are the measured best, see profiles/)
	const char* lv = std::getenv("EAR_B200_LEAF_VOTE");
	s->dev.leaf_vote = lv ? std::max(1, std::min(32, std::atoi(lv))) : kLeafVote;
	const char* mb = std::getenv("EAR_B200_MIN_BLOCKS");
	s->min_blocks = mb ? std::atoi(mb) : 5;
	const char* fv = std::getenv("EAR_B200_FETCH_VOTE");
	s->dev.fetch_vote = fv ? std::max(1, std::min(32, std::atoi(fv))) : 8;
	read_slot_knob(s);
	s->dev.vis_cap = kVisMaxList;
	if (const char* vc = std::getenv("EAR_B200_VISMAP_CAP")) s->dev.vis_cap = std::max(0, std::min(4096, std::atoi(vc)));
	s->dev.vis_prefix = 32;
	if (const char* vp = std::getenv("EAR_B200_VISMAP_PREFIX")) s->dev.vis_prefix = std::max(0, std::min(4096, std::atoi(vp)));
	if (const char* vr = std::getenv("EAR_B200_VISMAP_RES")) s->vismap_res = std::max(0, std::min(2048, std::atoi(vr)));
	if (const char* vb = std::getenv("EAR_B200_VISMAP_BUILD")) s->vismap_build = std::string(vb) == "atomic" ? 1 : 0;
	if (const char* vs = std::getenv("EAR_B200_VISMAP_SORT")) s->vismap_sort = std::max(-1, std::min(1, std::atoi(vs)));
	if (const char* gw = std::getenv("EAR_B200_GRID_WAVE")) s->grid_wave = std::atoi(gw) != 0 ? 1 : 0;
	if (const char* sq = std::getenv("EAR_B200_SORT_QUERIES")) s->sort_queries = std::atoi(sq) != 0 ? 1 : 0;
	if (const char* sm = std::getenv("EAR_B200_SPLAT")) s->splat_mode = std::string(sm) == "window" ? 1 : 0;
	if (const char* pg = std::getenv("EAR_B200_GENERATIONS")) s->pool_generations = std::max(0, std::atoi(pg));
	if (const char* rk = std::getenv("EAR_B200_RAY_KEY")) s->ray_key = std::max(0, std::min(4, std::atoi(rk)));
	const char* ce = std::getenv("EAR_B200_CHECK_EVERY");
	if (ce) s->check_every = std::max(1, std::atoi(ce));
	CUDA_TRY(cudaMallocHost(&s->h_counts, 16 * sizeof(int)));   // 8 pool counters + the queue head
	CUDA_TRY(dev_alloc(&s->d_scratch_counters, 8 * sizeof(unsigned long long)));
	return 0;
}

static int32_t pick_device(int32_t device, const char* who) {
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail("no CUDA device: ear_b200 has no CPU fallback");
	}
	if (device < 0 || device >= ndev) return fail(std::string(who) + ": device index out of range");
	CUDA_TRY(cudaSetDevice(device));
	return 0;
}

extern "C" int32_t ear_b200_scene_create(const float* verts, const int32_t* tri_material, int32_t n_tris,
                                         const float* materials, int32_t n_materials, int32_t n_bands, int32_t device,
                                         ear_b200_scene** out) {
	if (!out) return fail("scene_create: out is null");
	*out = nullptr;
	if (n_tris < 0 || n_materials <= 0 || n_bands <= 0 || n_bands > EAR_B200_MAX_BANDS)
		return fail("scene_create: bad sizes (need n_materials >= 1, 1 <= n_bands <= 8)");
	if ((n_tris > 0 && !verts) || !materials) return fail("scene_create: null input");
	if (n_tris >= (1 << 28)) return fail("scene_create: too many triangles (limit 2^28)");
	for (int32_t i = 0; i < n_tris && tri_material; ++i)
		if (tri_material[i] < 0 || tri_material[i] >= n_materials) return fail("scene_create: material index out of range");
	if (int32_t rc = pick_device(device, "scene_create")) return rc;
	std::unique_ptr<ear_b200_scene, void (*)(ear_b200_scene*)> s(new ear_b200_scene(), ear_b200_scene_destroy);
	s->device = device;
	const auto t0 = std::chrono::steady_clock::now();
	// K0: the BVH is built on the device (bvh_device.cuh) unless the scene is tiny or EAR_B200_BUILD=host asks for
	// the host builder (bvh_build.cpp, kept for A/B and for the CPU-side tests of the traversal logic)
	const char* bk = std::getenv("EAR_B200_BUILD");
	const bool on_device = bk ? std::string(bk) == "device" : n_tris >= 2048;
	ImageHeader h{};
	h.magic = kImageMagic; h.version = EAR_B200_ABI_VERSION;
	h.n_tris = n_tris; h.n_materials = n_materials; h.n_bands = n_bands;
	s->materials.assign(materials, materials + (size_t)n_materials * n_bands * 4);
	if (on_device) {
		DevBuf<float> d_verts;
		DevBuf<int32_t> d_mat;
		CUDA_TRY(d_verts.alloc((size_t)n_tris * 9));
		if (n_tris) CUDA_TRY(cudaMemcpyAsync(d_verts, verts, (size_t)n_tris * 36, cudaMemcpyHostToDevice, 0));
		if (tri_material) { CUDA_TRY(d_mat.alloc((size_t)n_tris)); if (n_tris) CUDA_TRY(cudaMemcpyAsync(d_mat, tri_material, (size_t)n_tris * 4, cudaMemcpyHostToDevice, 0)); }
		dbvh::Result r;
		std::string err;
		const bool ok = dbvh::build(d_verts, tri_material ? d_mat.p : nullptr, n_tris, 0, r, err);
		DevBuf<Node> own_nodes; own_nodes.p = r.d_nodes;          // released on every path below
		DevBuf<TriRecord> own_tris; own_tris.p = r.d_tris;
		if (!ok) return fail(err);
		h.n_nodes = r.n_nodes; h.depth = r.depth; h.diagonal = r.diagonal; h.s0 = r.s0;
		for (int k = 0; k < 3; ++k) {
			h.lo[k] = r.lo[k]; h.hi[k] = r.hi[k];
			h.maxabs = std::max(h.maxabs, std::max(std::fabs(r.lo[k]), std::fabs(r.hi[k])));
		}
		image_layout(h);
		CUDA_TRY(dev_alloc(&s->d_image, (size_t)h.bytes));
		CUDA_TRY(cudaMemcpyAsync(s->d_image, &h, sizeof(h), cudaMemcpyHostToDevice, 0));
		CUDA_TRY(cudaMemcpyAsync(s->d_image + h.off_nodes, r.d_nodes, (size_t)r.n_nodes * sizeof(Node), cudaMemcpyDeviceToDevice, 0));
		if (n_tris) CUDA_TRY(cudaMemcpyAsync(s->d_image + h.off_tris, r.d_tris, (size_t)n_tris * sizeof(TriRecord), cudaMemcpyDeviceToDevice, 0));
		CUDA_TRY(cudaMemcpyAsync(s->d_image + h.off_materials, s->materials.data(), s->materials.size() * sizeof(float), cudaMemcpyHostToDevice, 0));
		CUDA_TRY(cudaStreamSynchronize(0));
	} else {
		Bvh bvh;
		build_bvh(verts, tri_material, n_tris, bvh);
		h.n_nodes = (int32_t)bvh.nodes.size(); h.depth = bvh.depth; h.diagonal = bvh.diagonal; h.s0 = bvh.s0;
		for (int k = 0; k < 3; ++k) {
			h.lo[k] = bvh.lo[k]; h.hi[k] = bvh.hi[k];
			h.maxabs = std::max(h.maxabs, std::max(std::fabs(bvh.lo[k]), std::fabs(bvh.hi[k])));
		}
		image_layout(h);
		CUDA_TRY(dev_alloc(&s->d_image, (size_t)h.bytes));
		CUDA_TRY(cudaMemcpy(s->d_image, &h, sizeof(h), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(s->d_image + h.off_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(Node), cudaMemcpyHostToDevice));
		if (!bvh.tris.empty())
			CUDA_TRY(cudaMemcpy(s->d_image + h.off_tris, bvh.tris.data(), bvh.tris.size() * sizeof(TriRecord), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(s->d_image + h.off_materials, s->materials.data(), s->materials.size() * sizeof(float), cudaMemcpyHostToDevice));
	}
	s->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (int32_t rc = scene_finish(s.get(), h)) return rc;
	if (std::getenv("EAR_B200_DEBUG"))
		std::fprintf(stderr, "[ear_b200] scene_create: %d triangles, %s build %.2f ms, total %.2f ms\n", n_tris, on_device ? "device" : "host",
		             s->bvh_build_ms, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	*out = s.release();
	return 0;
}

// Emitter triangles of mesh sources.  Record: (v0, area) (v1, 0) (v2, 0) (unit normal, 0); normal and area in the
// float order of Triangle's constructor (src/Triangle.cpp:26-34: gmtl::normal = normalize(edge(0) x (v2 - v0))... as
// fixed by oracle/shim/gmtl/gmtl.h; area = |edge(0) x edge(1)| / 2 with edge(i) = v[(i+1)%3] - v[i]).  This file is
// compiled with -ffp-contract=off, so the host arithmetic below rounds once per operation.
extern "C" int32_t ear_b200_scene_set_emitters(ear_b200_scene* s, const float* verts, int32_t n) {
	if (!s) return fail("scene_set_emitters: null scene");
	if (n < 0 || (n > 0 && !verts)) return fail("scene_set_emitters: bad arguments");
	CUDA_TRY(cudaSetDevice(s->device));
	dev_free(s->d_emitters); s->d_emitters = nullptr; s->n_emitters = 0; s->emitter_area.clear(); s->emitter_verts.clear();
	s->dev.emitters = nullptr;
	if (n == 0) return 0;
	std::vector<float> rec((size_t)n * 16, 0.0f);
	s->emitter_area.resize((size_t)n);
	s->emitter_verts.assign(verts, verts + 9 * (size_t)n);
	for (int32_t i = 0; i < n; ++i) {
		const float* p = verts + 9 * (size_t)i;
		float* r = rec.data() + 16 * (size_t)i;
		float e1[3], e2[3], e12[3];
		for (int k = 0; k < 3; ++k) { e1[k] = p[3 + k] - p[k]; e2[k] = p[6 + k] - p[k]; e12[k] = p[6 + k] - p[3 + k]; }
		// gmtl::normal(tri): cross(v1 - v0, v2 - v0), normalised by division
		const float cx = (e1[1] * e2[2]) - (e1[2] * e2[1]);
		const float cy = (e1[2] * e2[0]) - (e1[0] * e2[2]);
		const float cz = (e1[0] * e2[1]) - (e1[1] * e2[0]);
		float l2 = cx * cx; l2 = l2 + cy * cy; l2 = l2 + cz * cz;
		const float len = std::sqrt(l2);
		float nx = cx, ny = cy, nz = cz;
		if (len != 0.0f) { nx = cx / len; ny = cy / len; nz = cz / len; }
		// Triangle::calcArea: cross(edge(0), edge(1)) = (v1 - v0) x (v2 - v1)
		const float ax = (e1[1] * e12[2]) - (e1[2] * e12[1]);
		const float ay = (e1[2] * e12[0]) - (e1[0] * e12[2]);
		const float az = (e1[0] * e12[1]) - (e1[1] * e12[0]);
		float a2 = ax * ax; a2 = a2 + ay * ay; a2 = a2 + az * az;
		const float area = std::sqrt(a2) / 2.0f;
		for (int k = 0; k < 3; ++k) { r[k] = p[k]; r[4 + k] = p[3 + k]; r[8 + k] = p[6 + k]; }
		r[3] = area; r[12] = nx; r[13] = ny; r[14] = nz;
		s->emitter_area[(size_t)i] = area;
	}
	CUDA_TRY(dev_alloc(&s->d_emitters, rec.size() * sizeof(float)));
	CUDA_TRY(cudaMemcpy(s->d_emitters, rec.data(), rec.size() * sizeof(float), cudaMemcpyHostToDevice));
	s->n_emitters = n;
	s->dev.emitters = s->d_emitters;
	return 0;
}

extern "C" int32_t ear_b200_scene_image_size(ear_b200_scene* s, uint64_t* bytes) {
	if (!s || !bytes) return fail("scene_image_size: null argument");
	*bytes = (uint64_t)s->image_bytes;
	return 0;
}

extern "C" int32_t ear_b200_scene_image_write(ear_b200_scene* s, void* dst_device, uint64_t bytes) {
	if (!s || !dst_device) return fail("scene_image_write: null argument");
	if (bytes < (uint64_t)s->image_bytes) return fail("scene_image_write: destination smaller than the image");
	CUDA_TRY(cudaSetDevice(s->device));
	CUDA_TRY(cudaMemcpy(dst_device, s->d_image, s->image_bytes, cudaMemcpyDefault));
	return 0;
}

extern "C" int32_t ear_b200_scene_create_from_image(const void* src_device, uint64_t bytes, int32_t device, ear_b200_scene** out) {
	if (!out) return fail("scene_create_from_image: out is null");
	*out = nullptr;
	if (!src_device || bytes < kImageAlign) return fail("scene_create_from_image: no image");
	if (int32_t rc = pick_device(device, "scene_create_from_image")) return rc;
	ImageHeader h{};
	CUDA_TRY(cudaMemcpy(&h, src_device, sizeof(h), cudaMemcpyDefault));
	ImageHeader want = h;
	if (h.magic != kImageMagic || h.version != (uint32_t)EAR_B200_ABI_VERSION)
		return fail("scene_create_from_image: not a scene image of this library version");
	if (h.n_tris < 0 || h.n_nodes < 1 || h.n_materials <= 0 || h.n_bands <= 0 || h.n_bands > EAR_B200_MAX_BANDS)
		return fail("scene_create_from_image: corrupt header");
	image_layout(want);
	if (want.off_nodes != h.off_nodes || want.off_tris != h.off_tris || want.off_materials != h.off_materials || want.bytes != h.bytes ||
	    h.bytes > bytes)
		return fail("scene_create_from_image: image size does not match its header");
	std::unique_ptr<ear_b200_scene, void (*)(ear_b200_scene*)> s(new ear_b200_scene(), ear_b200_scene_destroy);
	s->device = device;
	CUDA_TRY(dev_alloc(&s->d_image, (size_t)h.bytes));
	CUDA_TRY(cudaMemcpy(s->d_image, src_device, (size_t)h.bytes, cudaMemcpyDefault));
	s->materials.resize((size_t)h.n_materials * h.n_bands * 4);
	CUDA_TRY(cudaMemcpy(s->materials.data(), s->d_image + h.off_materials, s->materials.size() * sizeof(float), cudaMemcpyDeviceToHost));
	if (int32_t rc = scene_finish(s.get(), h)) return rc;
	*out = s.release();
	return 0;
}

extern "C" int32_t ear_b200_scene_clone(ear_b200_scene* s, int32_t device, ear_b200_scene** out) {
	if (!s) return fail("scene_clone: null scene");
	return ear_b200_scene_create_from_image(s->d_image, (uint64_t)s->image_bytes, device, out);
}

extern "C" void ear_b200_scene_destroy(ear_b200_scene* s) {
	if (!s) return;
	cudaSetDevice(s->device);
	cudaDeviceSynchronize();   // released blocks go back to the cache: nothing of this scene may still be in flight
	dev_free(s->d_image); dev_free(s->d_spill); dev_free(s->d_emitters);
	dev_free(s->d_ctx); dev_free(s->d_rec); dev_free(s->d_prefix); dev_free(s->d_queue); dev_free(s->d_counter_parts); dev_free(s->d_ctx_area);
	dev_free(s->pool.ro); dev_free(s->pool.rd); dev_free(s->pool.rm); dev_free(s->pool.hit);
	dev_free(s->pool.sh0); dev_free(s->pool.sh1); dev_free(s->pool.sh2); dev_free(s->pool.trav_list);
	dev_free(s->pool.q_list); dev_free(s->pool.vis_list); dev_free(s->pool.counts);
	dev_free(s->pool.trav_tmp); dev_free(s->pool.q_tmp); dev_free(s->pool.bins); dev_free(s->pool.ctx_log2af);
	dev_free(s->pool.pair_count); dev_free(s->pool.pair_base); dev_free(s->pool.trav_rank); dev_free(s->pool.q_rank);
	dev_free(s->d_scratch_counters);
	if (s->h_counts) cudaFreeHost(s->h_counts);
	for (auto& e : s->ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
	for (auto& m : s->vismaps) { dev_free(m.d_offsets); dev_free(m.d_items); }
	dev_free(s->d_maps); dev_free(s->d_map_of); dev_free(s->d_q_bvh); dev_free(s->d_vis_counts); dev_free(s->d_vis_sums); dev_free(s->d_post);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
}

static int32_t ensure_pool(ear_b200_scene* s, size_t slots, size_t queries);
static int32_t prepare_vismaps(ear_b200_scene* s, int n_ctx, int n_rec, cudaStream_t stream, int* n_mapped);

extern "C" int32_t ear_b200_first_hit(ear_b200_scene* s, const float* origins, const float* dirs, int64_t n,
                                      int32_t* tri_index, float* t) {
	if (!s) return fail("first_hit: null scene");
	if (n <= 0) return 0;
	CUDA_TRY(cudaSetDevice(s->device));
	const int64_t chunk = std::min<int64_t>(n, s->max_slots);
	DevBuf<float> d_o, d_d, d_t;
	DevBuf<int32_t> d_i;
	CUDA_TRY(d_o.alloc((size_t)chunk * 3)); CUDA_TRY(d_d.alloc((size_t)chunk * 3));
	CUDA_TRY(d_t.alloc((size_t)chunk)); CUDA_TRY(d_i.alloc((size_t)chunk));
	if (int32_t rc = ensure_pool(s, (size_t)chunk, 1)) return rc;
	RenderParams p{};
	for (int64_t at = 0; at < n; at += chunk) {
		const int m = (int)std::min<int64_t>(chunk, n - at);
		const unsigned grid = (unsigned)((m + kBlock - 1) / kBlock);
		CUDA_TRY(cudaMemcpyAsync(d_o, origins + 3 * at, (size_t)m * 12, cudaMemcpyHostToDevice, s->stream));
		CUDA_TRY(cudaMemcpyAsync(d_d, dirs + 3 * at, (size_t)m * 12, cudaMemcpyHostToDevice, s->stream));
		{
			// the production closest-hit kernel (persistent, dynamic fetch) over an explicit ray list
			WfPool pl = s->pool;
			wf_load_rays_kernel<<<grid, kBlock, 0, s->stream>>>(pl, d_o, d_d, m);
			int bps = 0;
			void (*closest)(SceneDev, WfPool, RenderParams) = s->dev.exact ? wf_traverse_kernel<false, true> : wf_traverse_kernel<false, false>;
			CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, closest, kBlock, kStackBytes));
			closest<<<s->sm_count * std::max(1, bps), kBlock, kStackBytes, s->stream>>>(s->dev, pl, p);
			wf_store_hits_kernel<<<grid, kBlock, 0, s->stream>>>(s->dev, pl, m, d_i, d_t);
		}
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(tri_index + at, d_i, (size_t)m * 4, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaMemcpyAsync(t + at, d_t, (size_t)m * 4, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
	}
	return 0;
}

extern "C" int32_t ear_b200_occluded(ear_b200_scene* s, const float* p_in, const float* x, int64_t n, uint8_t* out) {
	if (!s) return fail("occluded: null scene");
	if (n <= 0) return 0;
	CUDA_TRY(cudaSetDevice(s->device));
	const int64_t chunk = std::min<int64_t>(n, s->max_slots);
	DevBuf<float> d_p, d_x;
	DevBuf<uint8_t> d_out;
	DevBuf<float4> d_qx;
	CUDA_TRY(d_p.alloc((size_t)chunk * 3)); CUDA_TRY(d_x.alloc((size_t)chunk * 3)); CUDA_TRY(d_out.alloc((size_t)chunk));
	CUDA_TRY(d_qx.alloc((size_t)chunk));
	if (int32_t rc = ensure_pool(s, (size_t)chunk, (size_t)chunk)) return rc;
	RenderParams p{};
	// all segments end at one point (the render loop's case): answer through that point's visibility map,
	// exactly as the render does, with the BVH any-hit kernel for the texels whose lists are too long
	bool same_x = true;
	for (int64_t i = 1; i < n && same_x; ++i) same_x = std::memcmp(x, x + 3 * i, 12) == 0;
	int n_mapped = 0;
	if (same_x) {
		ear_b200_recorder one{};
		one.kind = EAR_B200_MONO;
		std::memcpy(one.position, x, 12);
		s->h_rec.assign(1, one);
		if (s->rec_cap < 1) { dev_free(s->d_rec); CUDA_TRY(dev_alloc(&s->d_rec, sizeof(ear_b200_recorder))); s->rec_cap = 1; }
		CUDA_TRY(cudaMemcpyAsync(s->d_rec, &one, sizeof(one), cudaMemcpyHostToDevice, s->stream));
		if (int32_t rc = prepare_vismaps(s, 1, 1, s->stream, &n_mapped)) return rc;
		if (n_mapped && (size_t)chunk > s->q_bvh_cap) {
			dev_free(s->d_q_bvh);
			CUDA_TRY(dev_alloc(&s->d_q_bvh, (size_t)chunk * sizeof(uint2)));
			s->q_bvh_cap = (size_t)chunk;
		}
		p.rec = s->d_rec; p.n_rec = 1; p.n_ctx = 1;
	}
	for (int64_t at = 0; at < n; at += chunk) {
		const int m = (int)std::min<int64_t>(chunk, n - at);
		const unsigned grid = (unsigned)((m + kBlock - 1) / kBlock);
		CUDA_TRY(cudaMemcpyAsync(d_p, p_in + 3 * at, (size_t)m * 12, cudaMemcpyHostToDevice, s->stream));
		CUDA_TRY(cudaMemcpyAsync(d_x, x + 3 * at, (size_t)m * 12, cudaMemcpyHostToDevice, s->stream));
		{
			WfPool pl = s->pool;
			pl.qx = d_qx;
			wf_load_segments_kernel<<<grid, kBlock, 0, s->stream>>>(pl, d_qx, d_p, d_x, m, d_out);
			int bps = 0;
			void (*anyhit)(SceneDev, WfPool, RenderParams) = s->dev.exact ? wf_traverse_kernel<true, true> : wf_traverse_kernel<true, false>;
			CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, anyhit, kBlock, kStackBytes));
			if (n_mapped) {
				WfPool pl_fb = pl;
				pl_fb.qx = nullptr; pl_fb.q_list = s->d_q_bvh; pl_fb.q_count_idx = 5; pl_fb.q_cursor_idx = 6;
				wf_vismap_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>(s->dev, pl, p, s->d_maps, s->d_map_of, s->d_q_bvh);
				anyhit<<<s->sm_count * std::max(1, bps), kBlock, kStackBytes, s->stream>>>(s->dev, pl_fb, p);
			} else anyhit<<<s->sm_count * std::max(1, bps), kBlock, kStackBytes, s->stream>>>(s->dev, pl, p);
			wf_mark_visible_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>(pl, d_out);
		}
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(out + at, d_out, (size_t)m, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
	}
	return 0;
}

// bins per track: the longest path the bounce loop can produce, in samples, plus the widest ramp
static int32_t default_bins(const ear_b200_scene* s, int32_t max_bounces) {
	if (max_bounces <= 0) max_bounces = 1000;
	// energy bound: a ray dies once intensity < 1e-8 (src/Scene.cpp:275); kept <= k_max per bounce
	float k_max = 0.0f;
	for (size_t i = 0; i < s->materials.size() / 4; ++i) k_max = std::max(k_max, s->materials[4 * i + 2]);
	double bounces = max_bounces;
	if (k_max > 0.0f && k_max < 1.0f) bounces = std::min(bounces, std::ceil(std::log(1e-8) / std::log((double)k_max)) + 2.0);
	const double reach = 2.0 * (double)s->diagonal + 1.0;  // source / recorder may sit outside the mesh bounds
	const double path = (bounces + 1.0) * reach;
	double bins = path / 343.0 * 44100.0 + std::sqrt(path) + 64.0;
	bins = std::min(bins, 64.0 * 1024.0 * 1024.0);
	bins = std::max(bins, 3.0 * 44100.0);               // FloatBuffer starts at 3 s (src/Recorder.h:36)
	return (int32_t)bins;
}
extern "C" int32_t ear_b200_tracks_per_recorder(const ear_b200_recorder* rec, int32_t n) {
	for (int32_t i = 0; rec && i < n; ++i) if (rec[i].kind == EAR_B200_STEREO) return 2;
	return 1;
}
extern "C" int32_t ear_b200_default_bins(ear_b200_scene* s, const ear_b200_options* opt) {
	if (!s) return 0;
	if (opt && opt->n_bins > 0) return opt->n_bins;
	return default_bins(s, opt ? opt->max_bounces : 1000);
}

static int32_t upload_params(ear_b200_scene* s, const ear_b200_context* ctx, int32_t n_ctx, const ear_b200_recorder* rec,
                             int32_t n_rec, const ear_b200_options* opt, cudaStream_t stream, RenderParams& p,
                             const long long* prefix_override = nullptr) {
	if (n_ctx <= 0 || n_rec <= 0 || !ctx || !rec) return fail("render: need at least one context and one recorder");
	for (int32_t c = 0; c < n_ctx; ++c) {
		if (ctx[c].band < 0 || ctx[c].band >= s->n_bands) return fail("render: context band outside the material table");
		if (ctx[c].num_samples < 0) return fail("render: negative num_samples");
		if (ctx[c].stream_id < 0) return fail("render: negative stream_id");
		if (ctx[c].source_kind != EAR_B200_POINT_SOURCE && ctx[c].source_kind != EAR_B200_MESH_SOURCE) return fail("render: unknown source kind");
		if (ctx[c].source_kind == EAR_B200_MESH_SOURCE &&
		    (ctx[c].emitter_first < 0 || ctx[c].emitter_count < 0 || (long long)ctx[c].emitter_first + ctx[c].emitter_count > s->n_emitters))
			return fail("render: a mesh source names triangles outside the emitter table (ear_b200_scene_set_emitters)");
	}
	for (int32_t i = 0; i < n_ctx * n_rec; ++i)
		if (rec[i].kind != EAR_B200_MONO && rec[i].kind != EAR_B200_STEREO) return fail("render: unknown recorder kind");
	if ((size_t)n_ctx > s->ctx_cap) {
		dev_free(s->d_ctx); dev_free(s->d_prefix); dev_free(s->d_ctx_area);
		s->d_ctx = nullptr; s->d_prefix = nullptr; s->d_ctx_area = nullptr; s->ctx_cap = 0;
		CUDA_TRY(dev_alloc(&s->d_ctx, sizeof(ear_b200_context) * n_ctx));
		CUDA_TRY(dev_alloc(&s->d_prefix, sizeof(long long) * (n_ctx + 1)));
		CUDA_TRY(dev_alloc(&s->d_ctx_area, sizeof(float) * n_ctx));
		s->ctx_cap = n_ctx;
	}
	// Mesh::total_area of every mesh source: areas added one by one in file order, in float (src/Mesh.cpp:90)
	std::vector<float> emit_area((size_t)n_ctx, 0.0f);
	for (int32_t c = 0; c < n_ctx; ++c)
		if (ctx[c].source_kind == EAR_B200_MESH_SOURCE) {
			float total = 0.0f;
			for (int32_t i = 0; i < ctx[c].emitter_count; ++i) total += s->emitter_area[(size_t)ctx[c].emitter_first + i];
			emit_area[(size_t)c] = total;
		}
	if ((size_t)n_ctx * n_rec > s->rec_cap) {
		dev_free(s->d_rec);
		CUDA_TRY(dev_alloc(&s->d_rec, sizeof(ear_b200_recorder) * (size_t)n_ctx * n_rec));
		s->rec_cap = (size_t)n_ctx * n_rec;
	}
	std::vector<long long> prefix(n_ctx + 1, 0);
	const long long first = opt ? opt->first_ray : 0;
	for (int32_t c = 0; c < n_ctx; ++c) {
		long long cnt = (opt && opt->ray_count >= 0) ? opt->ray_count : (long long)ctx[c].num_samples - first;
		cnt = std::max<long long>(0, std::min<long long>(cnt, (long long)ctx[c].num_samples - first));
		prefix[c + 1] = prefix[c] + cnt;
	}
	if (prefix_override) for (int32_t c = 0; c <= n_ctx; ++c) prefix[c] = prefix_override[c];
	s->h_rec.assign(rec, rec + (size_t)n_ctx * n_rec);
	CUDA_TRY(cudaMemcpyAsync(s->d_ctx, ctx, sizeof(ear_b200_context) * n_ctx, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s->d_rec, rec, sizeof(ear_b200_recorder) * (size_t)n_ctx * n_rec, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s->d_prefix, prefix.data(), sizeof(long long) * (n_ctx + 1), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s->d_ctx_area, emit_area.data(), sizeof(float) * n_ctx, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));  // `prefix` is a stack temporary
	p.ctx = s->d_ctx; p.rec = s->d_rec; p.work_prefix = s->d_prefix; p.ctx_emit_area = s->d_ctx_area;
	p.n_ctx = n_ctx; p.n_rec = n_rec;
	p.tpr = ear_b200_tracks_per_recorder(rec, n_ctx * n_rec);
	p.max_bounces = (opt && opt->max_bounces > 0) ? opt->max_bounces : 1000;
	// the pool keeps a ray's bounce number in 16 bits (wavefront.cuh, rm.z)
	if (p.max_bounces > 65535) return fail("render: max_bounces above 65535 is not supported");
	p.seed = opt ? opt->seed : 1;
	p.first_ray = first;
	p.total_work = prefix[n_ctx];
	p.next_work = s->d_queue;
	p.counter_parts = s->d_counter_parts;
	return 0;
}

static void harvest_events(ear_b200_scene* s) {
	for (size_t i = 0; i < s->ev_used; ++i) {
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, s->ev_pool[i].a, s->ev_pool[i].b) == cudaSuccess) s->stats.ms[s->ev_pool[i].cls] += ms;
	}
	s->ev_used = 0;
}
struct LaunchTimer {   // records an event pair around one launch
	ear_b200_scene* s; cudaStream_t stream; size_t slot;
	LaunchTimer(ear_b200_scene* sc, cudaStream_t st, int cls) : s(sc), stream(st) {
		if (s->ev_used == s->ev_pool.size()) {
			ear_b200_scene::Timed t; cudaEventCreate(&t.a); cudaEventCreate(&t.b); t.cls = cls; s->ev_pool.push_back(t);
		}
		slot = s->ev_used++;
		s->ev_pool[slot].cls = cls;
		++s->stats.launches[cls];
		s->last_stream = stream;
		cudaEventRecord(s->ev_pool[slot].a, stream);
	}
	~LaunchTimer() { cudaEventRecord(s->ev_pool[slot].b, stream); }
};

static int32_t ensure_pool(ear_b200_scene* s, size_t slots, size_t queries) {
	WfPool& pl = s->pool;
	if (slots > s->pool_slots) {
		dev_free(pl.ro); dev_free(pl.rd); dev_free(pl.rm); dev_free(pl.hit); dev_free(pl.sh0); dev_free(pl.sh1); dev_free(pl.sh2);
		dev_free(pl.trav_list); dev_free(pl.trav_tmp); dev_free(pl.trav_rank);
		CUDA_TRY(dev_alloc(&pl.ro, slots * sizeof(float4))); CUDA_TRY(dev_alloc(&pl.rd, slots * sizeof(float4)));
		CUDA_TRY(dev_alloc(&pl.rm, slots * sizeof(uint4))); CUDA_TRY(dev_alloc(&pl.hit, slots * sizeof(int2)));
		CUDA_TRY(dev_alloc(&pl.sh0, slots * sizeof(float4))); CUDA_TRY(dev_alloc(&pl.sh1, slots * sizeof(float4)));
		CUDA_TRY(dev_alloc(&pl.sh2, slots * sizeof(float4))); CUDA_TRY(dev_alloc(&pl.trav_list, slots * sizeof(int)));
		CUDA_TRY(dev_alloc(&pl.trav_tmp, slots * sizeof(uint2)));
		CUDA_TRY(dev_alloc(&pl.trav_rank, slots * sizeof(int)));
		s->pool_slots = slots;
	}
	if (queries > s->pool_queries) {
		dev_free(pl.q_list); dev_free(pl.vis_list); dev_free(pl.q_tmp); dev_free(pl.q_rank);
		CUDA_TRY(dev_alloc(&pl.q_list, queries * sizeof(uint2))); CUDA_TRY(dev_alloc(&pl.vis_list, queries * sizeof(uint2)));
		CUDA_TRY(dev_alloc(&pl.q_tmp, queries * sizeof(uint2)));
		CUDA_TRY(dev_alloc(&pl.q_rank, queries * sizeof(int)));
		s->pool_queries = queries;
	}
	if (!pl.counts) CUDA_TRY(dev_alloc(&pl.counts, 8 * sizeof(int)));
	if (!pl.pair_count) {
		CUDA_TRY(dev_alloc(&pl.pair_count, kPrivMaxPairs * sizeof(int)));
		CUDA_TRY(dev_alloc(&pl.pair_base, (kPrivMaxPairs + 1) * sizeof(int)));
	}
	pl.vis_sorted = pl.q_tmp;   // the pre-binning query list is dead once the queries are scattered
	pl.q_count_idx = 1; pl.q_cursor_idx = 4;
	pl.slot_bits = kMaxSlotBits;   // harness calls: one implicit recorder
	if (!pl.bins) CUDA_TRY(dev_alloc(&pl.bins, kBinsTotal * sizeof(int)));
	pl.ray_key = s->ray_key;
	pl.sort_queries = s->sort_queries != 0 ? 1 : 0;   // the render loop decides per call when the setting is -1
	for (int k = 0; k < 3; ++k) {
		pl.cell_origin[k] = s->lo[k];
		const float ext = s->hi[k] - s->lo[k];
		pl.cell_scale[k] = ext > 0.0f ? 16.0f / ext : 0.0f;
	}
	return 0;
}

// Builds (or finds) the visibility map of one recorder position; returns its index or -1 when maps are disabled.
static int32_t build_vismap_sorted(ear_b200_scene* s, const float x[3], int res, cudaStream_t stream, int* index);

static int32_t get_vismap(ear_b200_scene* s, const float x[3], cudaStream_t stream, int* index) {
	*index = -1;
	int res = s->vismap_res;
	if (res == 0 || s->n_tris == 0) return 0;
	if (res < 0) {
		res = 64;
		while (res < 1024 && (double)res * res < (double)s->n_tris) res *= 2;   // ~6 texels per triangle
	}
	for (size_t i = 0; i < s->vismaps.size(); ++i)
		if (s->vismaps[i].res == res && std::memcmp(s->vismaps[i].x, x, 12) == 0) { *index = (int)i; return 0; }
	if (s->vismaps.size() >= 64 || s->vis_budget_spent) return 0;   // the remaining recorder positions use the BVH
	if (s->vismap_build == 0) return build_vismap_sorted(s, x, res, stream, index);
	ear_b200_scene::VisMapHost m{};
	std::memcpy(m.x, x, 12); m.res = res;
	const int n_tex = 6 * res * res;
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!dbg) return;
		cudaDeviceSynchronize();
		const auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[ear_b200] vismap: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	// scratch (texel counters + block sums) is kept with the scene: every map of a scene has the same size
	if ((size_t)n_tex > s->vis_scratch_cap) {
		dev_free(s->d_vis_counts); dev_free(s->d_vis_sums);
		s->d_vis_counts = nullptr; s->d_vis_sums = nullptr; s->vis_scratch_cap = 0;
		CUDA_TRY(dev_alloc(&s->d_vis_counts, (size_t)n_tex * sizeof(int)));
		CUDA_TRY(dev_alloc(&s->d_vis_sums, kVisScanBlocks * sizeof(long long)));
		s->vis_scratch_cap = (size_t)n_tex;
	}
	int* d_counts = s->d_vis_counts;
	CUDA_TRY(dev_alloc(&m.d_offsets, ((size_t)n_tex + 1) * sizeof(int)));
	struct MapGuard {   // the two allocations of a map are released on every error return until the map is adopted
		ear_b200_scene::VisMapHost* m;
		~MapGuard() { if (m) { dev_free(m->d_offsets); dev_free(m->d_items); } }
	} guard{&m};
	CUDA_TRY(cudaMemsetAsync(d_counts, 0, (size_t)n_tex * sizeof(int), stream));
	lap("alloc + clear");
	const double reach = 2.0 * (double)s->diagonal + 1.0;
	int id_bits = 7;   // thread ids cover [0, 2^id_bits) >= 6 * n_tris, visited in bit-reversed order
	while ((1LL << id_bits) < 6LL * s->n_tris) ++id_bits;
	const unsigned grid = (unsigned)((1LL << id_bits) / 128);
	const int cap = s->dev.vis_cap;
	const int per = (n_tex + kVisScanBlocks - 1) / kVisScanBlocks;
	vis_build_kernel<0><<<grid, 128, 0, stream>>>(s->dev, x[0], x[1], x[2], res, reach, s->maxabs, d_counts, nullptr, nullptr, id_bits);
	lap("count pass");
	vis_scan_sums_kernel<<<kVisScanBlocks, 1024, 0, stream>>>(d_counts, s->d_vis_sums, n_tex, per, cap);
	vis_scan_top_kernel<<<1, kVisScanBlocks, 0, stream>>>(s->d_vis_sums, m.d_offsets, n_tex);
	vis_scan_offsets_kernel<<<kVisScanBlocks, 1024, 0, stream>>>(d_counts, s->d_vis_sums, m.d_offsets, n_tex, per, cap, cap);
	int total = 0;
	CUDA_TRY(cudaMemcpyAsync(&total, m.d_offsets + n_tex, sizeof(int), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	lap("scan");
	if (dbg) {
		std::vector<int> hc((size_t)n_tex);
		cudaMemcpy(hc.data(), d_counts, (size_t)n_tex * sizeof(int), cudaMemcpyDeviceToHost);
		long long sum = 0, over = 0, over_sum = 0; int mx = 0;
		for (int v : hc) { sum += v; mx = std::max(mx, v); if (v > cap) { ++over; over_sum += v; } }
		std::fprintf(stderr, "[ear_b200] vismap: %lld entries, longest list %d, %lld texels over the cap hold %lld entries, stored %d\n", sum, mx, over, over_sum, total);
		t_prev = std::chrono::steady_clock::now();
	}
	// memory guard: all maps of a scene together may take a quarter of the device memory that is free when the first
	// one is built (a 1e7-triangle scene with 64 recorders would otherwise ask for hundreds of GB); beyond that, and
	// when a single map's entries overflow 31 bits, the recorder's queries walk the BVH instead
	if (s->vis_budget == 0) {
		size_t free_b = 0, total_b = 0;
		CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
		s->vis_budget = std::max<size_t>(free_b / 4, 1);
		if (const char* vb = std::getenv("EAR_B200_VISMAP_BUDGET")) s->vis_budget = std::max<size_t>((size_t)std::atof(vb), 1);   // bytes (tests: force the BVH route)
	}
	const size_t map_bytes = ((size_t)n_tex + 1) * sizeof(int) + (total < 0 ? 0 : (size_t)total * sizeof(int));
	if (total < 0 || s->vis_bytes + map_bytes > s->vis_budget) {
		s->vis_budget_spent = true;
		return 0;   // the guard frees the offsets
	}
	s->vis_bytes += map_bytes;
	m.n_items = (size_t)total;
	CUDA_TRY(dev_alloc(&m.d_items, std::max<size_t>(m.n_items, 1) * sizeof(int)));
	CUDA_TRY(cudaMemsetAsync(d_counts, 0, (size_t)n_tex * sizeof(int), stream));
	vis_build_kernel<1><<<grid, 128, 0, stream>>>(s->dev, x[0], x[1], x[2], res, reach, s->maxabs, d_counts, m.d_offsets, m.d_items, id_bits);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaStreamSynchronize(stream));
	lap("fill pass");
	s->vismaps.push_back(m);
	guard.m = nullptr;
	*index = (int)s->vismaps.size() - 1;
	return 0;
}

// Sort-based build of one visibility map (vis_emit_kernel + one radix sort): no returning atomics, lists nearest-first.
static int32_t build_vismap_sorted(ear_b200_scene* s, const float x[3], int res, cudaStream_t stream, int* index) {
	ear_b200_scene::VisMapHost m{};
	std::memcpy(m.x, x, 12); m.res = res;
	const int n_tex = 6 * res * res;
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!dbg) return;
		cudaDeviceSynchronize();
		const auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[ear_b200] vismap (sort build): %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	CUDA_TRY(dev_alloc(&m.d_offsets, ((size_t)n_tex + 1) * sizeof(int)));
	struct MapGuard {
		ear_b200_scene::VisMapHost* m;
		~MapGuard() { if (m) { dev_free(m->d_offsets); dev_free(m->d_items); } }
	} guard{&m};
	int id_bits = 7;
	while ((1LL << id_bits) < 6LL * s->n_tris) ++id_bits;
	const size_t n_ids = (size_t)1 << id_bits;
	const unsigned grid = (unsigned)(n_ids / 128);
	DevBuf<int> d_pair_count, d_pair_base, d_pair_sums;
	DevBuf<unsigned long long> d_total;
	CUDA_TRY(d_pair_count.alloc(n_ids)); CUDA_TRY(d_pair_base.alloc(n_ids + 1));
	CUDA_TRY(d_pair_sums.alloc(n_ids / (dbvh::kScanBlock * dbvh::kScanPer) + 2));
	CUDA_TRY(d_total.alloc(1));
	CUDA_TRY(cudaMemsetAsync(d_total.p, 0, sizeof(unsigned long long), stream));
	const double reach = 2.0 * (double)s->diagonal + 1.0;
	int texel_bits = 1;
	while ((1 << texel_bits) < n_tex) ++texel_bits;
	const int dist_bits = 32 - texel_bits;
	const float dist_scale = (float)((double)(1u << dist_bits) / reach);
	const int cap = s->dev.vis_cap;
	lap("alloc + clear");
	vis_emit_kernel<0><<<grid, 128, 0, stream>>>(s->dev, x[0], x[1], x[2], res, reach, s->maxabs, id_bits, d_pair_count, nullptr, nullptr, nullptr, dist_bits, dist_scale);
	lap("count pass");
	vis_pair_total_kernel<<<s->sm_count * 4, 256, 0, stream>>>(d_pair_count, n_ids, d_total);
	dbvh::exclusive_scan(d_pair_count, (int)n_ids, d_pair_base, d_pair_sums, stream);
	unsigned long long total64 = 0;
	CUDA_TRY(cudaMemcpyAsync(&total64, d_total.p, sizeof(total64), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	lap("scan");
	if (s->vis_budget == 0) {
		size_t free_b = 0, total_b = 0;
		CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
		s->vis_budget = std::max<size_t>(free_b / 4, 1);
		if (const char* vb = std::getenv("EAR_B200_VISMAP_BUDGET")) s->vis_budget = std::max<size_t>((size_t)std::atof(vb), 1);
	}
	const bool too_many = total64 >= 0x7fffffffull;
	const int total = too_many ? 0 : (int)total64;
	const size_t map_bytes = ((size_t)n_tex + 1) * sizeof(int) + (size_t)total * sizeof(int);
	if (too_many || s->vis_bytes + map_bytes > s->vis_budget) { s->vis_budget_spent = true; return 0; }
	s->vis_bytes += map_bytes;
	m.n_items = (size_t)total;
	CUDA_TRY(dev_alloc(&m.d_items, std::max<size_t>(m.n_items, 1) * sizeof(int)));
	if (total > 0) {
		DevBuf<uint32_t> d_keys_a, d_keys_b;
		DevBuf<int> d_vals_a;
		DevBuf<unsigned char> d_tmp;
		CUDA_TRY(d_keys_a.alloc((size_t)total)); CUDA_TRY(d_keys_b.alloc((size_t)total)); CUDA_TRY(d_vals_a.alloc((size_t)total));
		vis_emit_kernel<1><<<grid, 128, 0, stream>>>(s->dev, x[0], x[1], x[2], res, reach, s->maxabs, id_bits, d_pair_count, d_pair_base, d_keys_a, d_vals_a, dist_bits, dist_scale);
		CUDA_TRY(cudaGetLastError());
		lap("emit pass");
		size_t tmp_bytes = 0;
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys_a.p, d_keys_b.p, d_vals_a.p, m.d_items, total, 0, 32, stream));
		CUDA_TRY(d_tmp.alloc(tmp_bytes));
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys_a.p, d_keys_b.p, d_vals_a.p, m.d_items, total, 0, 32, stream));
		lap("radix sort");
		vis_offsets_from_keys_kernel<<<s->sm_count * 8, 256, 0, stream>>>(d_keys_b, total, dist_bits, n_tex, m.d_offsets);
		vis_flag_overlong_kernel<<<(n_tex + 255) / 256, 256, 0, stream>>>(m.d_offsets, n_tex, cap);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaStreamSynchronize(stream));   // the key buffers go back to the cache when this scope ends
		lap("offsets");
	} else {
		CUDA_TRY(cudaMemsetAsync(m.d_offsets, 0, ((size_t)n_tex + 1) * sizeof(int), stream));
	}
	if (dbg) std::fprintf(stderr, "[ear_b200] vismap (sort build): %d entries, %d distance bits\n", total, dist_bits);
	m.sorted = true;
	s->vismaps.push_back(m);
	guard.m = nullptr;
	*index = (int)s->vismaps.size() - 1;
	return 0;
}

// orders the texel lists of every map of the scene by distance from its recorder (vis_sort_kernel), once
static int32_t sort_vismaps(ear_b200_scene* s, cudaStream_t stream) {
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	if (dbg) cudaStreamSynchronize(stream);
	int n_sorted = 0;
	for (auto& m : s->vismaps) {
		if (m.sorted) continue;
		const int n_tex = 6 * m.res * m.res;
		LaunchTimer t(s, stream, 7);
		vis_sort_kernel<<<s->sm_count * 8, 256, 0, stream>>>(s->dev, m.x[0], m.x[1], m.x[2], m.d_offsets, m.d_items, n_tex);
		m.sorted = true;
		++n_sorted;
	}
	CUDA_TRY(cudaGetLastError());
	if (dbg && n_sorted) {
		cudaStreamSynchronize(stream);
		std::fprintf(stderr, "[ear_b200] vismap: %d map(s) ordered by distance in %.2f ms\n", n_sorted, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
	return 0;
}

// uploads the map table for the recorders of the current call; returns the number of recorders that have a map
static int32_t prepare_vismaps(ear_b200_scene* s, int n_ctx, int n_rec, cudaStream_t stream, int* n_mapped) {
	*n_mapped = 0;
	const size_t n = (size_t)n_ctx * n_rec;
	if (n == 0 || s->h_rec.size() < n) return 0;
	std::vector<int> map_of(n, -1);
	for (size_t i = 0; i < n; ++i) {
		int idx = -1;
		if (int32_t rc = get_vismap(s, s->h_rec[i].position, stream, &idx)) return rc;
		map_of[i] = idx;
		if (idx >= 0) ++*n_mapped;
	}
	if (*n_mapped == 0) return 0;
	if (n > s->map_of_cap) { dev_free(s->d_map_of); CUDA_TRY(dev_alloc(&s->d_map_of, n * sizeof(int))); s->map_of_cap = n; }
	if (!s->d_maps) CUDA_TRY(dev_alloc(&s->d_maps, 64 * sizeof(VisMapDev)));
	std::vector<VisMapDev> maps(s->vismaps.size());
	for (size_t i = 0; i < maps.size(); ++i) {
		maps[i].offsets = s->vismaps[i].d_offsets; maps[i].items = s->vismaps[i].d_items; maps[i].res = s->vismaps[i].res;
		std::memcpy(maps[i].x, s->vismaps[i].x, 12);
	}
	CUDA_TRY(cudaMemcpyAsync(s->d_map_of, map_of.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s->d_maps, maps.data(), maps.size() * sizeof(VisMapDev), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return 0;
}

// Wavefront engine: shade -> closest -> any-hit -> splat per iteration until no ray is left.
// Asynchronous except for one tiny host read every `check_every` iterations (the loop must know when to stop).
static int32_t launch_wavefront(ear_b200_scene* s, RenderParams& p, cudaStream_t stream) {
	if (p.n_rec > 255) return fail("render: more than 255 recorders per context are not supported");
	if (p.n_ctx > 65535) return fail("render: more than 65535 contexts per call are not supported");
	// pool size: large launches amortise the persistent kernels' tails and the ~9 launches of an iteration (8 Mi slots:
	// +4 % over 4 Mi, 16 Mi: +3 % more).  With a short bounce cap nearly every ray lives for exactly max_bounces
	// iterations, the pool turns over in lockstep and one generation is best (8 GPUs x 1.25e7 work items each ran
	// 200 iterations of 3.1 M rays under the old work/4 rule and were 10 % slower per segment than one GPU).  With
	// long paths rays die at random times; the shade kernel walks every slot, so several generations keep it full.
	long long slots = std::min<long long>(std::max<long long>(p.total_work, 1), s->max_slots);
	const long long generations = (s->pool_generations > 0) ? s->pool_generations : (p.max_bounces <= 100 ? 1 : 4);
	if (!s->slots_forced) slots = std::min<long long>(slots, std::max<long long>(1 << 18, p.total_work / generations));
	// a query word holds slot | recorder << slot_bits
	int rec_bits = 0;
	while ((1 << rec_bits) < p.n_rec) ++rec_bits;
	const int slot_bits = std::min(kMaxSlotBits, 32 - rec_bits);
	slots = std::min<long long>(slots, 1LL << slot_bits);
	// the query lists are indexed with 32-bit ints and cost 24 bytes per (slot, recorder): at most 2^29 queries in flight
	slots = std::min<long long>(slots, std::max<long long>(256, (1LL << 29) / std::max(1, p.n_rec)));
	slots = (slots + 255) / 256 * 256;
	if (int32_t rc = ensure_pool(s, (size_t)slots, (size_t)slots * std::max(1, p.n_rec))) return rc;
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	auto t_a = std::chrono::steady_clock::now();
	if (dbg) cudaDeviceSynchronize();
	int n_mapped = 0;
	if (p.n_rec > 0) { if (int32_t rc = prepare_vismaps(s, p.n_ctx, p.n_rec, stream, &n_mapped)) return rc; }
	if (n_mapped && s->vismap_sort == 1) { if (int32_t rc = sort_vismaps(s, stream)) return rc; }
	if (dbg) { cudaDeviceSynchronize(); std::fprintf(stderr, "[ear_b200] pool + visibility maps: %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count()); }
	const size_t n_queries = (size_t)slots * std::max(1, p.n_rec);
	if (n_mapped && n_queries > s->q_bvh_cap) {
		dev_free(s->d_q_bvh);
		CUDA_TRY(dev_alloc(&s->d_q_bvh, n_queries * sizeof(uint2)));
		s->q_bvh_cap = n_queries;
	}
	WfPool pl = s->pool;
	pl.n_slots = (int)slots;
	pl.slot_bits = slot_bits;
	pl.q_count_idx = 1; pl.q_cursor_idx = 4;
	// Query order.  Sorted by (recorder, cell of the hit point): neighbouring lookups read the same texel lists.  Unsorted:
	// the shade kernel's order, a warp's queries for one recorder side by side.  One recorder (C4, per 4e7 rays): sorted
	// 1027 ms, unsorted 1056.  64 recorders (C5, per 3e6 rays): sorted 1012 ms (sort 198, rank atomics in the shade kernel
	// 140), unsorted 796 -- lookups 466 -> 620, but shade 175 -> 36, sort 198 -> 2, splat 99 -> 66.
	if (s->sort_queries < 0) pl.sort_queries = p.n_rec <= 4 ? 1 : 0;
	CUDA_TRY(cudaMemsetAsync(pl.rm, 0, (size_t)slots * sizeof(uint4), stream));
	if ((size_t)p.n_ctx > s->log2af_cap) {
		dev_free(s->pool.ctx_log2af);
		s->pool.ctx_log2af = nullptr; s->log2af_cap = 0;
		CUDA_TRY(dev_alloc(&s->pool.ctx_log2af, sizeof(double) * p.n_ctx));
		s->log2af_cap = p.n_ctx;
		pl.ctx_log2af = s->pool.ctx_log2af;
	}
	WfPool pl_fb = pl;   // the map fallback list, traced by the BVH any-hit kernel
	pl_fb.q_list = s->d_q_bvh; pl_fb.q_count_idx = 5; pl_fb.q_cursor_idx = 6;
	wf_ctx_table_kernel<<<(p.n_ctx + 127) / 128, 128, 0, stream>>>(pl, p);
	void (*closest)(SceneDev, WfPool, RenderParams) = s->dev.exact ? wf_traverse_kernel<false, true> : wf_traverse_kernel<false, false>;
	void (*anyhit)(SceneDev, WfPool, RenderParams) = s->dev.exact ? wf_traverse_kernel<true, true> : wf_traverse_kernel<true, false>;
	int bps = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, closest, kBlock, kStackBytes));
	const int trav_grid = s->sm_count * std::max(1, bps);
	const int shade_grid = (int)(slots / 256);
	// Grid-stride kernels.  Lookups: one wave of resident blocks (5 per SM at 48 registers) instead of a fixed 8 per SM,
	// which left a partial second wave: 108 -> 99 ms per 4e7 rays at C4, 489 -> 453 ms per 3e6 rays at C5.  The splat
	// kernel is the other way round (8 per SM: 97 / 99 ms, one wave: 100 / 110 ms).  Capping either kernel at 40 or 32
	// registers for 6 or 8 resident blocks spills and loses (profiles/r2_ab_closest.txt, call 15).
	int vismap_bps = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&vismap_bps, wf_vismap_kernel, 256, 0));
	const int splat_grid = s->sm_count * 8;
	const int vismap_grid = s->sm_count * (s->grid_wave ? std::max(1, vismap_bps) : 8);   // EAR_B200_GRID_WAVE=0: fixed 8
	const int n_pairs = p.n_ctx * p.n_rec;
	const bool windowed = s->splat_mode == 1 && n_pairs <= kPrivMaxPairs && p.n_rec > 0;
	if (windowed) CUDA_TRY(cudaFuncSetAttribute(wf_splat_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kPrivWindow * sizeof(float))));
	// upper bound on iterations: every slot hosts ceil(work/slots) rays of at most max_bounces iterations each
	const auto t_loop = std::chrono::steady_clock::now();
	const long long max_iter = ((p.total_work + slots - 1) / slots + 1) * (long long)(p.max_bounces + 2) + 4 + s->check_every;
	bool finished = false;
	for (long long it = 0; it < max_iter;) {
		for (int k = 0; k < s->check_every && it < max_iter; ++k, ++it) {
			wf_clear_kernel<<<(kBinsTotal + 255) / 256, 256, 0, stream>>>(pl);
			{ LaunchTimer t(s, stream, 0); wf_shade_kernel<<<shade_grid, 256, 0, stream>>>(s->dev, pl, p); }
			{
				LaunchTimer t(s, stream, 4);
				s->stats.launches[4] += 2;
				wf_scan_kernel<<<kScanBlocks, 1024, 0, stream>>>(pl);
				wf_scan_top_kernel<<<1, kScanBlocks, 0, stream>>>(pl);
				wf_scatter_kernel<<<s->sm_count * 8, 256, 0, stream>>>(pl);
			}
			{ LaunchTimer t(s, stream, 1); closest<<<trav_grid, kBlock, kStackBytes, stream>>>(s->dev, pl, p); }
			if (p.n_rec > 0) {
				if (n_mapped) {
					{ LaunchTimer t(s, stream, 6); wf_vismap_kernel<<<vismap_grid, 256, 0, stream>>>(s->dev, pl, p, s->d_maps, s->d_map_of, s->d_q_bvh); }
					{ LaunchTimer t(s, stream, 2); anyhit<<<trav_grid, kBlock, kStackBytes, stream>>>(s->dev, pl_fb, p); }
				} else { LaunchTimer t(s, stream, 2); anyhit<<<trav_grid, kBlock, kStackBytes, stream>>>(s->dev, pl, p); }
				if (windowed) {
					LaunchTimer t(s, stream, 3);
					s->stats.launches[3] += 3;
					CUDA_TRY(cudaMemsetAsync(pl.pair_count, 0, (size_t)n_pairs * sizeof(int), stream));
					wf_vis_count_kernel<<<splat_grid, 256, 0, stream>>>(pl, p);
					wf_vis_scan_kernel<<<1, 1024, 0, stream>>>(pl, n_pairs);
					wf_vis_scatter_kernel<<<splat_grid, 256, 0, stream>>>(pl, p);
					wf_splat_window_kernel<<<s->sm_count * 3, 256, kPrivWindow * sizeof(float), stream>>>(pl, p);
				} else { LaunchTimer t(s, stream, 3); wf_splat_kernel<<<splat_grid, 256, 0, stream>>>(pl, p); }
			}
			++s->stats.iterations;
		}
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaMemcpyAsync(s->h_counts, pl.counts, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
		CUDA_TRY(cudaMemcpyAsync(s->h_counts + 8, s->d_queue, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
		CUDA_TRY(cudaStreamSynchronize(stream));
		unsigned long long queue_head;
		std::memcpy(&queue_head, s->h_counts + 8, sizeof(queue_head));
		if (dbg)
			std::fprintf(stderr, "[ear_b200] it %lld: trav %d queries %d visible %d map-fallback %d queue %llu/%lld\n", it, s->h_counts[0], s->h_counts[1], s->h_counts[2], s->h_counts[5], queue_head, p.total_work);
		harvest_events(s);
		// Done when the last shade left no ray to trace AND the shard's queue is dry.  A slot whose ray ends in launch k
		// is refilled in launch k + 1, so "no live ray" alone also holds between two generations of a pool that turns
		// over in lockstep (every ray reaching the bounce cap in the same iteration).
		if (s->h_counts[0] == 0 && (long long)queue_head >= p.total_work) { finished = true; break; }
		// most occlusion queries of the last iteration were blocked: from now on the maps answer them nearest-first
		if (n_mapped && s->vismap_sort < 0 && s->h_counts[1] > 4096 && (long long)s->h_counts[2] * 2 < s->h_counts[1])
			if (int32_t rc = sort_vismaps(s, stream)) return rc;
	}
	wf_fold_counters_kernel<<<1, kCounterParts, 0, stream>>>(p);
	CUDA_TRY(cudaGetLastError());
	if (!finished) return fail("render: the wavefront loop hit its iteration bound with rays left (internal error)");
	if (dbg) std::fprintf(stderr, "[ear_b200] wavefront loop: %.2f ms after set-up\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_loop).count());
	return 0;
}

static int32_t launch_trace(ear_b200_scene* s, RenderParams& p, cudaStream_t stream) {
	CUDA_TRY(cudaMemsetAsync(s->d_queue, 0, sizeof(unsigned long long), stream));
	CUDA_TRY(cudaMemsetAsync(s->d_counter_parts, 0, kCounterParts * 4 * sizeof(unsigned long long), stream));   // (a failed call may have left some)
	if (p.total_work <= 0) return 0;
	// every ray of the shard must have been emitted when the engine returns: the caller's `rays` counter (which may
	// already hold earlier shards) has to advance by exactly total_work
	unsigned long long rays_before = 0, rays_after = 0;
	CUDA_TRY(cudaMemcpyAsync(&rays_before, p.counters, sizeof(rays_before), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	if (int32_t rc = launch_wavefront(s, p, stream)) return rc;
	CUDA_TRY(cudaMemcpyAsync(&rays_after, p.counters, sizeof(rays_after), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	if ((long long)(rays_after - rays_before) != p.total_work)
		return fail("render: traced " + std::to_string(rays_after - rays_before) + " of " + std::to_string(p.total_work) + " rays (internal error)");
	return 0;
}

static int32_t launch_finalise(ear_b200_scene* s, RenderParams& p, cudaStream_t stream) {
	const int n_tracks = p.n_ctx * p.n_rec * p.tpr;
	dim3 grid(64, (unsigned)std::min(n_tracks, 32768));
	const int direct_grid = std::max(1, std::min((p.n_ctx * p.n_rec + kBlock - 1) / kBlock, s->dev.spill_threads / kBlock));
	LaunchTimer t(s, stream, 5);
	s->stats.launches[5] += 2;
	scale_kernel<<<grid, 256, 0, stream>>>(p, 0);
	if (s->dev.exact) direct_kernel<true><<<direct_grid, kBlock, kStackBytes, stream>>>(s->dev, p);
	else direct_kernel<false><<<direct_grid, kBlock, kStackBytes, stream>>>(s->dev, p);
	scale_kernel<<<grid, 256, 0, stream>>>(p, 1);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

extern "C" int32_t ear_b200_trace_device(ear_b200_scene* s, const ear_b200_context* ctx, int32_t n_ctx,
                                         const ear_b200_recorder* rec, int32_t n_rec, const ear_b200_options* opt,
                                         int32_t n_bins, float* d_hist, uint32_t* d_range, uint64_t* d_counters,
                                         void* stream) {
	if (!s) return fail("trace_device: null scene");
	if (!d_hist || !d_range || !d_counters || n_bins <= 0) return fail("trace_device: null device buffer");
	CUDA_TRY(cudaSetDevice(s->device));
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	RenderParams p{};
	if (int32_t rc = upload_params(s, ctx, n_ctx, rec, n_rec, opt, (cudaStream_t)stream, p)) return rc;
	p.n_bins = n_bins; p.hist = d_hist; p.range = d_range; p.counters = (unsigned long long*)d_counters;
	const auto t1 = std::chrono::steady_clock::now();
	const int32_t rc = launch_trace(s, p, (cudaStream_t)stream);
	if (dbg) {
		cudaStreamSynchronize((cudaStream_t)stream);
		std::fprintf(stderr, "[ear_b200] trace_device: upload %.2f ms, trace (pool, maps, wavefront loop) %.2f ms\n",
		             std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
	}
	return rc;
}

extern "C" int32_t ear_b200_finalise_device(ear_b200_scene* s, const ear_b200_context* ctx, int32_t n_ctx,
                                            const ear_b200_recorder* rec, int32_t n_rec, int32_t n_bins, float* d_hist,
                                            uint32_t* d_range, void* stream) {
	if (!s) return fail("finalise_device: null scene");
	CUDA_TRY(cudaSetDevice(s->device));
	RenderParams p{};
	if (int32_t rc = upload_params(s, ctx, n_ctx, rec, n_rec, nullptr, (cudaStream_t)stream, p)) return rc;
	p.n_bins = n_bins; p.hist = d_hist; p.range = d_range;
	return launch_finalise(s, p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// post chain on device tracks (post.cuh)
// ------------------------------------------------------------------------------------------
static int32_t upload_recorders(ear_b200_scene* s, const ear_b200_recorder* rec, size_t n, cudaStream_t stream) {
	for (size_t i = 0; i < n; ++i)
		if (rec[i].kind != EAR_B200_MONO && rec[i].kind != EAR_B200_STEREO) return fail("post: unknown recorder kind");
	if (n > s->rec_cap) {
		dev_free(s->d_rec); s->d_rec = nullptr; s->rec_cap = 0;
		CUDA_TRY(dev_alloc(&s->d_rec, sizeof(ear_b200_recorder) * n));
		s->rec_cap = n;
	}
	CUDA_TRY(cudaMemcpyAsync(s->d_rec, rec, sizeof(ear_b200_recorder) * n, cudaMemcpyHostToDevice, stream));
	return 0;
}

static int32_t ensure_post_scratch(ear_b200_scene* s, size_t n_tracks) {
	if (n_tracks > s->post_cap) {
		dev_free(s->d_post); s->d_post = nullptr; s->post_cap = 0;
		CUDA_TRY(dev_alloc(&s->d_post, n_tracks * 3 * sizeof(float)));   // [track_max or t60 | track_len | real_length snapshot]
		s->post_cap = n_tracks;
	}
	return 0;
}

extern "C" int32_t ear_b200_post_power_device(ear_b200_scene* s, const ear_b200_recorder* rec, int32_t n_ctx, int32_t n_rec,
                                              int32_t n_bins, float* d_hist, const uint32_t* d_range, float exponent,
                                              float* maximum, float* track_maximum, void* stream) {
	if (!s) return fail("post_power_device: null scene");
	if (!rec || n_ctx <= 0 || n_rec <= 0 || n_bins <= 0 || !d_hist || !d_range) return fail("post_power_device: bad arguments");
	CUDA_TRY(cudaSetDevice(s->device));
	cudaStream_t st = (cudaStream_t)stream;
	const int tpr = ear_b200_tracks_per_recorder(rec, n_ctx * n_rec);
	const size_t n_tracks = (size_t)n_ctx * n_rec * tpr;
	if (int32_t rc = upload_recorders(s, rec, (size_t)n_ctx * n_rec, st)) return rc;
	if (int32_t rc = ensure_post_scratch(s, n_tracks)) return rc;
	post_power_kernel<<<(unsigned)n_tracks, kPostBlock, 0, st>>>(d_hist, d_range, s->d_rec, n_bins, tpr, exponent, s->d_post);
	CUDA_TRY(cudaGetLastError());
	std::vector<float> mx(n_tracks);
	CUDA_TRY(cudaMemcpyAsync(mx.data(), s->d_post, n_tracks * sizeof(float), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	float m = 0.0f;
	for (size_t t = 0; t < n_tracks; ++t) { if (mx[t] > m) m = mx[t]; if (track_maximum) track_maximum[t] = mx[t]; }
	if (maximum) *maximum = m;
	return 0;
}

extern "C" int32_t ear_b200_post_truncate_device(ear_b200_scene* s, const ear_b200_recorder* rec, int32_t n_ctx, int32_t n_rec,
                                                 int32_t n_bins, const float* d_hist, uint32_t* d_range, float threshold,
                                                 float* t60, void* stream) {
	if (!s) return fail("post_truncate_device: null scene");
	if (!rec || n_ctx <= 0 || n_rec <= 0 || n_bins <= 0 || !d_hist || !d_range) return fail("post_truncate_device: bad arguments");
	CUDA_TRY(cudaSetDevice(s->device));
	cudaStream_t st = (cudaStream_t)stream;
	const int tpr = ear_b200_tracks_per_recorder(rec, n_ctx * n_rec);
	const size_t n_tracks = (size_t)n_ctx * n_rec * tpr;
	if (int32_t rc = upload_recorders(s, rec, (size_t)n_ctx * n_rec, st)) return rc;
	if (int32_t rc = ensure_post_scratch(s, n_tracks)) return rc;
	uint32_t* d_len = (uint32_t*)(s->d_post + n_tracks);
	uint32_t* d_real = (uint32_t*)(s->d_post + 2 * n_tracks);
	post_length_kernel<<<(unsigned)n_tracks, kPostBlock, 0, st>>>(d_hist, d_range, s->d_rec, n_bins, tpr, threshold, d_len, d_real);
	post_truncate_t60_kernel<<<(unsigned)n_tracks, kPostBlock, 0, st>>>(d_hist, d_range, s->d_rec, n_bins, tpr, d_len, d_real, s->d_post);
	CUDA_TRY(cudaGetLastError());
	std::vector<float> h(n_tracks);
	CUDA_TRY(cudaMemcpyAsync(h.data(), s->d_post, n_tracks * sizeof(float), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (t60) std::memcpy(t60, h.data(), n_tracks * sizeof(float));
	return 0;
}

// ------------------------------------------------------------------------------------------
// Scene::Render fan-out over the GPUs of one process (src/EAR.cpp:196-207 is the seam): ray sharding + peer reduce
// ------------------------------------------------------------------------------------------
struct ear_b200_group {
	std::vector<ear_b200_scene*> scenes;   // [0] belongs to the caller, the others are clones owned by the group
	bool peer = true;                      // every other GPU can address GPU 0's memory and vice versa
};

// hist0[i] = sum over k of part[k][i] for i in [begin, end): run by GPU g on its slice, reading its peers' partial
// histograms over NVLink and writing the sum straight into GPU 0's buffer (part[0]).  Fixed order: deterministic.
struct PeerParts { const float* part[16]; int n; };
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerParts parts, float* dst, size_t begin, size_t end) {
	const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
	for (size_t i = begin + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < end; i += stride) {
		if (i + 4 <= end) {
			float4 acc = *reinterpret_cast<const float4*>(parts.part[0] + i);
			for (int k = 1; k < parts.n; ++k) {
				const float4 v = *reinterpret_cast<const float4*>(parts.part[k] + i);
				acc.x = fadd(acc.x, v.x); acc.y = fadd(acc.y, v.y); acc.z = fadd(acc.z, v.z); acc.w = fadd(acc.w, v.w);
			}
			*reinterpret_cast<float4*>(dst + i) = acc;
		} else {
			for (size_t j = i; j < end; ++j) {
				float acc = parts.part[0][j];
				for (int k = 1; k < parts.n; ++k) acc = fadd(acc, parts.part[k][j]);
				dst[j] = acc;
			}
		}
	}
}
__global__ void add_into_kernel(float* dst, const float* src, size_t n) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = fadd(dst[i], src[i]);
}

extern "C" int32_t ear_b200_group_create(ear_b200_scene* scene, const int32_t* devices, int32_t n_devices, ear_b200_group** out) {
	if (!scene || !out || !devices || n_devices < 1) return fail("group_create: bad arguments");
	*out = nullptr;
	if (n_devices > 16) return fail("group_create: at most 16 GPUs");
	if (devices[0] != scene->device) return fail("group_create: devices[0] must be the scene's own device");
	std::unique_ptr<ear_b200_group, void (*)(ear_b200_group*)> g(new ear_b200_group(), ear_b200_group_destroy);
	g->scenes.push_back(scene);
	for (int32_t k = 1; k < n_devices; ++k) {
		for (int32_t j = 0; j < k; ++j) if (devices[j] == devices[k]) return fail("group_create: a device is listed twice");
		ear_b200_scene* c = nullptr;
		if (int32_t rc = ear_b200_scene_clone(scene, devices[k], &c)) return rc;
		g->scenes.push_back(c);
		if (!scene->emitter_verts.empty())
			if (int32_t rc = ear_b200_scene_set_emitters(c, scene->emitter_verts.data(), (int32_t)(scene->emitter_verts.size() / 9))) return rc;
	}
	// peer access between GPU 0 and every other GPU, both directions (the sliced reduce reads all partials)
	for (int32_t a = 0; a < n_devices && g->peer; ++a)
		for (int32_t b = 0; b < n_devices && g->peer; ++b) {
			if (a == b) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) { g->peer = false; cudaGetLastError(); break; }
			CUDA_TRY(cudaSetDevice(devices[a]));
			const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { g->peer = false; }
			cudaGetLastError();
		}
	*out = g.release();
	return 0;
}
extern "C" void ear_b200_group_destroy(ear_b200_group* g) {
	if (!g) return;
	for (size_t k = 1; k < g->scenes.size(); ++k) ear_b200_scene_destroy(g->scenes[k]);
	delete g;
}
extern "C" int32_t ear_b200_group_size(ear_b200_group* g) { return g ? (int32_t)g->scenes.size() : 0; }

// What ear_b200_render hands out is the head of this box; ear_b200_result_free finds the track block through it.
constexpr uint64_t kResultMagic = 0x6561725f72657331ull;
struct ResultBox {
	ear_b200_result pub;
	uint64_t magic;
	void* block;          // page-locked block holding every track of the result, or null (tracks calloc'ed one by one)
	size_t block_bytes;
};

// host tracks out of GPU 0's buffers (shared by the single- and multi-GPU renders)
static int32_t assemble_result(ear_b200_scene* s, const ear_b200_recorder* rec, int32_t n_ctx, int32_t n_rec, int tpr, int32_t n_bins,
                               const float* d_hist, const uint32_t* d_range, const unsigned long long counters[8], double ms,
                               float maximum, const float* t60, ear_b200_result** out) {
	const size_t n_tracks = (size_t)n_ctx * n_rec * tpr;
	std::vector<uint32_t> range(n_tracks * 2);
	CUDA_TRY(cudaMemcpyAsync(range.data(), d_range, range.size() * 4, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	// the result is the caller's once it is handed out; until then every error path frees it
	std::unique_ptr<ear_b200_result, void (*)(ear_b200_result*)> holder((ear_b200_result*)calloc(1, sizeof(ResultBox)),
	                                                                    ear_b200_result_free);
	ear_b200_result* res = holder.get();
	if (!res) return fail("render: out of host memory");
	ResultBox* box = (ResultBox*)res;
	box->magic = kResultMagic;
	res->n_contexts = n_ctx; res->n_recorders = n_rec;
	// the result keeps the [context][recorder][2] shape whatever the device layout was (mono: slot 0 only)
	res->tracks = (ear_b200_track*)calloc((size_t)n_ctx * n_rec * 2, sizeof(ear_b200_track));
	if (!res->tracks) return fail("render: out of host memory");
	if (t60) {
		res->t60 = (float*)calloc((size_t)n_ctx * n_rec * 2, sizeof(float));
		if (!res->t60) return fail("render: out of host memory");
	}
	// all tracks of the result live in ONE page-locked block (whole rows come down: the device rows are zero beyond what
	// was recorded, so no host-side clearing is needed); pageable memory only if the block cannot be had
	size_t n_used = 0;
	for (size_t j = 0; j < (size_t)n_ctx * n_rec; ++j) n_used += rec[j].kind == EAR_B200_STEREO ? 2 : 1;
	const size_t row = ((size_t)n_bins + 63) / 64 * 64;   // floats; rows start on 256-byte boundaries
	box->block = hostcache::take(std::max<size_t>(n_used * row, 1) * sizeof(float), &box->block_bytes);
	size_t at = 0;
	for (size_t j = 0; j < (size_t)n_ctx * n_rec; ++j)
		for (int k = 0; k < 2; ++k) {
			ear_b200_track& tr = res->tracks[2 * j + k];
			const bool used = k == 0 || rec[j].kind == EAR_B200_STEREO;
			tr.first_sample = 3 * EAR_B200_SAMPLE_RATE - 1; tr.real_length = 0; tr.length = 0;
			if (!used) continue;
			const size_t t = j * tpr + k;   // device track
			tr.first_sample = range[2 * t]; tr.real_length = range[2 * t + 1]; tr.length = (uint32_t)n_bins;
			if (box->block) {
				tr.data = (float*)box->block + at * row;
				++at;
				// (after the post chain the device row still holds samples beyond the truncated length: those are cleared here)
				const size_t live = t60 ? std::min<size_t>((size_t)tr.real_length + 1, (size_t)n_bins) : (size_t)n_bins;
				CUDA_TRY(cudaMemcpyAsync(tr.data, d_hist + t * (size_t)n_bins, live * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
				if (live < (size_t)n_bins) std::memset(tr.data + live, 0, ((size_t)n_bins - live) * sizeof(float));
			} else {
				tr.data = (float*)calloc((size_t)n_bins, sizeof(float));
				if (!tr.data) return fail("render: out of host memory");
				const size_t live = std::min<size_t>((size_t)tr.real_length + 1, (size_t)n_bins);
				CUDA_TRY(cudaMemcpyAsync(tr.data, d_hist + t * (size_t)n_bins, live * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
			}
			if (t60) res->t60[2 * j + k] = t60[t];
		}
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	res->rays = counters[0]; res->segments = counters[1]; res->occlusion_queries = counters[2];
	res->contributions = counters[3]; res->bin_updates = counters[4]; res->dropped_updates = counters[5];
	res->device_ms = ms; res->bvh_build_ms = s->bvh_build_ms; res->maximum = maximum;
	*out = holder.release();
	return 0;
}

// Render()'s post chain on GPU 0's finalised tracks (opt->post_exponent > 0), see ear_b200_post_*_device
static int32_t run_post(ear_b200_scene* s, const ear_b200_recorder* rec, int32_t n_ctx, int32_t n_rec, int32_t n_bins, float* d_hist,
                        uint32_t* d_range, const ear_b200_options* opt, float& maximum, std::vector<float>& t60) {
	const size_t n_tracks = (size_t)n_ctx * n_rec * ear_b200_tracks_per_recorder(rec, n_ctx * n_rec);
	if (int32_t rc = ear_b200_post_power_device(s, rec, n_ctx, n_rec, n_bins, d_hist, d_range, opt->post_exponent, &maximum, nullptr, s->stream)) return rc;
	t60.assign(n_tracks, 0.0f);
	const float threshold = opt->post_divisor > 0.0f ? maximum / opt->post_divisor : -1.0f;
	return ear_b200_post_truncate_device(s, rec, n_ctx, n_rec, n_bins, d_hist, d_range, threshold, t60.data(), s->stream);
}

static int32_t render_on(std::vector<ear_b200_scene*>& scenes, bool peer, const ear_b200_context* ctx, int32_t n_ctx,
                         const ear_b200_recorder* rec, int32_t n_rec, const ear_b200_options* opt, ear_b200_result** out) {
	const int G = (int)scenes.size();
	ear_b200_scene* s0 = scenes[0];
	if (n_ctx <= 0 || n_rec <= 0 || !ctx || !rec) return fail("render: need at least one context and one recorder");
	const int32_t n_bins = ear_b200_default_bins(s0, opt);
	const int tpr = ear_b200_tracks_per_recorder(rec, n_ctx * n_rec);
	const size_t n_tracks = (size_t)n_ctx * n_rec * tpr;
	const size_t hist_floats = n_tracks * (size_t)n_bins;
	// ray-id range of this call, dealt to the GPUs in contiguous shares (src/EAR.cpp:196-207 deals whole contexts to
	// threads; rays shard finer and the Philox streams make the union independent of the deal)
	long long rays = 0;
	for (int32_t c = 0; c < n_ctx; ++c) rays = std::max<long long>(rays, ctx[c].num_samples);
	const long long first = opt ? opt->first_ray : 0;
	long long count = (opt && opt->ray_count >= 0) ? opt->ray_count : rays - first;
	count = std::max<long long>(0, std::min<long long>(count, rays - first));
	struct PerGpu {
		float* hist = nullptr; uint32_t* range = nullptr; unsigned long long* counters = nullptr;
		int32_t rc = 0; std::string err; double ms = 0.0;
		std::vector<uint32_t> h_range; unsigned long long h_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	};
	std::vector<PerGpu> per((size_t)G);
	struct Cleanup {
		std::vector<PerGpu>& per; std::vector<ear_b200_scene*>& scenes;
		~Cleanup() {
			for (size_t g = 0; g < per.size(); ++g) {
				cudaSetDevice(scenes[g]->device);
				dev_free(per[g].hist); dev_free(per[g].range); dev_free(per[g].counters);
			}
		}
	} cleanup{per, scenes};
	auto work = [&](int g) {
		PerGpu& me = per[(size_t)g];
		ear_b200_scene* s = scenes[(size_t)g];
		auto bad = [&](const std::string& what, cudaError_t e) { me.rc = 1; me.err = what + ": " + cudaGetErrorString(e); };
		cudaError_t e;
		if ((e = cudaSetDevice(s->device)) != cudaSuccess) return bad("cudaSetDevice", e);
		if ((e = dev_alloc(&me.hist, std::max<size_t>(hist_floats, 1) * sizeof(float))) != cudaSuccess) return bad("histogram allocation", e);
		if ((e = dev_alloc(&me.range, n_tracks * 2 * sizeof(uint32_t))) != cudaSuccess) return bad("range allocation", e);
		if ((e = dev_alloc(&me.counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return bad("counter allocation", e);
		cudaMemsetAsync(me.hist, 0, hist_floats * sizeof(float), s->stream);
		cudaMemsetAsync(me.counters, 0, 8 * sizeof(unsigned long long), s->stream);
		init_range_kernel<<<(unsigned)((n_tracks + 127) / 128), 128, 0, s->stream>>>(me.range, (int)n_tracks);
		ear_b200_options o{};
		if (opt) o = *opt;
		o.first_ray = first + count * g / G;
		o.ray_count = first + count * (g + 1) / G - o.first_ray;
		o.finalise = 0;
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0, s->stream);
		me.rc = ear_b200_trace_device(s, ctx, n_ctx, rec, n_rec, &o, n_bins, me.hist, me.range, (uint64_t*)me.counters, s->stream);
		if (me.rc) me.err = g_last_error;   // per-thread error text: carry it to the caller's thread
		cudaEventRecord(e1, s->stream);
		me.h_range.resize(n_tracks * 2);
		cudaMemcpyAsync(me.h_range.data(), me.range, n_tracks * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream);
		cudaMemcpyAsync(me.h_counters, me.counters, sizeof(me.h_counters), cudaMemcpyDeviceToHost, s->stream);
		if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess && !me.rc) bad("trace", e);
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		me.ms = ms;
		cudaEventDestroy(e0); cudaEventDestroy(e1);
	};
	if (G == 1) work(0);
	else {
		std::vector<std::thread> threads;
		for (int g = 0; g < G; ++g) threads.emplace_back(work, g);
		for (auto& t : threads) t.join();
	}
	for (int g = 0; g < G; ++g) if (per[(size_t)g].rc) return fail(per[(size_t)g].err);
	double ms = 0.0;
	unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (int g = 0; g < G; ++g) {
		ms = std::max(ms, per[(size_t)g].ms);
		for (int k = 0; k < 8; ++k) counters[k] += per[(size_t)g].h_counters[k];
	}
	if (G > 1) {
		// ---- ONE reduce of the partial histograms onto GPU 0 ----
		if (peer) {
			// sliced: GPU g sums slice g of all partials (peer reads over NVLink) and writes it into GPU 0's buffer
			PeerParts parts{};
			parts.n = G;
			for (int g = 0; g < G; ++g) parts.part[g] = per[(size_t)g].hist;
			const size_t quads = (hist_floats + 3) / 4;
			for (int g = 0; g < G; ++g) {
				const size_t begin = std::min(hist_floats, quads * g / G * 4), end = std::min(hist_floats, quads * (g + 1) / G * 4);
				if (end <= begin) continue;
				CUDA_TRY(cudaSetDevice(scenes[(size_t)g]->device));
				peer_reduce_kernel<<<scenes[(size_t)g]->sm_count * 4, 256, 0, scenes[(size_t)g]->stream>>>(parts, per[0].hist, begin, end);
				CUDA_TRY(cudaGetLastError());
			}
			for (int g = 0; g < G; ++g) {
				CUDA_TRY(cudaSetDevice(scenes[(size_t)g]->device));
				CUDA_TRY(cudaStreamSynchronize(scenes[(size_t)g]->stream));
			}
		} else {
			// no peer addressing: copy every partial to GPU 0 and add it there
			CUDA_TRY(cudaSetDevice(s0->device));
			DevBuf<float> stage;
			CUDA_TRY(stage.alloc(hist_floats));
			for (int g = 1; g < G; ++g) {
				CUDA_TRY(cudaMemcpyPeerAsync(stage, s0->device, per[(size_t)g].hist, scenes[(size_t)g]->device, hist_floats * sizeof(float), s0->stream));
				add_into_kernel<<<s0->sm_count * 8, 256, 0, s0->stream>>>(per[0].hist, stage, hist_floats);
			}
			CUDA_TRY(cudaStreamSynchronize(s0->stream));
		}
		// track ranges: min / max over the GPUs (a few words per track: on the host)
		std::vector<uint32_t>& r0 = per[0].h_range;
		for (int g = 1; g < G; ++g)
			for (size_t t = 0; t < n_tracks; ++t) {
				r0[2 * t] = std::min(r0[2 * t], per[(size_t)g].h_range[2 * t]);
				r0[2 * t + 1] = std::max(r0[2 * t + 1], per[(size_t)g].h_range[2 * t + 1]);
			}
		CUDA_TRY(cudaSetDevice(s0->device));
		CUDA_TRY(cudaMemcpyAsync(per[0].range, r0.data(), n_tracks * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s0->stream));
	}
	CUDA_TRY(cudaSetDevice(s0->device));
	float maximum = 0.0f;
	std::vector<float> t60;
	if (!opt || opt->finalise) {
		// the direct lobe is a Record() call like any other (src/Scene.cpp:308): GPU 0's counters take it
		RenderParams fp{};
		if (int32_t rc = upload_params(s0, ctx, n_ctx, rec, n_rec, nullptr, s0->stream, fp)) return rc;
		fp.n_bins = n_bins; fp.hist = per[0].hist; fp.range = per[0].range; fp.counters = per[0].counters;
		if (int32_t rc = launch_finalise(s0, fp, s0->stream)) return rc;
		unsigned long long after[8];
		CUDA_TRY(cudaMemcpyAsync(after, per[0].counters, sizeof(after), cudaMemcpyDeviceToHost, s0->stream));
		CUDA_TRY(cudaStreamSynchronize(s0->stream));
		for (int k = 0; k < 8; ++k) counters[k] += after[k] - per[0].h_counters[k];
		if (opt && opt->post_exponent > 0.0f)
			if (int32_t rc = run_post(s0, rec, n_ctx, n_rec, n_bins, per[0].hist, per[0].range, opt, maximum, t60)) return rc;
	}
	return assemble_result(s0, rec, n_ctx, n_rec, tpr, n_bins, per[0].hist, per[0].range, counters, ms, maximum, t60.empty() ? nullptr : t60.data(), out);
}

extern "C" int32_t ear_b200_render(ear_b200_scene* s, const ear_b200_context* ctx, int32_t n_ctx,
                                   const ear_b200_recorder* rec, int32_t n_rec, const ear_b200_options* opt,
                                   ear_b200_result** out) {
	if (!s || !out) return fail("render: null scene/out");
	*out = nullptr;
	std::vector<ear_b200_scene*> one(1, s);
	return render_on(one, true, ctx, n_ctx, rec, n_rec, opt, out);
}

extern "C" int32_t ear_b200_group_render(ear_b200_group* g, const ear_b200_context* ctx, int32_t n_ctx,
                                         const ear_b200_recorder* rec, int32_t n_rec, const ear_b200_options* opt,
                                         ear_b200_result** out) {
	if (!g || !out) return fail("group_render: null group/out");
	*out = nullptr;
	return render_on(g->scenes, g->peer, ctx, n_ctx, rec, n_rec, opt, out);
}

// ------------------------------------------------------------------------------------------
// impulse-response convolution (RecorderTrack::Process, src/Recorder.cpp:247-292)
// ------------------------------------------------------------------------------------------
// One thread per output sample k; the dry signal is walked in increasing i (the reference's outer loop), so
// each sample sees its terms in the reference's order.  Tiles of 256 dry samples and the matching 511-sample
// response window are staged through shared memory; global loads are coalesced, the inner loop reads a broadcast
// dry value and a stride-1 response value (conflict-free).
constexpr int kConvTile = 256;
template <bool FADE>
__global__ void __launch_bounds__(kConvTile) convolve_kernel(const float* __restrict__ r1, uint32_t len1, const float* __restrict__ r2,
                                                             uint32_t len2, uint32_t first, uint32_t len, const float* __restrict__ dry,
                                                             uint32_t n_dry, uint32_t offset, float* __restrict__ out, uint32_t out_len) {
	__shared__ float s_dry[kConvTile];
	__shared__ float s_r1[2 * kConvTile];
	__shared__ float s_r2[FADE ? 2 * kConvTile : 1];
	const long long k0 = (long long)blockIdx.x * kConvTile;          // first output sample of this block
	const long long k = k0 + threadIdx.x;
	const float inv_n = FADE ? fdiv(1.0f, (float)n_dry) : 0.0f;
	float acc = 0.0f;
	// i contributes to k iff first <= k - offset - i < len  <=>  k - offset - len < i <= k - offset - first
	long long i_lo = k0 - (long long)offset - (long long)len + 1;                       // lowest i any thread of the block needs
	long long i_hi = k0 + kConvTile - 1 - (long long)offset - (long long)first;         // highest
	if (i_lo < 0) i_lo = 0;
	if (i_hi > (long long)n_dry - 1) i_hi = (long long)n_dry - 1;
	for (long long ib = i_lo; ib <= i_hi; ib += kConvTile) {
		__syncthreads();
		const long long i_load = ib + threadIdx.x;
		s_dry[threadIdx.x] = i_load <= i_hi ? dry[i_load] : 0.0f;
		// response window: j = k - offset - i for k in [k0, k0+255], i in [ib, ib+255]  ->  j in [jb, jb + 510]
		const long long jb = k0 - (long long)offset - (ib + kConvTile - 1);
		for (int w = threadIdx.x; w < 2 * kConvTile; w += kConvTile) {
			const long long j = jb + w;
			const bool in = j >= (long long)first && j < (long long)len;
			s_r1[w] = (in && j < (long long)len1) ? r1[j] : 0.0f;
			if (FADE) s_r2[w] = (in && j < (long long)len2) ? r2[j] : 0.0f;
		}
		__syncthreads();
		const int n_i = (int)((i_hi - ib + 1 < kConvTile) ? (i_hi - ib + 1) : kConvTile);
		for (int ii = 0; ii < n_i; ++ii) {
			// j - jb = (k - offset - (ib + ii)) - jb = threadIdx.x + 255 - ii
			const int w = (int)threadIdx.x + kConvTile - 1 - ii;
			const long long j = jb + w;
			if (j >= (long long)first && j < (long long)len) {
				float p = s_r1[w];
				if (FADE) {
					const float i1 = fmul((float)(uint32_t)(ib + ii), inv_n), i2 = fsub(1.0f, i1);
					p = fadd(fmul(i2, s_r1[w]), fmul(i1, s_r2[w]));
				}
				acc = fadd(acc, fmul(s_dry[ii], p));
			}
		}
	}
	if (k < (long long)out_len) out[k] = acc;
}

extern "C" int32_t ear_b200_convolve(int32_t device, const float* response, uint32_t length, uint32_t first_sample,
                                     uint32_t real_length, const float* response2, uint32_t length2, uint32_t first_sample2,
                                     uint32_t real_length2, const float* dry, uint32_t n_dry, uint32_t offset, float* out,
                                     uint32_t out_len, uint32_t* out_first, uint32_t* out_real) {
	if (!response || !out || (n_dry && !dry)) return fail("convolve: null argument");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: ear_b200 has no CPU fallback"); }
	if (device < 0 || device >= ndev) return fail("convolve: device index out of range");
	CUDA_TRY(cudaSetDevice(device));
	const bool fade = response2 != nullptr;
	const uint32_t first = fade ? std::min(first_sample, first_sample2) : first_sample;
	const uint32_t len = fade ? std::max(real_length, real_length2) : real_length;
	const uint32_t init_first = 3 * EAR_B200_SAMPLE_RATE - 1;
	if (out_first) *out_first = init_first;
	if (out_real) *out_real = 0;
	if (out_len) std::memset(out, 0, (size_t)out_len * sizeof(float));
	if (n_dry == 0 || len <= first) return 0;       // the reference's loops do not run
	const unsigned long long last = (unsigned long long)(n_dry - 1) + offset + (len - 1);
	if (out_first) *out_first = std::min<uint32_t>(init_first, offset + first);
	if (out_real) *out_real = (uint32_t)last;
	if (last + 1 > out_len) return fail("convolve: output buffer shorter than n_dry - 1 + offset + real_length");
	DevBuf<float> d_r1, d_r2, d_dry, d_out;
	const uint32_t n1 = std::min(length, len), n2 = fade ? std::min(length2, len) : 0;
	CUDA_TRY(d_r1.alloc(n1)); CUDA_TRY(d_dry.alloc(n_dry));
	CUDA_TRY(d_out.alloc((size_t)(last + 1)));
	CUDA_TRY(cudaMemcpy(d_r1, response, (size_t)n1 * 4, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(d_dry, dry, (size_t)n_dry * 4, cudaMemcpyHostToDevice));
	if (fade) { CUDA_TRY(d_r2.alloc(n2)); CUDA_TRY(cudaMemcpy(d_r2, response2, (size_t)n2 * 4, cudaMemcpyHostToDevice)); }
	const unsigned grid = (unsigned)((last + 1 + kConvTile - 1) / kConvTile);
	if (fade) convolve_kernel<true><<<grid, kConvTile>>>(d_r1, n1, d_r2, n2, first, len, d_dry, n_dry, offset, d_out, (uint32_t)(last + 1));
	else convolve_kernel<false><<<grid, kConvTile>>>(d_r1, n1, (const float*)nullptr, 0, first, len, d_dry, n_dry, offset, d_out, (uint32_t)(last + 1));
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpy(out, d_out, (size_t)(last + 1) * 4, cudaMemcpyDeviceToHost));
	return 0;
}

// ------------------------------------------------------------------------------------------
// RecorderTrack::Process in the frequency domain (the reference's USE_FFTW build, src/Recorder.cpp:145-243)
// ------------------------------------------------------------------------------------------
// out = dry (*) response, or for keyframed scenes dry.(1 - i/n) (*) response + dry.(i/n) (*) response2 -- which is the SAME
// linear map as the direct form's per-sample interpolation of the two responses (the weight depends on the dry index only),
// so both forms agree up to float rounding.  One zero-padded real FFT of size N >= n_dry + len per signal (cuFFT, loaded on
// first use), a complex multiply-add, one inverse transform.  705 600 dry samples x 1e6 response samples: N = 2^21.
#include <dlfcn.h>
namespace fftconv {
typedef int cufftHandle;
typedef float2 cufftComplex;
struct Api {
	void* lib = nullptr;
	int (*plan1d)(cufftHandle*, int, int, int) = nullptr;
	int (*exec_r2c)(cufftHandle, float*, cufftComplex*) = nullptr;
	int (*exec_c2r)(cufftHandle, cufftComplex*, float*) = nullptr;
	int (*destroy)(cufftHandle) = nullptr;
	bool ok() const { return plan1d && exec_r2c && exec_c2r && destroy; }
};
static Api& api() {
	static Api a;
	static std::once_flag once;
	std::call_once(once, [] {
		for (const char* name : {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so"}) {
			a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
			if (a.lib) break;
		}
		if (!a.lib) return;
		a.plan1d = (int (*)(cufftHandle*, int, int, int))dlsym(a.lib, "cufftPlan1d");
		a.exec_r2c = (int (*)(cufftHandle, float*, cufftComplex*))dlsym(a.lib, "cufftExecR2C");
		a.exec_c2r = (int (*)(cufftHandle, cufftComplex*, float*))dlsym(a.lib, "cufftExecC2R");
		a.destroy = (int (*)(cufftHandle))dlsym(a.lib, "cufftDestroy");
	});
	return a;
}
constexpr int kR2C = 0x2a, kC2R = 0x2c;   // CUFFT_R2C, CUFFT_C2R
// padded, optionally faded copy of the dry signal: mode 0 plain, 1 x (1 - i/n), 2 x (i/n)   (src/Recorder.cpp:267-292)
__global__ void load_dry_kernel(const float* dry, uint32_t n_dry, float* dst, size_t n_fft, int mode) {
	const float inv_n = fdiv(1.0f, (float)n_dry);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fft; i += (size_t)gridDim.x * blockDim.x) {
		float v = 0.0f;
		if (i < n_dry) {
			v = dry[i];
			const float w = fmul((float)(uint32_t)i, inv_n);
			if (mode == 1) v = fmul(v, fsub(1.0f, w)); else if (mode == 2) v = fmul(v, w);
		}
		dst[i] = v;
	}
}
// response samples [first, len) shifted to start at 0, zero elsewhere
__global__ void load_response_kernel(const float* r, uint32_t have, uint32_t first, uint32_t len, float* dst, size_t n_fft) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fft; i += (size_t)gridDim.x * blockDim.x) {
		const size_t j = i + first;
		dst[i] = (j < len && j < have) ? r[j] : 0.0f;
	}
}
__global__ void multiply_kernel(const float2* a, const float2* b, const float2* c, const float2* d, float2* out, size_t n, float scale) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		float2 v = make_float2(a[i].x * b[i].x - a[i].y * b[i].y, a[i].x * b[i].y + a[i].y * b[i].x);
		if (c) { v.x += c[i].x * d[i].x - c[i].y * d[i].y; v.y += c[i].x * d[i].y + c[i].y * d[i].x; }
		out[i] = make_float2(v.x * scale, v.y * scale);
	}
}
}  // namespace fftconv

extern "C" int32_t ear_b200_convolve_fft(int32_t device, const float* response, uint32_t length, uint32_t first_sample,
                                         uint32_t real_length, const float* response2, uint32_t length2, uint32_t first_sample2,
                                         uint32_t real_length2, const float* dry, uint32_t n_dry, uint32_t offset, float* out,
                                         uint32_t out_len, uint32_t* out_first, uint32_t* out_real) {
	using namespace fftconv;
	if (!response || !out || (n_dry && !dry)) return fail("convolve_fft: null argument");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: ear_b200 has no CPU fallback"); }
	if (device < 0 || device >= ndev) return fail("convolve_fft: device index out of range");
	CUDA_TRY(cudaSetDevice(device));
	const bool fade = response2 != nullptr;
	const uint32_t first = fade ? std::min(first_sample, first_sample2) : first_sample;
	const uint32_t len = fade ? std::max(real_length, real_length2) : real_length;
	const uint32_t init_first = 3 * EAR_B200_SAMPLE_RATE - 1;
	if (out_first) *out_first = init_first;
	if (out_real) *out_real = 0;
	if (out_len) std::memset(out, 0, (size_t)out_len * sizeof(float));
	if (n_dry == 0 || len <= first) return 0;
	const unsigned long long last = (unsigned long long)(n_dry - 1) + offset + (len - 1);
	if (out_first) *out_first = std::min<uint32_t>(init_first, offset + first);
	if (out_real) *out_real = (uint32_t)last;
	if (last + 1 > out_len) return fail("convolve_fft: output buffer shorter than n_dry - 1 + offset + real_length");
	Api& f = api();
	if (!f.ok()) return fail("convolve_fft: libcufft could not be loaded");
	const size_t span = (size_t)n_dry + (len - first);        // linear convolution length (+1)
	size_t n_fft = 1;
	while (n_fft < span) n_fft <<= 1;
	if (n_fft > ((size_t)1 << 30)) return fail("convolve_fft: signal too long for one transform");
	const size_t n_c = n_fft / 2 + 1;
	DevBuf<float> d_in, d_dry, d_sig;
	DevBuf<float2> d_fd, d_fr, d_fd2, d_fr2;
	const uint32_t n1 = std::min(length, len), n2 = fade ? std::min(length2, len) : 0;
	CUDA_TRY(d_in.alloc(std::max<size_t>(std::max(n1, n2), 1))); CUDA_TRY(d_dry.alloc(n_dry)); CUDA_TRY(d_sig.alloc(n_fft));
	CUDA_TRY(d_fd.alloc(n_c)); CUDA_TRY(d_fr.alloc(n_c));
	if (fade) { CUDA_TRY(d_fd2.alloc(n_c)); CUDA_TRY(d_fr2.alloc(n_c)); }
	cufftHandle fwd = 0, inv = 0;
	if (f.plan1d(&fwd, (int)n_fft, kR2C, 1) != 0) return fail("convolve_fft: cufftPlan1d failed");
	if (f.plan1d(&inv, (int)n_fft, kC2R, 1) != 0) { f.destroy(fwd); return fail("convolve_fft: cufftPlan1d failed"); }
	struct Plans { Api& f; cufftHandle a, b; ~Plans() { f.destroy(a); f.destroy(b); } } plans{f, fwd, inv};
	CUDA_TRY(cudaMemcpy(d_dry, dry, (size_t)n_dry * 4, cudaMemcpyHostToDevice));
	const int grid = 148 * 8;
	auto transform = [&](int dry_mode, const float* resp, uint32_t have, float2* fd, float2* fr) -> int32_t {
		load_dry_kernel<<<grid, 256>>>(d_dry, n_dry, d_sig, n_fft, dry_mode);
		if (f.exec_r2c(fwd, d_sig, fd) != 0) return fail("convolve_fft: forward transform failed");
		CUDA_TRY(cudaMemcpy(d_in, resp, (size_t)have * 4, cudaMemcpyHostToDevice));
		load_response_kernel<<<grid, 256>>>(d_in, have, first, len, d_sig, n_fft);
		if (f.exec_r2c(fwd, d_sig, fr) != 0) return fail("convolve_fft: forward transform failed");
		return 0;
	};
	if (int32_t rc = transform(fade ? 1 : 0, response, n1, d_fd, d_fr)) return rc;
	if (fade) { if (int32_t rc = transform(2, response2, n2, d_fd2, d_fr2)) return rc; }
	multiply_kernel<<<grid, 256>>>(d_fd, d_fr, fade ? d_fd2.p : nullptr, fade ? d_fr2.p : nullptr, d_fd, n_c, 1.0f / (float)n_fft);
	if (f.exec_c2r(inv, d_fd, d_sig) != 0) return fail("convolve_fft: inverse transform failed");
	CUDA_TRY(cudaGetLastError());
	// sample k of the linear convolution lands at out[offset + first + k]
	CUDA_TRY(cudaMemcpy(out + offset + first, d_sig, (size_t)(span - 1) * 4, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int32_t ear_b200_scene_stats(ear_b200_scene* s, ear_b200_stats* out) {
	if (!s || !out) return fail("scene_stats: null argument");
	CUDA_TRY(cudaSetDevice(s->device));
	if (s->ev_used) { CUDA_TRY(cudaStreamSynchronize(s->last_stream)); harvest_events(s); }
	*out = s->stats;
	return 0;
}
extern "C" void ear_b200_scene_stats_reset(ear_b200_scene* s) {
	if (!s) return;
	cudaSetDevice(s->device);
	if (s->ev_used) { cudaStreamSynchronize(s->last_stream); harvest_events(s); }
	s->stats = ear_b200_stats{};
}

extern "C" void ear_b200_result_free(ear_b200_result* r) {
	if (!r) return;
	ResultBox* box = (ResultBox*)r;   // results only ever come from assemble_result
	const bool boxed = box->magic == kResultMagic && box->block != nullptr;
	if (r->tracks) {
		const size_t n = (size_t)r->n_contexts * r->n_recorders * 2;
		if (!boxed) for (size_t k = 0; k < n; ++k) free(r->tracks[k].data);
		free(r->tracks);
	}
	if (boxed) hostcache::give(box->block, box->block_bytes);
	free(r->t60);
	free(r);
}

extern "C" int32_t ear_b200_trace_paths(ear_b200_scene* s, const ear_b200_context* ctx, int32_t ctx_index,
                                        const ear_b200_options* opt, int64_t n, int32_t* hits, float* final_state) {
	if (!s || !ctx || !opt || !hits) return fail("trace_paths: null argument");
	if (n <= 0) return 0;
	if (ctx_index < 0 || ctx_index > 65535) return fail("trace_paths: context index out of range");
	CUDA_TRY(cudaSetDevice(s->device));
	const int max_b = opt->max_bounces > 0 ? opt->max_bounces : 1000;
	// the Philox stream is keyed by the context's index: place ctx there, give it all the work
	std::vector<ear_b200_context> cs(ctx_index + 1, *ctx);
	ear_b200_recorder dummy{}; dummy.kind = EAR_B200_MONO;
	std::vector<ear_b200_recorder> rs(ctx_index + 1, dummy);
	std::vector<long long> prefix(ctx_index + 2, 0);
	prefix[ctx_index + 1] = n;
	RenderParams p{};
	if (int32_t rc = upload_params(s, cs.data(), ctx_index + 1, rs.data(), 1, opt, s->stream, p, prefix.data())) return rc;
	p.first_ray = opt->first_ray;
	DevBuf<int32_t> d_hits;
	DevBuf<float> d_state;
	CUDA_TRY(d_hits.alloc((size_t)n * max_b)); CUDA_TRY(d_state.alloc((size_t)n * 8));
	CUDA_TRY(cudaMemsetAsync(d_state, 0, (size_t)n * 32, s->stream));
	wf_fill_int_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>(d_hits, (long long)n * max_b, -2);
	CUDA_TRY(cudaMemsetAsync(s->d_scratch_counters, 0, 8 * sizeof(unsigned long long), s->stream));
	p.n_rec = 0;   // paths only: no occlusion queries, nothing recorded
	p.hits = d_hits; p.final_state = d_state; p.counters = s->d_scratch_counters;
	if (int32_t rc = launch_trace(s, p, s->stream)) return rc;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(hits, d_hits, (size_t)n * max_b * 4, cudaMemcpyDeviceToHost, s->stream));
	if (final_state) CUDA_TRY(cudaMemcpyAsync(final_state, d_state, (size_t)n * 32, cudaMemcpyDeviceToHost, s->stream));
	CUDA_TRY(cudaStreamSynchronize(s->stream));
	return 0;
}
