// Binned-SAH BVH2 builder (host, multi-threaded) -> 64-byte two-box nodes + 48-byte triangle records.
//
// Exactness contract ("first hit must be the reference's"): the GPU evaluates the
// reference's float32 Moeller-Trumbore expression for every triangle it reaches, so the only way
// to differ from Mesh::RayIntersection (src/Mesh.cpp:33-56) is to cull a triangle the float test
// would have accepted.  Float rounding lets that test accept rays that miss the exact triangle:
//   lateral miss   h   <= ~8 eps |O - v0| / sin(phi)                       (phi = angle(e1, e2))
//   along the ray  dt  <= h / sin(theta), theta >= asin(1e-5 / |e1||e2| sin(phi))   (|det| >= 1e-5)
// so each leaf box is padded by `pad` (lateral) and each child carries a `slack` by which the ray
// interval is widened at both ends before that child is culled.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace earb {
namespace {

// four-lane float vectors (GCC vector extensions; min/max through SSE where the host has it)
typedef float f4 __attribute__((vector_size(16)));
typedef int32_t i4 __attribute__((vector_size(16)));
inline f4 vmin(f4 a, f4 b) {
#if defined(__SSE2__)
	return (f4)_mm_min_ps((__m128)a, (__m128)b);
#else
	return a < b ? a : b;
#endif
}
inline f4 vmax(f4 a, f4 b) {
#if defined(__SSE2__)
	return (f4)_mm_max_ps((__m128)a, (__m128)b);
#else
	return a > b ? a : b;
#endif
}
inline f4 splat(float v) { return (f4){v, v, v, v}; }

// lanes 0..2 are x, y, z; lane 3 is free (see Prim)
struct Box {
	f4 lo, hi;
	void reset() { lo = splat(INFINITY); hi = splat(-INFINITY); }
	void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
	float half_area() const {
		const f4 d = hi - lo;
		return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
	}
};

struct BuildNode {
	Box box;
	int32_t left, right;   // -1 for leaves
	int32_t first, count;  // leaf range in `prims`
	float slack;
};

// One primitive of the build: padded box; lane 3 of `lo` holds the original triangle index (as bits), lane 3 of
// `hi` the interval slack -- so a running max over `hi` also yields a subtree's slack.  The array is partitioned
// in place: every pass over a node streams a contiguous range.
struct Prim {
	f4 lo, hi;
	int32_t index() const { int32_t i; std::memcpy(&i, reinterpret_cast<const char*>(&lo) + 12, 4); return i; }
	void set_index(int32_t i) { std::memcpy(reinterpret_cast<char*>(&lo) + 12, &i, 4); }
	f4 centroid() const {   // lane 3: meaningless but finite (the index bits are masked: they may be denormal)
		const i4 mask = {-1, -1, -1, 0};
		return 0.5f * ((f4)((i4)lo & mask) + hi);
	}
};

// Worker pool for the builder.  Tasks are fire-and-forget closures that may submit more tasks; whoever waits
// (drain(), or a parallel pass waiting for its chunks) runs queued tasks itself instead of blocking.
class Pool {
public:
	explicit Pool(int threads) {
		for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { work(); });
	}
	~Pool() {
		{ std::lock_guard<std::mutex> g(m_); stop_ = true; }
		cv_.notify_all();
		for (auto& t : workers_) t.join();
	}
	int threads() const { return (int)workers_.size() + 1; }
	void submit(std::function<void()> fn) {
		pending_.fetch_add(1);
		{ std::lock_guard<std::mutex> g(m_); queue_.push_back(std::move(fn)); }
		cv_.notify_one();
	}
	// runs queued tasks on the calling thread until `done()` holds
	template <class Done>
	void help_until(Done done) {
		int idle = 0;
		while (!done()) {
			std::function<void()> fn;
			{
				std::lock_guard<std::mutex> g(m_);
				if (!queue_.empty()) { fn = std::move(queue_.front()); queue_.pop_front(); }
			}
			if (fn) { fn(); pending_.fetch_sub(1); idle = 0; }
			else if (++idle < 64) std::this_thread::yield();
			else std::this_thread::sleep_for(std::chrono::microseconds(20));   // oversubscribed host: do not spin
		}
	}
	void drain() { help_until([this] { return pending_.load() == 0; }); }
	// fn(part, begin, end) over [0, n) in `parts` contiguous ranges; returns when all are done
	template <class Fn>
	void ranges(int64_t n, int parts, Fn fn) {
		parts = (int)std::max<int64_t>(1, std::min<int64_t>(parts, n));
		if (parts <= 1) { fn(0, (int64_t)0, n); return; }
		const int64_t step = (n + parts - 1) / parts;
		std::atomic<int> left{parts - 1};
		for (int k = 1; k < parts; ++k)
			submit([&fn, &left, k, n, step] { fn(k, std::min(n, k * step), std::min(n, (k + 1) * step)); left.fetch_sub(1); });
		fn(0, (int64_t)0, std::min(n, step));
		help_until([&left] { return left.load() == 0; });
	}

private:
	void work() {
		for (;;) {
			std::function<void()> fn;
			{
				std::unique_lock<std::mutex> g(m_);
				cv_.wait(g, [this] { return stop_ || !queue_.empty(); });
				if (queue_.empty()) return;   // stop_ and nothing left
				fn = std::move(queue_.front());
				queue_.pop_front();
			}
			fn();
			pending_.fetch_sub(1);
		}
	}
	std::vector<std::thread> workers_;
	std::deque<std::function<void()>> queue_;
	std::mutex m_;
	std::condition_variable cv_;
	std::atomic<int> pending_{0};
	bool stop_ = false;
};

struct Builder {
	int32_t n = 0;
	RawVector<Prim> prims;
	RawVector<BuildNode> nodes;
	std::atomic<int32_t> next_node{0};
	static constexpr int kBins = 16;
	static constexpr int32_t kSmallNode = 4;              // nodes up to this size (4: -14 % build time, 8: -11 %, 16: -2 %) evaluate their split candidates directly
	static constexpr int32_t kTaskThreshold = 1 << 14;     // subtrees at least this big may run as their own task
	static constexpr int32_t kChunkThreshold = 1 << 17;    // passes over at least this many primitives are split over threads
	Pool* pool = nullptr;
	RawVector<Prim> scratch;   // for the parallel partition of big nodes
	int max_leaf = kMaxLeaf;   // triangles per leaf (<= 8: three bits in the leaf reference)
	float sah_ct = 1.0f;       // cost of one node step in triangle tests

	int32_t alloc() { return next_node.fetch_add(1); }

	void make_leaf(int32_t id, int32_t first, int32_t count) {
		BuildNode& nd = nodes[id];
		nd.left = nd.right = -1;
		nd.first = first; nd.count = count;
	}

	struct Bins {
		Box box[3][kBins];
		int32_t count[3][kBins];
		void reset() { for (int a = 0; a < 3; ++a) for (int b = 0; b < kBins; ++b) { box[a][b].reset(); count[a][b] = 0; } }
		void merge(const Bins& o) {
			for (int a = 0; a < 3; ++a) for (int b = 0; b < kBins; ++b) { box[a][b].grow(o.box[a][b]); count[a][b] += o.count[a][b]; }
		}
	};
	struct Extent {
		Box bounds, cbounds;
		void reset() { bounds.reset(); cbounds.reset(); }
		void merge(const Extent& o) { bounds.grow(o.bounds); cbounds.grow(o.cbounds); }
	};

	// one pass over [first, first+count) producing a mergeable partial; big ranges are split over idle threads
	// (min / max and integer counts: the merged result does not depend on the split)
	template <class Partial, class Fn>
	void reduce_pass(int32_t first, int32_t count, Partial& total, Fn fn) {
		const int parts = count >= kChunkThreshold ? pool->threads() : 1;
		if (parts <= 1) { fn(first, first + count, total); return; }
		std::vector<Partial> partial((size_t)parts, total);
		pool->ranges(count, parts, [&](int k, int64_t b, int64_t e) { fn(first + (int32_t)b, first + (int32_t)e, partial[(size_t)k]); });
		for (int k = 0; k < parts; ++k) total.merge(partial[(size_t)k]);
	}

	// depth guard: past kSahDepth levels fall back to median splits so the total depth (and the
	// traversal stack) stays bounded by kSahDepth + log2(n) even for adversarial inputs
	static constexpr int kSahDepth = 30;

	void build(int32_t id, int32_t first, int32_t count, int depth = 0) {
		BuildNode& nd = nodes[id];
		Extent ext;
		ext.reset();
		reduce_pass(first, count, ext, [this](int32_t b, int32_t e, Extent& x) {
			Box bb = x.bounds, cb = x.cbounds;
			for (int32_t i = b; i < e; ++i) {
				const Prim& p = prims[i];
				const f4 c = p.centroid();
				bb.lo = vmin(bb.lo, p.lo); bb.hi = vmax(bb.hi, p.hi);
				cb.lo = vmin(cb.lo, c); cb.hi = vmax(cb.hi, c);
			}
			x.bounds = bb; x.cbounds = cb;
		});
		const Box& bounds = ext.bounds;
		const Box& cbounds = ext.cbounds;
		nd.box = bounds;
		nd.slack = bounds.hi[3];   // max over the primitives' slack lanes
		if (count <= 1) { make_leaf(id, first, count); return; }

		// binned SAH, the three axes in one pass over the primitives
		bool usable[3];
		f4 cmin = cbounds.lo, scale = splat(0.0f);
		cmin[3] = 0.0f;
		for (int a = 0; a < 3; ++a) {
			usable[a] = cbounds.hi[a] > cbounds.lo[a];
			scale[a] = usable[a] ? (float)kBins / (cbounds.hi[a] - cbounds.lo[a]) : 0.0f;
		}
		float best_cost = INFINITY;
		int best_axis = -1, best_split = -1;
		if (count <= kSmallNode) {
			// Few primitives (two thirds of all nodes): the binned sweep below spends its time clearing and walking 48
			// mostly empty bins.  The same candidates, evaluated directly: a split position only matters where the set
			// of primitives on the left changes, i.e. at the occupied bins; costs, visiting order and the strict '<' are
			// those of the sweep, so the chosen (axis, bin) is identical.
			for (int axis = 0; axis < 3; ++axis) {
				if (!usable[axis]) continue;
				int bin_of[kSmallNode];
				uint32_t occupied = 0;
				for (int32_t i = 0; i < count; ++i) {
					const Prim& p = prims[first + i];
					int bin = (int)((0.5f * (p.lo[axis] + p.hi[axis]) - cmin[axis]) * scale[axis]);
					bin = std::min(std::max(bin, 0), kBins - 1);
					bin_of[i] = bin;
					occupied |= 1u << bin;
				}
				uint32_t rest = occupied;
				while (rest & (rest - 1)) {   // every occupied bin but the last is a candidate boundary
					const int b = __builtin_ctz(rest);
					rest &= rest - 1;
					Box left, right;
					left.reset(); right.reset();
					int32_t cl = 0, cr = 0;
					for (int32_t i = 0; i < count; ++i) {
						const Prim& p = prims[first + i];
						if (bin_of[i] <= b) { left.lo = vmin(left.lo, p.lo); left.hi = vmax(left.hi, p.hi); ++cl; }
						else { right.lo = vmin(right.lo, p.lo); right.hi = vmax(right.hi, p.hi); ++cr; }
					}
					const float cost = left.half_area() * (float)cl + right.half_area() * (float)cr;
					if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
				}
			}
		} else {
		Bins bins;
		bins.reset();
		reduce_pass(first, count, bins, [this, cmin, scale](int32_t b, int32_t e, Bins& x) {
			for (int32_t i = b; i < e; ++i) {
				const Prim& p = prims[i];
				const i4 bi = __builtin_convertvector((p.centroid() - cmin) * scale, i4);   // truncates like (int)
				for (int a = 0; a < 3; ++a) {
					const int bin = std::min(std::max((int)bi[a], 0), kBins - 1);
					Box& bx = x.box[a][bin];
					bx.lo = vmin(bx.lo, p.lo); bx.hi = vmax(bx.hi, p.hi);
					++x.count[a][bin];
				}
			}
		});
		for (int axis = 0; axis < 3; ++axis) {
			if (!usable[axis]) continue;
			float right_area[kBins];
			int32_t right_count[kBins];
			Box acc; acc.reset();
			int32_t cnt = 0;
			for (int b = kBins - 1; b > 0; --b) {
				acc.grow(bins.box[axis][b]);
				cnt += bins.count[axis][b];
				right_area[b] = cnt ? acc.half_area() : 0.0f;
				right_count[b] = cnt;
			}
			acc.reset(); cnt = 0;
			for (int b = 0; b < kBins - 1; ++b) {
				acc.grow(bins.box[axis][b]);
				cnt += bins.count[axis][b];
				if (cnt == 0 || right_count[b + 1] == 0) continue;
				const float cost = acc.half_area() * (float)cnt + right_area[b + 1] * (float)right_count[b + 1];
				if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
			}
		}
		}
		// leaf cost (1 per triangle) vs split cost (traversal step ~ sah_ct triangle tests)
		const float parent_area = bounds.half_area();
		if (count <= max_leaf) {
			const float leaf_cost = (float)count;
			const float split_cost = best_axis < 0 ? INFINITY : sah_ct + best_cost / std::max(parent_area, 1e-30f);
			if (!(split_cost < leaf_cost)) { make_leaf(id, first, count); return; }
		}
		int32_t mid;
		Prim* pb = &prims[first];
		if (depth >= kSahDepth && count > max_leaf) {
			int ax = 0;
			for (int k = 1; k < 3; ++k) if (cbounds.hi[k] - cbounds.lo[k] > cbounds.hi[ax] - cbounds.lo[ax]) ax = k;
			mid = first + count / 2;
			std::nth_element(pb, pb + count / 2, pb + count, [ax](const Prim& x, const Prim& y) { return x.lo[ax] + x.hi[ax] < y.lo[ax] + y.hi[ax]; });
		} else if (best_axis < 0) {
			mid = first + count / 2;  // coincident centroids: split by position
		} else {
			const float c0 = cmin[best_axis], sc = scale[best_axis];
			const int ax = best_axis, split = best_split;
			auto goes_left = [c0, sc, ax, split](const Prim& p) {
				int bin = (int)((0.5f * (p.lo[ax] + p.hi[ax]) - c0) * sc);
				bin = std::min(std::max(bin, 0), kBins - 1);
				return bin <= split;
			};
			if (count >= kChunkThreshold) mid = first + stable_partition_big(first, count, goes_left);
			else mid = first + (int32_t)(std::partition(pb, pb + count, goes_left) - pb);
			if (mid == first || mid == first + count) mid = first + count / 2;
		}
		const int32_t l = alloc(), r = alloc();
		nd.left = l; nd.right = r; nd.first = first; nd.count = count;
		const int32_t lc = mid - first, rc = first + count - mid;
		// nothing of a node depends on its finished children, so subtrees are forked and never joined
		if (std::min(lc, rc) >= kTaskThreshold) {
			pool->submit([this, l, first, lc, depth] { build(l, first, lc, depth + 1); });
			build(r, mid, rc, depth + 1);
		} else {
			build(l, first, lc, depth + 1);
			build(r, mid, rc, depth + 1);
		}
	}

	// Stable two-way partition of a big node through `scratch`, chunked over the pool (count flags per chunk,
	// prefix, scatter, copy back).  The result does not depend on the number of chunks.  Returns the left count.
	template <class Pred>
	int32_t stable_partition_big(int32_t first, int32_t count, Pred goes_left) {
		const int parts = pool->threads();
		std::vector<int32_t> lefts((size_t)parts + 1, 0);
		pool->ranges(count, parts, [&](int k, int64_t b, int64_t e) {
			int32_t c = 0;
			for (int64_t i = b; i < e; ++i) c += goes_left(prims[first + i]) ? 1 : 0;
			lefts[(size_t)k + 1] = c;
		});
		for (int k = 0; k < parts; ++k) lefts[(size_t)k + 1] += lefts[(size_t)k];
		const int32_t n_left = lefts[(size_t)parts];
		const int64_t step = (count + parts - 1) / parts;   // the same split Pool::ranges makes
		pool->ranges(count, parts, [&](int k, int64_t b, int64_t e) {
			int32_t l = first + lefts[(size_t)k];
			int32_t r = first + n_left + (int32_t)(std::min<int64_t>(count, k * step) - lefts[(size_t)k]);
			for (int64_t i = b; i < e; ++i) {
				const Prim& p = prims[first + i];
				if (goes_left(p)) scratch[l++] = p; else scratch[r++] = p;
			}
		});
		pool->ranges(count, parts, [&](int, int64_t b, int64_t e) {
			std::memcpy(&prims[first + b], &scratch[first + b], (size_t)(e - b) * sizeof(Prim));
		});
		return n_left;
	}
};

int32_t encode_leaf(int32_t first, int32_t count) { return ~((first << 3) | (count - 1)); }


uint16_t half_up(float f) {   // smallest fp16 >= f (f >= 0)
	if (!(f > 0.0f)) return 0;
	if (f >= 65504.0f) return 0x7bff;
	uint32_t bits; std::memcpy(&bits, &f, 4);
	const int e = (int)((bits >> 23) & 0xff) - 127;
	if (e < -14) {   // subnormal half: step 2^-24
		const uint32_t q = (uint32_t)std::ceil(f * 16777216.0f);
		return (uint16_t)std::min<uint32_t>(q, 0x400);
	}
	const uint32_t man = bits & 0x7fffffu;
	uint32_t h = (uint32_t)((e + 15) << 10) | (man >> 13);
	if (man & 0x1fffu) ++h;   // round up; a carry walks into the exponent correctly
	return (uint16_t)std::min<uint32_t>(h, 0x7bff);
}

// One output node: the (up to four) children of a collapsed subtree, quantised outward inside `bounds`.
void emit_node(Node& nd, const Box& bounds, const BuildNode* const* kids, int n_kids, const int32_t* refs) {
	std::memset(&nd, 0, sizeof(nd));
	double step[3];
	for (int a = 0; a < 3; ++a) {
		nd.lo[a] = bounds.lo[a];
		const double ext = (double)bounds.hi[a] - (double)bounds.lo[a];
		// e = ceil(log2(ext / 255)), from the binary exponent; the loop guards the boundary cases
		int e = -126;
		if (ext > 0.0) {
			int ex2;
			const double m = std::frexp(ext / 255.0, &ex2);   // ext/255 = m * 2^ex2, m in [0.5, 1)
			e = m == 0.5 ? ex2 - 1 : ex2;
		}
		e = std::max(-126, std::min(127, e));
		while (e < 127 && std::ldexp(255.0, e) < ext) ++e;
		nd.ex[a] = (uint8_t)(e + 127);
		step[a] = std::ldexp(1.0, e);
	}
	for (int k = 0; k < 4; ++k) {
		if (k >= n_kids || refs[k] == kEmptyChild) {
			nd.child[k] = kEmptyChild;
			for (int a = 0; a < 3; ++a) { nd.q[a][k] = 255; nd.q[3 + a][k] = 0; }
			nd.slack[k] = 0;
			continue;
		}
		nd.child[k] = refs[k];
		nd.slack[k] = half_up(kids[k]->slack);
		for (int a = 0; a < 3; ++a) {
			const double l = ((double)kids[k]->box.lo[a] - (double)nd.lo[a]) / step[a];
			const double h = ((double)kids[k]->box.hi[a] - (double)nd.lo[a]) / step[a];
			nd.q[a][k] = (uint8_t)std::max(0.0, std::min(255.0, std::floor(l)));
			nd.q[3 + a][k] = (uint8_t)std::max(0.0, std::min(255.0, std::ceil(h)));
		}
	}
}

}  // namespace

void build_bvh(const float* verts, const int32_t* tri_material, int32_t n, Bvh& out) {
	const bool dbg = std::getenv("EAR_B200_DEBUG") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!dbg) return;
		const auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[ear_b200] bvh: %-20s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	Builder b;
	b.n = n;
	int threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
	if (const char* e = std::getenv("EAR_B200_BUILD_THREADS")) threads = std::max(1, std::atoi(e));
	if (n < (1 << 14)) threads = 1;
	Pool pool(threads);
	b.pool = &pool;
	if (const char* e = std::getenv("EAR_B200_MAX_LEAF")) b.max_leaf = std::max(1, std::min(8, std::atoi(e)));
	if (const char* e = std::getenv("EAR_B200_SAH_CT")) b.sah_ct = (float)std::atof(e);
	const int wide = threads;   // chunks of the flat per-triangle passes

	// scene bounds (unpadded)
	float slo[3] = {INFINITY, INFINITY, INFINITY}, shi[3] = {-INFINITY, -INFINITY, -INFINITY};
	{
		std::vector<float> part((size_t)wide * 6);
		pool.ranges(n, wide, [&](int k, int64_t i0, int64_t i1) {
			float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
			for (int64_t i = i0; i < i1; ++i)
				for (int v = 0; v < 3; ++v)
					for (int a = 0; a < 3; ++a) {
						const float x = verts[9 * (size_t)i + 3 * v + a];
						lo[a] = std::min(lo[a], x); hi[a] = std::max(hi[a], x);
					}
			for (int a = 0; a < 3; ++a) { part[(size_t)k * 6 + a] = lo[a]; part[(size_t)k * 6 + 3 + a] = hi[a]; }
		});
		const int used = (int)std::max<int64_t>(1, std::min<int64_t>(wide, n));
		for (int k = 0; k < used && n > 0; ++k)
			for (int a = 0; a < 3; ++a) { slo[a] = std::min(slo[a], part[(size_t)k * 6 + a]); shi[a] = std::max(shi[a], part[(size_t)k * 6 + 3 + a]); }
	}
	if (n == 0) { for (int i = 0; i < 3; ++i) { slo[i] = 0; shi[i] = 0; } }
	float diag2 = 0, maxabs = 0;
	for (int i = 0; i < 3; ++i) {
		const float d = shi[i] - slo[i];
		diag2 += d * d;
		maxabs = std::max(maxabs, std::max(std::fabs(slo[i]), std::fabs(shi[i])));
		out.lo[i] = slo[i]; out.hi[i] = shi[i];
	}
	out.diagonal = std::sqrt(diag2);
	const float eps = 5.9604645e-8f;                  // 2^-24
	// test knob: scales pad and slack (0 = bare boxes) so the parity tests can show that the
	// adversarial ray set actually needs them; the product never sets it
	const char* knob = std::getenv("EAR_B200_BVH_MARGIN_SCALE");
	const float margin_scale = knob ? (float)std::atof(knob) : 1.0f;
	const float reach = 2.0f * out.diagonal + 1.0f;   // ray origins may sit outside the bounds
	out.s0 = margin_scale * 64.0f * eps * (maxabs + reach);

	// per-triangle padded box and interval slack (header comment)
	b.prims.resize(n);
	pool.ranges(n, wide, [&](int, int64_t i0, int64_t i1) {
		for (int64_t i = i0; i < i1; ++i) {
			const float* p = verts + 9 * (size_t)i;
			float e1[3], e2[3], cr[3];
			for (int k = 0; k < 3; ++k) { e1[k] = p[3 + k] - p[k]; e2[k] = p[6 + k] - p[k]; }
			cr[0] = e1[1] * e2[2] - e1[2] * e2[1]; cr[1] = e1[2] * e2[0] - e1[0] * e2[2]; cr[2] = e1[0] * e2[1] - e1[1] * e2[0];
			const float l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
			const float l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
			const float cl = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
			float sinphi = (l1 > 0 && l2 > 0) ? cl / (l1 * l2) : 1.0f;
			sinphi = std::max(sinphi, 1e-3f);
			const float pad = margin_scale * (16.0f * eps * reach / sinphi + 8.0f * eps * (maxabs + reach));
			Prim& pr = b.prims[(size_t)i];
			for (int k = 0; k < 3; ++k) {
				pr.lo[k] = std::min(p[k], std::min(p[3 + k], p[6 + k])) - pad;
				pr.hi[k] = std::max(p[k], std::max(p[3 + k], p[6 + k])) + pad;
			}
			pr.hi[3] = std::min(margin_scale * 16.0f * eps * reach * l1 * l2 / 1e-5f + pad, 4.0f * reach);
			pr.set_index((int32_t)i);
		}
	});
	lap("setup");

	b.nodes.resize(std::max<size_t>(2 * (size_t)n, 4));
	const int32_t root = b.alloc();
	if (n >= Builder::kChunkThreshold) b.scratch.resize(n);
	if (n > 0) { b.build(root, 0, n); pool.drain(); }
	else { b.nodes[root].left = b.nodes[root].right = -1; b.nodes[root].first = 0; b.nodes[root].count = 0; b.nodes[root].slack = 0; b.nodes[root].box.reset(); }
	lap("sah build");

	// triangle records in leaf order
	out.tris.resize(n);
	pool.ranges(n, wide, [&](int, int64_t i0, int64_t i1) {
		for (int64_t i = i0; i < i1; ++i) {
			const int32_t t = b.prims[(size_t)i].index();
			const float* p = verts + 9 * (size_t)t;
			TriRecord& r = out.tris[(size_t)i];
			for (int k = 0; k < 3; ++k) { r.v0[k] = p[k]; r.e1[k] = p[3 + k] - p[k]; r.e2[k] = p[6 + k] - p[k]; }
			r.index = t; r.material = tri_material ? tri_material[t] : 0; r.pad0 = 0; r.pad1 = 0;
			// gmtl::normal(tri): cross, length, divide each component (no FMA: see the build flags)
			const float cx = (r.e1[1] * r.e2[2]) - (r.e1[2] * r.e2[1]);
			const float cy = (r.e1[2] * r.e2[0]) - (r.e1[0] * r.e2[2]);
			const float cz = (r.e1[0] * r.e2[1]) - (r.e1[1] * r.e2[0]);
			float l2 = cx * cx; l2 = l2 + cy * cy; l2 = l2 + cz * cz;
			const float len = std::sqrt(l2);
			r.normal[0] = cx; r.normal[1] = cy; r.normal[2] = cz;
			if (len != 0.0f) { r.normal[0] = cx / len; r.normal[1] = cy / len; r.normal[2] = cz / len; }
		}
	});
	lap("triangle records");

	// Collapse to 4-wide nodes laid out depth-first with siblings adjacent: a node's inner children take
	// consecutive numbers, then follow the descendants of its last inner child, of the one before it, ...
	// Pass 1 counts the wide descendants of every node that becomes a wide node (recursive, big subtrees as pool
	// tasks); with those known, pass 2 numbers, quantises and writes each subtree independently.
	out.nodes.clear();
	out.depth = 0;
	const BuildNode* root_node = &b.nodes[root];
	if (root_node->left < 0) {
		// single-leaf (or empty) scene: one node, one child, so traversal can always start with a node fetch
		const BuildNode* kids[1] = {root_node};
		const int32_t refs[1] = {root_node->count ? encode_leaf(root_node->first, root_node->count) : kEmptyChild};
		Box bx = root_node->box;
		if (!root_node->count) { bx.lo = splat(0.0f); bx.hi = splat(0.0f); }
		out.nodes.resize(1);
		emit_node(out.nodes[0], bx, kids, 1, refs);
		out.depth = 1;
		return;
	}
	// expands the inner child with the largest surface area until there are four children
	auto expand = [&b](const BuildNode& bn, int32_t kid[4]) -> int {
		kid[0] = bn.left; kid[1] = bn.right; kid[2] = kid[3] = -1;
		int n_kids = 2;
		while (n_kids < 4) {
			int best = -1; float best_area = -1.0f;
			for (int k = 0; k < n_kids; ++k) {
				const BuildNode& c = b.nodes[kid[k]];
				if (c.left < 0) continue;
				const float area = c.box.half_area();
				if (area > best_area) { best_area = area; best = k; }
			}
			if (best < 0) break;
			const BuildNode& c = b.nodes[kid[best]];
			kid[best] = c.left;
			kid[n_kids++] = c.right;
		}
		return n_kids;
	};
	const int32_t n_build = b.next_node.load();
	RawVector<int32_t> desc((size_t)n_build);   // written for the nodes that become wide nodes only
	std::function<int32_t(int32_t)> count_desc = [&](int32_t id) -> int32_t {
		const BuildNode& bn = b.nodes[id];
		int32_t kid[4];
		const int n_kids = expand(bn, kid);
		int32_t part[4] = {0, 0, 0, 0};
		if (bn.count >= Builder::kTaskThreshold) {
			// big subtree: the inner children are counted as separate tasks; this thread helps until they are done
			std::atomic<int> left{0};
			int last = -1;
			for (int k = 0; k < n_kids; ++k) if (b.nodes[kid[k]].left >= 0) last = k;
			for (int k = 0; k < n_kids; ++k) {
				if (b.nodes[kid[k]].left < 0 || k == last) continue;
				left.fetch_add(1);
				const int32_t c_id = kid[k];
				int32_t* slot = &part[k];
				pool.submit([&count_desc, &left, c_id, slot] { *slot = 1 + count_desc(c_id); left.fetch_sub(1); });
			}
			if (last >= 0) part[last] = 1 + count_desc(kid[last]);
			pool.help_until([&left] { return left.load() == 0; });
		} else {
			for (int k = 0; k < n_kids; ++k) if (b.nodes[kid[k]].left >= 0) part[k] = 1 + count_desc(kid[k]);
		}
		const int32_t d = part[0] + part[1] + part[2] + part[3];
		desc[(size_t)id] = d;
		return d;
	};
	count_desc(root);
	out.nodes.resize((size_t)desc[(size_t)root] + 1);
	std::atomic<int> max_depth{0};
	std::function<void(int32_t, int32_t, int32_t, int)> emit = [&](int32_t id, int32_t out_id, int32_t block, int depth) {
		const BuildNode& bn = b.nodes[id];
		int32_t kid[4], ref[4] = {kEmptyChild, kEmptyChild, kEmptyChild, kEmptyChild}, kid_block[4] = {0, 0, 0, 0};
		const int n_kids = expand(bn, kid);
		const BuildNode* kids[4] = {nullptr, nullptr, nullptr, nullptr};
		int32_t m = 0;
		for (int k = 0; k < n_kids; ++k) {
			const BuildNode& c = b.nodes[kid[k]];
			kids[k] = &c;
			if (c.left < 0) ref[k] = c.count ? encode_leaf(c.first, c.count) : kEmptyChild;
			else ref[k] = block + m++;
		}
		emit_node(out.nodes[(size_t)out_id], bn.box, kids, n_kids, ref);
		int32_t start = block + m;
		for (int k = n_kids - 1; k >= 0; --k)
			if (kids[k]->left >= 0) { kid_block[k] = start; start += desc[(size_t)kid[k]]; }
		if (m == 0) {
			int seen = max_depth.load();
			while (seen < depth && !max_depth.compare_exchange_weak(seen, depth)) {}
			return;
		}
		for (int k = 0; k < n_kids; ++k) {
			if (kids[k]->left < 0) continue;
			const int32_t c_id = kid[k], c_out = ref[k], c_block = kid_block[k];
			if (kids[k]->count >= Builder::kTaskThreshold) pool.submit([&emit, c_id, c_out, c_block, depth] { emit(c_id, c_out, c_block, depth + 1); });
			else emit(c_id, c_out, c_block, depth + 1);
		}
	};
	emit(root, 0, 1, 1);
	pool.drain();
	out.depth = max_depth.load();
	lap("collapse + quantise");
}

}  // namespace earb
