// TEST INFRASTRUCTURE: host build of the product's BVH builder + traversal headers (see cuda_shim.h).
// Exposes first-hit / occlusion over flat arrays so tests can compare them with the oracle here.
#include "cuda_shim.h"
#include "../../ear_b200/csrc/bvh_build.h"
#include "../../ear_b200/csrc/traverse.cuh"
#include <vector>
#include <chrono>
using namespace earb;
struct Emul { Bvh bvh; SceneDev dev; double build_ms; };
extern "C" {
void* emul_create(const float* verts, const int32_t* mats, int32_t n) {
	Emul* e = new Emul();
	auto t0 = std::chrono::steady_clock::now();
	build_bvh(verts, mats, n, e->bvh);
	e->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	e->dev.nodes = (const float4*)e->bvh.nodes.data();
	e->dev.tris = (const float4*)e->bvh.tris.data();
	e->dev.materials = nullptr; e->dev.n_tris = n; e->dev.n_materials = 0; e->dev.n_bands = 0;
	e->dev.s0 = e->bvh.s0; e->dev.exact = 0; e->dev.leaf_vote = kLeafVote; e->dev.fetch_vote = 8; e->dev.vis_cap = 64;
	return e;
}
void emul_destroy(void* h) { delete (Emul*)h; }
void emul_set_exact(void* h, int32_t exact) { ((Emul*)h)->dev.exact = exact; }
void emul_stats(void* h, int32_t* n_nodes, int32_t* depth, double* build_ms) {
	Emul* e = (Emul*)h; *n_nodes = (int32_t)e->bvh.nodes.size(); *depth = e->bvh.depth; *build_ms = e->build_ms;
}
void emul_first_hit(void* h, const float* o, const float* d, int64_t n, int32_t* idx, float* t) {
	Emul* e = (Emul*)h;
	for (int64_t i = 0; i < n; ++i) {
		float bt; int32_t slot, bi; int2 stack[kStackEntries];
		if (e->dev.exact) traverse_warp<false, true>(e->dev, stack, 1, true, mk(o[3*i], o[3*i+1], o[3*i+2]), mk(d[3*i], d[3*i+1], d[3*i+2]), bt, bi, slot);
		else traverse_warp<false, false>(e->dev, stack, 1, true, mk(o[3*i], o[3*i+1], o[3*i+2]), mk(d[3*i], d[3*i+1], d[3*i+2]), bt, bi, slot);
		idx[i] = bi; t[i] = bt;
	}
}
void emul_occluded(void* h, const float* p, const float* x, int64_t n, uint8_t* out) {
	Emul* e = (Emul*)h;
	for (int64_t i = 0; i < n; ++i) {
		float bt; int32_t slot, bi; int2 stack[kStackEntries];
		const V3 a = mk(p[3*i], p[3*i+1], p[3*i+2]), b = mk(x[3*i], x[3*i+1], x[3*i+2]);
		if (e->dev.exact) traverse_warp<true, true>(e->dev, stack, 1, true, a, vsub(b, a), bt, bi, slot);
		else traverse_warp<true, false>(e->dev, stack, 1, true, a, vsub(b, a), bt, bi, slot);
		out[i] = (uint8_t)bi;
	}
}
}
