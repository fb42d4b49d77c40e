// TEST INFRASTRUCTURE: host build of the product's BVH builder + traversal headers (see cuda_shim.h).
// Exposes first-hit / occlusion over flat arrays so tests can compare them with the oracle here.
#include "cuda_shim.h"
#include "../../ear_b200/csrc/bvh_build.h"
#include "../../ear_b200/csrc/traverse.cuh"
#include "../../ear_b200/csrc/vismap_geom.cuh"
#include <vector>
#include <cstdlib>
#include <chrono>
using namespace earb;
struct Emul { Bvh bvh; SceneDev dev; double build_ms; };
extern "C" {
void* emul_create(const float* verts, const int32_t* mats, int32_t n) {
	Emul* e = new Emul();
	if (const char* k = getenv("EAR_B200_BVH_MARGIN_SCALE")) g_emul_decode_bias = 0.00390625f * (float)atof(k);   // the margins' test knob covers the decode bias too
	auto t0 = std::chrono::steady_clock::now();
	build_bvh(verts, mats, n, e->bvh);
	e->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	e->dev.nodes = (const float4*)e->bvh.nodes.data();
	e->dev.tris = (const float4*)e->bvh.tris.data();
	e->dev.materials = nullptr; e->dev.n_tris = n; e->dev.n_materials = 0; e->dev.n_bands = 0;
	e->dev.s0 = e->bvh.s0; e->dev.exact = 0; e->dev.leaf_vote = kLeafVote; e->dev.fetch_vote = 8; e->dev.vis_cap = 64;
	e->dev.emitters = nullptr; e->dev.spill = nullptr; e->dev.spill_threads = 1; e->dev.spill_rows = kStackSpill;
	return e;
}
void emul_destroy(void* h) { delete (Emul*)h; }
void emul_set_exact(void* h, int32_t exact) { ((Emul*)h)->dev.exact = exact; }
void emul_stats(void* h, int32_t* n_nodes, int32_t* depth, double* build_ms) {
	Emul* e = (Emul*)h; *n_nodes = (int32_t)e->bvh.nodes.size(); *depth = e->bvh.depth; *build_ms = e->build_ms;
}
// FNV-1a over the node and triangle-record arrays: the layout the GPU would be handed
uint64_t emul_hash(void* h) {
	Emul* e = (Emul*)h;
	uint64_t x = 1469598103934665603ull;
	const unsigned char* p = (const unsigned char*)e->bvh.nodes.data();
	for (size_t i = 0; i < e->bvh.nodes.size() * sizeof(Node); ++i) { x ^= p[i]; x *= 1099511628211ull; }
	p = (const unsigned char*)e->bvh.tris.data();
	for (size_t i = 0; i < e->bvh.tris.size() * sizeof(TriRecord); ++i) { x ^= p[i]; x *= 1099511628211ull; }
	return x;
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

// Visibility-map superset property (vismap_geom.cuh): for each point P, every triangle the reference's float test
// accepts on the segment P -> X (1e-5 < t < 1) must be among the candidates of P's texel.  Returns the number of
// accepted triangles that are NOT candidates (must be 0); *accepted = accepted triangles, *listed = candidates seen.
int64_t emul_vismap_violations(void* h, const float* X, int32_t res, const float* P, int64_t n, int64_t* accepted, int64_t* listed) {
	Emul* e = (Emul*)h;
	const int T = e->dev.n_tris;
	float maxabs = 0.0f;
	for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::max(std::fabs(e->bvh.lo[k]), std::fabs(e->bvh.hi[k])));
	const double reach = 2.0 * (double)e->bvh.diagonal + 1.0;
	const double Xd[3] = {X[0], X[1], X[2]};
	struct Foot { int i0, i1, j0, j1; float edge[9]; bool valid, has_edges; };
	std::vector<Foot> foot((size_t)T * 6);
	for (int t = 0; t < T; ++t)
		for (int f = 0; f < 6; ++f) {
			Foot& ft = foot[(size_t)t * 6 + f];
			ft.valid = vis_footprint(e->dev, t, Xd, res, f, reach, (double)maxabs, ft.i0, ft.i1, ft.j0, ft.j1, ft.edge, ft.has_edges);
		}
	VisMapDev mp;
	mp.offsets = nullptr; mp.items = nullptr; mp.res = res;
	for (int k = 0; k < 3; ++k) mp.x[k] = X[k];
	int64_t bad = 0, acc = 0, lst = 0;
	const V3 x = mk(X[0], X[1], X[2]);
	for (int64_t q = 0; q < n; ++q) {
		const V3 pnt = mk(P[3 * q], P[3 * q + 1], P[3 * q + 2]);
		const int texel = vis_texel(mp, pnt.x - x.x, pnt.y - x.y, pnt.z - x.z);
		const int f = texel / (res * res), j = (texel / res) % res, i = texel % res;
		const V3 d = vsub(x, pnt);
		for (int t = 0; t < T; ++t) {
			const Foot& ft = foot[(size_t)t * 6 + f];
			const bool in_list = ft.valid && i >= ft.i0 && i <= ft.i1 && j >= ft.j0 && j <= ft.j1 && (!ft.has_edges || vis_covers(ft.edge, res, i, j));
			lst += in_list ? 1 : 0;
			const float4* rec = e->dev.tris + 4 * (size_t)t;
			float tt;
			const bool hit = moeller_trumbore(mk(rec[0].x, rec[0].y, rec[0].z), mk(rec[1].x, rec[1].y, rec[1].z), mk(rec[2].x, rec[2].y, rec[2].z), pnt, d, tt) &&
			                 tt > 1e-5f && tt < 1.0f;
			if (hit) { ++acc; if (!in_list) ++bad; }
		}
	}
	*accepted = acc; *listed = lst;
	return bad;
}
}
