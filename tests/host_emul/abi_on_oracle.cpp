// TEST INFRASTRUCTURE -- the slice of include/ear_b200.h that the two CLIs call (this repo's EAR and the reference
// linked through INTEGRATION.md's binding, oracle/_ref/EAR_ref_gpu), implemented on the CPU oracle.  LD_PRELOADed by
// tests/test_host_surface.py so that the HOST surface of both binaries -- .ear parsing, band split, Power / Truncate /
// T60, convolution bookkeeping, merge / normalise, WAV writing -- can be compared on a machine without a GPU: both
// receive identical tracks from this shim.  Never built into, linked with or loaded by the product.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ear_b200.h"

extern "C" {
void* oracle_scene_create(const float*, const int32_t*, int32_t, const float*, int32_t, int32_t);
void oracle_scene_destroy(void*);
void oracle_scene_set_emitters(void*, const float*, int32_t);
void* oracle_render(void*, const ear_b200_context*, int32_t, const ear_b200_recorder*, int32_t, int32_t, int32_t, uint64_t, int64_t,
                    int64_t, int32_t);
void oracle_render_free(void*);
void oracle_render_counters(void*, uint64_t*);
void oracle_render_track_info(void*, int32_t, int32_t, int32_t, uint32_t*, uint32_t*, uint32_t*);
void oracle_render_track_copy(void*, int32_t, int32_t, int32_t, float*, uint32_t);
void oracle_convolve(const float*, uint32_t, uint32_t, uint32_t, const float*, uint32_t, uint32_t, uint32_t, const float*, uint32_t,
                     uint32_t, float*);
}

struct ear_b200_scene { void* o; };
struct ear_b200_group { ear_b200_scene* s; };
static std::string g_err;

extern "C" {
const char* ear_b200_last_error(void) { return g_err.c_str(); }
int32_t ear_b200_abi_version(void) { return EAR_B200_ABI_VERSION; }
int32_t ear_b200_device_count(void) { return 1; }
int32_t ear_b200_scene_create(const float* verts, const int32_t* tri_material, int32_t n_tris, const float* materials,
                              int32_t n_materials, int32_t n_bands, int32_t, ear_b200_scene** out) {
	*out = new ear_b200_scene{oracle_scene_create(verts, tri_material, n_tris, materials, n_materials, n_bands)};
	return 0;
}
int32_t ear_b200_scene_set_emitters(ear_b200_scene* s, const float* verts, int32_t n) { oracle_scene_set_emitters(s->o, verts, n); return 0; }
void ear_b200_scene_destroy(ear_b200_scene* s) { if (s) { oracle_scene_destroy(s->o); delete s; } }
int32_t ear_b200_group_create(ear_b200_scene* s, const int32_t*, int32_t, ear_b200_group** out) { *out = new ear_b200_group{s}; return 0; }
void ear_b200_group_destroy(ear_b200_group* g) { delete g; }
int32_t ear_b200_group_size(ear_b200_group*) { return 1; }
int32_t ear_b200_render(ear_b200_scene* s, const ear_b200_context* ctx, int32_t n_ctx, const ear_b200_recorder* rec, int32_t n_rec,
                        const ear_b200_options* opt, ear_b200_result** out) {
	void* h = oracle_render(s->o, ctx, n_ctx, rec, n_rec, opt->max_bounces, /*PHILOX*/ 1, opt->seed, opt->first_ray, opt->ray_count, opt->finalise);
	ear_b200_result* r = (ear_b200_result*)calloc(1, sizeof(ear_b200_result));
	r->n_contexts = n_ctx; r->n_recorders = n_rec;
	r->tracks = (ear_b200_track*)calloc((size_t)n_ctx * n_rec * 2, sizeof(ear_b200_track));
	for (int32_t c = 0; c < n_ctx; ++c)
		for (int32_t k = 0; k < n_rec; ++k)
			for (int t = 0; t < (rec[c * n_rec + k].kind == EAR_B200_STEREO ? 2 : 1); ++t) {
				ear_b200_track& tr = r->tracks[((size_t)c * n_rec + k) * 2 + t];
				oracle_render_track_info(h, c, k, t, &tr.first_sample, &tr.real_length, &tr.length);
				tr.data = (float*)calloc(tr.length, sizeof(float));
				oracle_render_track_copy(h, c, k, t, tr.data, tr.length);
			}
	uint64_t cnt[5];
	oracle_render_counters(h, cnt);
	r->rays = cnt[0]; r->segments = cnt[1]; r->occlusion_queries = cnt[2]; r->contributions = cnt[3]; r->bin_updates = cnt[4];
	oracle_render_free(h);
	*out = r;
	return 0;
}
int32_t ear_b200_group_render(ear_b200_group* g, const ear_b200_context* ctx, int32_t n_ctx, const ear_b200_recorder* rec, int32_t n_rec,
                              const ear_b200_options* opt, ear_b200_result** out) {
	return ear_b200_render(g->s, ctx, n_ctx, rec, n_rec, opt, out);
}
void ear_b200_result_free(ear_b200_result* r) {
	if (!r) return;
	if (r->tracks) { for (size_t k = 0; k < (size_t)r->n_contexts * r->n_recorders * 2; ++k) free(r->tracks[k].data); free(r->tracks); }
	free(r->t60);
	free(r);
}
int32_t ear_b200_convolve(int32_t, const float* response, uint32_t length, uint32_t first_sample, uint32_t real_length,
                          const float* response2, uint32_t length2, uint32_t first_sample2, uint32_t real_length2, const float* dry,
                          uint32_t n_dry, uint32_t offset, float* out, uint32_t out_len, uint32_t* out_first, uint32_t* out_real) {
	const bool fade = response2 != nullptr;
	const uint32_t first = fade ? (first_sample < first_sample2 ? first_sample : first_sample2) : first_sample;
	const uint32_t len = fade ? (real_length > real_length2 ? real_length : real_length2) : real_length;
	const uint32_t init_first = 3 * EAR_B200_SAMPLE_RATE - 1;
	if (out_first) *out_first = init_first;
	if (out_real) *out_real = 0;
	memset(out, 0, (size_t)out_len * sizeof(float));
	if (n_dry == 0 || len <= first) return 0;
	const unsigned long long last = (unsigned long long)(n_dry - 1) + offset + (len - 1);
	if (out_first) *out_first = init_first < offset + first ? init_first : offset + first;
	if (out_real) *out_real = (uint32_t)last;
	if (last + 1 > out_len) { g_err = "convolve: output buffer too short"; return 1; }
	if (fade) {
		// the oracle walks [min first, max len) of both tracks itself
		oracle_convolve(response, length, first_sample, real_length, response2, length2, first_sample2, real_length2, dry, n_dry, offset, out);
	} else oracle_convolve(response, length, first_sample, real_length, nullptr, 0, 0, 0, dry, n_dry, offset, out);
	return 0;
}
}
