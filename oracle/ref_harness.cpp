/*
 * TEST INFRASTRUCTURE -- driver that links the UNMODIFIED reference objects
 * (compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/) and
 * exposes the hot path's inputs/outputs as flat binary files, so that
 *   (1) oracle/ear_oracle.cpp can be pinned bit-for-bit against the reference, and
 *   (2) bench.py --impl reference can time the reference's own CPU render.
 * Nothing under ear_b200/ links or executes this.
 *
 * It calls the reference's public classes only:
 *   Datatype/Settings (src/Datatype.cpp, src/Settings.cpp), Material, Mesh,
 *   Mono/StereoRecorder, SoundFile, Keyframes, Scene, SceneContext -- the same
 *   sequence `Render()` performs in src/EAR.cpp:55-207, minus post/convolution.
 *
 * Commands
 *   firsthit <scene.ear> <rays.bin> <out.bin>   rays: n x {ox,oy,oz,dx,dy,dz} f32
 *            out: n x {int32 tri, f32 t, f32 p[3], f32 n[3]} via Mesh::RayIntersection
 *   occluded <scene.ear> <segs.bin> <out.bin>   segs: n x {px,py,pz,xx,xy,xz} f32; out: n x u8
 *   render   <scene.ear> <seed> <out.bin> [t60] [threads=N] [budget=SECONDS]
 *            runs every SceneContext the CLI would create (EAR.cpp:170-191; with
 *            `t60` only the single calc-T60 context) and dumps the raw tracks.
 *            budget=S (single-threaded runs): after S seconds of Scene::Render the process
 *            prints the REF_RENDER line for the work done so far and exits -- a bounded
 *            throughput sample for scenes where 50 rays (the reference's minimum:
 *            DrawProgressBar divides by samples/50) take minutes of brute force.
 * rand() is seeded through a time() override: both srand(time) call sites
 * (src/EAR.cpp:58, src/Scene.cpp:116) see the requested seed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <sys/time.h>
#include <signal.h>
#include <unistd.h>
#include <string>
#include <vector>

#include <gmtl/gmtl.h>
#include <boost/thread/thread.hpp>

#include "Settings.h"
#include "Mesh.h"
#include "SoundFile.h"
#include "MonoRecorder.h"
#include "StereoRecorder.h"
#include "Recorder.h"
#include "Scene.h"
#include "Material.h"
#include "SceneContext.h"

static long g_fake_time = 12345;
extern "C" time_t time(time_t* out) {
	if (out) *out = (time_t)g_fake_time;
	return (time_t)g_fake_time;
}

static double wall_seconds() {
	struct timeval tv;
	gettimeofday(&tv, 0);
	return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

/* budgeted runs: SIGALRM reports what Scene::Render has done so far and ends the process */
static double g_t0 = 0.0;
static size_t g_triangles = 0;
static int g_contexts = 0, g_recorders = 0, g_rays = 0;
static void on_budget(int) {
	const gmtl::ShimCounters& c = gmtl::shim_counters();
	const double secs = wall_seconds() - g_t0;
	char line[512];
	const int n = snprintf(line, sizeof(line),
	    "\nREF_RENDER contexts=%d recorders=%d triangles=%zu rays_per_context=%d threads=1 seconds=%.6f "
	    "ray_tests=%llu seg_tests=%llu segments=%.0f budget=1\n",
	    g_contexts, g_recorders, g_triangles, g_rays, secs, c.ray_tests, c.seg_tests,
	    g_triangles ? (double)c.ray_tests / (double)g_triangles : 0.0);
	if (write(1, line, n) < 0) {}
	_exit(0);
}

struct LoadedScene {
	Scene* scene;
	gmtl::Vec3f air;
	float dry;
	int samples;
	int maxthreads;
};

/* Same block dispatch as src/EAR.cpp:60-119, written against the public API. */
static bool load_scene(const char* path, LoadedScene& out, bool want_sources) {
	out.scene = new Scene();
	if (!Datatype::SetInput(path)) { fprintf(stderr, "cannot read %s\n", path); return false; }
	Datatype* set = Datatype::Scan("SET ");
	if (!set) { fprintf(stderr, "no SET block\n"); return false; }
	Settings::init(set);
	delete set;
	out.air = Settings::GetVec("absorption");
	out.dry = Settings::GetFloat("drylevel");
	out.samples = Settings::GetInt("samples") / 10;
	out.maxthreads = Settings::IsSet("maxthreads") ? Settings::GetInt("maxthreads") : -1;
	while (Datatype::input_length) {
		const std::string id = Datatype::PeakId();
		if (id == "OUT1") out.scene->addListener(new MonoRecorder());
		else if (id == "OUT2") out.scene->addListener(new StereoRecorder());
		else if (id == "SSRC" && want_sources) out.scene->addSoundSource(new SoundFile());
		else if (id == "3SRC" && want_sources) out.scene->addSoundSource(new TripleBandSoundFile());
		else if (id == "MESH") out.scene->addMesh(new Mesh());
		else if (id == "MAT ") out.scene->addMaterial(new Material());
		else if (id == "KEYS") Keyframes::Init();
		else delete Datatype::Read();
	}
	if (out.scene->meshes.empty()) out.scene->addMesh(Mesh::Empty());
	return true;
}

static std::vector<float> read_floats(const char* path) {
	std::vector<float> v;
	FILE* f = fopen(path, "rb");
	if (!f) return v;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	v.resize(n / 4);
	if (n && fread(&v[0], 4, v.size(), f) != v.size()) v.clear();
	fclose(f);
	return v;
}

static int cmd_firsthit(const char* scene_path, const char* in_path, const char* out_path) {
	LoadedScene ls;
	if (!load_scene(scene_path, ls, false)) return 1;
	Mesh* mesh = ls.scene->meshes[0];
	std::vector<float> rays = read_floats(in_path);
	const size_t n = rays.size() / 6;
	FILE* out = fopen(out_path, "wb");
	size_t disagreements = 0;
	for (size_t i = 0; i < n; ++i) {
		const float* r = &rays[6 * i];
		gmtl::Rayf ray(gmtl::Point3f(r[0], r[1], r[2]), gmtl::Vec3f(r[3], r[4], r[5]));
		gmtl::Point3f* p = 0;
		gmtl::Vec3f* nrm = 0;
		Material* mat = 0;
		const bool hit = mesh->RayIntersection(&ray, p, nrm, mat);
		/* The reference API does not return the winning index or t; recover them by
		   asking each triangle the same question and matching the returned point. */
		int32_t idx = -1;
		float best = 1000000;
		for (size_t k = 0; k < mesh->tris.size(); ++k) {
			float u, v, t;
			if (gmtl::shim_moeller_trumbore(*mesh->tris[k], ray, u, v, t) && t > 0.001f && t < best) {
				best = t;
				idx = (int32_t)k;
			}
		}
		float rec[7] = {0, 0, 0, 0, 0, 0, 0};
		if (hit) {
			const gmtl::Point3f q = ray.mOrigin + ray.mDir * best;
			if (idx < 0 || memcmp(q.mData, p->mData, 12) != 0) ++disagreements;
			rec[0] = best;
			rec[1] = (*p)[0]; rec[2] = (*p)[1]; rec[3] = (*p)[2];
			rec[4] = (*nrm)[0]; rec[5] = (*nrm)[1]; rec[6] = (*nrm)[2];
		} else if (idx >= 0) ++disagreements;
		fwrite(&idx, 4, 1, out);
		fwrite(rec, 4, 7, out);
		delete p;
		delete nrm;
	}
	fclose(out);
	fprintf(stderr, "firsthit: %zu rays, %zu triangles, %zu harness/reference disagreements\n",
	        n, mesh->tris.size(), disagreements);
	return disagreements ? 2 : 0;
}

static int cmd_occluded(const char* scene_path, const char* in_path, const char* out_path) {
	LoadedScene ls;
	if (!load_scene(scene_path, ls, false)) return 1;
	Mesh* mesh = ls.scene->meshes[0];
	std::vector<float> segs = read_floats(in_path);
	const size_t n = segs.size() / 6;
	FILE* out = fopen(out_path, "wb");
	for (size_t i = 0; i < n; ++i) {
		const float* s = &segs[6 * i];
		gmtl::LineSegf seg(gmtl::Point3f(s[0], s[1], s[2]), gmtl::Point3f(s[3], s[4], s[5]));
		const uint8_t occ = mesh->LineIntersection(&seg) ? 1 : 0;
		fwrite(&occ, 1, 1, out);
	}
	fclose(out);
	return 0;
}

static int cmd_render(const char* scene_path, long seed, const char* out_path, bool t60_only, int threads, int budget) {
	g_fake_time = seed;
	gmtl::Math::seedRandom((unsigned int)time(0)); /* src/EAR.cpp:58 */
	LoadedScene ls;
	if (!load_scene(scene_path, ls, true)) return 1;
	Scene* scene = ls.scene;
	if (scene->sources.empty() || scene->listeners.empty()) { fprintf(stderr, "no source/listener\n"); return 1; }
	Keyframes* keys = Keyframes::Get();
	std::vector<SceneContext> ctxs;
	for (unsigned s = 0; s < scene->sources.size(); ++s) {
		const int kf_begin = keys ? 0 : -1;
		const int kf_end = keys ? (int)keys->keys.size() : 0;
		for (int kf = kf_begin; kf < kf_end; ++kf) {
			for (int band = 0; band < 3; ++band) {
				if (t60_only && band != 1) continue;
				ctxs.push_back(SceneContext(scene, band, (int)s, ls.samples, 1.0f - ls.air[band], ls.dry, kf));
			}
			if (t60_only) break;
		}
		if (t60_only) break;
	}
	const size_t T = scene->meshes[0]->tris.size();
	g_triangles = T; g_contexts = (int)ctxs.size(); g_recorders = (int)scene->listeners.size(); g_rays = ls.samples;
	fflush(stdout);
	const double t0 = wall_seconds();
	g_t0 = t0;
	if (budget > 0 && threads <= 1) { signal(SIGALRM, on_budget); alarm((unsigned)budget); }
	if (threads <= 1) {
		for (size_t i = 0; i < ctxs.size(); ++i) ctxs[i]();
		boost::shim_totals::fold();
	} else {
		size_t next = 0;
		while (next < ctxs.size()) {
			boost::thread_group group;
			for (int k = 0; k < threads && next < ctxs.size(); ++k) group.create_thread(ctxs[next++]);
			group.join_all();
		}
	}
	const double t1 = wall_seconds();
	alarm(0);
	const unsigned long long ray_tests = boost::shim_totals::ray_tests();
	const unsigned long long seg_tests = boost::shim_totals::seg_tests();
	FILE* out = fopen(out_path, "wb");
	const int32_t n_ctx = (int32_t)ctxs.size();
	const int32_t n_rec = (int32_t)scene->listeners.size();
	fwrite(&n_ctx, 4, 1, out);
	fwrite(&n_rec, 4, 1, out);
	unsigned long long bins = 0;
	for (int32_t c = 0; c < n_ctx; ++c) {
		int32_t hdr[3] = {ctxs[c].band, ctxs[c].soundfile_id, ctxs[c].keyframe_id};
		fwrite(hdr, 4, 3, out);
		for (int32_t r = 0; r < n_rec; ++r) {
			Recorder* rec = ctxs[c].recorders[r];
			const int32_t n_tracks = rec->trackCount();
			fwrite(&n_tracks, 4, 1, out);
			for (int32_t k = 0; k < n_tracks; ++k) {
				const RecorderTrack& tr = *rec->tracks[k];
				const uint32_t first = tr.first_sample, real = tr.real_length;
				fwrite(&first, 4, 1, out);
				fwrite(&real, 4, 1, out);
				for (uint32_t i = 0; i <= real; ++i) { const float v = tr[i]; fwrite(&v, 4, 1, out); }
				(void)bins;
			}
		}
	}
	fclose(out);
	const double segments = T ? (double)ray_tests / (double)T : 0.0;
	/* one machine-readable line for bench.py / tests */
	printf("\nREF_RENDER contexts=%d recorders=%d triangles=%zu rays_per_context=%d threads=%d seconds=%.6f "
	       "ray_tests=%llu seg_tests=%llu segments=%.0f\n",
	       n_ctx, n_rec, T, ls.samples, threads, t1 - t0, ray_tests, seg_tests, segments);
	return 0;
}

int main(int argc, char** argv) {
	std::cout << std::setprecision(3) << std::fixed;
	if (argc >= 5 && !strcmp(argv[1], "firsthit")) return cmd_firsthit(argv[2], argv[3], argv[4]);
	if (argc >= 5 && !strcmp(argv[1], "occluded")) return cmd_occluded(argv[2], argv[3], argv[4]);
	if (argc >= 5 && !strcmp(argv[1], "render")) {
		bool t60 = false;
		int threads = 1, budget = 0;
		for (int i = 5; i < argc; ++i) {
			if (!strcmp(argv[i], "t60")) t60 = true;
			else if (!strncmp(argv[i], "threads=", 8)) threads = atoi(argv[i] + 8);
			else if (!strncmp(argv[i], "budget=", 7)) budget = atoi(argv[i] + 7);
		}
		return cmd_render(argv[2], atol(argv[3]), argv[4], t60, threads, budget);
	}
	fprintf(stderr, "usage: ref_harness firsthit|occluded|render ...\n");
	return 64;
}
