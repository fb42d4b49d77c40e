// TEST INFRASTRUCTURE -- the reference-side binding of INTEGRATION.md, compiled for real.
//
// oracle/build_ref_gpu.sh links this file with the reference's own objects (oracle/_ref/obj/*.o, built from
// /root/reference by oracle/build_ref.sh) and libear_b200.so into oracle/_ref/EAR_ref_gpu: the reference's `main`,
// parser, band split, post chain, convolution, merge and WAV writer, with ONE change -- the boost::thread fan-out of
// Scene::Render (src/EAR.cpp:196-207) is replaced by a call to RenderContextsOnGpu() below.  tests/test_dropin_gpu.py
// then checks that `EAR_ref_gpu` and this repo's own `EAR` print the same T60 and write identical WAV files for the
// same seed: the drop-in proof, and a parity test of this repo's host surface against the reference's.
//
// Compiled with -Dprivate=public -Dprotected=public: StereoRecorder keeps right_ear / head_size / head_absorption
// private and AbstractSoundFile keeps its emitter mesh protected (a maintainer would add three accessors instead).
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <stdexcept>
#include <vector>

#include "Scene.h"
#include "SceneContext.h"
#include "StereoRecorder.h"
#include "Triangle.h"
#include "Material.h"
#include "Mesh.h"

#include "ear_b200.h"

void RenderContextsOnGpu(Scene* scene, std::vector<SceneContext>& scs) {
	// triangles in meshes[0]->tris order == file order == triangle index (src/Scene.cpp:103-106)
	Mesh* mesh = scene->meshes[0];
	std::map<Material*, int> mat_id;
	std::vector<float> table;                       // [M][3][4] refl, refr, kept, spec
	std::vector<float> verts;
	std::vector<int32_t> tri_mat;
	for (size_t i = 0; i < mesh->tris.size(); ++i) {
		Triangle* t = mesh->tris[i];
		if (!mat_id.count(t->m)) {
			const int id = (int)mat_id.size();
			mat_id[t->m] = id;
			for (int b = 0; b < 3; ++b) {
				table.push_back(t->m->reflection_coefficient[b]);
				table.push_back(t->m->refraction_coefficient[b]);
				table.push_back(t->m->absorption_coefficient[b]);   // surviving fraction (src/Material.cpp:59)
				table.push_back(t->m->specularity_coefficient[b]);
			}
		}
		tri_mat.push_back(mat_id[t->m]);
		for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) verts.push_back(t->mVerts[v][k]);
	}
	if (table.empty()) table.assign(12, 0.0f);
	ear_b200_scene* gpu = 0;
	if (ear_b200_scene_create(verts.empty() ? 0 : &verts[0], tri_mat.empty() ? 0 : &tri_mat[0], (int32_t)tri_mat.size(), &table[0],
	                          (int32_t)(mat_id.empty() ? 1 : mat_id.size()), 3, /*device*/ 0, &gpu))
		throw std::runtime_error(ear_b200_last_error());          // main prints "Error: <what>" (src/EAR.cpp:404-408)

	// emitter triangles of mesh sources, all of them concatenated (src/SoundFile.cpp:50-53)
	std::vector<float> emit;
	std::map<AbstractSoundFile*, std::pair<int, int> > emit_range;
	for (size_t s = 0; s < scene->sources.size(); ++s) {
		AbstractSoundFile* sf = scene->sources[s];
		if (!sf->mesh) continue;
		const int first = (int)(emit.size() / 9);
		for (size_t i = 0; i < sf->mesh->tris.size(); ++i)
			for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) emit.push_back(sf->mesh->tris[i]->mVerts[v][k]);
		emit_range[sf] = std::make_pair(first, (int)(emit.size() / 9) - first);
	}
	if (!emit.empty() && ear_b200_scene_set_emitters(gpu, &emit[0], (int32_t)(emit.size() / 9)))
		throw std::runtime_error(ear_b200_last_error());

	const int R = (int)scene->listeners.size();
	std::vector<ear_b200_context> ctx(scs.size());
	std::vector<ear_b200_recorder> rec(scs.size() * R);
	for (size_t c = 0; c < scs.size(); ++c) {
		AbstractSoundFile* sf = scene->sources[scs[c].soundfile_id];
		memset(&ctx[c], 0, sizeof(ctx[c]));
		ctx[c].band = scs[c].band;
		ctx[c].stream_id = (int32_t)c + 1;     // what this repo's EAR passes: streams keyed by the global context index
		ctx[c].num_samples = scs[c].samples;
		ctx[c].absorption_factor = scs[c].absorption;
		ctx[c].dry_level = scs[c].dry_level;
		ctx[c].gain = sf->getGain();
		if (sf->mesh) {
			ctx[c].source_kind = EAR_B200_MESH_SOURCE;
			ctx[c].emitter_first = emit_range[sf].first;
			ctx[c].emitter_count = emit_range[sf].second;
		} else {
			const gmtl::Point3f p = sf->getLocation(scs[c].keyframe_id);
			for (int k = 0; k < 3; ++k) ctx[c].source_position[k] = p[k];
		}
		for (int r = 0; r < R; ++r) {
			Recorder* l = scs[c].recorders[r];
			ear_b200_recorder& o = rec[c * R + r];
			memset(&o, 0, sizeof(o));
			o.kind = l->trackCount() == 2 ? EAR_B200_STEREO : EAR_B200_MONO;
			const gmtl::Point3f& lp = l->getLocation(scs[c].keyframe_id);
			for (int k = 0; k < 3; ++k) o.position[k] = lp[k];
			if (StereoRecorder* s = dynamic_cast<StereoRecorder*>(l)) {
				const gmtl::Vec3f& e = s->getRightEar(scs[c].keyframe_id);
				for (int k = 0; k < 3; ++k) o.right_ear[k] = e[k];
				for (int k = 0; k < EAR_B200_MAX_BANDS; ++k) o.head_absorption[k] = s->head_absorption[k < 3 ? k : 2];
				o.head_size = s->head_size;
			}
		}
	}
	ear_b200_options opt;
	memset(&opt, 0, sizeof(opt));
	opt.max_bounces = getenv("EAR_MAX_BOUNCES") ? atoi(getenv("EAR_MAX_BOUNCES")) : 1000;
	opt.seed = getenv("EAR_SEED") ? strtoull(getenv("EAR_SEED"), 0, 10) : (uint64_t)time(0);
	opt.ray_count = -1;
	opt.finalise = 1;
	ear_b200_result* res = 0;
	if (ear_b200_render(gpu, &ctx[0], (int32_t)ctx.size(), &rec[0], R, &opt, &res))
		throw std::runtime_error(ear_b200_last_error());
	// results come back the way Scene::Render returns them: by filling the contexts' recorder tracks.  Writing
	// through the non-const FloatBuffer::operator[] reproduces the reference's own bookkeeping of first_sample /
	// real_length (src/Recorder.cpp:52-59); untouched tracks stay exactly as the reference leaves them.
	for (size_t c = 0; c < scs.size(); ++c)
		for (int r = 0; r < R; ++r)
			for (int k = 0; k < scs[c].recorders[r]->trackCount(); ++k) {
				const ear_b200_track& t = res->tracks[(c * R + r) * 2 + k];
				RecorderTrack& dst = *scs[c].recorders[r]->tracks[k];
				if (t.real_length == 0 && t.first_sample != 0) continue;   // no sample ever recorded
				dst[t.first_sample] = t.data[t.first_sample];              // min touched index
				dst[t.real_length] = t.data[t.real_length];                // max touched index (grows the buffer once)
				for (uint32_t i = t.first_sample; i <= t.real_length; ++i) dst[i] = t.data[i];
				scs[c].recorders[r]->has_samples = true;
			}
	ear_b200_result_free(res);
	ear_b200_scene_destroy(gpu);
}
