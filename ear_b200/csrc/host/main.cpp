// EAR -- command line of the B200-native render path.  Same verbs and output contract as the
// reference's main() (src/EAR.cpp:395-431):
//     EAR render <file>        trace every (sound, keyframe, band) context, convolve, write the wavs
//     EAR calc T60 <file>      trace the mid band of the first sound, print T60_ear / Sabine / Eyring
//     EAR test                 exit 0 (the Blender add-on probes the binary with it)
// The orchestration mirrors Render() (src/EAR.cpp:55-393); the thread fan-out over SceneContexts
// (src/EAR.cpp:196-207) is one ear_b200_render() call per GPU.  Host post-processing (Power,
// Truncate, T60, convolution, merge) is kept on the CPU exactly as the reference has it.
//
// Differences that are deliberate: rays use the library's Philox streams (seed from the clock, or
// EAR_SEED); `render` does not wait for a key press unless EAR_WAIT_KEY=1 is set (the reference
// blocks on stdin, src/EAR.cpp:409-410); max_bounces can be lowered with EAR_MAX_BOUNCES.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <thread>

#include "../../../include/ear_b200.h"
#include "audio.h"
#include "ear_format.h"
#include "tracks.h"

using namespace earhost;

namespace {

const char* kBanner =
    "  ______                      _____  \n |  ____|         /\\         |  __ \\ \n | |__           /  \\        | |__) | \n"
    " |  __|         / /\\ \\       |  _  / \n | |____       / ____ \\      | | \\ \\ \n |______| (_) /_/    \\_\\ (_) |_|  \\_\\\n"
    " Evaluation of Acoustics using Ray-tracing\n ear_b200: B200-native render path (ABI v3)";

struct Context {
	int sound, keyframe, band;
	std::vector<std::unique_ptr<Track>> tracks;      // [listener][2]
	std::vector<std::unique_ptr<Track>> processed;   // [listener][2]
};

struct Sound {
	std::vector<float> band[3];
};

void fail_abi(const char* what) { throw std::runtime_error(std::string(what) + ": " + ear_b200_last_error()); }

std::string file_name(const std::string& p) {
	const size_t pos = p.find_last_of("\\/");
	if (pos == std::string::npos) throw std::runtime_error("Failed to interpret file path");
	return p.substr(pos + 1);
}

// Mesh::Volume / Area / TotalAbsorption / AverageAbsorption (src/Mesh.cpp:156-176) over the merged
// mesh.  Mesh::Combine adds areas but forgets the weighted area (src/Mesh.cpp:116-123): only the first
// MESH block contributes absorption -- kept for output parity.
void sabine_eyring(const SceneFile& sf, float air_mid, float* sabine, float* eyring) {
	float V = 0.0f, S = 0.0f, A = 0.0f;
	for (int i = 0; i < sf.triangle_count(); ++i) {
		const float* p = &sf.vertices[9 * (size_t)i];
		const float v321 = p[6] * p[4] * p[2], v231 = p[3] * p[7] * p[2], v312 = p[6] * p[1] * p[5];
		const float v132 = p[0] * p[7] * p[5], v213 = p[3] * p[1] * p[8], v123 = p[0] * p[4] * p[8];
		V += (1.0f / 6.0f) * (-v321 + v231 + v312 - v132 - v213 + v123);
	}
	for (size_t m = 0; m < sf.meshes.size(); ++m) {
		const float absorption = 1.0f - sf.materials[sf.meshes[m].material].kept[1];
		float area = 0.0f, weighted = 0.0f;
		for (int i = sf.meshes[m].first_triangle; i < sf.meshes[m].first_triangle + sf.meshes[m].triangle_count; ++i) {
			const float* p = &sf.vertices[9 * (size_t)i];
			const float e0[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
			const float e1[3] = {p[6] - p[3], p[7] - p[4], p[8] - p[5]};
			const float cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
			float l2 = cx * cx; l2 = l2 + cy * cy; l2 = l2 + cz * cz;
			const float a = sqrtf(l2) / 2.0f;
			area += a;
			weighted += a * absorption;
		}
		if (m == 0) { S = area; A = weighted; } else S += area;
	}
	const float a = A / S;
	*sabine = 0.1611f * V / (A + 4.0f * air_mid * V);
	*eyring = 0.1611f * V / (-S * logf(1.0f - a) + 4.0f * air_mid * V);
}

int run(const std::string& filename, float* calc_t60, float* t60_sabine, float* t60_eyring) {
	SceneFile sf = load_scene_file(filename, &std::cout);
	const std::array<float, 3> air = sf.get_vec("absorption");
	const float dry_level = sf.get_float("drylevel");
	const int num_samples = sf.get_int("samples") / 10;
	const bool has_debugdir = sf.is_set("debugdir");
	const std::string debugdir = has_debugdir ? sf.get_string("debugdir") + "/" : "";
	const char* lomihi[] = {"low", "mid", "high"};

	if (sf.sources.empty()) { std::cout << std::endl << "No sound sources defined" << std::endl << std::endl; return 1; }
	if (sf.listeners.empty()) { std::cout << std::endl << "No listeners defined" << std::endl << std::endl; return 1; }
	if (sf.triangle_count() == 0) std::cout << std::endl << "Warning: no reflective geometry" << std::endl << std::endl;

	// dry signals: the reference loads every wav while parsing and fails if one is missing
	std::vector<Sound> sounds(sf.sources.size());
	for (size_t s = 0; s < sf.sources.size(); ++s) {
		const Source& src = sf.sources[s];
		if (src.wavs.size() == 1) {
			std::vector<float> dry = load_wav_mono(src.wavs[0]);
			if (dry.empty()) throw std::runtime_error("Failed to open sound file " + src.wavs[0]);
			(void)file_name(src.wavs[0]);
			if (!calc_t60) split_bands(dry, sf.freq[0] * 1000.0f, sf.freq[1] * 1000.0f, sf.freq[2] * 1000.0f, sounds[s].band[0], sounds[s].band[1], sounds[s].band[2]);
		} else {
			for (int b = 0; b < 3; ++b) {
				sounds[s].band[b] = load_wav_mono(src.wavs[b]);
				(void)file_name(src.wavs[b]);
			}
		}
	}

	std::cout << "Rendering..." << std::endl;
	// contexts: sound x keyframe x band (src/EAR.cpp:170-191)
	std::vector<Context> ctxs;
	for (size_t s = 0; s < sf.sources.size(); ++s) {
		const int kf_begin = sf.has_keys ? 0 : -1, kf_end = sf.has_keys ? (int)sf.keys.size() : 0;
		for (int kf = kf_begin; kf < kf_end; ++kf) {
			for (int band = 0; band < 3; ++band) {
				if (calc_t60 && band != 1) continue;
				Context c; c.sound = (int)s; c.keyframe = kf; c.band = band;
				ctxs.push_back(std::move(c));
			}
			if (calc_t60) break;
		}
		if (calc_t60) break;
	}
	const int n_ctx = (int)ctxs.size(), n_rec = (int)sf.listeners.size();
	std::vector<ear_b200_context> cc((size_t)n_ctx);
	std::vector<ear_b200_recorder> rr((size_t)n_ctx * n_rec);
	for (int c = 0; c < n_ctx; ++c) {
		const Source& src = sf.sources[ctxs[c].sound];
		std::memset(&cc[c], 0, sizeof(cc[c]));
		cc[c].band = ctxs[c].band;
		cc[c].stream_id = c + 1;   // random streams keyed by the global context index: the result does not depend on how
		                           // the contexts are dealt to GPUs
		cc[c].num_samples = num_samples;
		cc[c].absorption_factor = 1.0f - air[ctxs[c].band];
		cc[c].dry_level = dry_level;
		cc[c].gain = src.gain;
		const std::array<float, 3>& sp = src.location.at(ctxs[c].keyframe);
		for (int k = 0; k < 3; ++k) cc[c].source_position[k] = sp[k];
		for (int r = 0; r < n_rec; ++r) {
			const Listener& l = sf.listeners[r];
			ear_b200_recorder& rec = rr[(size_t)c * n_rec + r];
			std::memset(&rec, 0, sizeof(rec));
			rec.kind = l.stereo ? EAR_B200_STEREO : EAR_B200_MONO;
			const std::array<float, 3>& lp = l.location.at(ctxs[c].keyframe);
			const std::array<float, 3>& re = l.right_ear.at(ctxs[c].keyframe);
			for (int k = 0; k < 3; ++k) { rec.position[k] = lp[k]; rec.right_ear[k] = re[k]; }
			rec.head_size = l.head_size;
			for (int k = 0; k < EAR_B200_MAX_BANDS; ++k) rec.head_absorption[k] = l.head_absorption[k < 3 ? k : 2];
		}
	}
	std::vector<float> table((size_t)std::max<size_t>(sf.materials.size(), 1) * 3 * 4, 0.0f);
	for (size_t m = 0; m < sf.materials.size(); ++m)
		for (int b = 0; b < 3; ++b) {
			float* row = &table[(m * 3 + b) * 4];
			row[0] = sf.materials[m].refl[b]; row[1] = sf.materials[m].refr[b]; row[2] = sf.materials[m].kept[b]; row[3] = sf.materials[m].spec[b];
		}

	// one library call per GPU; contexts are independent, so they are dealt round-robin over devices
	int n_gpus = std::max(1, ear_b200_device_count());
	if (const char* e = std::getenv("EAR_GPUS")) n_gpus = std::max(1, std::min(n_gpus, std::atoi(e)));
	n_gpus = std::min(n_gpus, n_ctx);
	ear_b200_options opt;
	std::memset(&opt, 0, sizeof(opt));
	opt.max_bounces = std::getenv("EAR_MAX_BOUNCES") ? std::atoi(std::getenv("EAR_MAX_BOUNCES")) : 1000;
	opt.seed = std::getenv("EAR_SEED") ? std::strtoull(std::getenv("EAR_SEED"), 0, 10) : (uint64_t)std::time(0);
	opt.first_ray = 0; opt.ray_count = -1; opt.finalise = 1;
	std::vector<std::string> errors((size_t)n_gpus);
	std::vector<std::thread> workers;
	uint64_t segments = 0; double device_ms = 0.0;
	std::vector<uint64_t> seg_per_gpu((size_t)n_gpus, 0);
	std::vector<double> ms_per_gpu((size_t)n_gpus, 0.0);
	// one BVH build: the scene is built on GPU 0 and its device image copied to the other GPUs
	std::vector<ear_b200_scene*> scenes((size_t)n_gpus, nullptr);
	if (ear_b200_scene_create(sf.vertices.data(), sf.tri_material.data(), sf.triangle_count(), table.data(),
	                          (int32_t)std::max<size_t>(sf.materials.size(), 1), 3, 0, &scenes[0]))
		throw std::runtime_error(ear_b200_last_error());
	for (int g = 0; g < n_gpus; ++g) {
		workers.emplace_back([&, g]() {
			std::vector<int> mine;
			for (int c = g; c < n_ctx; c += n_gpus) mine.push_back(c);
			std::vector<ear_b200_context> lc; std::vector<ear_b200_recorder> lr;
			for (int c : mine) { lc.push_back(cc[c]); for (int r = 0; r < n_rec; ++r) lr.push_back(rr[(size_t)c * n_rec + r]); }
			if (g > 0 && ear_b200_scene_clone(scenes[0], g, &scenes[g])) { errors[g] = ear_b200_last_error(); return; }
			ear_b200_scene* scene = scenes[g];
			ear_b200_result* res = nullptr;
			ear_b200_options o = opt;   // same seed everywhere: the contexts carry their global stream ids
			if (ear_b200_render(scene, lc.data(), (int32_t)lc.size(), lr.data(), n_rec, &o, &res)) { errors[g] = ear_b200_last_error(); return; }
			for (size_t i = 0; i < mine.size(); ++i) {
				Context& c = ctxs[mine[i]];
				for (int r = 0; r < n_rec; ++r)
					for (int k = 0; k < 2; ++k) {
						const ear_b200_track& t = res->tracks[(i * n_rec + r) * 2 + k];
						std::unique_ptr<Track> tr;
						if (t.data) { tr.reset(new Track()); tr->assign(t.data, t.length, t.first_sample, t.real_length); }
						c.tracks.push_back(std::move(tr));
					}
			}
			seg_per_gpu[g] = res->segments; ms_per_gpu[g] = res->device_ms;
			if (res->dropped_updates) errors[g] = "histogram too short: bin updates were dropped";
			ear_b200_result_free(res);
		});
	}
	for (auto& w : workers) w.join();
	for (ear_b200_scene* sc : scenes) ear_b200_scene_destroy(sc);
	for (int g = 0; g < n_gpus; ++g) {
		if (!errors[g].empty()) throw std::runtime_error(errors[g]);
		segments += seg_per_gpu[g]; device_ms = std::max(device_ms, ms_per_gpu[g]);
	}
	std::cout << "[" << std::string(49, '=') << "]" << std::endl;
	std::cout << "Traced " << segments << " ray-bounce segments on " << n_gpus << " GPU(s) in " << device_ms << " ms" << std::endl;

	// ---- post: Power, global max, truncate (src/EAR.cpp:209-244) ----
	float max = 0.0f;
	for (Context& c : ctxs)
		for (auto& t : c.tracks) if (t) { t->power(0.335f); const float m = t->maximum(); if (m > max) max = m; }
	const float threshold = max / 256.0f;
	for (Context& c : ctxs) {
		for (int r = 0; r < n_rec; ++r) {
			unsigned len = 0;
			bool has_samples = false;
			for (int k = 0; k < 2; ++k) if (c.tracks[r * 2 + k] && c.tracks[r * 2 + k]->real_length > 0) has_samples = true;
			if (has_samples)
				for (int k = 0; k < 2; ++k) if (c.tracks[r * 2 + k]) len = std::max(len, c.tracks[r * 2 + k]->length(threshold));
			for (int k = 0; k < 2; ++k) if (c.tracks[r * 2 + k]) c.tracks[r * 2 + k]->truncate(len);
			if (has_debugdir) {
				std::stringstream ss;
				ss << debugdir << "response-" << r << ".sound-" << c.sound;
				if (c.keyframe != -1) ss << ".frame-" << std::setw(2) << std::setfill('0') << c.keyframe;
				ss << ".band-" << c.band << lomihi[c.band];
				const Track& t0 = *c.tracks[r * 2];
				if (c.tracks[r * 2 + 1]) save_wav_stereo(ss.str() + ".wav", t0.data(), t0.length(), c.tracks[r * 2 + 1]->data(), c.tracks[r * 2 + 1]->length(), true);
				else save_wav_mono(ss.str() + ".wav", t0.data(), t0.length(), true, max);
				t0.write_raw(ss.str() + ".bin");
			}
		}
	}

	const bool noprocess = sf.is_set("noprocessing") && sf.get_int("noprocessing") > 0;
	if (noprocess || calc_t60) {
		std::cout << std::endl << "Not processing data" << std::endl;
		if (calc_t60) {
			*calc_t60 = ctxs[0].tracks[0]->t60();
			sabine_eyring(sf, air[1], t60_sabine, t60_eyring);
		}
		return 0;
	}

	std::cout << std::endl << "Processing data..." << std::endl;
	// ---- convolution with the dry signal (src/EAR.cpp:296-355, src/Recorder.cpp:343-363) ----
	{
		std::vector<std::thread> pool;
		const unsigned hw = (unsigned)std::max(1, 2 * n_gpus);   // two host threads per GPU keep its copy engines busy
		std::atomic<int> next(0);
		std::string conv_error;
		for (Context& c : ctxs) c.processed.resize((size_t)n_rec * 2);
		auto work = [&]() {
			for (;;) {
				const int job = next.fetch_add(1);
				if (job >= n_ctx * n_rec) break;
				const int ci = job / n_rec, r = job % n_rec;
				Context& c = ctxs[ci];
				const Source& src = sf.sources[c.sound];
				const std::vector<float>& dry = sounds[c.sound].band[c.band];
				const unsigned total = (unsigned)dry.size();
				auto section = [&](float start_s, float length_s, const float*& ptr, unsigned& n, unsigned& offset) {
					const unsigned start = (unsigned)(int)(start_s * 44100.0f);
					const unsigned want = length_s < 0 ? total - start : (unsigned)(int)(length_s * 44100.0f);
					if (start >= total) { ptr = nullptr; n = 0; offset = 0; return; }
					ptr = dry.data() + start; n = std::min(want, total - start); offset = src.offset + start;
				};
				for (int k = 0; k < 2; ++k) {
					const Track* tr = c.tracks[r * 2 + k].get();
					if (!tr) continue;
					const float* ptr; unsigned n, off;
					const Track* next = nullptr;
					if (sf.has_keys) {
						const float offset_s = sf.keys[(size_t)c.keyframe];
						const int lastkey = (int)sf.keys.size() - 1;
						if (c.keyframe == lastkey) section(offset_s, -1.0f, ptr, n, off);
						else {
							next = ctxs[(size_t)ci + 3].tracks[r * 2 + k].get();   // same sound/band, next keyframe
							section(offset_s, sf.keys[(size_t)c.keyframe + 1] - offset_s, ptr, n, off);
						}
					} else section(0.0f, -1.0f, ptr, n, off);
					// RecorderTrack::Process on the GPU (include/ear_b200.h: ear_b200_convolve)
					const unsigned len = next ? std::max(tr->real_length, next->real_length) : tr->real_length;
					std::vector<float> out((size_t)std::max<unsigned>(3 * kSampleRate, n + off + len), 0.0f);
					uint32_t of = 0, orl = 0;
					if (ear_b200_convolve(job % n_gpus, tr->data(), tr->allocated(), tr->first_sample, tr->real_length,
					                      next ? next->data() : nullptr, next ? next->allocated() : 0, next ? next->first_sample : 0,
					                      next ? next->real_length : 0, n ? ptr : nullptr, n, off, out.data(), (uint32_t)out.size(), &of, &orl)) {
						conv_error = ear_b200_last_error();
						return;
					}
					c.processed[r * 2 + k].reset(new Track());
					c.processed[r * 2 + k]->assign(out.data(), (uint32_t)out.size(), of, orl);
				}
			}
		};
		for (unsigned t = 0; t < std::min<unsigned>(hw, (unsigned)(n_ctx * n_rec)); ++t) pool.emplace_back(work);
		for (auto& t : pool) t.join();
		if (!conv_error.empty()) throw std::runtime_error(conv_error);
	}

	std::cout << "Merging result..." << std::endl;
	// ---- merge, normalise, truncate, save (src/EAR.cpp:357-386) ----
	for (int r = 0; r < n_rec; ++r) {
		const Listener& l = sf.listeners[r];
		const int n_tracks = l.stereo ? 2 : 1;
		Track total[2];
		for (Context& c : ctxs) {
			if (has_debugdir) {
				std::stringstream ss;
				ss << debugdir << "rec-" << r << ".sound-" << c.sound;
				if (c.keyframe != -1) ss << ".frame-" << std::setw(2) << std::setfill('0') << c.keyframe;
				ss << ".band-" << c.band << ".wav";
				const Track& p0 = *c.processed[r * 2];
				if (l.stereo) save_wav_stereo(ss.str(), p0.data(), p0.length(), c.processed[r * 2 + 1]->data(), c.processed[r * 2 + 1]->length(), false);
				else save_wav_mono(ss.str(), p0.data(), p0.length(), false, -1.0f);
			}
			for (int k = 0; k < n_tracks; ++k) total[k].add(*c.processed[r * 2 + k]);
		}
		float mx = -1e9f;
		for (int k = 0; k < n_tracks; ++k) mx = std::max(mx, total[k].maximum());
		for (int k = 0; k < n_tracks; ++k) total[k].normalize(0.8f, mx);
		unsigned len = 0;
		for (int k = 0; k < n_tracks; ++k) len = std::max(len, total[k].length());   // processed tracks: plain real_length
		for (int k = 0; k < n_tracks; ++k) total[k].truncate(len);
		if (l.stereo) save_wav_stereo(l.filename, total[0].data(), total[0].length(), total[1].data(), total[1].length(), false);
		else save_wav_mono(l.filename, total[0].data(), total[0].length(), false, -1.0f);
		std::cout << "Saved " << l.filename << " (" << len << " samples)" << std::endl;
	}
	return 0;
}

}  // namespace

int main(int argc, char** argv) {
	std::cout << kBanner << std::endl << std::endl << std::endl;
	std::cout << std::setprecision(3) << std::fixed;
	for (int i = 1; i < argc; ++i) {
		const std::string cmd(argv[i]);
		const std::string arg1 = (i + 1 < argc) ? argv[i + 1] : "";
		const std::string arg2 = (i + 2 < argc) ? argv[i + 2] : "";
		if (cmd == "render" && !arg1.empty()) {
			int ret = 1;
			try { ret = run(arg1, nullptr, nullptr, nullptr); }
			catch (std::exception& e) { std::cout << std::endl << "Error: " << e.what() << std::endl << std::endl; }
			if (std::getenv("EAR_WAIT_KEY")) { std::cout << "Press a key to exit..." << std::endl; std::cin.get(); }
			return ret;
		} else if (cmd == "calc" && arg1 == "T60" && !arg2.empty()) {
			float t60 = 0, sabine = 0, eyring = 0;
			int ret = 1;
			try {
				ret = run(arg2, &t60, &sabine, &eyring);
				std::cout << "T60_ear   : " << std::setprecision(9) << std::fixed << t60 << "s" << std::endl;
				std::cout << "T60_sabine: " << std::setprecision(9) << std::fixed << sabine << "s" << std::endl;
				std::cout << "T60_eyring: " << std::setprecision(9) << std::fixed << eyring << "s" << std::endl;
			} catch (std::exception& e) { std::cout << std::endl << "Error: " << e.what() << std::endl << std::endl; }
			return ret;
		} else if (cmd == "test") return 0;
	}
	std::cout << "Usage:" << std::endl << " EAR render <filename>" << std::endl << " EAR calc T60 <filename>" << std::endl;
	return 0;
}
