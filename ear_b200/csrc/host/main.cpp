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
#include <mutex>
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

	// equalizer output saved for debugging (src/EAR.cpp:123-148); the reference splits every source here, also for calc T60
	if (has_debugdir) {
		for (size_t sid = 0; sid < sf.sources.size(); ++sid) {
			if (calc_t60 && sounds[sid].band[1].empty() && sf.sources[sid].wavs.size() == 1) {
				std::vector<float> dry = load_wav_mono(sf.sources[sid].wavs[0]);
				split_bands(dry, sf.freq[0] * 1000.0f, sf.freq[1] * 1000.0f, sf.freq[2] * 1000.0f, sounds[sid].band[0], sounds[sid].band[1], sounds[sid].band[2]);
			}
			for (int band = 0; band < 3; ++band) {
				if (calc_t60 && band != 1) continue;
				std::stringstream ss;
				ss << debugdir << "sound-" << sid << ".band-" << band << lomihi[band] << ".wav";
				save_wav_mono(ss.str(), sounds[sid].band[band].data(), sounds[sid].band[band].size(), false, -1.0f);
			}
			if (calc_t60) break;
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
		if (src.is_mesh) {   // AbstractSoundFile::mesh: rays start on its triangles (src/SoundFile.cpp:216-221)
			cc[c].source_kind = EAR_B200_MESH_SOURCE;
			cc[c].emitter_first = src.emitter_first;
			cc[c].emitter_count = src.emitter_count;
		} else {
			const std::array<float, 3>& sp = src.location.at(ctxs[c].keyframe);
			for (int k = 0; k < 3; ++k) cc[c].source_position[k] = sp[k];
		}
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

	// One library call for the whole fan-out (src/EAR.cpp:196-207): the scene is built on GPU 0 and peer-copied to the
	// other GPUs of the box; every GPU traces a contiguous share of the ray ids of EVERY context, the partial histograms
	// meet on GPU 0 over NVLink (ear_b200_group_render).  `calc T60` (one context) uses all GPUs this way too.
	int n_gpus = std::max(1, ear_b200_device_count());
	if (const char* e = std::getenv("EAR_GPUS")) n_gpus = std::max(1, std::min(n_gpus, std::atoi(e)));
	n_gpus = std::min(n_gpus, 16);
	ear_b200_options opt;
	std::memset(&opt, 0, sizeof(opt));
	opt.max_bounces = std::getenv("EAR_MAX_BOUNCES") ? std::atoi(std::getenv("EAR_MAX_BOUNCES")) : 1000;
	opt.seed = std::getenv("EAR_SEED") ? std::strtoull(std::getenv("EAR_SEED"), 0, 10) : (uint64_t)std::time(0);
	opt.first_ray = 0; opt.ray_count = -1; opt.finalise = 1;
	uint64_t segments = 0; double device_ms = 0.0;
	{
		struct Handles {   // released on every path, exceptions included
			ear_b200_scene* scene = nullptr; ear_b200_group* group = nullptr; ear_b200_result* res = nullptr;
			~Handles() { ear_b200_result_free(res); ear_b200_group_destroy(group); ear_b200_scene_destroy(scene); }
		} h;
		if (ear_b200_scene_create(sf.vertices.data(), sf.tri_material.data(), sf.triangle_count(), table.data(),
		                          (int32_t)std::max<size_t>(sf.materials.size(), 1), 3, 0, &h.scene))
			throw std::runtime_error(ear_b200_last_error());
		if (!sf.emitter_vertices.empty() &&
		    ear_b200_scene_set_emitters(h.scene, sf.emitter_vertices.data(), (int32_t)(sf.emitter_vertices.size() / 9)))
			throw std::runtime_error(ear_b200_last_error());
		std::vector<int32_t> devices((size_t)n_gpus);
		for (int g = 0; g < n_gpus; ++g) devices[(size_t)g] = g;
		if (ear_b200_group_create(h.scene, devices.data(), n_gpus, &h.group)) throw std::runtime_error(ear_b200_last_error());
		if (ear_b200_group_render(h.group, cc.data(), n_ctx, rr.data(), n_rec, &opt, &h.res)) throw std::runtime_error(ear_b200_last_error());
		if (h.res->dropped_updates) throw std::runtime_error("histogram too short: bin updates were dropped");
		for (int c = 0; c < n_ctx; ++c)
			for (int r = 0; r < n_rec; ++r)
				for (int k = 0; k < 2; ++k) {
					const ear_b200_track& t = h.res->tracks[((size_t)c * n_rec + r) * 2 + k];
					std::unique_ptr<Track> tr;
					if (t.data) { tr.reset(new Track()); tr->assign(t.data, t.length, t.first_sample, t.real_length); }
					ctxs[(size_t)c].tracks.push_back(std::move(tr));
				}
		segments = h.res->segments; device_ms = h.res->device_ms;
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
		const char* conv_mode = std::getenv("EAR_CONVOLUTION");
		const bool use_fft = conv_mode && std::string(conv_mode) == "fft";
		std::string conv_error;
		std::mutex conv_error_lock;
		std::atomic<bool> conv_failed(false);
		for (Context& c : ctxs) c.processed.resize((size_t)n_rec * 2);
		auto work = [&]() {
			for (;;) {
				const int job = next.fetch_add(1);
				if (job >= n_ctx * n_rec || conv_failed.load()) break;
				const int ci = job / n_rec, r = job % n_rec;
				Context& c = ctxs[ci];
				const Source& src = sf.sources[c.sound];
				const std::vector<float>& dry = sounds[c.sound].band[c.band];
				const unsigned total = (unsigned)dry.size();
				auto section = [&](float start_s, float length_s, const float*& ptr, unsigned& n, unsigned& offset) {
					const unsigned start = (unsigned)(int)(start_s * 44100.0f);
					const unsigned want = length_s < 0 ? total - start : (unsigned)(int)(length_s * 44100.0f);
					if (start >= total) { ptr = nullptr; n = 0; offset = 0; return; }
					ptr = dry.data() + start; n = std::min(want, total - start);
					// The per-band SoundFiles that Band() hands to the convolution are constructed with offset 0
					// (src/SoundFile.cpp:83,150-152): the source's own `offset` setting never reaches Section() -- kept.
					offset = start;
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
					const unsigned first = next ? std::min(tr->first_sample, next->first_sample) : tr->first_sample;
					// The reference's result track is a FloatBuffer written through operator[] in increasing index order
					// (src/Recorder.cpp:247-292): it starts at 3 s and grows to "index + 1 s" whenever an index falls outside
					// (src/Recorder.cpp:52-59).  Its allocated length matters later: RecorderTrack::Add walks the WHOLE
					// allocation (getLength(0.0f), src/Recorder.cpp:293-300), which sets the length of the saved file.
					size_t alloc = 3 * kSampleRate;
					if (n > 0 && len > first) {
						const size_t i0 = (size_t)off + first, last = (size_t)(n - 1) + off + (len - 1);
						if (i0 >= alloc) alloc = i0 + kSampleRate;
						while (last >= alloc) alloc += kSampleRate;
					}
					std::vector<float> out(alloc, 0.0f);
					uint32_t of = 0, orl = 0;
					// EAR_CONVOLUTION=fft: the frequency-domain form (the reference's USE_FFTW build); default: the direct
					// form, bit-identical with the reference's default build
					if ((use_fft ? ear_b200_convolve_fft : ear_b200_convolve)(job % n_gpus, tr->data(), tr->allocated(), tr->first_sample, tr->real_length,
					                      next ? next->data() : nullptr, next ? next->allocated() : 0, next ? next->first_sample : 0,
					                      next ? next->real_length : 0, n ? ptr : nullptr, n, off, out.data(), (uint32_t)out.size(), &of, &orl)) {
						std::lock_guard<std::mutex> lock(conv_error_lock);
						if (conv_error.empty()) conv_error = ear_b200_last_error();
						conv_failed.store(true);
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
				// other->Save(fn) through a Recorder*: the defaults of the BASE declaration apply, norm = true, norm_max = -1
				// (src/Recorder.h:181) -- the peak of the track lands at 0.8 (lib/wave/WaveFile.cpp:194-201)
				if (l.stereo) save_wav_stereo(ss.str(), p0.data(), p0.length(), c.processed[r * 2 + 1]->data(), c.processed[r * 2 + 1]->length(), true);
				else save_wav_mono(ss.str(), p0.data(), p0.length(), true, -1.0f);
			}
			for (int k = 0; k < n_tracks; ++k) total[k].add(*c.processed[r * 2 + k]);
		}
		float mx = -1e9f;
		for (int k = 0; k < n_tracks; ++k) mx = std::max(mx, total[k].maximum());
		for (int k = 0; k < n_tracks; ++k) total[k].normalize(0.8f, mx);
		// total->Truncate(total->getLength(1e-6f)) (src/EAR.cpp:382): getLength of a processed recorder is the longest
		// real_length (src/Recorder.cpp:412-418) and Truncate acts on the recorder's RESPONSE tracks (:399-404), which
		// are blank here -- the summed tracks are saved with their own real_length
		unsigned len = 0;
		for (int k = 0; k < n_tracks; ++k) len = std::max(len, total[k].length());
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
