#include "ear_format.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace earhost {
namespace {

class Cursor {
public:
	Cursor(const char* b, size_t n) : p_(b), end_(b + n) {}
	bool more() const { return p_ < end_; }
	size_t left() const { return (size_t)(end_ - p_); }
	std::string peek() const { return left() >= 4 ? std::string(p_, 4) : std::string(); }
	bool next_is(const char* tag) const { return left() >= 4 && std::memcmp(p_, tag, 4) == 0; }
	void expect(const char* tag) {
		if (!next_is(tag)) throw FormatError("Found '" + peek() + "' while expecting '" + tag + "'");
		p_ += 4;
	}
	int32_t raw_i32() { need(4); int32_t v; std::memcpy(&v, p_, 4); p_ += 4; return v; }
	float raw_f32() { need(4); float v; std::memcpy(&v, p_, 4); p_ += 4; return v; }
	int32_t read_int() { expect("int4"); return raw_i32(); }
	float read_float() { expect("flt4"); return raw_f32(); }
	std::array<float, 3> read_vec() { expect("vec3"); std::array<float, 3> v; for (int i = 0; i < 3; ++i) v[i] = read_float(); return v; }
	std::string read_string() {
		expect("str ");
		size_t n = 0;
		while (p_ + n < end_ && p_[n]) ++n;
		std::string s(p_, n);
		const size_t skip = n + (4 - n % 4);   // padded to a multiple of 4, always >= 1 NUL
		need(std::min(skip, left()));
		p_ += std::min(skip, left());
		return s;
	}
	// container: id, payload length, payload -> sub-cursor; this cursor moves past it
	Cursor open(const char* tag) {
		expect(tag);
		const int32_t n = raw_i32();
		if (n < 0 || (size_t)n > left()) throw FormatError(std::string("Truncated '") + tag + "' block");
		Cursor sub(p_, (size_t)n);
		p_ += n;
		return sub;
	}
	void skip_block() {
		need(8);
		p_ += 4;
		const int32_t n = raw_i32();
		if (n < 0 || (size_t)n > left()) throw FormatError("Truncated block");
		p_ += n;
	}
	const char* pos() const { return p_; }
private:
	void need(size_t n) const { if (left() < n) throw FormatError("Unexpected end of file"); }
	const char* p_;
	const char* end_;
};

std::string fmt_vec(const std::array<float, 3>& v) {
	std::ostringstream ss;
	ss << std::fixed << std::setprecision(3) << "(" << v[0] << ", " << v[1] << ", " << v[2] << ")";
	return ss.str();
}
std::string describe(const Placement& p) {
	if (!p.animated) return fmt_vec(p.value);
	return "< Animated " + fmt_vec(p.frames.front()) + " -> " + fmt_vec(p.frames.back()) + " >";
}

Placement read_placement(Cursor& c, const SceneFile& sf) {
	Placement p;
	if (c.next_is("anim")) {
		if (!sf.has_keys || sf.keys.empty()) throw FormatError("Keyframe data not read");
		Cursor a = c.open("anim");
		while (a.more()) p.frames.push_back(a.read_vec());
		if (p.frames.size() != sf.keys.size()) throw FormatError("Keyframe count does not match");
		p.animated = true;
	} else p.value = c.read_vec();
	return p;
}

void read_settings(Cursor c, SceneFile& sf, std::ostream* log) {
	if (log) *log << "Settings" << std::endl;
	while (c.next_is("str ")) {
		const std::string key = c.read_string();
		Setting s;
		if (c.next_is("int4")) { s.kind = Setting::INT; s.i = c.read_int(); }
		else if (c.next_is("flt4")) { s.kind = Setting::FLOAT; s.f = c.read_float(); }
		else if (c.next_is("vec3")) { s.kind = Setting::VEC; s.v = c.read_vec(); }
		else if (c.next_is("str ")) { s.kind = Setting::STRING; s.s = c.read_string(); }
		else throw FormatError("Setting '" + key + "' has an unknown value type");
		if (log) {
			*log << " +- " << key << ": ";
			if (s.kind == Setting::INT) *log << s.i;
			else if (s.kind == Setting::FLOAT) *log << s.f;
			else if (s.kind == Setting::VEC) *log << "[" << s.v[0] << ", " << s.v[1] << ", " << s.v[2] << "]";
			else *log << s.s;
			*log << std::endl;
		}
		sf.settings[key] = s;
	}
}

void read_material(Cursor c, SceneFile& sf, std::ostream* log) {
	Material m;
	m.name = c.read_string();
	float a[3] = {1.0f, 1.0f, 1.0f};
	for (int i = 0; i < 3; ++i) { m.refl[i] = c.read_float(); a[i] -= m.refl[i] - 1e-9f; }
	if (c.next_is("flt4")) for (int i = 0; i < 3; ++i) { m.refr[i] = c.read_float(); a[i] -= m.refr[i] - 1e-9f; }
	for (int i = 0; i < 3; ++i) {
		if (a[i] < 0.0f) throw FormatError("Invalid material settings");
		m.kept[i] = 1.0f - a[i];
	}
	if (c.next_is("flt4")) for (int i = 0; i < 3; ++i) m.spec[i] = c.read_float();
	if (log) {
		*log << "Material '" << m.name << "'" << std::endl;
		*log << " +- refl:   [" << m.refl[0] << ", " << m.refl[1] << ", " << m.refl[2] << "]" << std::endl;
		*log << " +- trans:  [" << m.refr[0] << ", " << m.refr[1] << ", " << m.refr[2] << "]" << std::endl;
		*log << " +- absorp: [" << a[0] << ", " << a[1] << ", " << a[2] << "]" << std::endl;
		*log << " +- spec:   [" << m.spec[0] << ", " << m.spec[1] << ", " << m.spec[2] << "]" << std::endl;
	}
	for (size_t i = 0; i < sf.materials.size(); ++i)
		if (sf.materials[i].name == m.name) { sf.materials[i] = m; return; }   // std::map semantics: last wins
	sf.materials.push_back(m);
}

void read_mesh(Cursor c, SceneFile& sf, std::ostream* log) {
	const std::string name = c.read_string();
	MeshBlock mb;
	mb.material = -1;
	for (size_t i = 0; i < sf.materials.size(); ++i) if (sf.materials[i].name == name) mb.material = (int)i;
	if (mb.material < 0) throw FormatError("Mesh refers to undefined material '" + name + "'");
	mb.first_triangle = sf.triangle_count();
	float lo[3] = {1e9f, 1e9f, 1e9f}, hi[3] = {-1e9f, -1e9f, -1e9f};
	while (c.next_is("tri ")) {
		c.expect("tri ");
		for (int v = 0; v < 3; ++v) {
			const std::array<float, 3> p = c.read_vec();
			for (int k = 0; k < 3; ++k) { sf.vertices.push_back(p[k]); lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
		}
		sf.tri_material.push_back(mb.material);
		++mb.triangle_count;
	}
	sf.meshes.push_back(mb);
	if (log) {
		*log << "Mesh" << std::endl << " +- faces: " << mb.triangle_count << std::endl << " +- material: '" << name << "'" << std::endl;
		*log << " +- bounds: (" << lo[0] << ", " << lo[1] << ", " << lo[2] << ") - (" << hi[0] << ", " << hi[1] << ", " << hi[2] << ")" << std::endl;
	}
}

void read_source(Cursor c, SceneFile& sf, int n_files, std::ostream* log) {
	Source s;
	for (int i = 0; i < n_files; ++i) s.wavs.push_back(c.read_string());
	if (c.next_is("mesh")) {
		// Mesh's constructor inside a source (src/SoundFile.cpp:50-53, src/Mesh.cpp:76-91): [str material][tri ...].  The
		// material must exist (the reference dereferences it); the triangles emit, they are not part of the geometry.
		Cursor m = c.open("mesh");
		const std::string name = m.read_string();
		bool known = false;
		for (size_t i = 0; i < sf.materials.size(); ++i) if (sf.materials[i].name == name) known = true;
		if (!known) throw FormatError("Mesh refers to undefined material '" + name + "'");
		s.is_mesh = true;
		s.emitter_first = (int)(sf.emitter_vertices.size() / 9);
		while (m.next_is("tri ")) {
			m.expect("tri ");
			for (int v = 0; v < 3; ++v) {
				const std::array<float, 3> p = m.read_vec();
				for (int k = 0; k < 3; ++k) sf.emitter_vertices.push_back(p[k]);
			}
			++s.emitter_count;
		}
	} else s.location = read_placement(c, sf);
	if (c.more() && c.next_is("flt4")) s.gain = c.read_float();
	if (c.more() && c.next_is("flt4")) s.offset = (unsigned int)(c.read_float() * 44100.0f);
	if (log) {
		*log << "Sound source" << std::endl << " +- location: " << describe(s.location) << std::endl;
		for (size_t i = 0; i < s.wavs.size(); ++i) *log << " +- data" << (s.wavs.size() > 1 ? std::to_string(i + 1) : "") << ": " << s.wavs[i] << std::endl;
		*log << " +- offset: " << s.offset << std::endl;
	}
	sf.sources.push_back(s);
}

void read_listener(Cursor c, SceneFile& sf, bool stereo, std::ostream* log) {
	Listener l;
	l.stereo = stereo;
	l.filename = c.read_string();
	(void)c.read_float();   // exporter writes 35.0; read and ignored (src/MonoRecorder.cpp:45)
	l.location = read_placement(c, sf);
	if (stereo) {
		l.right_ear = read_placement(c, sf);
		l.head_size = c.read_float();
		const std::array<float, 3> ab = c.read_vec();
		for (int i = 0; i < 3; ++i) l.head_absorption[i] = std::max(0.0f, powf(1.0f - ab[i], 4));
	}
	if (log) {
		*log << "Recorder" << std::endl << " +- " << (stereo ? "stereo" : "mono") << std::endl << " +- location: " << describe(l.location) << std::endl;
		if (stereo)
			*log << " +- right: " << describe(l.right_ear) << std::endl << " +- head size: " << l.head_size << std::endl
			     << " +- head absorption: (" << l.head_absorption[0] << ", " << l.head_absorption[1] << ", " << l.head_absorption[2] << ")" << std::endl;
	}
	sf.listeners.push_back(l);
}

}  // namespace

int SceneFile::get_int(const std::string& k) const {
	auto it = settings.find(k);
	if (it == settings.end()) throw FormatError("Setting '" + k + "' not found");
	return it->second.i;
}
float SceneFile::get_float(const std::string& k) const {
	auto it = settings.find(k);
	if (it == settings.end()) throw FormatError("Setting '" + k + "' not found");
	return it->second.f;
}
std::array<float, 3> SceneFile::get_vec(const std::string& k) const {
	auto it = settings.find(k);
	if (it == settings.end()) throw FormatError("Setting '" + k + "' not found");
	return it->second.v;
}
std::string SceneFile::get_string(const std::string& k) const {
	auto it = settings.find(k);
	if (it == settings.end()) throw FormatError("Setting '" + k + "' not found");
	return it->second.s;
}

SceneFile load_scene_file(const std::string& path, std::ostream* log) {
	std::ifstream f(path.c_str(), std::ios::binary);
	if (!f.good()) throw FormatError("Failed to read file");
	std::string image((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	if (image.size() < 4 || image.compare(0, 4, ".EAR") != 0) throw FormatError("Failed to read file");
	SceneFile sf;
	// the settings block is located first, wherever it sits (Datatype::Scan, src/EAR.cpp:67)
	{
		Cursor scan(image.data() + 4, image.size() - 4);
		bool found = false;
		while (scan.more()) {
			if (scan.next_is("SET ")) { read_settings(scan.open("SET "), sf, log); found = true; break; }
			scan.skip_block();
		}
		if (!found) throw FormatError("No settings block found in file");
	}
	Cursor c(image.data() + 4, image.size() - 4);
	while (c.more()) {
		if (c.next_is("OUT1")) read_listener(c.open("OUT1"), sf, false, log);
		else if (c.next_is("OUT2")) read_listener(c.open("OUT2"), sf, true, log);
		else if (c.next_is("SSRC")) read_source(c.open("SSRC"), sf, 1, log);
		else if (c.next_is("3SRC")) read_source(c.open("3SRC"), sf, 3, log);
		else if (c.next_is("MESH")) read_mesh(c.open("MESH"), sf, log);
		else if (c.next_is("MAT ")) read_material(c.open("MAT "), sf, log);
		else if (c.next_is("SET ") || c.next_is("VRSN")) c.skip_block();
		else if (c.next_is("KEYS")) {
			Cursor k = c.open("KEYS");
			while (k.more()) sf.keys.push_back(k.read_float());
			sf.has_keys = true;
		} else if (c.next_is("FREQ")) {
			Cursor q = c.open("FREQ");
			for (int i = 0; i < 3; ++i) sf.freq[i] = q.read_float();
		} else {
			if (log) *log << "Unknown block '" << c.peek() << "'" << std::endl;
			c.skip_block();
		}
	}
	return sf;
}

}  // namespace earhost
