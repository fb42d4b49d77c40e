// `.ear` scene files -> flat scene description (host side of the drop-in).
//
// Grammar and quirks follow the reference reader (src/Datatype.cpp:25-163, src/EAR.cpp:60-119,
// src/Material.cpp:26-72, src/Mesh.cpp:77-93, src/SoundFile.cpp:33-117, src/MonoRecorder.cpp:38-59,
// src/StereoRecorder.cpp:33-67, src/Animated.h:35-108, src/Settings.cpp:23-103) and the exporter
// that defines the encoding (blender/render_EAR/__init__.py:154-196).  The reader is a bounds-checked
// cursor over the file image instead of the reference's static global cursor.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace earhost {

struct FormatError : public std::runtime_error {
	explicit FormatError(const std::string& m) : std::runtime_error(m) {}
};

struct Material {
	std::string name;
	float refl[3] = {0, 0, 0};
	float refr[3] = {0, 0, 0};
	float kept[3] = {0, 0, 0};   // absorption_coefficient: surviving fraction per bounce
	float spec[3] = {0, 0, 0};
};

struct MeshBlock {
	int material = 0;
	int first_triangle = 0, triangle_count = 0;
};

// A position (or direction) that is either constant or one value per keyframe (no interpolation,
// src/Animated.h:89-91).
struct Placement {
	bool animated = false;
	std::array<float, 3> value = {0, 0, 0};
	std::vector<std::array<float, 3>> frames;
	const std::array<float, 3>& at(int keyframe) const {
		return (keyframe >= 0 && animated) ? frames[(size_t)keyframe] : value;
	}
};

struct Source {
	std::vector<std::string> wavs;   // 1 = SSRC (band-split by the equalizer), 3 = 3SRC
	Placement location;
	float gain = 1.0f;
	unsigned offset = 0;             // start offset in samples
	// mesh source (`mesh` sub-block instead of a location, src/SoundFile.cpp:50-53): its triangles are
	// [emitter_first, emitter_first + emitter_count) of SceneFile::emitter_vertices
	bool is_mesh = false;
	int emitter_first = 0, emitter_count = 0;
};

struct Listener {
	std::string filename;
	bool stereo = false;
	Placement location;
	Placement right_ear;
	float head_size = 0.0f;
	float head_absorption[3] = {0, 0, 0};   // already max(0, (1-a)^4)
};

struct Setting {
	enum Kind { INT, FLOAT, VEC, STRING } kind = INT;
	int i = 0;
	float f = 0.0f;
	std::array<float, 3> v = {0, 0, 0};
	std::string s;
};

struct SceneFile {
	std::map<std::string, Setting> settings;
	std::vector<Material> materials;
	std::vector<MeshBlock> meshes;
	std::vector<float> vertices;         // [T][3][3] in file order == triangle index
	std::vector<int32_t> tri_material;   // [T]
	std::vector<float> emitter_vertices; // [E][3][3]: the triangles of all mesh sources, file order
	std::vector<Source> sources;
	std::vector<Listener> listeners;
	std::vector<float> keys;
	bool has_keys = false;
	float freq[3] = {0.2f, 1.0f, 2.0f};  // SoundFile::f1..f3 defaults (src/SoundFile.cpp:240-242)

	bool is_set(const std::string& k) const { return settings.count(k) != 0; }
	int get_int(const std::string& k) const;
	float get_float(const std::string& k) const;
	std::array<float, 3> get_vec(const std::string& k) const;
	std::string get_string(const std::string& k) const;
	int triangle_count() const { return (int)tri_material.size(); }
};

// Reads and echoes the scene like the reference does while parsing (to `log`, may be null).
SceneFile load_scene_file(const std::string& path, std::ostream* log);

}  // namespace earhost
