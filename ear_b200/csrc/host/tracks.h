// Recorder tracks on the host: the FloatBuffer / RecorderTrack / Recorder semantics the post chain
// relies on (src/Recorder.h:55-221, src/Recorder.cpp:33-118, 247-340, 365-460).  The GPU returns
// raw tracks (data, first_sample, real_length); everything after Scene::Render -- Power, Truncate,
// T60, convolution with the dry signal, merge, normalise, save -- happens here, bit-compatible with
// the reference's host code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace earhost {

constexpr unsigned kSampleRate = 44100;

// Auto-growing float array that remembers the lowest / highest index written.
class Track {
public:
	Track();
	// adopt a rendered track: `length` allocated bins, of which [0, real_length] may be non-zero
	void assign(const float* data, uint32_t length, uint32_t first_sample, uint32_t real_length);
	float& at(unsigned i);                                   // non-const operator[]: grows, tracks range
	float get(unsigned i) const { return i < data_.size() ? data_[i] : 0.0f; }
	const float* data() const { return data_.data(); }
	unsigned allocated() const { return (unsigned)data_.size(); }
	unsigned first_sample, real_length;

	float maximum() const;                                   // max |x| over [first_sample, real_length)
	float root_mean_square() const;
	void multiply(float f);                                  // over [first_sample, real_length)
	void normalize(float m, float max);
	void truncate(unsigned l);
	void power(float a);                                     // sign(x) |x|^a
	unsigned length(float threshold = -1.0f) const;          // getLength
	float t60() const;
	void add(const Track& other);
	// convolution with the dry signal (RecorderTrack::Process) runs on the GPU: ear_b200_convolve (include/ear_b200.h)
	void write_raw(const std::string& path) const;
private:
	std::vector<float> data_;
};

}  // namespace earhost
