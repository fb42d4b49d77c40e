#include "tracks.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>

namespace earhost {

Track::Track() : first_sample(3 * kSampleRate - 1), real_length(0), data_(3 * kSampleRate, 0.0f) {}

void Track::assign(const float* data, uint32_t length, uint32_t first, uint32_t real) {
	data_.assign(std::max<uint32_t>(length, 3 * kSampleRate), 0.0f);
	const uint32_t live = std::min<uint32_t>(real + 1, length);
	if (data && live) std::memcpy(data_.data(), data, (size_t)live * sizeof(float));
	first_sample = first;
	real_length = real;
}

float& Track::at(unsigned i) {
	if (i >= data_.size()) data_.resize((size_t)i + kSampleRate, 0.0f);
	if (i > real_length) real_length = i;
	if (i < first_sample) first_sample = i;
	return data_[i];
}

float Track::maximum() const {
	float x = 0.0f;
	for (unsigned i = first_sample; i < real_length; ++i) { const float a = std::fabs(data_[i]); if (a > x) x = a; }
	return x;
}

float Track::root_mean_square() const {
	if (!real_length) return 0.0f;
	float x = 0.0f;
	for (unsigned i = first_sample; i < real_length; ++i) x += data_[i] * data_[i];
	return std::sqrt(x / (float)real_length);
}

void Track::multiply(float f) { for (unsigned i = first_sample; i < real_length; ++i) data_[i] *= f; }

void Track::normalize(float m, float max) { multiply(m / (max < 0 ? maximum() : max)); }

void Track::truncate(unsigned l) {
	if (l == 0) l = 1;
	if (l >= data_.size()) data_.resize(l, 0.0f);
	real_length = l;
}

void Track::power(float a) {
	for (unsigned i = first_sample; i < real_length; ++i) {
		const float f = powf(std::fabs(data_[i]), a);
		data_[i] = data_[i] < 0 ? (f * -1.0f) : f;
	}
}

unsigned Track::length(float threshold) const {
	if (threshold < 0.0f) return real_length;
	unsigned mx = 0;
	for (unsigned i = first_sample; i < data_.size(); ++i) if (std::fabs(data_[i]) >= threshold) mx = i;
	return mx + 1;
}

// T60 from the (Power-compressed, truncated) response: the first sample that drops below its
// predecessor ends the direct lobe; the tail ends at the last sample above direct/1000.
float Track::t60() const {
	const float attenuation_gain = powf(10.0f, 60.0f / 20.0f);
	float min_gain = 0.0f, previous = -1.0f;
	int last_significant = 0, direct_offset = 0;
	bool in_tail = false;
	for (unsigned j = first_sample; j < real_length; ++j) {
		const float s = get(j);
		if (in_tail) { if (s > min_gain) last_significant = (int)j; }
		else if (s < previous) { in_tail = true; min_gain = previous / attenuation_gain; direct_offset = (int)j; }
		previous = s;
	}
	return (float)(last_significant - direct_offset) / 44100.0f;
}

void Track::add(const Track& other) {
	const unsigned len = other.length(0.0f);
	for (unsigned i = 0; i < len; ++i) at(i) += other.get(i);
}

void Track::write_raw(const std::string& path) const {
	std::ofstream f(path.c_str(), std::ios::binary);
	f.write((const char*)data_.data(), sizeof(float) * ((size_t)real_length + 1));
}

}  // namespace earhost
