#include "audio.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace earhost {
namespace {

#pragma pack(push, 1)
struct RiffHeader { char riff[4]; uint32_t size; char wave[4]; };
struct FmtChunk { char id[4]; uint32_t size; int16_t format, channels; uint32_t rate, byte_rate; int16_t align, bits; };
#pragma pack(pop)

bool write_wav(const std::string& path, const std::vector<int16_t>& pcm, int channels) {
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f) return false;
	const uint32_t bytes = (uint32_t)(pcm.size() * 2);
	RiffHeader h; std::memcpy(h.riff, "RIFF", 4); std::memcpy(h.wave, "WAVE", 4); h.size = bytes + 36;
	FmtChunk c; std::memcpy(c.id, "fmt ", 4); c.size = 16; c.format = 1; c.channels = (int16_t)channels; c.rate = 44100;
	c.byte_rate = 88200u * channels; c.align = (int16_t)(2 * channels); c.bits = 16;
	std::fwrite(&h, sizeof(h), 1, f); std::fwrite(&c, sizeof(c), 1, f);
	std::fwrite("data", 1, 4, f); std::fwrite(&bytes, 4, 1, f);
	if (bytes) std::fwrite(pcm.data(), 1, bytes, f);
	std::fclose(f);
	return true;
}

}  // namespace

std::vector<float> load_wav_mono(const std::string& path) {
	std::vector<float> out;
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) return out;
	RiffHeader h; FmtChunk c;
	std::vector<unsigned char> data;
	bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::strncmp(h.wave, "WAVE", 4) == 0;
	ok = ok && std::fread(&c, sizeof(c), 1, f) == 1 && std::strncmp(c.id, "fmt", 3) == 0 && c.format == 1;
	if (ok) {
		// the reference reads the fixed 16-byte fmt body and then walks chunks until `riff.size`
		char id[4]; uint32_t n = 0;
		while (std::fread(id, 1, 4, f) == 4 && std::fread(&n, 4, 1, f) == 1) {
			if (std::strncmp(id, "data", 4) == 0) {
				const size_t at = data.size();
				data.resize(at + n);
				const size_t got = std::fread(data.data() + at, 1, n, f);
				data.resize(at + got);
				if (got != n) break;
			} else if (std::fseek(f, (long)n, SEEK_CUR) != 0) break;
			if ((uint32_t)std::ftell(f) >= h.size) break;
		}
	}
	std::fclose(f);
	if (!ok || data.empty() || c.channels <= 0) return out;
	if (c.bits != 8 && c.bits != 16 && c.bits != 24) return out;
	const int bps = c.bits >> 3;
	const size_t frames = data.size() / (size_t)bps / (size_t)c.channels;
	const float max_sample = (float)(2 << (c.bits - 2));
	out.resize(frames);
	const unsigned char* p = data.data();
	for (size_t i = 0; i < frames; ++i) {
		float acc = 0.0f;
		for (int ch = 0; ch < c.channels; ++ch) {
			float s = 0.0f;
			if (bps == 1) s = (float)p[0] - 128.0f;
			else if (bps == 2) { int16_t v; std::memcpy(&v, p, 2); s = (float)v; }
			else { const int32_t v = (int32_t)(p[0] | (p[1] << 8) | (p[2] << 16) | ((p[2] & 0x80) ? (0xffu << 24) : 0u)); s = (float)v; }
			p += bps;
			acc += s / max_sample;
		}
		out[i] = acc / (float)c.channels;
	}
	return out;
}

bool save_wav_mono(const std::string& path, const float* data, size_t n, bool norm, float norm_max) {
	float mx = 1.0f;
	if (norm) {
		if (norm_max < 0) { mx = -1e9f; for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(data[i])); mx /= 0.8f; }
		else mx = norm_max / 0.95f;
	}
	std::vector<int16_t> pcm(n);
	for (size_t i = 0; i < n; ++i) pcm[i] = (int16_t)(data[i] / mx * 32768.0f);
	return write_wav(path, pcm, 1);
}

bool save_wav_stereo(const std::string& path, const float* left, size_t n_left, const float* right, size_t n_right, bool norm) {
	const size_t n = std::max(n_left, n_right);
	float mx = 1.0f;
	if (norm) {
		mx = -1e9f;   // signed maximum, as in WaveFile::FromFloat(left, right, ...)
		for (size_t i = 0; i < n_left; ++i) mx = std::max(mx, left[i]);
		for (size_t i = 0; i < n_right; ++i) mx = std::max(mx, right[i]);
		mx /= 0.8f;
	}
	std::vector<int16_t> pcm(2 * n);
	for (size_t i = 0; i < n; ++i) {
		pcm[2 * i] = (int16_t)(i < n_left ? (left[i] / mx * 32768.0f) : 0);
		pcm[2 * i + 1] = (int16_t)(i < n_right ? (right[i] / mx * 32768.0f) : 0);
	}
	return write_wav(path, pcm, 2);
}

namespace {
// One 4th-order Linkwitz-Riley section (direct form I).  Coefficient formulas and the constants
// (pi = 22/7, wc built from the SAMPLE RATE, not the cutoff) are the reference's, bugs included,
// because the band signals feed the convolution output.
struct Lr4 {
	float a[5], b[5], x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
	Lr4(float fc, bool highpass) {
		const float srate = 44100.0f, pi = 3.14285714285714f;
		const float wc = 2.0f * pi * srate, wc2 = wc * wc, wc3 = wc2 * wc, wc4 = wc2 * wc2;
		const float k = wc / std::tan(pi * fc / srate) /* the float overload, as the reference gets through <math.h> */, k2 = k * k, k3 = k2 * k, k4 = k2 * k2;
		const float sqrt2 = sqrtf(2.0f), t1 = sqrt2 * wc3 * k, t2 = sqrt2 * wc * k3;
		const float at = 4.0f * wc2 * k2 + 2.0f * t1 + k4 + 2.0f * t2 + wc4;
		b[0] = 0.0f;
		b[1] = (4.0f * (wc4 + t1 - k4 - t2)) / at;
		b[2] = (6.0f * wc4 - 8.0f * wc2 * k2 + 6.0f * k4) / at;
		b[3] = (4.0f * (wc4 - t1 + t2 - k4)) / at;
		b[4] = (k4 - 2.0f * t1 + wc4 - 2.0f * t2 + 4.0f * wc2 * k2) / at;
		const float g = highpass ? k4 : wc4;
		a[0] = g / at; a[1] = (highpass ? -4.0f : 4.0f) * g / at; a[2] = 6.0f * g / at; a[3] = a[1]; a[4] = a[0];
	}
	float step(float in) {
		const float out = a[0] * in + a[1] * x[0] + a[2] * x[1] + a[3] * x[2] + a[4] * x[3] - b[1] * y[0] - b[2] * y[1] -
		                  b[3] * y[2] - b[4] * y[3];
		x[3] = x[2]; x[2] = x[1]; x[1] = x[0]; x[0] = in;
		y[3] = y[2]; y[2] = y[1]; y[1] = y[0]; y[0] = out;
		return out;
	}
};
}  // namespace

void split_bands(const std::vector<float>& in, float f1, float f2, float f3, std::vector<float>& low,
                 std::vector<float>& mid, std::vector<float>& high) {
	const float fc1 = (f1 + f2) / 2.0f, fc2 = (f2 + f3) / 2.0f;
	Lr4 hp1(fc1, true), hp2(fc2, true), lp1(fc1, false), lp2(fc2, false);
	low.resize(in.size()); mid.resize(in.size()); high.resize(in.size());
	for (size_t i = 0; i < in.size(); ++i) {
		high[i] = hp2.step(in[i]);
		low[i] = lp1.step(in[i]);
		mid[i] = hp1.step(lp2.step(in[i]));
	}
}

}  // namespace earhost
