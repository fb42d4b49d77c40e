// WAV I/O and the 3-band split of dry source signals (host side; disk / sequential IIR work that
// stays off the GPU).  Behaviour follows lib/wave/WaveFile.cpp:50-256 (PCM 8/16/24-bit -> mono
// float in [-1,1), 16-bit mono/stereo save) and lib/equalizer/Equalizer.cpp:27-96 (4th-order
// Linkwitz-Riley sections, including its constants as shipped).
#pragma once
#include <string>
#include <vector>

namespace earhost {

// Loads a RIFF/WAVE PCM file and mixes its channels down to mono floats. Empty on failure.
std::vector<float> load_wav_mono(const std::string& path);

// 16-bit 44.1 kHz writers. `norm`: scale so that `norm_max` (or the peak when < 0) lands near full scale.
bool save_wav_mono(const std::string& path, const float* data, size_t n, bool norm, float norm_max);
bool save_wav_stereo(const std::string& path, const float* left, size_t n_left, const float* right, size_t n_right, bool norm);

// Splits `in` into low / mid / high bands around the centre frequencies f1 < f2 < f3 (Hz).
void split_bands(const std::vector<float>& in, float f1, float f2, float f3, std::vector<float>& low,
                 std::vector<float>& mid, std::vector<float>& high);

}  // namespace earhost
