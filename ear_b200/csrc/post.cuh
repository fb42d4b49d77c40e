// the post chain on device-resident tracks.
//   Render() after the thread fan-out (src/EAR.cpp:209-228): Recorder::Power(0.335) on every track, the global
//   maximum over all tracks, Recorder::Truncate(getLength(max / 256)); calc T60 then reads RecorderTrack::T60()
//   of the first track (src/EAR.cpp:258-261).  FloatBuffer semantics (src/Recorder.cpp:76-118): Power, Maximum
//   and Multiply run over [first_sample, real_length) -- the last touched bin is excluded -- while getLength scans
//   [first_sample, length).  `length` is the buffer size: n_bins here ("behavioural differences").
// Tracks are short (<= ~1e6 bins) and few: one block per track, every pass a strided loop; nothing here is
// performance-critical, the point is that C5-sized histograms never have to visit the host.
#pragma once
#include "device_exact.cuh"

namespace earb {

constexpr int kPostBlock = 1024;

__device__ __forceinline__ float block_max_f(float v, float* scratch) {
	for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	v = scratch[threadIdx.x & 31];
	for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	__syncthreads();
	return v;
}
__device__ __forceinline__ int block_max_i(int v, int* scratch) {
	for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	v = scratch[threadIdx.x & 31];
	for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	__syncthreads();
	return v;
}
__device__ __forceinline__ int block_min_i(int v, int* scratch) { return -block_max_i(-v, scratch); }

// track t = (context * n_rec + recorder) * tpr + k exists iff it is the first track of its recorder or the recorder is stereo
__device__ __forceinline__ bool track_exists(const ear_b200_recorder* rec, int t, int tpr) {
	return (t % tpr) == 0 || rec[t / tpr].kind == EAR_B200_STEREO;
}

// FloatBuffer::Power (src/Recorder.cpp:101-106) in place + FloatBuffer::Maximum (:76-83) of the result.
// track_max[t] = max |x| over [first_sample, real_length), 0 for tracks that do not exist.
__global__ void __launch_bounds__(kPostBlock) post_power_kernel(float* hist, const uint32_t* range, const ear_b200_recorder* rec,
                                                                int n_bins, int tpr, float exponent, float* track_max) {
	__shared__ float scratch[32];
	const int t = blockIdx.x;
	float mx = 0.0f;
	if (track_exists(rec, t, tpr)) {
		float* x = hist + (size_t)t * n_bins;
		const uint32_t first = range[2 * t], real = min(range[2 * t + 1], (uint32_t)n_bins);
		for (uint32_t i = first + threadIdx.x; i < real; i += kPostBlock) {
			const float v = x[i];
			const float f = pow_ref(fabsf(v), exponent);
			x[i] = v < 0.0f ? fmul(f, -1.0f) : f;
			mx = fmaxf(mx, f);   // |sign * f| = f; a NaN sample never raises the maximum (a > x is false for NaN)
		}
	}
	mx = block_max_f(mx, scratch);
	if (threadIdx.x == 0) track_max[t] = mx;
}

// FloatBuffer::getLength(threshold) (src/Recorder.cpp:108-118): 1 + the last i in [first_sample, length) with
// |x[i]| >= threshold (1 when there is none); real_length when threshold < 0.
__global__ void __launch_bounds__(kPostBlock) post_length_kernel(const float* hist, const uint32_t* range, const ear_b200_recorder* rec,
                                                                 int n_bins, int tpr, float threshold, uint32_t* track_len,
                                                                 uint32_t* track_real) {
	__shared__ int scratch[32];
	const int t = blockIdx.x;
	int last = 0;
	const bool exists = track_exists(rec, t, tpr);
	if (exists && !(threshold < 0.0f)) {
		const float* x = hist + (size_t)t * n_bins;
		for (uint32_t i = range[2 * t] + threadIdx.x; i < (uint32_t)n_bins; i += kPostBlock)
			if (fabsf(x[i]) >= threshold) last = (int)i;
	}
	last = block_max_i(last, scratch);
	if (threadIdx.x == 0) {
		track_len[t] = !exists ? 0u : (threshold < 0.0f ? range[2 * t + 1] : (uint32_t)last + 1u);
		track_real[t] = range[2 * t + 1];   // snapshot: the truncate kernel overwrites real_length while sibling blocks still need it
	}
}

// Recorder::getLength / Recorder::Truncate (src/Recorder.cpp:399-430): the recorder's length is the longest of its
// tracks' (0 when no track holds a sample), every track is truncated to it (0 -> 1); then RecorderTrack::T60
// (src/Recorder.cpp:303-340) of the truncated track: the first sample that drops below its predecessor ends the direct
// lobe, min_gain = predecessor / 10^(60/20), the last later sample above min_gain ends the tail.  When a track has no
// such samples the reference reads uninitialised offsets; like the host port (host/tracks.cpp) this uses 0 for both.
__global__ void __launch_bounds__(kPostBlock) post_truncate_t60_kernel(const float* hist, uint32_t* range, const ear_b200_recorder* rec,
                                                                       int n_bins, int tpr, const uint32_t* track_len,
                                                                       const uint32_t* track_real, float* t60) {
	__shared__ int scratch[32];
	__shared__ float s_min_gain;
	const int t = blockIdx.x, t0 = t - t % tpr;
	if (!track_exists(rec, t, tpr)) { if (threadIdx.x == 0) t60[t] = 0.0f; return; }
	const bool stereo = rec[t / tpr].kind == EAR_B200_STEREO;
	bool has_samples = track_real[t0] > 0 || (stereo && track_real[t0 + 1] > 0);
	uint32_t len = 0;
	if (has_samples) len = max(track_len[t0], stereo ? track_len[t0 + 1] : 0u);
	if (len == 0) len = 1;
	len = min(len, (uint32_t)n_bins);
	const uint32_t first = range[2 * t];
	const float* x = hist + (size_t)t * n_bins;
	// pass 1: end of the direct lobe
	int direct = 0x7fffffff;
	for (uint32_t j = first + threadIdx.x; j < len; j += kPostBlock) {
		const float prev = j == first ? -1.0f : x[j - 1];
		if (x[j] < prev) { direct = (int)j; break; }   // this thread's later candidates are larger
	}
	direct = block_min_i(direct, scratch);
	const bool found = direct != 0x7fffffff;
	if (threadIdx.x == 0) {
		const float prev = !found ? 0.0f : ((uint32_t)direct == first ? -1.0f : x[direct - 1]);
		s_min_gain = fdiv(prev, pow_ref(10.0f, fdiv(60.0f, 20.0f)));
	}
	__syncthreads();
	// pass 2: last significant sample after it
	int last = 0;
	if (found) {
		const float min_gain = s_min_gain;
		for (uint32_t j = (uint32_t)direct + 1u + threadIdx.x; j < len; j += kPostBlock)
			if (x[j] > min_gain) last = (int)j;
	}
	last = block_max_i(last, scratch);
	if (threadIdx.x == 0) {
		t60[t] = fdiv((float)(last - (found ? direct : 0)), 44100.0f);
		range[2 * t + 1] = len;   // Truncate: real_length = l (the buffer already spans n_bins)
	}
}

}  // namespace earb
