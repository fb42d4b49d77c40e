// TEST INFRASTRUCTURE: lets the device headers (device_exact.cuh, traverse.cuh) compile as plain
// C++ so the BVH builder + traversal LOGIC (padding, slack, tie-breaks) can be checked against the
// oracle on a machine without a GPU.  Never part of the product; built only by tests/.
// Compile with -ffp-contract=off so a*b+c is not fused, like the explicit *_rn intrinsics.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#define __device__
#define __forceinline__ inline
#define __host__
struct float4 { float x, y, z, w; };
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int32_t __float_as_int(float f) { int32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldg(const T* p) { return *p; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }   // a one-lane warp
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __double2float_ru(double x) { float f = (float)x; return (double)f < x ? nextafterf(f, INFINITY) : f; }
static inline float __double2float_rd(double x) { float f = (float)x; return (double)f > x ? nextafterf(f, -INFINITY) : f; }
using std::min;
using std::max;
#define EARB_HOST_EMULATION 1
