/*
 * TEST INFRASTRUCTURE -- stand-in for GMTL 0.6.1 (Generic Math Template Library,
 * ggt.sourceforge.net), which the reference depends on (README.md:28-30) but does
 * not vendor.  Only the surface the reference touches is provided.  This header is
 * used for ONE purpose: compiling the unmodified reference sources under
 * /root/reference into oracle/_ref/ so the CPU restatement (oracle/ear_oracle.cpp)
 * can be pinned against the reference's own control flow.  It is never part of the
 * product (ear_b200/), which has no CPU path at all.
 *
 * Arithmetic contract (float32, no FMA, left-to-right):
 *   dot(a,b)      = (a0*b0 + a1*b1) + a2*b2
 *   cross(a,b)    = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
 *   lengthSquared = (a0*a0 + a1*a1) + a2*a2 ; length = sqrtf(lengthSquared)
 *   normalize(v)  : len = length(v); if len != 0 each component is DIVIDED by len
 *   reflect(v,n)  = v - 2*(dot(v,n)*n)
 *   intersectDoubleSided = Moeller-Trumbore, non-culling branch, EPSILON 1e-5,
 *                          accepts t >= 0  (reference call sites: src/Mesh.cpp:40,65)
 * The exact summation order inside GMTL 0.6.1 could not be checked on disk; this
 * file DEFINES the contract that shim-built reference, oracle and GPU all follow.
 *
 * Instrumentation: intersectDoubleSided counts calls per static argument type
 * (Ray = closest-hit loop of Mesh::RayIntersection, LineSeg = occlusion loop of
 * Mesh::LineIntersection) so the harness can report ray-bounce segments without
 * patching the reference: segments = ray_tests / triangle_count.
 */
#ifndef EAR_B200_GMTL_SHIM_H
#define EAR_B200_GMTL_SHIM_H

#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <iostream>
#include <sstream>
#include <iomanip>
#include <algorithm>
#include <vector>
#include <string>

namespace gmtl {

struct ShimCounters { unsigned long long ray_tests, seg_tests; };
inline ShimCounters& shim_counters() { static __thread ShimCounters c = {0, 0}; return c; }

template <class T, unsigned N> struct VecBase {
	T mData[N];
	VecBase() { for (unsigned i = 0; i < N; ++i) mData[i] = T(0); }
	T& operator[](unsigned i) { return mData[i]; }
	const T& operator[](unsigned i) const { return mData[i]; }
};

template <class T, unsigned N> struct Vec : public VecBase<T, N> {
	Vec() {}
	Vec(const VecBase<T, N>& o) : VecBase<T, N>(o) {}
	Vec(T x, T y, T z) { this->mData[0] = x; this->mData[1] = y; this->mData[2] = z; }
	void set(T x, T y, T z) { this->mData[0] = x; this->mData[1] = y; this->mData[2] = z; }
};

template <class T, unsigned N> struct Point : public VecBase<T, N> {
	Point() {}
	Point(const VecBase<T, N>& o) : VecBase<T, N>(o) {}
	Point(T x, T y, T z) { this->mData[0] = x; this->mData[1] = y; this->mData[2] = z; }
};

typedef Vec<float, 3> Vec3f;
typedef Point<float, 3> Point3f;

template <class T, unsigned N>
inline VecBase<T, N> operator+(const VecBase<T, N>& a, const VecBase<T, N>& b) {
	VecBase<T, N> r; for (unsigned i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r;
}
template <class T, unsigned N>
inline VecBase<T, N> operator-(const VecBase<T, N>& a, const VecBase<T, N>& b) {
	VecBase<T, N> r; for (unsigned i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r;
}
template <class T, unsigned N>
inline VecBase<T, N> operator-(const VecBase<T, N>& a) {
	VecBase<T, N> r; for (unsigned i = 0; i < N; ++i) r[i] = -a[i]; return r;
}
template <class T, unsigned N>
inline VecBase<T, N> operator*(const VecBase<T, N>& a, const T& s) {
	VecBase<T, N> r; for (unsigned i = 0; i < N; ++i) r[i] = a[i] * s; return r;
}
template <class T, unsigned N>
inline VecBase<T, N> operator*(const T& s, const VecBase<T, N>& a) {
	VecBase<T, N> r; for (unsigned i = 0; i < N; ++i) r[i] = a[i] * s; return r;
}
template <class T, unsigned N, class S>
inline VecBase<T, N>& operator/=(VecBase<T, N>& a, const S& s) {
	for (unsigned i = 0; i < N; ++i) a[i] /= (T)s; return a;
}

template <class T, unsigned N>
inline T dot(const VecBase<T, N>& a, const VecBase<T, N>& b) {
	T r = a[0] * b[0];
	for (unsigned i = 1; i < N; ++i) r = r + a[i] * b[i];
	return r;
}
template <class T>
inline Vec<T, 3>& cross(Vec<T, 3>& result, const VecBase<T, 3>& a, const VecBase<T, 3>& b) {
	result.set((a[1] * b[2]) - (a[2] * b[1]),
	           (a[2] * b[0]) - (a[0] * b[2]),
	           (a[0] * b[1]) - (a[1] * b[0]));
	return result;
}
template <class T, unsigned N>
inline T lengthSquared(const VecBase<T, N>& a) {
	T r = a[0] * a[0];
	for (unsigned i = 1; i < N; ++i) r = r + a[i] * a[i];
	return r;
}
template <class T, unsigned N>
inline T length(const VecBase<T, N>& a) { return (T)sqrtf(lengthSquared(a)); }

template <class T, unsigned N>
inline T normalize(VecBase<T, N>& a) {
	const T len = length(a);
	if (len != T(0)) { for (unsigned i = 0; i < N; ++i) a[i] /= len; }
	return len;
}
template <class T, unsigned N>
inline Vec<T, N> makeNormal(const VecBase<T, N>& a) {
	Vec<T, N> r(a); normalize(r); return r;
}
template <class T, unsigned N>
inline VecBase<T, N>& reflect(VecBase<T, N>& result, const VecBase<T, N>& v, const VecBase<T, N>& n) {
	const T d = dot(v, n);
	T tmp[N];
	for (unsigned i = 0; i < N; ++i) tmp[i] = v[i] - T(2) * (d * n[i]);
	for (unsigned i = 0; i < N; ++i) result[i] = tmp[i];
	return result;
}

template <class T> struct Ray {
	Point<T, 3> mOrigin;
	Vec<T, 3> mDir;
	Ray() {}
	Ray(const Point<T, 3>& o, const Vec<T, 3>& d) : mOrigin(o), mDir(d) {}
	const Point<T, 3>& getOrigin() const { return mOrigin; }
	const Vec<T, 3>& getDir() const { return mDir; }
};
template <class T> struct LineSeg : public Ray<T> {
	LineSeg() {}
	LineSeg(const Point<T, 3>& p, const Point<T, 3>& q) : Ray<T>(p, Vec<T, 3>(q - p)) {}
	T getLength() const { return length(this->mDir); }
};
typedef Ray<float> Rayf;
typedef LineSeg<float> LineSegf;

template <class T> struct Tri {
	Point<T, 3> mVerts[3];
	Tri() {}
	Tri(const Point<T, 3>& a, const Point<T, 3>& b, const Point<T, 3>& c) { mVerts[0] = a; mVerts[1] = b; mVerts[2] = c; }
	Point<T, 3>& operator[](int i) { return mVerts[i]; }
	const Point<T, 3>& operator[](int i) const { return mVerts[i]; }
	Vec<T, 3> edge(int i) const { return Vec<T, 3>(mVerts[(i + 1) % 3] - mVerts[i]); }
};
typedef Tri<float> Trif;

template <class T>
inline Vec<T, 3> normal(const Tri<T>& tri) {
	Vec<T, 3> n;
	cross(n, tri[1] - tri[0], tri[2] - tri[0]);
	normalize(n);
	return n;
}

template <class T>
inline bool shim_moeller_trumbore(const Tri<T>& tri, const Ray<T>& ray, float& u, float& v, float& t) {
	const float EPSILON = 0.00001f;
	const Vec<T, 3> edge1(tri[1] - tri[0]);
	const Vec<T, 3> edge2(tri[2] - tri[0]);
	Vec<T, 3> pvec, qvec;
	cross(pvec, ray.getDir(), edge2);
	const float det = dot(edge1, pvec);
	if (det > -EPSILON && det < EPSILON) return false;
	const float inv_det = 1.0f / det;
	const Vec<T, 3> tvec(ray.getOrigin() - tri[0]);
	u = dot(tvec, pvec) * inv_det;
	if (u < 0.0f || u > 1.0f) return false;
	cross(qvec, tvec, edge1);
	v = dot(ray.getDir(), qvec) * inv_det;
	if (v < 0.0f || u + v > 1.0f) return false;
	t = dot(edge2, qvec) * inv_det;
	return t >= 0.0f;
}
template <class T>
inline bool intersectDoubleSided(const Tri<T>& tri, const Ray<T>& ray, float& u, float& v, float& t) {
	++shim_counters().ray_tests;
	return shim_moeller_trumbore(tri, ray, u, v, t);
}
template <class T>
inline bool intersectDoubleSided(const Tri<T>& tri, const LineSeg<T>& seg, float& u, float& v, float& t) {
	++shim_counters().seg_tests;
	return shim_moeller_trumbore(tri, static_cast<const Ray<T>&>(seg), u, v, t);
}

namespace Math {
inline void seedRandom(unsigned int s) { srand(s); }
inline float unitRandom() { return float(rand()) / float(RAND_MAX); }
inline float rangeRandom(float lo, float hi) { const float r = unitRandom(); const float size = hi - lo; return r * size + lo; }
inline float sqrt(float x) { return sqrtf(x); }
}

template <class T, unsigned N>
inline std::ostream& operator<<(std::ostream& o, const VecBase<T, N>& v) {
	o << "(";
	for (unsigned i = 0; i < N; ++i) { if (i) o << ", "; o << v[i]; }
	o << ")";
	return o;
}

} // namespace gmtl

#endif
