// Geometry of the recorder visibility maps (vismap.cuh): which texel a direction falls into, and the conservative
// footprint of a triangle on a cube face.  Kept free of kernel machinery so that tests/host_emul can compile it for
// the host and check the superset property (every triangle the float test accepts for a segment P -> X is in the
// list of P's texel) against the O(T) loop without a GPU.
#pragma once
#include "traverse.cuh"

namespace earb {

struct VisMapDev {
	const int* offsets;    // [6 * res * res + 1]
	const int* items;      // triangle record slots
	float x[3];            // recorder position
	int res;
};

// face f: major axis m = f >> 1, sign = +1 (even) / -1 (odd); the other two axes in cyclic order
__device__ __forceinline__ int vis_texel(const VisMapDev& mp, float dx, float dy, float dz) {
	const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
	int m = 0; float w = ax;
	if (ay > w) { m = 1; w = ay; }
	if (az > w) { m = 2; w = az; }
	const float major = m == 0 ? dx : m == 1 ? dy : dz;
	const float b = m == 0 ? dy : m == 1 ? dz : dx;
	const float c = m == 0 ? dz : m == 1 ? dx : dy;
	const int face = 2 * m + (major < 0.0f ? 1 : 0);
	if (!(w > 0.0f)) return 0;
	const float u = b / w, v = c / w;   // [-1, 1]
	const int i = min(mp.res - 1, max(0, (int)((u + 1.0f) * 0.5f * (float)mp.res)));
	const int j = min(mp.res - 1, max(0, (int)((v + 1.0f) * 0.5f * (float)mp.res)));
	return (face * mp.res + j) * mp.res + i;
}

// Conservative footprint of triangle `t` on face `face` of the cube map around X: texel rectangle [i0,i1]x[j0,j1].
// When all three vertices lie in front of the face, `edge` also receives the three edge functions of the projected
// triangle, pushed outward by the margin plus half a texel diagonal-wise: texel centre (u, v) can only matter if
// edge[3k] * u + edge[3k+1] * v + edge[3k+2] >= 0 for k = 0..2 (conservative rasterisation; about half of the
// bounding rectangle of a triangle is empty).  has_edges = false means "take the whole rectangle".
__device__ __forceinline__ bool vis_footprint(const SceneDev& sc, int t, const double X[3], int res, int face, double reach,
                                              double maxabs, int& i0, int& i1, int& j0, int& j1, float edge[9], bool& has_edges) {
	has_edges = false;
	const float4 r0 = sc.tris[4 * (size_t)t], r1 = sc.tris[4 * (size_t)t + 1], r2 = sc.tris[4 * (size_t)t + 2];
	double a[3][3] = {{(double)r0.x - X[0], (double)r0.y - X[1], (double)r0.z - X[2]}, {0, 0, 0}, {0, 0, 0}};
	const double e1[3] = {r1.x, r1.y, r1.z}, e2[3] = {r2.x, r2.y, r2.z};
	for (int k = 0; k < 3; ++k) { a[1][k] = a[0][k] + e1[k]; a[2][k] = a[0][k] + e2[k]; }
	// lateral slop of the float test (same bound as the BVH leaf padding), doubled
	const double l1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]), l2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
	const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
	double sinphi = (l1 > 0 && l2 > 0) ? sqrt(cx * cx + cy * cy + cz * cz) / (l1 * l2) : 1.0;
	if (sinphi < 1e-3) sinphi = 1e-3;
	const double eps = 5.9604645e-8;
	const double h = 2.0 * (16.0 * eps * reach / sinphi + 8.0 * eps * (maxabs + reach)) + 4.0 * eps * (l1 + l2);
	// lower bound of the distance from X to the triangle: centroid distance minus the largest centroid-vertex distance
	double g[3] = {(a[0][0] + a[1][0] + a[2][0]) / 3.0, (a[0][1] + a[1][1] + a[2][1]) / 3.0, (a[0][2] + a[1][2] + a[2][2]) / 3.0};
	double rmax = 0.0;
	for (int v = 0; v < 3; ++v) {
		const double dx = a[v][0] - g[0], dy = a[v][1] - g[1], dz = a[v][2] - g[2];
		rmax = fmax(rmax, sqrt(dx * dx + dy * dy + dz * dz));
	}
	const double lb = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) - rmax;
	if (lb < 16.0 * h) { i0 = 0; j0 = 0; i1 = res - 1; j1 = res - 1; return true; }   // X (almost) touches the triangle
	const double margin = 3.1 * (h / lb) + 1e-5;     // angular slop h/lb in projected (tangent) coordinates, |u|,|v| <= 1
	const int m = face >> 1, b = (m + 1) % 3, c = (m + 2) % 3;
	const double sgn = (face & 1) ? -1.0 : 1.0;
	const double wmin = 1e-9 * reach;
	// clip the triangle against w = sgn * a[m] >= wmin (Sutherland-Hodgman), project, bound
	double u0 = 1e30, u1 = -1e30, v0 = 1e30, v1 = -1e30;
	bool any = false;
	for (int k = 0; k < 3; ++k) {
		const double* p = a[k];
		const double* q = a[(k + 1) % 3];
		const double wp = sgn * p[m], wq = sgn * q[m];
		if (wp >= wmin) {
			const double u = p[b] / wp, v = p[c] / wp;
			u0 = fmin(u0, u); u1 = fmax(u1, u); v0 = fmin(v0, v); v1 = fmax(v1, v); any = true;
		}
		if ((wp >= wmin) != (wq >= wmin)) {
			const double s = (wmin - wp) / (wq - wp);
			const double ub = (p[b] + s * (q[b] - p[b])) / wmin, vc = (p[c] + s * (q[c] - p[c])) / wmin;
			u0 = fmin(u0, ub); u1 = fmax(u1, ub); v0 = fmin(v0, vc); v1 = fmax(v1, vc); any = true;
		}
	}
	if (!any) return false;
	{
		const double w0 = sgn * a[0][m], w1 = sgn * a[1][m], w2 = sgn * a[2][m];
		if (w0 >= wmin && w1 >= wmin && w2 >= wmin) {
			const double pu[3] = {a[0][b] / w0, a[1][b] / w1, a[2][b] / w2}, pv[3] = {a[0][c] / w0, a[1][c] / w1, a[2][c] / w2};
			const double orient = (pu[1] - pu[0]) * (pv[2] - pv[0]) - (pv[1] - pv[0]) * (pu[2] - pu[0]);
			const double span = fmax(fmax(fabs(pu[1] - pu[0]), fabs(pu[2] - pu[0])), fmax(fabs(pv[1] - pv[0]), fabs(pv[2] - pv[0])));
			if (fabs(orient) > 1e-12 * span * span) {   // not edge-on
				const double flip = orient > 0.0 ? 1.0 : -1.0, grow = margin + 1.0 / (double)res;
				has_edges = true;
				for (int k = 0; k < 3; ++k) {
					const int k1 = (k + 1) % 3;
					double A = -(pv[k1] - pv[k]) * flip, B = (pu[k1] - pu[k]) * flip;
					const double mx = fmax(fabs(A), fabs(B));
					if (!(mx > 0.0)) { has_edges = false; break; }
					A /= mx; B /= mx;
					double C = -(A * pu[k] + B * pv[k]) + (fabs(A) + fabs(B)) * grow;
					C += 1e-5 + 1e-6 * fabs(C);   // float evaluation of the test below
					edge[3 * k] = (float)A; edge[3 * k + 1] = (float)B; edge[3 * k + 2] = __double2float_ru(C);
				}
			}
		}
	}
	u0 -= margin; u1 += margin; v0 -= margin; v1 += margin;
	if (u1 < -1.0 || u0 > 1.0 || v1 < -1.0 || v0 > 1.0) return false;
	i0 = max(0, min(res - 1, (int)floor((fmax(u0, -1.0) + 1.0) * 0.5 * res)));
	i1 = max(0, min(res - 1, (int)floor((fmin(u1, 1.0) + 1.0) * 0.5 * res)));
	j0 = max(0, min(res - 1, (int)floor((fmax(v0, -1.0) + 1.0) * 0.5 * res)));
	j1 = max(0, min(res - 1, (int)floor((fmin(v1, 1.0) + 1.0) * 0.5 * res)));
	return true;
}

// texel (i, j) of the face belongs to the footprint iff its centre passes the three pushed-out edge functions
__device__ __forceinline__ bool vis_covers(const float edge[9], int res, int i, int j) {
	const float texel = 2.0f / (float)res;
	const float u = ((float)i + 0.5f) * texel - 1.0f, v = ((float)j + 0.5f) * texel - 1.0f;
	return edge[0] * u + edge[1] * v + edge[2] >= 0.0f && edge[3] * u + edge[4] * v + edge[5] >= 0.0f &&
	       edge[6] * u + edge[7] * v + edge[8] >= 0.0f;
}

}  // namespace earb
