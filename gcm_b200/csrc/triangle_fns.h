// TriangleInterpolator of the reference (util/math/interpolation/TriangleInterpolator.hpp:8-130) for scalar values: the
// 2-D member of the simplex interpolators.  __host__ __device__ like the rest of the per-thread
// functions; the arithmetic follows the reference expression by expression (linal/geometry.hpp:108-116,
// linal/linearSystems.hpp:60-88, linal/determinants.hpp:20-35, linal/functions.hpp:327-334,640-679), no FMA contraction.
#pragma once
#include "internal.cuh"

namespace gcmb {
namespace tri2 {

constexpr double TOL = 1e-9;  // EQUALITY_TOLERANCE (util/infrastructure/Types.hpp:10)

struct P2 { double x, y; };

/// linal::barycentricCoordinates(a, b, c, q) in 2-D: false where the reference throws "SLE determinant is zero"
GCMB_HD bool barycentric(P2 a, P2 b, P2 c, P2 q, double (&l)[3]) {
	const double T00 = a.x - c.x, T01 = b.x - c.x;
	const double T10 = a.y - c.y, T11 = b.y - c.y;
	const double r0 = q.x - c.x, r1 = q.y - c.y;
	const double det = T00 * T11 - T01 * T10;
	if (det == 0) { return false; }
	const double det1 = r0 * T11 - T01 * r1;
	const double det2 = T00 * r1 - r0 * T10;
	l[0] = det1 / det;
	l[1] = det2 / det;
	l[2] = 1 - l[0] - l[1];
	return true;
}

/// std::min / std::max as linal's variadic min / max fold them (linal/functions.hpp:560-598)
GCMB_HD double stdmin(double a, double b) { return b < a ? b : a; }
GCMB_HD double stdmax(double a, double b) { return a < b ? b : a; }

GCMB_HD bool isInterpolation(const double (&l)[3]) { return l[0] > -TOL && l[1] > -TOL && l[2] > -TOL; }

/// mode 0: interpolate (linear); 1: interpolate (quadratic); 2: minMaxInterpolate; 3: hybridInterpolate.
/// c, v: the three points and values; g: the three gradients (modes 1-3).  Returns false where the reference throws
/// (degenerate triangle, or q outside it: assert_true(isInterpolation(lambda))).
GCMB_HD bool interpolate(int mode, const P2 (&c)[3], const double (&v)[3], const P2 (&g)[3], P2 q, double& out) {
	double l[3];
	out = 0;
	if (!barycentric(c[0], c[1], c[2], q, l)) { return false; }
	if (!isInterpolation(l)) { return false; }
	const double linear = l[0] * v[0] + l[1] * v[1] + l[2] * v[2];
	if (mode == 0) { out = linear; return true; }
	double quadratic = 0;
	for (int i = 0; i < 3; i++) {
		const double dx = q.x - c[i].x, dy = q.y - c[i].y;
		double dot = g[i].x * dx;
		dot += g[i].y * dy;
		const double term = l[i] * (v[i] + dot / 2.0);
		quadratic = i == 0 ? term : quadratic + term;
	}
	if (mode == 1) { out = quadratic; return true; }
	// limiterMinMax(u, v0, v1, v2) = min(max(u, min(v0, v1, v2)), max(v0, v1, v2))  (linal/functions.hpp:676-679)
	const double lo = stdmin(stdmin(v[0], v[1]), v[2]);
	const double hi = stdmax(stdmax(v[0], v[1]), v[2]);
	const double limited = stdmin(stdmax(quadratic, lo), hi);
	if (mode == 2) { out = limited; return true; }
	out = quadratic == limited ? quadratic : linear;
	return true;
}

/// interpolateInOwner: the first of the triangles (0,1,2), (0,1,3), (0,2,3), (1,2,3) that contains q, linear inside it
GCMB_HD bool interpolateInOwner(const P2 (&c)[4], const double (&v)[4], P2 q, double& out) {
	const int T[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
	out = 0;
	for (int i = 0; i < 4; i++) {
		double l[3];
		if (!barycentric(c[T[i][0]], c[T[i][1]], c[T[i][2]], q, l)) { return false; }   // the reference's solver throws
		if (isInterpolation(l)) {
			out = l[0] * v[T[i][0]] + l[1] * v[T[i][1]] + l[2] * v[T[i][2]];
			return true;
		}
	}
	return false;  // "Containing triangle is not found"
}

/// one query of the test hook gcmb_triangle_interpolate
GCMB_HD void query_thread(int mode, long long i, const double* points, const double* values, const double* grads,
                          const double* queries, double* out, int* status) {
	const P2 q = {queries[2 * i], queries[2 * i + 1]};
	double r = 0;
	bool ok;
	if (mode == 4) {
		P2 c[4];
		double v[4];
		for (int k = 0; k < 4; k++) { c[k] = {points[(4 * i + k) * 2], points[(4 * i + k) * 2 + 1]}; v[k] = values[4 * i + k]; }
		ok = interpolateInOwner(c, v, q, r);
	} else {
		P2 c[3], g[3];
		double v[3];
		for (int k = 0; k < 3; k++) {
			c[k] = {points[(3 * i + k) * 2], points[(3 * i + k) * 2 + 1]};
			v[k] = values[3 * i + k];
			g[k] = mode == 0 ? P2{0, 0} : P2{grads[(3 * i + k) * 2], grads[(3 * i + k) * 2 + 1]};
		}
		ok = interpolate(mode, c, v, g, q, r);
	}
	out[i] = r;
	status[i] = ok ? 0 : 1;
}

}  // namespace tri2
}  // namespace gcmb
