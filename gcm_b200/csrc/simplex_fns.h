// Per-thread work functions of the simplex (tetrahedral) path: characteristic-foot cell location by line
// walk, least-squares gradients, hybrid quadratic interpolation, space-time interpolation at border facets and
// the outer-wave border correction.  Like thread_fns.h these are __host__ __device__ so that the host-logic
// tests can step through them; the arithmetic follows the reference expression by expression (fp64, no FMA):
// the integer results of the cell location must equal the reference's bit for bit.
//
// Reference (relative to src/libgcm): grid/simplex/SimplexGrid.cpp:61-164, grid/simplex/cgal/LineWalker.hpp:26-89,
// grid/simplex/cgal/Cgal3DTriangulation.hpp:160-292, linal/geometry.hpp, linal/linearSystems.hpp:46-158,
// util/math/Differentiation.hpp:33-63, util/math/interpolation/TetrahedronInterpolator.hpp:15-155,
// engine/simplex/GridCharacteristicMethodInRiemannInvariants.hpp:44-198, engine/simplex/common.hpp:48-260,
// engine/simplex/BorderCorrector.hpp:122-286, rheology/models/ElasticModel.hpp:111-232, AcousticModel.hpp:95-147.
#pragma once
#include "internal.cuh"

namespace gcmb {
namespace sx {

constexpr double TOL = 1e-9;       // EQUALITY_TOLERANCE (util/infrastructure/Types.hpp:10)
constexpr int MAX_NEIGHBORS = 20;  // Cgal3DTriangulation.hpp:53
constexpr int EMPTY = -1;          // CellInfo::EmptySpaceFlag

struct V3 {
	double x, y, z;
	GCMB_HD double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
GCMB_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
GCMB_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
GCMB_HD V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
GCMB_HD V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
GCMB_HD double dot(V3 a, V3 b) { double r = a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
GCMB_HD double length(V3 a) { return sqrt(dot(a, a)); }
GCMB_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

GCMB_HD double det2(double a, double b, double c, double d) { return a * d - b * c; }
GCMB_HD double det3(double a11, double a12, double a13, double a21, double a22, double a23,
                    double a31, double a32, double a33) {
	return a11 * (a22 * a33 - a23 * a32) - a12 * (a21 * a33 - a23 * a31) + a13 * (a21 * a32 - a22 * a31);
}

// Cramer's rule; false when the reference throws "SLE determinant is zero"
GCMB_HD bool solve3(const double (&A)[3][3], const double (&b)[3], double (&x)[3]) {
	const double det = det3(A[0][0], A[0][1], A[0][2], A[1][0], A[1][1], A[1][2], A[2][0], A[2][1], A[2][2]);
	if (det == 0) { return false; }
	x[0] = det3(b[0], A[0][1], A[0][2], b[1], A[1][1], A[1][2], b[2], A[2][1], A[2][2]) / det;
	x[1] = det3(A[0][0], b[0], A[0][2], A[1][0], b[1], A[1][2], A[2][0], b[2], A[2][2]) / det;
	x[2] = det3(A[0][0], A[0][1], b[0], A[1][0], A[1][1], b[1], A[2][0], A[2][1], b[2]) / det;
	return true;
}

// least squares with unit weights: 3x2 system (columns p, q) and 3x1 system (column p)
GCMB_HD bool leastSquares2(V3 p, V3 q, V3 rhs, double (&x)[2]) {
	const V3 c[2] = {p, q};
	double N[2][2], b[2];
	for (int i = 0; i < 2; i++) {
		for (int j = 0; j < 2; j++) {
			double r = c[i].x * (1.0 * c[j].x); r += c[i].y * (1.0 * c[j].y); r += c[i].z * (1.0 * c[j].z);
			N[i][j] = r;
		}
		double r = c[i].x * (1.0 * rhs.x); r += c[i].y * (1.0 * rhs.y); r += c[i].z * (1.0 * rhs.z);
		b[i] = r;
	}
	const double det = det2(N[0][0], N[0][1], N[1][0], N[1][1]);
	if (det == 0) { return false; }
	x[0] = det2(b[0], N[0][1], b[1], N[1][1]) / det;
	x[1] = det2(N[0][0], b[0], N[1][0], b[1]) / det;
	return true;
}
GCMB_HD bool leastSquares1(V3 p, V3 rhs, double& x) {
	double a = p.x * (1.0 * p.x); a += p.y * (1.0 * p.y); a += p.z * (1.0 * p.z);
	double b = p.x * (1.0 * rhs.x); b += p.y * (1.0 * rhs.y); b += p.z * (1.0 * rhs.z);
	if (a == 0) { return false; }
	x = b / a;
	return true;
}

GCMB_HD double orientedVolume(V3 a, V3 b, V3 c, V3 d) {
	const V3 ba = b - a, ca = c - a, da = d - a;
	return det3(ba.x, ba.y, ba.z, ca.x, ca.y, ca.z, da.x, da.y, da.z) / 6;
}
GCMB_HD bool barycentric(V3 a, V3 b, V3 c, V3 d, V3 q, double (&l)[4]) {
	const double T[3][3] = {{a.x - d.x, b.x - d.x, c.x - d.x}, {a.y - d.y, b.y - d.y, c.y - d.y}, {a.z - d.z, b.z - d.z, c.z - d.z}};
	const double rhs[3] = {q.x - d.x, q.y - d.y, q.z - d.z};
	double x[3];
	if (!solve3(T, rhs, x)) { return false; }
	l[0] = x[0]; l[1] = x[1]; l[2] = x[2]; l[3] = 1 - x[0] - x[1] - x[2];
	return true;
}
GCMB_HD double area(V3 a, V3 b, V3 c) { return length(cross(b - a, c - a)) / 2; }
GCMB_HD double volume(V3 a, V3 b, V3 c, V3 d) { return fabs(orientedVolume(a, b, c, d)); }
GCMB_HD bool degenerateTetra(V3 a, V3 b, V3 c, V3 d, double eps) {
	const double V = volume(a, b, c, d);
	const double A = area(b, c, d), B = area(c, d, a), C = area(d, a, b), D = area(a, b, c);
	const double h = 3 * V / fmax(A, fmax(B, fmax(C, D)));
	const double l = (length(a - b) + length(a - c) + length(a - d) + length(d - b) + length(d - c) + length(b - c)) / 6;
	return h <= eps * l;
}
GCMB_HD bool degenerateTriangle(V3 a, V3 b, V3 c, double eps) {
	const double S = area(a, b, c);
	const double ab = length(a - b), ac = length(a - c), bc = length(b - c);
	const double h = 2 * S / fmax(ab, fmax(ac, bc));
	const double l = (length(a - b) + length(a - c) + length(b - c)) / 3;
	return h <= eps * l;
}
GCMB_HD bool segmentContains(V3 a, V3 b, V3 q, double eps, double degEps, int& err) {
	if (!degenerateTriangle(a, b, q, degEps)) { return false; }
	double x;
	if (!leastSquares1(a - b, q - b, x)) { err = 1; return false; }
	return x >= -eps && 1 - x >= -eps;
}
GCMB_HD bool triangleContains(V3 a, V3 b, V3 c, V3 q, double eps, double degEps, int& err) {
	if (!degenerateTetra(a, b, c, q, degEps)) { return false; }
	double x[2];
	if (!leastSquares2(a - c, b - c, q - c, x)) { err = 1; return false; }
	return x[0] >= -eps && x[1] >= -eps && 1 - x[0] - x[1] >= -eps;
}
GCMB_HD bool tetrahedronContains(V3 a, V3 b, V3 c, V3 d, V3 q, double eps, int& err) {
	double l[4];
	if (!barycentric(a, b, c, d, q, l)) { err = 1; return false; }
	return l[0] >= -eps && l[1] >= -eps && l[2] >= -eps && l[3] >= -eps;
}
GCMB_HD bool solidAngleContains(V3 a, V3 b, V3 c, V3 d, V3 q, double eps, int& err) {
	double l[4];
	if (!barycentric(a, b, c, d, q, l)) { err = 1; return false; }
	return l[0] <= 1 + eps && l[1] >= -eps && l[2] >= -eps && l[3] >= -eps;
}
GCMB_HD bool lineFlatIntersection(V3 f1, V3 f2, V3 f3, V3 l1, V3 l2, V3& out) {
	const V3 tau = l2 - l1, p = f2 - f1, q = f3 - f1;
	const double A[3][3] = {{tau.x, -p.x, -q.x}, {tau.y, -p.y, -q.y}, {tau.z, -p.z, -q.z}};
	const double b[3] = {f1.x - l1.x, f1.y - l1.y, f1.z - l1.z};
	double x[3];
	if (!solve3(A, b, x)) { return false; }
	out = l1 + tau * x[0];
	return true;
}

// ---------------------------------------------------------------------------------------------
// the body's view of the flat triangulation (all arrays in device memory)
// ---------------------------------------------------------------------------------------------
struct Tri {
	int nV, nC, nLocal, gridId;
	const double* xyz;      // [nV][3]
	const int* cellV;       // [nC][4] global vertex ids
	const int* cellN;       // [nC][4]
	const int* cellGrid;    // [nC]
	const int* incOff;      // [nV+1]
	const int* incCell;
	const int* localOf;     // [nV]
	const int* globalOf;    // [nLocal]
	const int* nbOff;       // [nLocal+1] neighbour vertices (ascending local id, at most MAX_NEIGHBORS)
	const int* nbIdx;
	const uint8_t* state;   // [nLocal] 0 inner, 1 border, 2 contact, 3 multicontact
	const int* locOff;      // [nLocal+1] the body's own cells around a local vertex, ascending cell id ...
	const int4* locABC;     // ... as {cell, a, b, c}: the other three vertices in the order findCrossedIncidentCell takes them
	// direction masks (k_s_dir_masks): bit k of dirMask[lv * DIR_BUCKETS + bucket] = entry locOff[lv] + k may contain a ray
	// from the vertex whose direction falls into the bucket; every entry whose bit is clear is PROVEN not to contain it
	const unsigned long long* dirMask = nullptr;  // or null
	const float* dirMinShift = nullptr;  // [nLocal] shortest ray the masks hold for; < 0: no masks for this vertex (> 64 cells)

	GCMB_HD V3 point(int g) const { return {xyz[3 * (long long) g], xyz[3 * (long long) g + 1], xyz[3 * (long long) g + 2]}; }
	GCMB_HD V3 localPoint(int l) const { return point(globalOf[l]); }
	GCMB_HD bool isLocal(int cell) const { return cell >= 0 && cellGrid[cell] == gridId; }
	GCMB_HD int otherVertexIndex(int cell, int a, int b, int c) const {
		for (int i = 0; i < 4; i++) {
			const int d = cellV[4 * (long long) cell + i];
			if (d != a && d != b && d != c) { return i; }
		}
		return 0;
	}
	GCMB_HD int otherVertex(int cell, int a, int b, int c) const { return cellV[4 * (long long) cell + otherVertexIndex(cell, a, b, c)]; }
};

// n = 4: a cell, with the barycentric coordinates of the query that the containment test computed (the
// interpolation needs exactly these: same points in the same order, same query)
struct Found { int n; int p[4]; double l[4]; };
GCMB_HD Found nothing() { return {0, {-1, -1, -1, -1}, {0, 0, 0, 0}}; }
GCMB_HD Found wholeCell(const Tri& t, int cell, const double (&l)[4]) {
	Found f;
	f.n = 4;
	for (int i = 0; i < 4; i++) { f.p[i] = t.localOf[t.cellV[4 * (long long) cell + i]]; f.l[i] = l[i]; }
	return f;
}

// the reference keeps the list of crossed cells; only its size, the last two cells and the facet the walk
// left the previous cell through are ever used
struct Walked { int count, last, prev, exitSlot; };

// exact value of (d / det < 0) without dividing: the quotient of two finite non-zero doubles has the sign of
// d * det unless it underflows to zero, which the magnitude guard excludes
GCMB_HD bool quotientNegative(double d, double det) {
	if (d == 0) { return false; }
	if (fabs(d) >= 1e-200 && fabs(det) <= 1e+100) { return (d < 0) != (det < 0); }
	return d / det < 0;
}

// solidAngleContains(apex, b, c, d, q, eps = 0) (linal/geometry.hpp:326-334 through barycentricCoordinates
// :201-217 and the Cramer solve linearSystems.hpp:104-129).  Same determinants as `barycentric`; the three
// divisions are only carried out when the sign conditions that do not need them already hold, which leaves the
// outcome unchanged: l0 <= 1 <=> d0 <= det (det > 0) or d0 >= det (det < 0) because rounding is monotonic and
// 1 is representable; l1, l2 >= -0 <=> not (d / det < 0).
GCMB_HD bool solidAngleContainsExact0(V3 a, V3 b, V3 c, V3 d, V3 q, int& err) {
	const double T00 = a.x - d.x, T01 = b.x - d.x, T02 = c.x - d.x;
	const double T10 = a.y - d.y, T11 = b.y - d.y, T12 = c.y - d.y;
	const double T20 = a.z - d.z, T21 = b.z - d.z, T22 = c.z - d.z;
	const double r0 = q.x - d.x, r1 = q.y - d.y, r2 = q.z - d.z;
	const double det = det3(T00, T01, T02, T10, T11, T12, T20, T21, T22);
	if (det == 0) { err = 1; return false; }
	const double d1 = det3(T00, r0, T02, T10, r1, T12, T20, r2, T22);
	if (quotientNegative(d1, det)) { return false; }
	const double d2 = det3(T00, T01, r0, T10, T11, r1, T20, T21, r2);
	if (quotientNegative(d2, det)) { return false; }
	const double d0 = det3(r0, T01, T02, r1, T11, T12, r2, T21, T22);
	if (det > 0 ? !(d0 <= det) : !(d0 >= det)) { return false; }
	const double x0 = d0 / det, x1 = d1 / det, x2 = d2 / det;
	const double l3 = 1 - x0 - x1 - x2;
	return x0 <= 1 + 0.0 && x1 >= -0.0 && x2 >= -0.0 && l3 >= -0.0;
}

// ---- direction buckets ------------------------------------------------------------------------------------------------
// The sphere of ray directions around a vertex is cut into 6 x DIR_N x DIR_N buckets (faces of a cube, a regular grid on
// each).  For every (vertex, bucket) a bit mask over the vertex's incident cells says which cells' solid angles can meet
// the bucket at all; findCrossedIncidentCell's loop -- "the first incident cell, in the table's order, whose solid angle
// contains the ray" -- then runs its exact test only on the cells of the mask, in the same order.  A cell is left out only
// when one face plane of its solid angle has ALL FOUR corner rays of the (padded) bucket strictly on its outer side by a
// relative margin of DIR_MARGIN: then every ray of the bucket has a barycentric coordinate below -DIR_MARGIN * |row| * |ray|
// in that cell, ten orders of magnitude beyond what rounding in the exact fp64 test (relative 1e-15 of |row| * (cell size +
// |ray|)) could turn into "inside" -- for rays no shorter than dirMinShift = DIR_MIN_RAY x the largest edge around the vertex.
// So the answer, cell for cell, is the reference's.  (The locate protocol of 428 544 queries and all engine fixtures stay
// bit-identical; profiles/r2_simplex_dir_masks.md has the timings.)
constexpr int DIR_N = 4;
constexpr int DIR_BUCKETS = 6 * DIR_N * DIR_N;
constexpr double DIR_PAD = 1e-3;      // the buckets overlap by this much (in face coordinates): rounding of dirBucket, ties
constexpr double DIR_MARGIN = 1e-4;
constexpr double DIR_MIN_RAY = 1e-6;

GCMB_HD int dirBucket(V3 d) {
	const double ax = fabs(d.x), ay = fabs(d.y), az = fabs(d.z);
	const int m = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
	const double dm = m == 0 ? d.x : (m == 1 ? d.y : d.z);
	const double da = m == 0 ? d.y : d.x;                       // the other two axes in ascending order
	const double db = m == 2 ? d.y : d.z;
	const double inv = 1.0 / fabs(dm);
	int iu = (int) ((da * inv + 1.0) * (0.5 * DIR_N));
	int iv = (int) ((db * inv + 1.0) * (0.5 * DIR_N));
	iu = iu < 0 ? 0 : (iu > DIR_N - 1 ? DIR_N - 1 : iu);
	iv = iv < 0 ? 0 : (iv > DIR_N - 1 ? DIR_N - 1 : iv);
	return ((2 * m + (dm < 0 ? 1 : 0)) * DIR_N + iu) * DIR_N + iv;
}

// corner ray `c` (0..3) of a bucket, padded
GCMB_HD V3 dirBucketCorner(int bucket, int c) {
	const int iv = bucket % DIR_N, iu = (bucket / DIR_N) % DIR_N, face = bucket / (DIR_N * DIR_N);
	const int m = face / 2;
	const double sgn = (face & 1) ? -1.0 : 1.0;
	const double u = (c & 1) ? (iu + 1) * (2.0 / DIR_N) - 1.0 + DIR_PAD : iu * (2.0 / DIR_N) - 1.0 - DIR_PAD;
	const double v = (c & 2) ? (iv + 1) * (2.0 / DIR_N) - 1.0 + DIR_PAD : iv * (2.0 / DIR_N) - 1.0 - DIR_PAD;
	return m == 0 ? V3{sgn, u, v} : (m == 1 ? V3{u, sgn, v} : V3{u, v, sgn});
}

// one thread per (local vertex, bucket)
GCMB_HD void dirMaskThread(const Tri& t, int lv, int bucket, unsigned long long* masks, float* minShift) {
	const V3 apex = t.localPoint(lv);
	const int begin = t.locOff[lv], end = t.locOff[lv + 1];
	if (end - begin > 64) {
		if (bucket == 0) { minShift[lv] = -1.0f; }
		masks[(long long) lv * DIR_BUCKETS + bucket] = ~0ull;
		return;
	}
	V3 corner[4];
	double cl[4];
	for (int c = 0; c < 4; c++) { corner[c] = dirBucketCorner(bucket, c); cl[c] = length(corner[c]); }
	unsigned long long mask = 0;
	double hmax = 0;
	for (int i = begin; i < end; i++) {
		const int4 e = t.locABC[i];
		const V3 pa = t.point(e.y), pb = t.point(e.z), pc = t.point(e.w);
		const V3 e1 = pa - apex, e2 = pb - apex, e3 = pc - apex;
		const V3 n[3] = {cross(e2, e3), cross(e3, e1), cross(e1, e2)};  // rows of the inverse of [e1 e2 e3], times det
		const double det = dot(e1, n[0]);
		const double edges[6] = {length(e1), length(e2), length(e3), length(pa - pb), length(pb - pc), length(pc - pa)};
		for (int k = 0; k < 6; k++) { hmax = edges[k] > hmax ? edges[k] : hmax; }
		bool out = false;
		if (det != 0 && det == det && det - det == 0) {   // (finite and non-zero; a degenerate cell is always tested exactly)
			const double sgn = det > 0 ? 1.0 : -1.0;
			for (int r = 0; r < 3 && !out; r++) {
				const double nl = length(n[r]);
				bool all = true;
				for (int c = 0; c < 4; c++) { all = all && (sgn * dot(n[r], corner[c]) < -DIR_MARGIN * nl * cl[c]); }
				out = all;
			}
		}
		if (!out) { mask |= 1ull << (i - begin); }
	}
	masks[(long long) lv * DIR_BUCKETS + bucket] = mask;
	if (bucket == 0) { minShift[lv] = (float) (DIR_MIN_RAY * hmax) * 1.0001f; }
}

GCMB_HD int lowestBit(unsigned long long m) {
#if defined(__CUDA_ARCH__)
	return __ffsll((long long) m) - 1;
#else
	return __builtin_ffsll((long long) m) - 1;
#endif
}

// Cgal3DTriangulation::findCrossedIncidentCell over the body's own cells around the vertex (table built at
// body creation); returns the table entry {cell, a, b, c} or cell = -2.  The loop is software-pipelined: the
// entry two cells ahead and the points of the next cell are in flight while the current cell is tested.
GCMB_HD int4 crossedIncidentCell(const Tri& t, int lv, V3 query, double eps, int& err) {
	const V3 apex = t.localPoint(lv);
	const int begin = t.locOff[lv], end = t.locOff[lv + 1];
	if (eps == 0 && t.dirMask != nullptr) {
		const V3 ray = query - apex;
		const double ms = (double) t.dirMinShift[lv];
		if (ms >= 0 && dot(ray, ray) > ms * ms) {
			// the same loop over the cells the ray can lie in at all (see "direction buckets" above)
			unsigned long long m = t.dirMask[(long long) lv * DIR_BUCKETS + dirBucket(ray)];
			while (m) {
				const int4 e = t.locABC[begin + lowestBit(m)];
				m &= m - 1;
				if (solidAngleContainsExact0(apex, t.point(e.y), t.point(e.z), t.point(e.w), query, err)) { return e; }
			}
			return make_int4(-2, -1, -1, -1);
		}
	}
	int4 e = t.locABC[begin];
	int4 e1 = begin + 1 < end ? t.locABC[begin + 1] : e;
	V3 pa = t.point(e.y), pb = t.point(e.z), pc = t.point(e.w);
	for (int i = begin; i < end; i++) {
		const int4 e2 = i + 2 < end ? t.locABC[i + 2] : e1;
		const V3 na = t.point(e1.y), nb = t.point(e1.z), nc = t.point(e1.w);
		const bool inside = eps == 0 ? solidAngleContainsExact0(apex, pa, pb, pc, query, err)
		                             : solidAngleContains(apex, pa, pb, pc, query, eps, err);
		if (inside) { return e; }
		e = e1; e1 = e2;
		pa = na; pb = nb; pc = nc;
	}
	return make_int4(-2, -1, -1, -1);
}

GCMB_HD Walked collectCells(const Tri& t, V3 q, V3 p, int cell, int u, int v, int w) {
	Walked ans = {1, cell, -2, -1};
	for (int guard = 0; guard < 100000; guard++) {
		if (!(orientedVolume(t.point(u), t.point(v), t.point(w), p) < 0)) { break; }
		ans.exitSlot = t.otherVertexIndex(cell, u, v, w);
		cell = t.cellN[4 * (long long) cell + ans.exitSlot];
		ans.prev = ans.last; ans.last = cell; ans.count++;
		if (!t.isLocal(cell)) { break; }
		const int s = t.otherVertex(cell, u, v, w);
		const V3 ps = t.point(s);
		if (orientedVolume(t.point(u), ps, q, p) > 0) {
			if (orientedVolume(t.point(v), ps, q, p) > 0) { u = s; } else { w = s; }
		} else {
			if (orientedVolume(t.point(w), ps, q, p) > 0) { v = s; } else { u = s; }
		}
	}
	return ans;
}

GCMB_HD Walked walkFromVertex(const Tri& t, int lv, int gv, V3 p, int& err) {
	const int4 e = crossedIncidentCell(t, lv, p, 0, err);
	const int cell = e.x;
	if (cell == -2) { return {0, -2, -2, -1}; }
	int u = e.y, v = e.z;   // otherVertex(cell, gv, gv, gv), (cell, gv, gv, u), (cell, gv, u, v)
	const int w = e.w;
	if (orientedVolume(t.point(u), t.point(v), t.point(w), t.point(gv)) < 0) { const int x = u; u = v; v = x; }
	return collectCells(t, t.point(gv), p, cell, u, v, w);
}

GCMB_HD Walked walkFromCell(const Tri& t, int cell, V3 q, V3 p, int& err) {
	int u = -1, v = -1, w = -1;
	for (int attempt = 0; attempt < 2 && u < 0; attempt++) {
		const double eps = attempt == 0 ? 0.0 : TOL;
		for (int i = 0; i < 4; i++) {
			const int a1 = t.cellV[4 * (long long) cell + (i + 1) % 4];
			const int b1 = t.cellV[4 * (long long) cell + (i + 2) % 4];
			const int c1 = t.cellV[4 * (long long) cell + (i + 3) % 4];
			if (solidAngleContains(q, t.point(a1), t.point(b1), t.point(c1), p, eps, err)) { u = a1; v = b1; w = c1; break; }
		}
	}
	if (u < 0) { return {0, -2, -2, -1}; }
	if (orientedVolume(t.point(u), t.point(v), t.point(w), q) < 0) { const int x = u; u = v; v = x; }
	return collectCells(t, q, p, cell, u, v, w);
}

GCMB_HD bool cellContains(const Tri& t, int cell, V3 q, double (&l)[4], int& err) {
	const int* v = t.cellV + 4 * (long long) cell;
	if (!barycentric(t.point(v[0]), t.point(v[1]), t.point(v[2]), t.point(v[3]), q, l)) { err = 1; return false; }
	return l[0] >= -TOL && l[1] >= -TOL && l[2] >= -TOL && l[3] >= -TOL;   // tetrahedronContains
}

GCMB_HD Found borderFacet(const Tri& t, int prev, int exitSlot, V3 start, V3 query, int& err) {
	int face[3], n = 0;
	for (int i = 0; i < 4; i++) { if (i != exitSlot) { face[n++] = t.cellV[4 * (long long) prev + i]; } }
	Found f = nothing();
	const V3 p[3] = {t.point(face[0]), t.point(face[1]), t.point(face[2])};
	V3 x;
	if (!lineFlatIntersection(p[0], p[1], p[2], start, query, x)) { err = 1; return f; }
	if (triangleContains(p[0], p[1], p[2], x, TOL, TOL, err)) {
		f.n = 3;
		for (int i = 0; i < 3; i++) { f.p[i] = t.localOf[face[i]]; }
		return f;
	}
	for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++) {
		if (segmentContains(p[i], p[j], x, TOL, TOL, err)) { f.n = 2; f.p[0] = t.localOf[face[i]]; f.p[1] = t.localOf[face[j]]; return f; }
	}
	for (int i = 0; i < 3; i++) {
		if (segmentContains(start, query, p[i], TOL, TOL, err)) { f.n = 1; f.p[0] = t.localOf[face[i]]; return f; }
	}
	return f;
}

GCMB_HD Found checkWalk(const Tri& t, bool inner, Walked w, V3 start, V3 query, int& err) {
	if (w.count == 0) { return nothing(); }
	double l[4];
	// a walk of one cell ends in its start cell, which belongs to the body
	if ((w.count == 1 || t.isLocal(w.last)) && cellContains(t, w.last, query, l, err)) { return wholeCell(t, w.last, l); }
	if (w.count == 1) { if (inner) { err = 1; } return nothing(); }
	if (cellContains(t, w.prev, query, l, err)) { return wholeCell(t, w.prev, l); }
	if (!inner) { return nothing(); }
	if (!t.isLocal(w.last)) { return borderFacet(t, w.prev, w.exitSlot, start, query, err); }
	return nothing();
}

// SimplexGrid::findCellCrossedByTheRay
GCMB_HD Found locate(const Tri& t, int lv, V3 shift, int& err) {
	const int gv = t.globalOf[lv];
	const bool inner = t.state[lv] == 0;
	const V3 start = t.point(gv);
	const V3 query = start + shift;
	Found f = checkWalk(t, inner, walkFromVertex(t, lv, gv, query, err), start, query, err);
	if (f.n > 0) { return f; }
	int startCell = crossedIncidentCell(t, lv, query, 0, err).x;
	if (startCell == -2) { startCell = crossedIncidentCell(t, lv, query, TOL, err).x; }
	if (startCell == -2) { startCell = t.locABC[t.locOff[lv]].x; }
	const int* cv = t.cellV + 4 * (long long) startCell;
	const V3 center = (t.point(cv[0]) + t.point(cv[1]) + t.point(cv[2]) + t.point(cv[3])) / 4;
	const double w = 1e-3;
	const V3 startPoint = center * w + start * (1 - w);
	f = checkWalk(t, inner, walkFromCell(t, startCell, startPoint, query, err), start, query, err);
	if (f.n > 0) { return f; }
	if (inner) { err = 1; }
	return nothing();
}

// border state and normals of a local vertex (SimplexGrid.hpp:385-444); which: 0 border, 1 common
GCMB_HD int borderState(const Tri& t, int lv) {
	const int g = t.globalOf[lv];
	bool empty = false, other = false, multi = false;
	int otherId = 0;
	for (int i = t.incOff[g]; i < t.incOff[g + 1]; i++) {
		const int c = t.incCell[i];
		for (int k = 0; k < 4; k++) { if (t.cellN[4 * (long long) c + k] < 0 && t.cellV[4 * (long long) c + k] != g) { empty = true; } }
		const int id = t.cellGrid[c];
		if (id == t.gridId) { continue; }
		if (id == EMPTY) { empty = true; continue; }
		if (!other) { other = true; otherId = id; } else if (id != otherId) { multi = true; }
	}
	if (!empty && !other) { return 0; }
	if (multi || (empty && other)) { return 3; }
	return empty ? 1 : 2;
}
// which: 0 faces towards empty space, 1 towards anything that is not this body, 2 towards the body `neighbor`
GCMB_HD bool vertexNormal(const Tri& t, int lv, int which, V3& out, int neighbor = 0) {
	const int g = t.globalOf[lv];
	V3 sum = {0, 0, 0};
	int count = 0;
	for (int i = t.incOff[g]; i < t.incOff[g + 1]; i++) {
		const int cell = t.incCell[i];
		if (!t.isLocal(cell)) { continue; }
		for (int k = 0; k < 4; k++) {
			const int oc = t.cellN[4 * (long long) cell + k];
			const int og = oc < 0 ? EMPTY : t.cellGrid[oc];
			if (og == t.gridId || (which == 0 && og != EMPTY) || (which == 2 && og != neighbor)) { continue; }
			if (t.cellV[4 * (long long) cell + k] == g) { continue; }
			const V3 opposite = t.point(t.cellV[4 * (long long) cell + k]);
			const V3 a = t.point(t.cellV[4 * (long long) cell + (k + 1) % 4]);
			const V3 b = t.point(t.cellV[4 * (long long) cell + (k + 2) % 4]);
			const V3 c = t.point(t.cellV[4 * (long long) cell + (k + 3) % 4]);
			V3 n = cross(a - b, c - b);
			n = n / length(n);
			if (!(dot(n, a - opposite) > 0)) { n = n * -1.0; }
			sum = sum + n;
			count++;
		}
	}
	if (!count) { out = {0, 0, 0}; return false; }
	out = sum / length(sum);
	return true;
}

// ---------------------------------------------------------------------------------------------
// stage pieces
// ---------------------------------------------------------------------------------------------
template<int M>
GCMB_HD void matVec(const double* A, const double* x, double* y) {
	for (int i = 0; i < M; i++) {
		double r = A[i * M] * x[0];
		for (int n = 1; n < M; n++) { r += A[i * M + n] * x[n]; }
		y[i] = r;
	}
}

// Differentiation::estimateGradient for one vertex: values [nLocal][M] -> grad [3][M] of that vertex.
// One pass over the neighbours: every sum of the normal equations (A^T W A and A^T W b for all M right-hand
// sides) is accumulated neighbour by neighbour in the reference's order, so nothing is staged in local memory.
template<int M>
GCMB_HD void gradientThread(const Tri& t, const double* values, int it, double4* vg /* [M] of this vertex: {value, gradient} */, int& err) {
	const int n = t.nbOff[it + 1] - t.nbOff[it];
	const int* nb = t.nbIdx + t.nbOff[it];
	const V3 x0 = t.localPoint(it);
	double v0[M], b[M][3], N[3][3];
	for (int c = 0; c < M; c++) { v0[c] = values[(long long) it * M + c]; }
	for (int k = 0; k < n; k++) {
		const V3 d = t.localPoint(nb[k]) - x0;
		const double w = 1.0 / length(d);
		const double A[3] = {d.x, d.y, d.z};
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
			const double term = A[i] * (w * A[j]);
			N[i][j] = k == 0 ? term : N[i][j] + term;
		}
		const double* vk = values + (long long) nb[k] * M;
		for (int c = 0; c < M; c++) {
			const double bw = (vk[c] - v0[c]) * w;
			for (int i = 0; i < 3; i++) {
				const double term = bw * A[i];
				b[c][i] = k == 0 ? term : b[c][i] + term;
			}
		}
	}
	if (n == 0) {
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { N[i][j] = 0; }
		for (int c = 0; c < M; c++) for (int i = 0; i < 3; i++) { b[c][i] = 0; }
	}
	const double det = det3(N[0][0], N[0][1], N[0][2], N[1][0], N[1][1], N[1][2], N[2][0], N[2][1], N[2][2]);
	if (det == 0) { err = 1; }
	for (int c = 0; c < M; c++) {
		double g[3] = {0, 0, 0};
		if (det != 0) {
			g[0] = det3(b[c][0], N[0][1], N[0][2], b[c][1], N[1][1], N[1][2], b[c][2], N[2][1], N[2][2]) / det;
			g[1] = det3(N[0][0], b[c][0], N[0][2], N[1][0], b[c][1], N[1][2], N[2][0], b[c][2], N[2][2]) / det;
			g[2] = det3(N[0][0], N[0][1], b[c][0], N[1][0], N[1][1], b[c][1], N[2][0], N[2][1], b[c][2]) / det;
		}
		// value and gradient of a component side by side: one 32-byte sector per (vertex, component) for the
		// interpolation's gathers
		vg[c] = make_double4(v0[c], g[0], g[1], g[2]);
	}
}

GCMB_HD bool isInterpolation(const double (&l)[4]) { return l[0] > -TOL && l[1] > -TOL && l[2] > -TOL && l[3] > -TOL; }

// TetrahedronInterpolator::hybridInterpolate, split so that characteristics with the same foot share the
// geometry: the barycentric coordinates and q - c_i depend on the cell and the point only
struct HybridGeom {
	double l[4];
	V3 d[4];       // q - c_i
};
GCMB_HD void hybridGeometry(const Tri& t, const int (&cell)[4], V3 q, HybridGeom& h, int& err) {
	V3 c[4];
	for (int i = 0; i < 4; i++) { c[i] = t.localPoint(cell[i]); h.d[i] = q - c[i]; }
	if (!barycentric(c[0], c[1], c[2], c[3], q, h.l) || !isInterpolation(h.l)) { err = 1; }
}
// component k; grad is [nLocal][3][M]
template<int M>
GCMB_HD double hybridValue(const HybridGeom& h, const double4* vg, const int (&cell)[4], int k) {
	V3 g[4];
	double v[4];
	for (int i = 0; i < 4; i++) {
		const double4 x = vg[(long long) cell[i] * M + k];
		v[i] = x.x;
		g[i] = {x.y, x.z, x.w};
	}
	const double (&l)[4] = h.l;
	double quadratic = l[0] * (v[0] + dot(g[0], h.d[0]) / 2.0);
	for (int i = 1; i < 4; i++) { quadratic = quadratic + l[i] * (v[i] + dot(g[i], h.d[i]) / 2.0); }
	const double lo = fmin(fmin(v[0], v[1]), fmin(v[2], v[3]));
	const double hi = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
	const double limited = fmin(fmax(quadratic, lo), hi);
	if (quadratic == limited) { return quadratic; }
	return l[0] * v[0] + l[1] * v[1] + l[2] * v[2] + l[3] * v[3];
}
template<int M>
GCMB_HD double hybridInterpolate(const Tri& t, const double4* vg, const int (&cell)[4], int k, V3 q, int& err) {
	HybridGeom h;
	hybridGeometry(t, cell, q, h, err);
	return hybridValue<M>(h, vg, cell, k);
}

// TetrahedronInterpolator::interpolateInOwner over the 15 tetrahedra of 6 points, split the same way: which
// tetrahedron owns q and with what weights depends on the points only
struct OwnerGeom {
	int p[4];
	double l[4];
	bool found;
};
GCMB_HD void ownerGeometry(const V3 (&c)[6], V3 q, OwnerGeom& o, int& err) {
	const int T[15][4] = {{0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 2, 5}, {0, 1, 3, 4}, {0, 1, 3, 5}, {0, 1, 4, 5}, {0, 2, 3, 4},
			{0, 2, 3, 5}, {0, 2, 4, 5}, {0, 3, 4, 5}, {1, 2, 3, 4}, {1, 2, 3, 5}, {1, 2, 4, 5}, {1, 3, 4, 5}, {2, 3, 4, 5}};
	o.found = false;
	for (int i = 0; i < 15; i++) {
		const int* p = T[i];
		if (volume(c[p[0]], c[p[1]], c[p[2]], c[p[3]]) != 0) {
			double l[4];
			if (!barycentric(c[p[0]], c[p[1]], c[p[2]], c[p[3]], q, l)) { err = 1; continue; }
			if (isInterpolation(l)) {
				for (int j = 0; j < 4; j++) { o.p[j] = p[j]; o.l[j] = l[j]; }
				o.found = true;
				return;
			}
		}
	}
	err = 1;
}
GCMB_HD double ownerValue(const OwnerGeom& o, const double (&v)[6]) {
	if (!o.found) { return 0; }
	return o.l[0] * v[o.p[0]] + o.l[1] * v[o.p[1]] + o.l[2] * v[o.p[2]] + o.l[3] * v[o.p[3]];
}

// interpolateInSpaceTime (engine/simplex/common.hpp:70-108): where the ray leaves through the border facet
// r[0..2], in the prism between the current layer (w = 0) and the next one (w = 1)
GCMB_HD bool spaceTimeGeometry(V3 shift, V3 r0, const V3 (&r)[3], OwnerGeom& o, int& err) {
	V3 rc;
	o.found = false;
	if (!lineFlatIntersection(r[0], r[1], r[2], r0, r0 + shift, rc)) { err = 1; return false; }
	double w[2];
	if (!leastSquares2(r[1] - r[0], r[2] - r[0], rc - r[0], w)) { err = 1; return false; }
	const V3 c[6] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}};
	const V3 q = {w[0], w[1], 1 - length(rc - r0) / length(shift)};
	ownerGeometry(c, q, o, err);
	return true;
}

// arguments of the stage kernels
struct StageS {
	Tri t;
	int model;              // 0 elastic (M = 9, 3 outer waves), 1 acoustic (M = 4, 1 outer wave)
	int s;                  // stage
	double tau;
	const double* U;        // [M*M] of this stage
	const double* U1;
	const double* L;        // [M]
	int nFeet;              // distinct NON-ZERO eigenvalues of this stage, in order of first appearance
	double footLambda[9];
	unsigned footMask[9];   // characteristics sharing the eigenvalue
	// feet of the inner vertices found in an earlier step with the same time step, basis and eigenvalues (the
	// mesh does not move, so the cell location of a foot is the same in every step): 0 off, 1 fill, 2 use
	int cacheMode;
	int4* cacheCell;        // [foot][inner index] local vertex ids of the cell, x = -1: not a plain cell hit
	double* cacheLam;       // [foot][inner index][4] barycentric coordinates of the foot
	// BorderCalcMode::LOCAL_BASIS: border and contact vertices carry eigen-systems of their own, written in a basis
	// whose first axis is their normal (engine/simplex/DefaultMesh.hpp:245-266); null = every vertex uses U/U1/dir
	const int* slotOf;          // [nLocal] index of the vertex' tables or -1
	const double* nodeTables;   // [slot][3 stages][U 81 | U1 81 | L 9]
	const double* nodeDirs;     // [slot][3 stages][3] calculation direction of the vertex at each stage
	int pdeMode;            // GcmType::ADVECT_PDE_VECTORS: `riem` is the current layer itself, `next` receives U*V rows
	unsigned zeroMask;      // characteristics with a zero eigenvalue: the invariant is carried over
	int footMajor;          // thread mapping of the inner pass: 1 = a warp shares the foot, 0 = adjacent lanes share the vertex
	double dir[3];          // calculation direction = column s of the basis
	const double* cur;      // PDE vectors [nLocal][M]
	double* riem;           // Riemann invariants of the current layer
	double4* vg;            // [nLocal][M] {invariant, its gradient} (the values repeat riem)
	double* next;           // next layer (invariants until the last kernel)
	unsigned* waves;        // outer invariants of border vertices (bit k)
	int* errors;            // counter of "the reference would have thrown"
};

// eigen-system and calculation direction of a vertex at the stage in flight
GCMB_HD const double* vertexU(const StageS& a, int it) {
	const int slot = a.slotOf ? a.slotOf[it] : -1;
	return slot < 0 ? a.U : a.nodeTables + ((long long) slot * 3 + a.s) * 171;
}
GCMB_HD const double* vertexU1(const StageS& a, int it) { return vertexU(a, it) + 81; }
GCMB_HD V3 vertexDirection(const StageS& a, int it) {
	const int slot = a.slotOf ? a.slotOf[it] : -1;
	if (slot < 0) { return {a.dir[0], a.dir[1], a.dir[2]}; }
	const double* d = a.nodeDirs + ((long long) slot * 3 + a.s) * 3;
	return {d[0], d[1], d[2]};
}

// SimplexGrid::findCellCrossedByTheRay through the cache of feet: slot < 0 = no caching for this call
GCMB_HD Found locateFoot(const StageS& a, int it, V3 shift, long long slot, int& err) {
	// cache entry: x >= 0: a cell (x, y, z, w) with its barycentrics; x <= -10: the other outcomes, n = -(x + 10)
	// vertices in (y, z, w); x = -1: not cached (the location raised an error: recomputed, and counted, every time)
	if (slot >= 0 && a.cacheMode == 2) {
		const int4 c = a.cacheCell[slot];
		if (c.x >= 0) {
			Found f;
			f.n = 4;
			f.p[0] = c.x; f.p[1] = c.y; f.p[2] = c.z; f.p[3] = c.w;
			for (int i = 0; i < 4; i++) { f.l[i] = a.cacheLam[4 * slot + i]; }
			return f;
		}
		if (c.x <= -10) {
			Found f = nothing();
			f.n = -(c.x + 10);
			f.p[0] = c.y; f.p[1] = c.z; f.p[2] = c.w;
			return f;
		}
	}
	const Found f = locate(a.t, it, shift, err);
	if (slot >= 0 && a.cacheMode == 1) {
		int4 c = make_int4(-1, -1, -1, -1);
		if (!err) { c = f.n == 4 ? make_int4(f.p[0], f.p[1], f.p[2], f.p[3]) : make_int4(-10 - f.n, f.p[0], f.p[1], f.p[2]); }
		a.cacheCell[slot] = c;
		for (int i = 0; i < 4; i++) { a.cacheLam[4 * slot + i] = f.l[i]; }
	}
	return f;
}

// Cell location of ONE foot of a vertex, stored in the cache of feet and nothing else: the first of the two passes of a
// stage whose feet are not cached yet (a new calculation basis every step, the reference's default).  The pass that
// follows interpolates from the stored cells (cacheMode 2).  Split like this, the location runs without the
// interpolation's registers (the fused kernel spills 0.9-1.2 KB per thread at 128 registers), and the results are the
// same bits: the stored cell and barycentric coordinates are exactly what the fused pass would have computed.
// A location that raised an error is stored as "not cached" and recomputed -- and counted -- by the second pass.
GCMB_HD void locateFootOnly(const StageS& a, int it, int foot, long long slot) {
	const double dx = -a.tau * a.footLambda[foot];
	if (dx == 0) { return; }
	int err = 0;
	locateFoot(a, it, vertexDirection(a, it) * dx, slot, err);
}

// interpolateValuesAround (…InRiemannInvariants.hpp:146-198) for the characteristics `same` of vertex `it`
// that share the eigenvalue lambda, hence the foot x0 - tau*lambda*direction: the cell location and the
// interpolation geometry are computed once per distinct foot (the reference recomputes them identically for
// each characteristic).  Writes out[j] for every j in `same`; returns the characteristics that turned out outer.
template<int M>
GCMB_HD unsigned footCharacteristics(const StageS& a, int it, double lambda, unsigned same, bool borderPass, double* out, int& err,
                                     long long slot = -1) {
	const Tri& t = a.t;
	const double dx = -a.tau * lambda;
	if (dx == 0) {
		for (int j = 0; j < M; j++) { if ((same >> j) & 1u) { out[j] = a.riem[(long long) it * M + j]; } }
		return 0;
	}
	const V3 x0 = t.localPoint(it);
	const V3 shift = vertexDirection(a, it) * dx;
	const Found f = locateFoot(a, it, shift, slot, err);
	if (f.n == 4) {
		const int cell[4] = {f.p[0], f.p[1], f.p[2], f.p[3]};
		// hybridGeometry(t, cell, x0 + shift, h, err) with the coordinates the cell location already has
		HybridGeom h;
		const V3 q = x0 + shift;
		for (int i = 0; i < 4; i++) { h.l[i] = f.l[i]; h.d[i] = q - t.localPoint(cell[i]); }
		if (!isInterpolation(h.l)) { err = 1; }
		for (int j = 0; j < M; j++) { if ((same >> j) & 1u) { out[j] = hybridValue<M>(h, a.vg, cell, j); } }
		return 0;
	}
	if (f.n == 3 && !borderPass) {
		const V3 r[3] = {t.localPoint(f.p[0]), t.localPoint(f.p[1]), t.localPoint(f.p[2])};
		OwnerGeom o;
		spaceTimeGeometry(shift, x0, r, o, err);
		for (int j = 0; j < M; j++) {
			if (!((same >> j) & 1u)) { continue; }
			double v[6];
			for (int i = 0; i < 3; i++) {
				v[i] = a.riem[(long long) f.p[i] * M + j];
				v[3 + i] = a.next[(long long) f.p[i] * M + j];
			}
			out[j] = ownerValue(o, v);
		}
		return 0;
	}
	if (f.n == 2 && !borderPass) { err = 1; }   // THROW_UNSUPPORTED in 3-D
	for (int j = 0; j < M; j++) { if ((same >> j) & 1u) { out[j] = 0; } }
	return (f.n == 0 || (borderPass && f.n >= 2)) ? same : 0u;
}

// The same for GcmType::ADVECT_PDE_VECTORS (…InPdeVectors.hpp:103-169 + localGcmStep): the whole PDE vector is
// interpolated at the foot (TetrahedronInterpolator<PdeVector>::hybridInterpolate: linear as soon as ANY
// component of the quadratic interpolant leaves the [min, max] of the four values), and the characteristics k
// sharing the foot get r_k = sum_j U(k,j) * v_j -- row k of diagonalMultiply(U, V) (linal/functions.hpp:254-267).
template<int M>
GCMB_HD unsigned footVectors(const StageS& a, int it, double lambda, unsigned same, bool borderPass, double* out, int& err,
                             long long slot = -1) {
	const Tri& t = a.t;
	const double dx = -a.tau * lambda;
	double v[M];
	unsigned outers = 0;
	for (int j = 0; j < M; j++) { v[j] = 0; }
	if (dx == 0) {
		for (int j = 0; j < M; j++) { v[j] = a.riem[(long long) it * M + j]; }
	} else {
		const V3 x0 = t.localPoint(it);
		const V3 shift = vertexDirection(a, it) * dx;
		const Found f = locateFoot(a, it, shift, slot, err);
		if (f.n == 4) {
			const int cell[4] = {f.p[0], f.p[1], f.p[2], f.p[3]};
			HybridGeom h;
			const V3 q = x0 + shift;
			for (int i = 0; i < 4; i++) { h.l[i] = f.l[i]; h.d[i] = q - t.localPoint(cell[i]); }
			if (!isInterpolation(h.l)) { err = 1; }
			bool quadraticHolds = true;
			for (int j = 0; j < M; j++) {
				double w[4];
				V3 g[4];
				for (int i = 0; i < 4; i++) {
					const double4 x = a.vg[(long long) cell[i] * M + j];
					w[i] = x.x;
					g[i] = {x.y, x.z, x.w};
				}
				double quadratic = h.l[0] * (w[0] + dot(g[0], h.d[0]) / 2.0);
				for (int i = 1; i < 4; i++) { quadratic = quadratic + h.l[i] * (w[i] + dot(g[i], h.d[i]) / 2.0); }
				const double lo = fmin(fmin(w[0], w[1]), fmin(w[2], w[3]));
				const double hi = fmax(fmax(w[0], w[1]), fmax(w[2], w[3]));
				if (!(quadratic == fmin(fmax(quadratic, lo), hi))) { quadraticHolds = false; }
				v[j] = quadratic;
			}
			if (!quadraticHolds) {
				for (int j = 0; j < M; j++) {
					double w[4];
					for (int i = 0; i < 4; i++) { w[i] = a.riem[(long long) cell[i] * M + j]; }
					v[j] = h.l[0] * w[0] + h.l[1] * w[1] + h.l[2] * w[2] + h.l[3] * w[3];
				}
			}
		} else if (f.n == 3 && !borderPass) {
			const V3 r[3] = {t.localPoint(f.p[0]), t.localPoint(f.p[1]), t.localPoint(f.p[2])};
			OwnerGeom o;
			spaceTimeGeometry(shift, x0, r, o, err);
			for (int j = 0; j < M; j++) {
				double w[6];
				for (int i = 0; i < 3; i++) {
					w[i] = a.riem[(long long) f.p[i] * M + j];
					w[3 + i] = a.next[(long long) f.p[i] * M + j];
				}
				v[j] = ownerValue(o, w);
			}
		} else {
			if (f.n == 2 && !borderPass) { err = 1; }
			if (f.n == 0 || (borderPass && f.n >= 2)) { outers = same; }
		}
	}
	for (int k = 0; k < M; k++) {
		if (!((same >> k) & 1u)) { continue; }
		const double* U = vertexU(a, it);
		double r = U[k * M] * v[0];
		for (int j = 1; j < M; j++) { r += U[k * M + j] * v[j]; }
		out[k] = r;
	}
	return outers;
}

template<int M>
GCMB_HD unsigned footAny(const StageS& a, int it, double lambda, unsigned same, bool borderPass, double* out, int& err,
                         long long slot = -1) {
	return a.pdeMode ? footVectors<M>(a, it, lambda, same, borderPass, out, err, slot)
	                 : footCharacteristics<M>(a, it, lambda, same, borderPass, out, err, slot);
}

GCMB_HD void countError(int* errors) {
#ifdef __CUDA_ARCH__
	atomicAdd(errors, 1);
#else
	(*errors)++;
#endif
}

// border and contact vertices (…InRiemannInvariants.hpp:59-96), one thread per (vertex, distinct foot): the
// invariants go straight to the next layer, the outer ones are collected in waves[it] (zeroed before the pass)
template<int M>
GCMB_HD void borderFootThread(const StageS& a, int it, int foot, long long slot) {
	int err = 0;
	double* out = a.next + (long long) it * M;
	if (foot == 0) { footAny<M>(a, it, 0.0, a.zeroMask, true, out, err); }
	const unsigned outers = footAny<M>(a, it, a.footLambda[foot], a.footMask[foot], true, out, err, slot);
	if (outers) {
#ifdef __CUDA_ARCH__
		atomicOr(a.waves + it, outers);
#else
		a.waves[it] |= outers;
#endif
	}
	if (err) { countError(a.errors); }
}

// ... then, per vertex, the bookkeeping of its outer invariants (…InRiemannInvariants.hpp:73-85): a vertex with
// only part of a family outer gets the whole family marked and zeroed
template<int M>
GCMB_HD void borderFinishThread(const StageS& a, int it) {
	if (a.pdeMode) {
		// …InPdeVectors.hpp:52-72: the outer invariants stay as found; u = U1 * (rows of U*V)
		double r[M], u[M];
		for (int k = 0; k < M; k++) { r[k] = a.next[(long long) it * M + k]; }
		matVec<M>(vertexU1(a, it), r, u);
		for (int k = 0; k < M; k++) { a.next[(long long) it * M + k] = u[k]; }
		return;
	}
	const unsigned LEFT = a.model == 0 ? 0x15u : 0x1u, RIGHT = a.model == 0 ? 0x2au : 0x2u;
	unsigned outers = a.waves[it];
	if (outers != RIGHT && outers != LEFT && outers != (LEFT | RIGHT) && outers != 0) {
		if (outers & RIGHT) { outers |= RIGHT; }
		if (outers & LEFT) { outers |= LEFT; }
		for (int k = 0; k < M; k++) { if ((outers >> k) & 1u) { a.next[(long long) it * M + k] = 0; } }
		a.waves[it] = outers;
	}
}

// inner vertices (…InRiemannInvariants.hpp:99-113): one thread per (vertex, distinct foot)
template<int M>
GCMB_HD void innerFootThread(const StageS& a, int it, int foot, long long slot) {
	int err = 0;
	if (foot == 0) { footAny<M>(a, it, 0.0, a.zeroMask, false, a.next + (long long) it * M, err); }
	footAny<M>(a, it, a.footLambda[foot], a.footMask[foot], false, a.next + (long long) it * M, err, slot);
	if (err) { countError(a.errors); }
}

// ---- border conditions -----------------------------------------------------------------------
GCMB_HD int symIndex3(int i, int j) { if (i > j) { const int x = i; i = j; j = x; } return i * 3 - ((i - 1) * i) / 2 + j - i; }

GCMB_HD void localBasis(V3 n, double (&S)[3][3]) {
	V3 a = {n.y, -n.x, 0};
	if (n.x == 0 && n.y == 0) { a = {n.z, 0, 0}; }
	const V3 t1 = (a * length(n)) / length(a);
	const V3 t2 = cross(n, t1);
	S[0][0] = t1.x; S[1][0] = t1.y; S[2][0] = t1.z;
	S[0][1] = t2.x; S[1][1] = t2.y; S[2][1] = t2.z;
	S[0][2] = n.x; S[1][2] = n.y; S[2][2] = n.z;
}

template<int M>
GCMB_HD void borderMatrix(int model, int type, V3 p, double* B) {
	const int outer = model == 0 ? 3 : 1;
	for (int i = 0; i < outer * M; i++) { B[i] = 0; }
	if (model == 1) {
		if (type == 0) { B[3] = 1; } else { B[0] = p.x; B[1] = p.y; B[2] = p.z; }
		return;
	}
	double S[3][3];
	localBasis(p, S);
	for (int k = 0; k < 3; k++) {
		if (type == 0) {
			double G[6] = {0, 0, 0, 0, 0, 0};
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { G[symIndex3(i, j)] += S[i][k] * p[j]; }
			for (int i = 0; i < 6; i++) { B[k * M + 3 + i] = G[i]; }
		} else {
			for (int i = 0; i < 3; i++) { B[k * M + i] = S[i][k]; }
		}
	}
}

GCMB_HD void mul33(const double (&A)[3][3], const double (&B)[3][3], double (&C)[3][3]) {
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		double r = A[i][0] * B[0][j]; r += A[i][1] * B[1][j]; r += A[i][2] * B[2][j];
		C[i][j] = r;
	}
}

GCMB_HD void plainBorder(int model, int type, V3 normal, const double* value, double* u) {
	double S[3][3], St[3][3];
	localBasis(normal, S);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { St[i][j] = S[j][i]; }
	if (model == 1) {
		if (type == 0) { u[3] = value[0]; return; }
		double vl[3], vg[3];
		for (int i = 0; i < 3; i++) { double r = St[i][0] * u[0]; r += St[i][1] * u[1]; r += St[i][2] * u[2]; vl[i] = r; }
		vl[2] = value[0];
		for (int i = 0; i < 3; i++) { double r = St[0][i] * vl[0]; r += St[1][i] * vl[1]; r += St[2][i] * vl[2]; vg[i] = r; }
		for (int i = 0; i < 3; i++) { u[i] = vg[i]; }
		return;
	}
	if (type == 1) {
		for (int i = 0; i < 3; i++) { double r = S[i][0] * value[0]; r += S[i][1] * value[1]; r += S[i][2] * value[2]; u[i] = r; }
		return;
	}
	double sg[3][3], t1[3][3], sl[3][3], t2[3][3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { sg[i][j] = u[3 + symIndex3(i, j)]; }
	mul33(St, sg, t1);
	mul33(t1, S, sl);
	for (int i = 0; i < 3; i++) { sl[i][2] = value[i]; }
	for (int j = 0; j < 3; j++) { sl[2][j] = value[j]; }
	mul33(S, sl, t2);
	mul33(t2, St, sg);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { u[3 + symIndex3(i, j)] = sg[i][j]; }
}

// calculateOuterWaveCorrection (common.hpp:187-207); Omega [M][outer], B [outer][M]
template<int M>
GCMB_HD bool outerWaveCorrection(int outer, const double* u, const double* Omega, const double* B, const double* b,
                                 double minDet, double* value, double& detFabs) {
	double X[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	for (int i = 0; i < outer; i++) for (int j = 0; j < outer; j++) {
		double r = B[i * M] * Omega[j];
		for (int n = 1; n < M; n++) { r += B[i * M + n] * Omega[n * outer + j]; }
		X[i][j] = r;
	}
	const double det = outer == 1 ? X[0][0] : det3(X[0][0], X[0][1], X[0][2], X[1][0], X[1][1], X[1][2], X[2][0], X[2][1], X[2][2]);
	detFabs = fabs(det);
	for (int i = 0; i < M; i++) { value[i] = 0; }
	if (!(detFabs > minDet)) { return false; }
	double rhs[3] = {0, 0, 0}, alpha[3] = {0, 0, 0};
	for (int i = 0; i < outer; i++) {
		double r = B[i * M] * u[0];
		for (int n = 1; n < M; n++) { r += B[i * M + n] * u[n]; }
		rhs[i] = b[i] - r;
	}
	if (outer == 1) { if (X[0][0] == 0) { return false; } alpha[0] = rhs[0] / X[0][0]; }
	else if (!solve3(X, rhs, alpha)) { return false; }
	for (int i = 0; i < M; i++) {
		double r = Omega[i * outer] * alpha[0];
		for (int n = 1; n < outer; n++) { r += Omega[i * outer + n] * alpha[n]; }
		value[i] = r;
	}
	return true;
}

template<int M>
GCMB_HD void outerColumns(int outer, const double* U1, unsigned mask, double* Omega) {
	int c = 0;
	for (int i = 0; i < M * outer; i++) { Omega[i] = 0; }
	for (int k = 0; k < M; k++) {
		if (!((mask >> k) & 1u)) { continue; }
		if (c < outer) { for (int i = 0; i < M; i++) { Omega[i * outer + c] = U1[i * M + k]; } }
		c++;
	}
}

struct BorderS {
	int local;              // applyInLocalBasis: the vertex' own stage-0 tables, right invariants, no validity threshold
	const int* slotOf;
	const double* nodeTables;
	int* errors;
	int pdeMode;            // the next layer already holds PDE variables (BorderCorrectorInPdeVectors)
	int model, type;        // condition type 0 FIXED_FORCE, 1 FIXED_VELOCITY
	const double* U;        // this stage
	const double* U1;
	double dir[3];          // calculation direction (the aligned case of getMaximalPossibleDeterminant)
	double b[3];            // border values at t + tau
	int n;                  // nodes of this condition
	const int* node;        // local vertex ids
	const double* normal;   // [n][3]
	const unsigned* waves;
	double* next;           // invariants of the next layer
};

// BorderCorrectorInRiemannInvariants::applyInGlobalBasis for one border node
// BorderCorrectorInRiemannInvariants / InPdeVectors ::applyInLocalBasis (BorderCorrector.hpp:101-120, 226-239)
template<int M>
GCMB_HD void borderCorrectLocalThread(const BorderS& a, int i) {
	const int outer = a.model == 0 ? 3 : 1;
	const unsigned RIGHT = a.model == 0 ? 0x2au : 0x2u;
	const int node = a.node[i];
	const int slot = a.slotOf[node];
	const double* U = slot < 0 ? a.U : a.nodeTables + (long long) slot * 3 * 171;   // stage 0
	const double* U1 = U + 81;
	const V3 normal = {a.normal[3 * i], a.normal[3 * i + 1], a.normal[3 * i + 2]};
	double Omega[M * 3], B[3 * M], u[M], w[M], value[M], det;
	outerColumns<M>(outer, U1, RIGHT, Omega);
	borderMatrix<M>(a.model, a.type, normal, B);
	if (a.pdeMode) { for (int k = 0; k < M; k++) { u[k] = a.next[(long long) node * M + k]; } }
	else { matVec<M>(U1, a.next + (long long) node * M, u); }
	if (!outerWaveCorrection<M>(outer, u, Omega, B, a.b, 0, value, det)) { countError(a.errors); }   // assert_true(isSuccessful)
	for (int k = 0; k < M; k++) { u[k] += value[k]; }
	if (a.pdeMode) { for (int k = 0; k < M; k++) { w[k] = u[k]; } }
	else { matVec<M>(U, u, w); }
	for (int k = 0; k < M; k++) { a.next[(long long) node * M + k] = w[k]; }
}

template<int M>
GCMB_HD void borderCorrectThread(const BorderS& a, int i) {
	if (a.local) { borderCorrectLocalThread<M>(a, i); return; }
	const int outer = a.model == 0 ? 3 : 1;
	const unsigned LEFT = a.model == 0 ? 0x15u : 0x1u, RIGHT = a.model == 0 ? 0x2au : 0x2u;
	double Omega[M * 3], B[3 * M], zero[M], value[M], det;
	outerColumns<M>(outer, a.U1, RIGHT, Omega);
	borderMatrix<M>(a.model, a.type, V3{a.dir[0], a.dir[1], a.dir[2]}, B);
	for (int k = 0; k < M; k++) { zero[k] = 0; }
	outerWaveCorrection<M>(outer, zero, Omega, B, a.b, 0, value, det);
	const double minDet = 1e-3 * det;
	const int node = a.node[i];
	const V3 normal = {a.normal[3 * i], a.normal[3 * i + 1], a.normal[3 * i + 2]};
	double u[M], w[M];
	if (a.pdeMode) { for (int k = 0; k < M; k++) { u[k] = a.next[(long long) node * M + k]; } }
	else { matVec<M>(a.U1, a.next + (long long) node * M, u); }
	borderMatrix<M>(a.model, a.type, normal, B);
	const unsigned outers = a.waves[node];
	if (outers == RIGHT || outers == LEFT) {
		outerColumns<M>(outer, a.U1, outers, Omega);
		if (outerWaveCorrection<M>(outer, u, Omega, B, a.b, minDet, value, det)) { for (int k = 0; k < M; k++) { u[k] += value[k]; } }
		else { plainBorder(a.model, a.type, normal, a.b, u); }
	} else {
		double vr[M], vl[M], d2;
		outerColumns<M>(outer, a.U1, RIGHT, Omega);
		const bool okr = outerWaveCorrection<M>(outer, u, Omega, B, a.b, minDet, vr, det);
		outerColumns<M>(outer, a.U1, LEFT, Omega);
		const bool okl = outerWaveCorrection<M>(outer, u, Omega, B, a.b, minDet, vl, d2);
		if (okr && okl) { for (int k = 0; k < M; k++) { u[k] += (vr[k] + vl[k]) / 2; } }
		else { plainBorder(a.model, a.type, normal, a.b, u); }
	}
	if (a.pdeMode) { for (int k = 0; k < M; k++) { w[k] = u[k]; } }
	else { matVec<M>(a.U, u, w); }
	for (int k = 0; k < M; k++) { a.next[(long long) node * M + k] = w[k]; }
}

// ---- contact of two bodies --------------------------------------------------------------------
// C[R][C] = A[R][N] * B[N][C] in the reference's accumulation order (linal/operators.hpp:109-123)
template<int R, int N, int C>
GCMB_HD void matMul(const double* A, const double* B, double* Cm) {
	for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) {
		double x = A[i * N] * B[j];
		for (int k = 1; k < N; k++) { x += A[i * N + k] * B[k * C + j]; }
		Cm[i * C + j] = x;
	}
}

// LU with partial pivoting (the reference calls GSL for N > 3: util/math/GslUtils.hpp:70-150)
template<int N>
GCMB_HD double luDecompose(double* a, int* perm) {
	double sign = 1;
	for (int i = 0; i < N; i++) { perm[i] = i; }
	for (int j = 0; j < N - 1; j++) {
		double best = fabs(a[j * N + j]);
		int piv = j;
		for (int i = j + 1; i < N; i++) { if (fabs(a[i * N + j]) > best) { best = fabs(a[i * N + j]); piv = i; } }
		if (piv != j) {
			for (int k = 0; k < N; k++) { const double x = a[j * N + k]; a[j * N + k] = a[piv * N + k]; a[piv * N + k] = x; }
			const int x = perm[j]; perm[j] = perm[piv]; perm[piv] = x;
			sign = -sign;
		}
		const double ajj = a[j * N + j];
		if (ajj != 0) {
			for (int i = j + 1; i < N; i++) {
				const double aij = a[i * N + j] / ajj;
				a[i * N + j] = aij;
				for (int k = j + 1; k < N; k++) { a[i * N + k] = a[i * N + k] - aij * a[j * N + k]; }
			}
		}
	}
	double det = sign;
	for (int i = 0; i < N; i++) { det = det * a[i * N + i]; }
	return det;
}

template<int N>
GCMB_HD double detN(const double* m) {
	if constexpr (N == 1) { return m[0]; }
	else if constexpr (N == 2) { return m[0] * m[3] - m[1] * m[2]; }
	else if constexpr (N == 3) { return det3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }
	else {
		double a[N * N];
		int perm[N];
		for (int i = 0; i < N * N; i++) { a[i] = m[i]; }
		return luDecompose<N>(a, perm);
	}
}

// linal::solveLinearSystem for N = 1, 2, 3 (Cramer) and N > 3 (LU); false when the reference would throw
template<int N>
GCMB_HD bool solveN(const double* m, const double* b, double* x) {
	if constexpr (N == 1) { if (m[0] == 0) { return false; } x[0] = b[0] / m[0]; return true; }
	else if constexpr (N == 2) {
		const double det = m[0] * m[3] - m[1] * m[2];
		if (det == 0) { return false; }
		const double d1 = b[0] * m[3] - m[1] * b[1];
		const double d2 = m[0] * b[1] - b[0] * m[2];
		x[0] = d1 / det; x[1] = d2 / det;
		return true;
	}
	else if constexpr (N == 3) {
		const double det = det3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
		if (det == 0) { return false; }
		x[0] = det3(b[0], m[1], m[2], b[1], m[4], m[5], b[2], m[7], m[8]) / det;
		x[1] = det3(m[0], b[0], m[2], m[3], b[1], m[5], m[6], b[2], m[8]) / det;
		x[2] = det3(m[0], m[1], b[0], m[3], m[4], b[1], m[6], m[7], b[2]) / det;
		return true;
	}
	else {
		double a[N * N];
		int perm[N];
		for (int i = 0; i < N * N; i++) { a[i] = m[i]; }
		luDecompose<N>(a, perm);
		for (int i = 0; i < N; i++) { x[i] = b[perm[i]]; }
		for (int i = 0; i < N; i++) { double t = x[i]; for (int j = 0; j < i; j++) { t -= a[i * N + j] * x[j]; } x[i] = t; }
		for (int i = N - 1; i >= 0; i--) { double t = x[i]; for (int j = i + 1; j < N; j++) { t -= a[i * N + j] * x[j]; } x[i] = t / a[i * N + i]; }
		return true;
	}
}

// border form of calculateOuterWaveCorrection with N outer waves (common.hpp:187-207)
template<int M, int N>
GCMB_HD bool outerWaveCorrectionN(const double* u, const double* Omega, const double* B, const double* b, double minDet,
                                  double* value, int& err) {
	double X[N * N], rhs[N], alpha[N], Bu[N];
	matMul<N, M, N>(B, Omega, X);
	const double detFabs = fabs(detN<N>(X));
	for (int i = 0; i < M; i++) { value[i] = 0; }
	if (!(detFabs > minDet)) { return false; }
	matMul<N, M, 1>(B, u, Bu);
	for (int i = 0; i < N; i++) { rhs[i] = b[i] - Bu[i]; }
	if (!solveN<N>(X, rhs, alpha)) { err = 1; return false; }
	matMul<M, N, 1>(Omega, alpha, value);
	return true;
}

// linal::invert for 1x1 and 3x3 (linal/functions.hpp:101-134)
template<int N>
GCMB_HD void invertN(const double* m, double* r) {
	if constexpr (N == 1) { r[0] = 1.0 / m[0]; }
	else {
		const double det = det3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
		const double adj[9] = {
			m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[4] * m[2],
			m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
			m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]};
		for (int i = 0; i < 9; i++) { r[i] = adj[i] / det; }
	}
}

// contact form of calculateOuterWaveCorrection (common.hpp:209-260); O outer waves per body
template<int M, int O>
GCMB_HD bool contactWaveCorrection(const double* uA, const double* OmA, const double* B1A, const double* B2A,
                                   const double* uB, const double* OmB, const double* B1B, const double* B2B,
                                   double min1, double min2, double* valueA, double* valueB, double& det1, double& det2, int& err) {
	double R1[O * O], R[O * O], t1[O], t2[O], d[O], p[O], BO[O * O], Q[O * O], B2Bo[O * O], B2Ao[O * O], BQ[O * O], A[O * O],
			f[O], Bp[O], alphaB[O], alphaA[O], Qa[O];
	det1 = det2 = 0;
	for (int i = 0; i < M; i++) { valueA[i] = 0; valueB[i] = 0; }
	matMul<O, M, O>(B1A, OmA, R1);
	det1 = fabs(detN<O>(R1));
	if (!(det1 > min1)) { return false; }
	invertN<O>(R1, R);
	matMul<O, M, 1>(B1B, uB, t1);
	matMul<O, M, 1>(B1A, uA, t2);
	for (int i = 0; i < O; i++) { d[i] = t1[i] - t2[i]; }
	matMul<O, O, 1>(R, d, p);
	matMul<O, M, O>(B1B, OmB, BO);
	matMul<O, O, O>(R, BO, Q);
	matMul<O, M, O>(B2B, OmB, B2Bo);
	matMul<O, M, O>(B2A, OmA, B2Ao);
	matMul<O, O, O>(B2Ao, Q, BQ);
	for (int i = 0; i < O * O; i++) { A[i] = B2Bo[i] - BQ[i]; }
	matMul<O, O, 1>(B2Ao, p, Bp);
	matMul<O, M, 1>(B2A, uA, t1);
	matMul<O, M, 1>(B2B, uB, t2);
	for (int i = 0; i < O; i++) { f[i] = (Bp[i] + t1[i]) - t2[i]; }
	det2 = fabs(detN<O>(A));
	if (!(det2 > min2)) { return false; }
	if (!solveN<O>(A, f, alphaB)) { err = 1; return false; }
	matMul<O, O, 1>(Q, alphaB, Qa);
	for (int i = 0; i < O; i++) { alphaA[i] = p[i] + Qa[i]; }
	matMul<M, O, 1>(OmA, alphaA, valueA);
	matMul<M, O, 1>(OmB, alphaB, valueB);
	return true;
}

// contact matrices B1 (which = 1) and B2 (which = 2): elastic ADHESION in the global basis
// (ElasticModel.hpp:156-189), acoustic SLIDE (AcousticModel.hpp:95-117); ContactCorrector.hpp:443-481
template<int M>
GCMB_HD void contactMatrix(int model, int which, V3 n, double* B) {
	constexpr int outer = M == 9 ? 3 : 1;   // elastic: 3 outer waves, acoustic: 1
	for (int i = 0; i < outer * M; i++) { B[i] = 0; }
	if (model == 1) {
		if (which == 1) { B[0] = n.x; B[1] = n.y; B[2] = n.z; } else { B[3] = 1; }
		return;
	}
	for (int i = 0; i < 3; i++) {
		if (which == 1) { B[i * M + i] = 1; }
		else { for (int j = 0; j < 3; j++) { B[i * M + 3 + symIndex3(i, j)] = n[j]; } }
	}
}

// Model::applyPlainContactCorrectionAsAverage (average) / applyPlainContactCorrection (A takes B's values)
// (ElasticModel.hpp:243-298, AcousticModel.hpp:158-210)
GCMB_HD void plainContact(int model, bool average, V3 normal, double* uA, double* uB) {
	double S[3][3], St[3][3];
	localBasis(normal, S);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { St[i][j] = S[j][i]; }
	if (model == 1) {
		if (average) { const double p = (uA[3] + uB[3]) / 2; uA[3] = p; uB[3] = p; } else { uA[3] = uB[3]; }
		double la[3], lb[3], g[3];
		for (int i = 0; i < 3; i++) { double r = St[i][0] * uA[0]; r += St[i][1] * uA[1]; r += St[i][2] * uA[2]; la[i] = r; }
		for (int i = 0; i < 3; i++) { double r = St[i][0] * uB[0]; r += St[i][1] * uB[1]; r += St[i][2] * uB[2]; lb[i] = r; }
		if (average) { const double vn = (la[2] + lb[2]) / 2; la[2] = vn; lb[2] = vn; } else { la[2] = lb[2]; }
		for (int i = 0; i < 3; i++) { double r = St[0][i] * la[0]; r += St[1][i] * la[1]; r += St[2][i] * la[2]; g[i] = r; }
		for (int i = 0; i < 3; i++) { uA[i] = g[i]; }
		if (average) {
			for (int i = 0; i < 3; i++) { double r = St[0][i] * lb[0]; r += St[1][i] * lb[1]; r += St[2][i] * lb[2]; g[i] = r; }
			for (int i = 0; i < 3; i++) { uB[i] = g[i]; }
		}
		return;
	}
	if (average) { for (int i = 0; i < 3; i++) { const double v = (uA[i] + uB[i]) / 2; uA[i] = v; uB[i] = v; } }
	else { for (int i = 0; i < 3; i++) { uA[i] = uB[i]; } }
	double ga[3][3], gb[3][3], t[3][3], la[3][3], lb[3][3], sn[3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { ga[i][j] = uA[3 + symIndex3(i, j)]; gb[i][j] = uB[3 + symIndex3(i, j)]; }
	mul33(St, ga, t); mul33(t, S, la);
	mul33(St, gb, t); mul33(t, S, lb);
	for (int i = 0; i < 3; i++) { sn[i] = average ? (la[i][2] + lb[i][2]) / 2 : lb[i][2]; }
	for (int i = 0; i < 3; i++) { la[i][2] = sn[i]; }
	for (int j = 0; j < 3; j++) { la[2][j] = sn[j]; }
	mul33(S, la, t); mul33(t, St, ga);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { uA[3 + symIndex3(i, j)] = ga[i][j]; }
	if (average) {
		for (int i = 0; i < 3; i++) { lb[i][2] = sn[i]; }
		for (int j = 0; j < 3; j++) { lb[2][j] = sn[j]; }
		mul33(S, lb, t); mul33(t, St, gb);
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { uB[3 + symIndex3(i, j)] = gb[i][j]; }
	}
}

// [M][2*O]: RIGHT columns then LEFT columns of U1 (ContactCorrector.hpp:186-190)
template<int M, int O>
GCMB_HD void outerColumnsBoth(const double* U1, unsigned RIGHT, unsigned LEFT, double* Omega) {
	double part[M * O];
	outerColumns<M>(O, U1, RIGHT, part);
	for (int i = 0; i < M; i++) for (int c = 0; c < O; c++) { Omega[i * 2 * O + c] = part[i * O + c]; }
	outerColumns<M>(O, U1, LEFT, part);
	for (int i = 0; i < M; i++) for (int c = 0; c < O; c++) { Omega[i * 2 * O + O + c] = part[i * O + c]; }
}

GCMB_HD int popcountU(unsigned x) { int c = 0; while (x) { c += (int) (x & 1u); x >>= 1; } return c; }

struct ContactS {
	int local;                           // applyInLocalBasis: the vertices' own stage-0 tables, right invariants
	const int *slotOfA, *slotOfB;
	const double *nodeTablesA, *nodeTablesB;
	int pdeMode;                         // ContactCorrectorInPdeVectors: no matching, no conversions
	int model, n;
	const double *UA, *U1A, *UB, *U1B;   // this stage, body A and body B
	double dir[3];
	const int *nodeA, *nodeB;            // local vertex ids of the pairs
	const double* normal;                // [n][3], from A to B
	unsigned *wavesA, *wavesB;
	double *nextA, *nextB;               // invariants of the next layers
	int* errors;
};

// ContactCorrectorInRiemannInvariants::applyInGlobalBasis for one pair of nodes (ContactCorrector.hpp:334-410
// around :133-253)
// ContactCorrectorInRiemannInvariants / InPdeVectors ::applyInLocalBasis (ContactCorrector.hpp:104-131, 320-332)
template<int M, int O>
GCMB_HD void contactCorrectLocalThread(const ContactS& a, int i) {
	const unsigned RIGHT = a.model == 0 ? 0x2au : 0x2u;
	int err = 0;
	const int slotA = a.slotOfA[a.nodeA[i]], slotB = a.slotOfB[a.nodeB[i]];
	const double* UA = slotA < 0 ? a.UA : a.nodeTablesA + (long long) slotA * 3 * 171;
	const double* UB = slotB < 0 ? a.UB : a.nodeTablesB + (long long) slotB * 3 * 171;
	double OmA[M * O], OmB[M * O], B1A[O * M], B1B[O * M], B2A[O * M], B2B[O * M], uA[M], uB[M], vA[M], vB[M], w[M], det1, det2;
	outerColumns<M>(O, UA + 81, RIGHT, OmA);
	outerColumns<M>(O, UB + 81, RIGHT, OmB);
	const V3 normal = {a.normal[3 * i], a.normal[3 * i + 1], a.normal[3 * i + 2]};
	contactMatrix<M>(a.model, 1, normal, B1A); contactMatrix<M>(a.model, 1, normal, B1B);
	contactMatrix<M>(a.model, 2, normal, B2A); contactMatrix<M>(a.model, 2, normal, B2B);
	double* ra = a.nextA + (long long) a.nodeA[i] * M;
	double* rb = a.nextB + (long long) a.nodeB[i] * M;
	if (a.pdeMode) { for (int k = 0; k < M; k++) { uA[k] = ra[k]; uB[k] = rb[k]; } }
	else { matVec<M>(UA + 81, ra, uA); matVec<M>(UB + 81, rb, uB); }
	if (!contactWaveCorrection<M, O>(uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, 0, 0, vA, vB, det1, det2, err)) { err = 1; }
	for (int k = 0; k < M; k++) { uA[k] += vA[k]; uB[k] += vB[k]; }
	if (a.pdeMode) { for (int k = 0; k < M; k++) { ra[k] = uA[k]; rb[k] = uB[k]; } }
	else {
		matVec<M>(UA, uA, w);
		for (int k = 0; k < M; k++) { ra[k] = w[k]; }
		matVec<M>(UB, uB, w);
		for (int k = 0; k < M; k++) { rb[k] = w[k]; }
	}
	if (err) { countError(a.errors); }
}

template<int M, int O>
GCMB_HD void contactCorrectThread(const ContactS& a, int i) {
	if (a.local) { contactCorrectLocalThread<M, O>(a, i); return; }
	const unsigned LEFT = a.model == 0 ? 0x15u : 0x1u, RIGHT = a.model == 0 ? 0x2au : 0x2u;
	int err = 0;
	double OmA[M * 2 * O], OmB[M * O], B1A[O * M], B1B[O * M], B2A[O * M], B2B[O * M], zero[M], vA[M], vB[M], det1, det2;
	const V3 direction = {a.dir[0], a.dir[1], a.dir[2]};
	// getMaximalPossibleDeterminants (:273-300)
	for (int k = 0; k < M; k++) { zero[k] = 0; }
	outerColumns<M>(O, a.U1A, LEFT, OmA);
	outerColumns<M>(O, a.U1B, RIGHT, OmB);
	contactMatrix<M>(a.model, 1, direction, B1A); contactMatrix<M>(a.model, 1, direction, B1B);
	contactMatrix<M>(a.model, 2, direction, B2A); contactMatrix<M>(a.model, 2, direction, B2B);
	if (!contactWaveCorrection<M, O>(zero, OmA, B1A, B2A, zero, OmB, B1B, B2B, 0, 0, vA, vB, det1, det2, err)) { err = 1; }
	const double min1 = 1e-3 * det1, min2 = 1e-3 * det2;
	double* ra = a.nextA + (long long) a.nodeA[i] * M;
	double* rb = a.nextB + (long long) a.nodeB[i] * M;
	unsigned wa = a.wavesA[a.nodeA[i]], wb = a.wavesB[a.nodeB[i]];
	// matchInnersAndOuters (:381-410)
	const int N = (popcountU(wa) + popcountU(wb)) / O;
	if (!a.pdeMode && N % 2 != 0) {
		if (N == 3) { wa = wb = LEFT | RIGHT; }
		else if (wa == 0) { if (wb == LEFT) { wa = RIGHT; } else { if (wb != RIGHT) { err = 1; } wa = LEFT; } }
		else { if (wb != 0) { err = 1; } if (wa == LEFT) { wb = RIGHT; } else { if (wa != RIGHT) { err = 1; } wb = LEFT; } }
		for (int k = 0; k < M; k++) { if ((wa >> k) & 1u) { ra[k] = 0; } if ((wb >> k) & 1u) { rb[k] = 0; } }
	}
	a.wavesA[a.nodeA[i]] = wa; a.wavesB[a.nodeB[i]] = wb;
	double uA[M], uB[M], w[M];
	if (a.pdeMode) { for (int k = 0; k < M; k++) { uA[k] = ra[k]; uB[k] = rb[k]; } }
	else { matVec<M>(a.U1A, ra, uA); matVec<M>(a.U1B, rb, uB); }
	const V3 normal = {a.normal[3 * i], a.normal[3 * i + 1], a.normal[3 * i + 2]};
	contactMatrix<M>(a.model, 1, normal, B1A); contactMatrix<M>(a.model, 1, normal, B1B);
	contactMatrix<M>(a.model, 2, normal, B2A); contactMatrix<M>(a.model, 2, normal, B2B);
	const int na = popcountU(wa), nb = popcountU(wb);
	if (na == O && nb == O) {
		outerColumns<M>(O, a.U1A, wa, OmA);
		outerColumns<M>(O, a.U1B, wb, OmB);
		if (contactWaveCorrection<M, O>(uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA, vB, det1, det2, err)) {
			for (int k = 0; k < M; k++) { uA[k] += vA[k]; uB[k] += vB[k]; }
		} else { plainContact(a.model, true, normal, uA, uB); }
	} else if ((na == 2 * O && nb == 0) || (nb == 2 * O && na == 0)) {
		// the node with both families outer is a border with two conditions taken from the other node
		const bool first = na == 2 * O;
		double* uX = first ? uA : uB;
		double* uY = first ? uB : uA;
		const double* U1X = first ? a.U1A : a.U1B;
		const double *B1X = first ? B1A : B1B, *B2X = first ? B2A : B2B, *B1Y = first ? B1B : B1A, *B2Y = first ? B2B : B2A;
		double Bc[2 * O * M], b12[2 * O], value[M];
		for (int k = 0; k < O * M; k++) { Bc[k] = B1X[k]; Bc[O * M + k] = B2X[k]; }
		matMul<O, M, 1>(B1Y, uY, b12);
		matMul<O, M, 1>(B2Y, uY, b12 + O);
		outerColumnsBoth<M, O>(U1X, RIGHT, LEFT, OmA);
		if (outerWaveCorrectionN<M, 2 * O>(uX, OmA, Bc, b12, min1, value, err)) { for (int k = 0; k < M; k++) { uX[k] += value[k]; } }
		else { plainContact(a.model, false, normal, uX, uY); }
	} else {
		double vA2[M], vB2[M], d1, d2;
		outerColumns<M>(O, a.U1A, RIGHT, OmA);
		outerColumns<M>(O, a.U1B, LEFT, OmB);
		const bool ok1 = contactWaveCorrection<M, O>(uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA, vB, d1, d2, err);
		outerColumns<M>(O, a.U1A, LEFT, OmA);
		outerColumns<M>(O, a.U1B, RIGHT, OmB);
		const bool ok2 = contactWaveCorrection<M, O>(uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA2, vB2, d1, d2, err);
		if (ok1 && ok2) { for (int k = 0; k < M; k++) { uA[k] += (vA[k] + vA2[k]) / 2; uB[k] += (vB[k] + vB2[k]) / 2; } }
		else { plainContact(a.model, true, normal, uA, uB); }
	}
	if (a.pdeMode) { for (int k = 0; k < M; k++) { ra[k] = uA[k]; rb[k] = uB[k]; } }
	else {
		matVec<M>(a.UA, uA, w);
		for (int k = 0; k < M; k++) { ra[k] = w[k]; }
		matVec<M>(a.UB, uB, w);
		for (int k = 0; k < M; k++) { rb[k] = w[k]; }
	}
	if (err) {
#ifdef __CUDA_ARCH__
		atomicAdd(a.errors, 1);
#else
		(*a.errors)++;
#endif
	}
}

}  // namespace sx
}  // namespace gcmb
