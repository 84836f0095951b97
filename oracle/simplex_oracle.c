/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's simplex (tetrahedral) GCM path.
 * PARITY UNPINNED (see simplex_oracle.h): the reference cannot be built here (CGAL) and pins no value itself.
 * Compiled with -ffp-contract=off: every operation rounds separately, like the reference build.
 * Paths below are relative to /root/reference/src/libgcm.
 */
#include "simplex_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EQUALITY_TOLERANCE 1e-9 /* util/infrastructure/Types.hpp:10 */

/* ------------------------------------------------------------------------------------------ */
/* linal                                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double v[3]; } R3;

static R3 r3(const double* p) { R3 a = {{p[0], p[1], p[2]}}; return a; }
static R3 sub(R3 a, R3 b) { R3 c = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return c; }
static R3 add(R3 a, R3 b) { R3 c = {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; return c; }
static R3 scale(R3 a, double x) { R3 c = {{a.v[0] * x, a.v[1] * x, a.v[2] * x}}; return c; }
static R3 divide(R3 a, double x) { R3 c = {{a.v[0] / x, a.v[1] / x, a.v[2] / x}}; return c; }
/* linal/functions.hpp:327-334 */
static double dot(R3 a, R3 b) { double r = a.v[0] * b.v[0]; r += a.v[1] * b.v[1]; r += a.v[2] * b.v[2]; return r; }
static double len(R3 a) { return sqrt(dot(a, a)); }
/* linal/geometry.hpp:13-17 */
static R3 cross(R3 a, R3 b) {
	R3 c = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
	return c;
}
/* linal/determinants.hpp:18-59 */
static double det2(double m11, double m12, double m21, double m22) { return m11 * m22 - m12 * m21; }
static double det3(double m11, double m12, double m13, double m21, double m22, double m23,
                   double m31, double m32, double m33) {
	return m11 * (m22 * m33 - m23 * m32) - m12 * (m21 * m33 - m23 * m31) + m13 * (m21 * m32 - m22 * m31);
}
/* linal/linearSystems.hpp:104-129 (Cramer); returns 1 when the reference throws (zero determinant) */
static int solve3(const double A[3][3], const double b[3], double x[3]) {
	const double det = det3(A[0][0], A[0][1], A[0][2], A[1][0], A[1][1], A[1][2], A[2][0], A[2][1], A[2][2]);
	if (det == 0) { return 1; }
	const double d1 = det3(b[0], A[0][1], A[0][2], b[1], A[1][1], A[1][2], b[2], A[2][1], A[2][2]);
	const double d2 = det3(A[0][0], b[0], A[0][2], A[1][0], b[1], A[1][2], A[2][0], b[2], A[2][2]);
	const double d3 = det3(A[0][0], A[0][1], b[0], A[1][0], A[1][1], b[1], A[2][0], A[2][1], b[2]);
	x[0] = d1 / det; x[1] = d2 / det; x[2] = d3 / det;
	return 0;
}
/* linal/linearSystems.hpp:78-96 */
static int solve2(const double A[2][2], const double b[2], double x[2]) {
	const double det = det2(A[0][0], A[0][1], A[1][0], A[1][1]);
	if (det == 0) { return 1; }
	const double d1 = det2(b[0], A[0][1], b[1], A[1][1]);
	const double d2 = det2(A[0][0], b[0], A[1][0], b[1]);
	x[0] = d1 / det; x[1] = d2 / det;
	return 0;
}
/* linearLeastSquares with identity weights for a 3x2 matrix with columns p, q (linearSystems.hpp:150-158) */
static int lls32(R3 p, R3 q, R3 rhs, double x[2]) {
	const R3 col[2] = {p, q};
	double A[2][2], b[2];
	for (int i = 0; i < 2; i++) {
		for (int j = 0; j < 2; j++) {
			double r = col[i].v[0] * (1.0 * col[j].v[0]);
			for (int n = 1; n < 3; n++) { r += col[i].v[n] * (1.0 * col[j].v[n]); }
			A[i][j] = r;
		}
		double r = col[i].v[0] * (1.0 * rhs.v[0]);
		for (int n = 1; n < 3; n++) { r += col[i].v[n] * (1.0 * rhs.v[n]); }
		b[i] = r;
	}
	return solve2(A, b, x);
}
/* same for a 3x1 matrix (column p): 1x1 system (linearSystems.hpp:46-60) */
static int lls31(R3 p, R3 rhs, double* x) {
	double a = p.v[0] * (1.0 * p.v[0]);
	double b = p.v[0] * (1.0 * rhs.v[0]);
	for (int n = 1; n < 3; n++) { a += p.v[n] * (1.0 * p.v[n]); b += p.v[n] * (1.0 * rhs.v[n]); }
	if (a == 0) { return 1; }
	*x = b / a;
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* linal/geometry.hpp                                                                          */
/* ------------------------------------------------------------------------------------------ */
/* :248-255 */
static double oriented_volume(R3 a, R3 b, R3 c, R3 d) {
	const R3 ba = sub(b, a), ca = sub(c, a), da = sub(d, a);
	return det3(ba.v[0], ba.v[1], ba.v[2], ca.v[0], ca.v[1], ca.v[2], da.v[0], da.v[1], da.v[2]) / 6;
}
double gcmo_oriented_volume(const double a[3], const double b[3], const double c[3], const double d[3]) {
	return oriented_volume(r3(a), r3(b), r3(c), r3(d));
}
/* :142-151 */
static int barycentric4(R3 a, R3 b, R3 c, R3 d, R3 q, double l[4]) {
	double T[3][3], rhs[3], x[3];
	for (int i = 0; i < 3; i++) {
		T[i][0] = a.v[i] - d.v[i]; T[i][1] = b.v[i] - d.v[i]; T[i][2] = c.v[i] - d.v[i];
		rhs[i] = q.v[i] - d.v[i];
	}
	if (solve3(T, rhs, x)) { return 1; }
	l[0] = x[0]; l[1] = x[1]; l[2] = x[2]; l[3] = 1 - x[0] - x[1] - x[2];
	return 0;
}
int gcmo_barycentric4(const double a[3], const double b[3], const double c[3], const double d[3],
                      const double q[3], double lambda[4]) {
	return barycentric4(r3(a), r3(b), r3(c), r3(d), r3(q), lambda);
}
/* :118-129 triangle in 3-D */
static int barycentric3(R3 a, R3 b, R3 c, R3 q, double l[3]) {
	double x[2];
	if (lls32(sub(a, c), sub(b, c), sub(q, c), x)) { return 1; }
	l[0] = x[0]; l[1] = x[1]; l[2] = 1 - x[0] - x[1];
	return 0;
}
/* :91-103 segment in 3-D */
static int barycentric2(R3 a, R3 b, R3 q, double l[2]) {
	double x;
	if (lls31(sub(a, b), sub(q, b), &x)) { return 1; }
	l[0] = x; l[1] = 1 - x;
	return 0;
}
/* :236-238, :240-243 */
static double area3(R3 a, R3 b, R3 c) { return len(cross(sub(b, a), sub(c, a))) / 2; }
static double volume(R3 a, R3 b, R3 c, R3 d) { return fabs(oriented_volume(a, b, c, d)); }
/* :266-274 */
static double min_height4(R3 a, R3 b, R3 c, R3 d) {
	const double V = volume(a, b, c, d);
	const double A = area3(b, c, d), B = area3(c, d, a), C = area3(d, a, b), D = area3(a, b, c);
	return 3 * V / fmax(A, fmax(B, fmax(C, D)));
}
/* :256-263 */
static double min_height3(R3 a, R3 b, R3 c) {
	const double S = area3(a, b, c);
	const double ab = len(sub(a, b)), ac = len(sub(a, c)), bc = len(sub(b, c));
	return 2 * S / fmax(ab, fmax(ac, bc));
}
/* :289-296 */
static int is_degenerate4(R3 a, R3 b, R3 c, R3 d, double eps) {
	const double h = min_height4(a, b, c, d);
	const double l = (len(sub(a, b)) + len(sub(a, c)) + len(sub(a, d)) + len(sub(d, b)) + len(sub(d, c)) + len(sub(b, c))) / 6;
	return h <= eps * l;
}
/* :279-285 */
static int is_degenerate3(R3 a, R3 b, R3 c, double eps) {
	const double h = min_height3(a, b, c);
	const double l = (len(sub(a, b)) + len(sub(a, c)) + len(sub(b, c))) / 3;
	return h <= eps * l;
}
/* :326-332; *err set when the reference would throw */
static int segment_contains(R3 a, R3 b, R3 q, double eps, double deg_eps, int* err) {
	if (!is_degenerate3(a, b, q, deg_eps)) { return 0; }
	double l[2];
	if (barycentric2(a, b, q, l)) { *err = 1; return 0; }
	return l[0] >= -eps && l[1] >= -eps;
}
/* :345-351 */
static int triangle_contains(R3 a, R3 b, R3 c, R3 q, double eps, double deg_eps, int* err) {
	if (!is_degenerate4(a, b, c, q, deg_eps)) { return 0; }
	double l[3];
	if (barycentric3(a, b, c, q, l)) { *err = 1; return 0; }
	return l[0] >= -eps && l[1] >= -eps && l[2] >= -eps;
}
/* :356-362 */
static int tetrahedron_contains(R3 a, R3 b, R3 c, R3 d, R3 q, double eps, int* err) {
	double l[4];
	if (barycentric4(a, b, c, d, q, l)) { *err = 1; return 0; }
	return l[0] >= -eps && l[1] >= -eps && l[2] >= -eps && l[3] >= -eps;
}
/* :410-416 */
static int solid_angle_contains(R3 a, R3 b, R3 c, R3 d, R3 q, double eps, int* err) {
	double l[4];
	if (barycentric4(a, b, c, d, q, l)) { *err = 1; return 0; }
	return l[0] <= 1 + eps && l[1] >= -eps && l[2] >= -eps && l[3] >= -eps;
}
/* :201-217 */
static int line_flat_intersection(R3 f1, R3 f2, R3 f3, R3 l1, R3 l2, R3* out) {
	const R3 tau = sub(l2, l1), p = sub(f2, f1), q = sub(f3, f1);
	double A[3][3], b[3], x[3];
	for (int i = 0; i < 3; i++) { A[i][0] = tau.v[i]; A[i][1] = -p.v[i]; A[i][2] = -q.v[i]; b[i] = f1.v[i] - l1.v[i]; }
	if (solve3(A, b, x)) { return 1; }
	*out = add(l1, scale(tau, x[0]));
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* flat triangulation helpers (the static API of grid/simplex/cgal/Cgal3DTriangulation.hpp:251-292) */
/* ------------------------------------------------------------------------------------------ */
static R3 point(const gcmo_tri* t, int v) { return r3(t->xyz + 3 * (size_t) v); }
static int is_local(const gcmo_tri* t, int cell) { return cell >= 0 && t->cell_grid[cell] == t->grid_id; }
static int other_vertex_index(const gcmo_tri* t, int cell, int a, int b, int c) {
	for (int i = 0; i < 4; i++) {
		const int d = t->cell_v[4 * cell + i];
		if (d != a && d != b && d != c) { return i; }
	}
	return -1;
}
static int other_vertex(const gcmo_tri* t, int cell, int a, int b, int c) {
	return t->cell_v[4 * cell + other_vertex_index(t, cell, a, b, c)];
}
static int neighbor_through(const gcmo_tri* t, int cell, int a, int b, int c) {
	return t->cell_n[4 * cell + other_vertex_index(t, cell, a, b, c)];
}
static int cell_contains(const gcmo_tri* t, int cell, R3 q, double eps, int* err) {
	const int* v = t->cell_v + 4 * cell;
	return tetrahedron_contains(point(t, v[0]), point(t, v[1]), point(t, v[2]), point(t, v[3]), q, eps, err);
}
static double orientation(const gcmo_tri* t, int a, int b, int c, R3 d) {
	return oriented_volume(point(t, a), point(t, b), point(t, c), d);
}
static double orientation_vq(const gcmo_tri* t, int a, int b, R3 c, R3 d) {
	return oriented_volume(point(t, a), point(t, b), c, d);
}

/* hull exterior counts as empty space: a vertex on the hull has a "virtual" empty neighbour */
static int on_hull(const gcmo_tri* t, int g) {
	for (int i = t->inc_off[g]; i < t->inc_off[g + 1]; i++) {
		const int c = t->inc_cell[i];
		for (int k = 0; k < 4; k++) {
			if (t->cell_n[4 * c + k] < 0 && t->cell_v[4 * c + k] != g) { return 1; }
		}
	}
	return 0;
}

int gcmo_simplex_border_state(const gcmo_tri* t, int lv) {
	/* CGAL's infinite cells carry EmptySpaceFlag (Cgal3DMesher.hpp:66-80); our flat triangulation has no
	 * infinite cells, so a hull vertex sees one extra "empty" grid */
	const int g = t->global_of[lv];
	int have_empty = on_hull(t, g), have_other = 0, other_id = 0, multi = 0;
	for (int i = t->inc_off[g]; i < t->inc_off[g + 1]; i++) {
		const int id = t->cell_grid[t->inc_cell[i]];
		if (id == t->grid_id) { continue; }
		if (id == GCMO_EMPTY_SPACE) { have_empty = 1; continue; }
		if (!have_other) { have_other = 1; other_id = id; } else if (id != other_id) { multi = 1; }
	}
	if (!have_empty && !have_other) { return 0; }
	if (multi || (have_empty && have_other)) { return 3; }
	return have_empty ? 1 : 2;
}

/* Cgal3DTriangulation.hpp:222-238 */
static int find_crossed_incident_cell(const gcmo_tri* t, int gv, R3 query, double eps, int* err) {
	for (int i = t->inc_off[gv]; i < t->inc_off[gv + 1]; i++) {
		const int cand = t->inc_cell[i];
		if (!is_local(t, cand)) { continue; }
		const int a = other_vertex(t, cand, gv, gv, gv);
		const int b = other_vertex(t, cand, gv, gv, a);
		const int c = other_vertex(t, cand, gv, a, b);
		if (solid_angle_contains(point(t, gv), point(t, a), point(t, b), point(t, c), query, eps, err)) { return cand; }
	}
	return -2; /* NULL */
}

/* Cgal3DTriangulation.hpp:247-260 */
static void find_crossed_inside_out_facet(const gcmo_tri* t, int cell, R3 q, R3 p, double eps,
		int* a, int* b, int* c, int* err) {
	*a = *b = *c = -1;
	for (int i = 0; i < 4; i++) {
		const int a1 = t->cell_v[4 * cell + (i + 1) % 4];
		const int b1 = t->cell_v[4 * cell + (i + 2) % 4];
		const int c1 = t->cell_v[4 * cell + (i + 3) % 4];
		if (solid_angle_contains(q, point(t, a1), point(t, b1), point(t, c1), p, eps, err)) {
			*a = a1; *b = b1; *c = c1;
			break;
		}
	}
}

/* result of a line walk: the reference keeps the whole list, only its size, last and previous are used */
typedef struct { int count, last, prev, exit_slot; } Walk;  /* exit_slot: facet of prev the walk left through */

/* grid/simplex/cgal/LineWalker.hpp:26-52 */
static Walk collect_cells(const gcmo_tri* t, R3 q, R3 p, int cell, int u, int v, int w) {
	Walk ans = {1, cell, -2, -1};
	int guard = 0;
	while (orientation(t, u, v, w, p) < 0) {
		ans.exit_slot = other_vertex_index(t, cell, u, v, w);
		cell = neighbor_through(t, cell, u, v, w);
		ans.prev = ans.last; ans.last = cell; ans.count++;
		if (!is_local(t, cell)) { break; }
		const int s = other_vertex(t, cell, u, v, w);
		if (orientation_vq(t, u, s, q, p) > 0) {
			if (orientation_vq(t, v, s, q, p) > 0) { u = s; } else { w = s; }
		} else {
			if (orientation_vq(t, w, s, q, p) > 0) { v = s; } else { u = s; }
		}
		if (++guard > 100000) { break; }
	}
	return ans;
}

/* LineWalker.hpp:55-69 */
static Walk cells_along_from_vertex(const gcmo_tri* t, int gv, R3 p, int* err) {
	Walk none = {0, -2, -2, -1};
	const int cell = find_crossed_incident_cell(t, gv, p, 0, err);
	if (cell == -2) { return none; }
	int u = other_vertex(t, cell, gv, gv, gv);
	int v = other_vertex(t, cell, gv, gv, u);
	const int w = other_vertex(t, cell, gv, u, v);
	if (orientation(t, u, v, w, point(t, gv)) < 0) { const int x = u; u = v; v = x; }
	return collect_cells(t, point(t, gv), p, cell, u, v, w);
}

/* LineWalker.hpp:71-89 */
static Walk cells_along_from_cell(const gcmo_tri* t, int cell, R3 q, R3 p, int* err) {
	Walk none = {0, -2, -2, -1};
	int u, v, w;
	find_crossed_inside_out_facet(t, cell, q, p, 0, &u, &v, &w, err);
	if (u < 0) { find_crossed_inside_out_facet(t, cell, q, p, EQUALITY_TOLERANCE, &u, &v, &w, err); }
	if (u < 0) { return none; }
	if (orientation(t, u, v, w, q) < 0) { const int x = u; u = v; v = x; }
	return collect_cells(t, q, p, cell, u, v, w);
}

static void set_cell(const gcmo_tri* t, int cell, int out[5]) {
	out[0] = 4;
	for (int i = 0; i < 4; i++) { out[1 + i] = t->local_of[t->cell_v[4 * cell + i]]; }
}
static void set_empty(int out[5]) { out[0] = 0; out[1] = out[2] = out[3] = out[4] = -1; }

/* Cgal3DTriangulation.hpp:183-213 on commonVertices(prev, last) (CgalTriangulation.hpp:89-105): the vertices of
 * prev, in prev's order, shared with the cell behind the facet the walk left through */
static void filter_face(const gcmo_tri* t, int prev, int exit_slot, R3 start, R3 query, double eps, int out[5], int* err) {
	int face[3], n = 0;
	for (int i = 0; i < 4; i++) { if (i != exit_slot) { face[n++] = t->cell_v[4 * prev + i]; } }
	set_empty(out);
	const R3 p[3] = {point(t, face[0]), point(t, face[1]), point(t, face[2])};
	R3 inter;
	if (line_flat_intersection(p[0], p[1], p[2], start, query, &inter)) { *err = 1; return; }
	if (triangle_contains(p[0], p[1], p[2], inter, EQUALITY_TOLERANCE, eps, err)) {
		out[0] = 3;
		for (int i = 0; i < 3; i++) { out[1 + i] = t->local_of[face[i]]; }
		return;
	}
	for (int i = 0; i < 3; i++) {
		for (int j = i + 1; j < 3; j++) {
			if (segment_contains(p[i], p[j], inter, EQUALITY_TOLERANCE, eps, err)) {
				out[0] = 2; out[1] = t->local_of[face[i]]; out[2] = t->local_of[face[j]];
				return;
			}
		}
	}
	for (int i = 0; i < 3; i++) {
		if (segment_contains(start, query, p[i], EQUALITY_TOLERANCE, eps, err)) {
			out[0] = 1; out[1] = t->local_of[face[i]];
			return;
		}
	}
}

/* grid/simplex/SimplexGrid.cpp:114-164 */
static void check_walk(const gcmo_tri* t, int inner, Walk w, R3 start, R3 query, int out[5], int* err) {
	set_empty(out);
	if (w.count == 0) { return; }
	if (is_local(t, w.last) && cell_contains(t, w.last, query, EQUALITY_TOLERANCE, err)) { set_cell(t, w.last, out); return; }
	if (w.count == 1) { if (inner) { *err = 1; } return; }
	if (cell_contains(t, w.prev, query, EQUALITY_TOLERANCE, err)) { set_cell(t, w.prev, out); return; }
	if (!inner) { return; }
	if (!is_local(t, w.last)) { filter_face(t, w.prev, w.exit_slot, start, query, EQUALITY_TOLERANCE, out, err); }
}

int gcmo_simplex_locate(const gcmo_tri* t, int lv, const double shift[3], int out[5]) {
	/* SimplexGrid.cpp:61-112 */
	int err = 0;
	const int gv = t->global_of[lv];
	const int inner = gcmo_simplex_border_state(t, lv) == 0;
	const R3 start = point(t, gv);
	const R3 query = add(start, r3(shift));
	Walk w = cells_along_from_vertex(t, gv, query, &err);
	check_walk(t, inner, w, start, query, out, &err);
	if (out[0] > 0) { return err; }
	int start_cell = find_crossed_incident_cell(t, gv, query, 0, &err);
	if (start_cell == -2) { start_cell = find_crossed_incident_cell(t, gv, query, EQUALITY_TOLERANCE, &err); }
	if (start_cell == -2) {
		for (int i = t->inc_off[gv]; i < t->inc_off[gv + 1]; i++) {
			if (is_local(t, t->inc_cell[i])) { start_cell = t->inc_cell[i]; break; }
		}
	}
	const double wgt = 1e-3;
	const int* cv = t->cell_v + 4 * start_cell;
	/* center(t) = (a + b + c + d) / 4 (Cgal3DTriangulation.hpp:264-270) */
	const R3 center = divide(add(add(add(point(t, cv[0]), point(t, cv[1])), point(t, cv[2])), point(t, cv[3])), 4);
	const R3 start_point = add(scale(center, wgt), scale(start, 1 - wgt));
	w = cells_along_from_cell(t, start_cell, start_point, query, &err);
	check_walk(t, inner, w, start, query, out, &err);
	if (out[0] > 0) { return err; }
	if (inner) { err = 1; }  /* assert_false(isInner(it)) */
	set_empty(out);
	return err;
}

/* SimplexGrid.hpp:427-444 + Cgal3DTriangulation.hpp:103-113 + geometry.hpp:423-427.
 * which: 0 = border normal (faces towards empty space only), 1 = common normal (faces towards anything that is
 * not this grid).  Returns 0 and a zero vector when there is no such face. */
static int normal_impl(const gcmo_tri* t, int lv, int which, int neighbor, double out[3]) {
	const int g = t->global_of[lv];
	R3 sum = {{0, 0, 0}};
	int count = 0;
	for (int i = t->inc_off[g]; i < t->inc_off[g + 1]; i++) {
		const int cell = t->inc_cell[i];
		if (!is_local(t, cell)) { continue; }
		for (int k = 0; k < 4; k++) {
			const int outerc = t->cell_n[4 * cell + k];
			const int outer_grid = outerc < 0 ? GCMO_EMPTY_SPACE : t->cell_grid[outerc];
			if (outer_grid == t->grid_id) { continue; }
			if (which == 0 && outer_grid != GCMO_EMPTY_SPACE) { continue; }
			if (which == 2 && outer_grid != neighbor) { continue; }
			if (t->cell_v[4 * cell + k] == g) { continue; }  /* the shared facet must contain the vertex */
			const R3 opposite = point(t, t->cell_v[4 * cell + k]);
			const R3 a = point(t, t->cell_v[4 * cell + (k + 1) % 4]);
			const R3 b = point(t, t->cell_v[4 * cell + (k + 2) % 4]);
			const R3 c = point(t, t->cell_v[4 * cell + (k + 3) % 4]);
			R3 nrm = cross(sub(a, b), sub(c, b));
			nrm = divide(nrm, len(nrm));
			if (!(dot(nrm, sub(a, opposite)) > 0)) { nrm = scale(nrm, -1.0); }
			sum = add(sum, nrm);
			count++;
		}
	}
	if (!count) { out[0] = out[1] = out[2] = 0; return 0; }
	sum = divide(sum, len(sum));
	out[0] = sum.v[0]; out[1] = sum.v[1]; out[2] = sum.v[2];
	return 1;
}

int gcmo_simplex_normal(const gcmo_tri* t, int lv, int which, double out[3]) { return normal_impl(t, lv, which, 0, out); }

/* SimplexGrid.hpp:141-144: faces towards the body `neighbor` only */
int gcmo_simplex_contact_normal(const gcmo_tri* t, int lv, int neighbor, double out[3]) { return normal_impl(t, lv, 2, neighbor, out); }

/* ------------------------------------------------------------------------------------------ */
/* neighbours, gradient, interpolation                                                         */
/* ------------------------------------------------------------------------------------------ */
static int cmp_int(const void* a, const void* b) { return *(const int*) a - *(const int*) b; }

int gcmo_simplex_neighbors(const gcmo_tri* t, int lv, int* out, int capacity) {
	const int g = t->global_of[lv];
	int tmp[512], n = 0;
	for (int i = t->inc_off[g]; i < t->inc_off[g + 1]; i++) {
		const int c = t->inc_cell[i];
		if (!is_local(t, c)) { continue; }
		for (int k = 0; k < 4; k++) {
			const int l = t->local_of[t->cell_v[4 * c + k]];
			if (l != lv && n < 512) { tmp[n++] = l; }
		}
	}
	qsort(tmp, (size_t) n, sizeof(int), cmp_int);
	int m = 0;
	for (int i = 0; i < n; i++) {
		if (i > 0 && tmp[i] == tmp[i - 1]) { continue; }
		if (m < capacity) { out[m] = tmp[i]; }
		m++;
	}
	return m;
}

int gcmo_simplex_gradient(const gcmo_tri* t, int M, const double* values, double* grad) {
	int errors = 0;
	for (int it = 0; it < t->n_local; it++) {
		int nb[512];
		int n = gcmo_simplex_neighbors(t, it, nb, 512);
		if (n > GCMO_MAX_NEIGHBORS) { n = GCMO_MAX_NEIGHBORS; }
		const R3 x0 = point(t, t->global_of[it]);
		double A[GCMO_MAX_NEIGHBORS][3], W[GCMO_MAX_NEIGHBORS];
		for (int i = 0; i < n; i++) {
			const R3 d = sub(point(t, t->global_of[nb[i]]), x0);
			for (int k = 0; k < 3; k++) { A[i][k] = d.v[k]; }
			W[i] = 1.0 / len(d);
		}
		/* A^T (W A), 3x3 (functions.hpp:220-234; rows beyond n are zeros and add nothing) */
		double N[3][3];
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
			double r = 0;
			for (int k = 0; k < n; k++) { const double term = A[k][i] * (W[k] * A[k][j]); r = k == 0 ? term : r + term; }
			N[i][j] = r;
		}
		const double det = det3(N[0][0], N[0][1], N[0][2], N[1][0], N[1][1], N[1][2], N[2][0], N[2][1], N[2][2]);
		for (int c = 0; c < M; c++) {
			double b[3];
			for (int i = 0; i < 3; i++) {
				double r = 0;
				for (int k = 0; k < n; k++) {
					const double bk = values[(size_t) nb[k] * M + c] - values[(size_t) it * M + c];
					const double term = (bk * W[k]) * A[k][i];
					r = k == 0 ? term : r + term;
				}
				b[i] = r;
			}
			double g[3] = {0, 0, 0};
			if (det == 0) { errors++; }
			else {
				const double d1 = det3(b[0], N[0][1], N[0][2], b[1], N[1][1], N[1][2], b[2], N[2][1], N[2][2]);
				const double d2 = det3(N[0][0], b[0], N[0][2], N[1][0], b[1], N[1][2], N[2][0], b[2], N[2][2]);
				const double d3 = det3(N[0][0], N[0][1], b[0], N[1][0], N[1][1], b[1], N[2][0], N[2][1], b[2]);
				g[0] = d1 / det; g[1] = d2 / det; g[2] = d3 / det;
			}
			for (int d = 0; d < 3; d++) { grad[((size_t) it * 3 + d) * M + c] = g[d]; }
		}
	}
	return errors;
}

static int is_interpolation(const double l[4]) {
	return l[0] > -EQUALITY_TOLERANCE && l[1] > -EQUALITY_TOLERANCE && l[2] > -EQUALITY_TOLERANCE && l[3] > -EQUALITY_TOLERANCE;
}

double gcmo_simplex_hybrid_interpolate(const gcmo_tri* t, int M, const double* values, const double* grad,
		const int cell[4], int k, const double q_[3], int* err) {
	/* util/math/interpolation/TetrahedronInterpolator.hpp:27-104 */
	const R3 q = r3(q_);
	R3 c[4], g[4];
	double v[4], l[4];
	for (int i = 0; i < 4; i++) {
		c[i] = point(t, t->global_of[cell[i]]);
		v[i] = values[(size_t) cell[i] * M + k];
		for (int d = 0; d < 3; d++) { g[i].v[d] = grad[((size_t) cell[i] * 3 + d) * M + k]; }
	}
	if (barycentric4(c[0], c[1], c[2], c[3], q, l) || !is_interpolation(l)) { *err = 1; }
	double quadratic = l[0] * (v[0] + dot(g[0], sub(q, c[0])) / 2.0);
	for (int i = 1; i < 4; i++) { quadratic = quadratic + l[i] * (v[i] + dot(g[i], sub(q, c[i])) / 2.0); }
	const double mn = fmin(fmin(v[0], v[1]), fmin(v[2], v[3]));
	const double mx = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
	const double limited = fmin(fmax(quadratic, mn), mx);
	if (quadratic == limited) { return quadratic; }
	return l[0] * v[0] + l[1] * v[1] + l[2] * v[2] + l[3] * v[3];
}

/* TetrahedronInterpolator.hpp:113-155: linear interpolation in the first of 15 tetrahedra (of 6 points) that
 * contains q */
static double interpolate_in_owner(const R3 c[6], const double v[6], R3 q, int* err) {
	static const int T[15][4] = {{0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 2, 5}, {0, 1, 3, 4}, {0, 1, 3, 5}, {0, 1, 4, 5},
			{0, 2, 3, 4}, {0, 2, 3, 5}, {0, 2, 4, 5}, {0, 3, 4, 5}, {1, 2, 3, 4}, {1, 2, 3, 5}, {1, 2, 4, 5},
			{1, 3, 4, 5}, {2, 3, 4, 5}};
	for (int i = 0; i < 15; i++) {
		const int* p = T[i];
		if (volume(c[p[0]], c[p[1]], c[p[2]], c[p[3]]) != 0) {
			double l[4];
			if (barycentric4(c[p[0]], c[p[1]], c[p[2]], c[p[3]], q, l)) { *err = 1; continue; }
			if (is_interpolation(l)) { return l[0] * v[p[0]] + l[1] * v[p[1]] + l[2] * v[p[2]] + l[3] * v[p[3]]; }
		}
	}
	*err = 1;
	return 0;
}

/* engine/simplex/common.hpp:106-133 */
static double interpolate_space_time(R3 shift, R3 r0, const R3 r[3], const double vcurr[3], const double vnext[3], int* err) {
	R3 rc;
	if (line_flat_intersection(r[0], r[1], r[2], r0, add(r0, shift), &rc)) { *err = 1; return 0; }
	const R3 e1 = sub(r[1], r[0]), e2 = sub(r[2], r[0]), a = sub(rc, r[0]);
	double w[2];
	if (lls32(e1, e2, a, w)) { *err = 1; return 0; }
	const R3 c[6] = {{{0, 0, 0}}, {{1, 0, 0}}, {{0, 1, 0}}, {{0, 0, 1}}, {{1, 0, 1}}, {{0, 1, 1}}};
	const double v[6] = {vcurr[0], vcurr[1], vcurr[2], vnext[0], vnext[1], vnext[2]};
	const R3 q = {{w[0], w[1], 1 - len(sub(rc, r0)) / len(shift)}};
	return interpolate_in_owner(c, v, q, err);
}

/* ------------------------------------------------------------------------------------------ */
/* border conditions                                                                           */
/* ------------------------------------------------------------------------------------------ */
static int sym_index3(int i, int j) { if (i > j) { const int x = i; i = j; j = x; } return i * 3 - ((i - 1) * i) / 2 + j - i; }

/* linal/basis.hpp:58-66 with geometry.hpp:46-52: columns tau1, tau2, n */
static void local_basis3(R3 n, double S[3][3]) {
	R3 a = {{n.v[1], -n.v[0], 0}};
	if (n.v[0] == 0 && n.v[1] == 0) { a.v[0] = n.v[2]; a.v[1] = 0; a.v[2] = 0; }
	const R3 t1 = divide(scale(a, len(n)), len(a));
	const R3 t2 = cross(n, t1);
	for (int i = 0; i < 3; i++) { S[i][0] = t1.v[i]; S[i][1] = t2.v[i]; S[i][2] = n.v[i]; }
}

/* border matrix B [outer][M] (ElasticModel.hpp:111-153, AcousticModel.hpp:95-117) */
static void border_matrix(int model, int M, int type, R3 p, double* B) {
	const int outer = model == 0 ? 3 : 1;
	memset(B, 0, (size_t) (outer * M) * sizeof(double));
	if (model == 1) {
		if (type == 0) { B[3] = 1; } else { for (int i = 0; i < 3; i++) { B[i] = p.v[i]; } }
		return;
	}
	double S[3][3];
	local_basis3(p, S);
	for (int k = 0; k < 3; k++) {
		if (type == 0) {
			/* G(i,j) += S(i,k) * p(j) into a SYMMETRIC matrix: (i,j) and (j,i) alias */
			double G[6] = {0, 0, 0, 0, 0, 0};
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { G[sym_index3(i, j)] += S[i][k] * p.v[j]; }
			for (int i = 0; i < 3; i++) for (int j = 0; j <= i; j++) { B[k * M + 3 + sym_index3(i, j)] = G[sym_index3(i, j)]; }
		} else {
			for (int i = 0; i < 3; i++) { B[k * M + i] = S[i][k]; }
		}
	}
}

/* plain corrections (ElasticModel.hpp:202-232, AcousticModel.hpp:126-147) */
static void plain_border(int model, int M, int type, R3 normal, const double* value, double* u) {
	(void) M;
	if (model == 1) {
		if (type == 0) { u[3] = value[0]; return; }
		/* createLocalBasisTranspose: rows tau1, tau2, n */
		double S[3][3], St[3][3];
		local_basis3(normal, S);
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { St[i][j] = S[j][i]; }
		double vl[3], vg[3];
		for (int i = 0; i < 3; i++) { double r = St[i][0] * u[0]; r += St[i][1] * u[1]; r += St[i][2] * u[2]; vl[i] = r; }
		vl[2] = value[0];
		for (int i = 0; i < 3; i++) { double r = St[0][i] * vl[0]; r += St[1][i] * vl[1]; r += St[2][i] * vl[2]; vg[i] = r; }
		for (int i = 0; i < 3; i++) { u[i] = vg[i]; }
		return;
	}
	double S[3][3];
	local_basis3(normal, S);
	if (type == 1) {
		for (int i = 0; i < 3; i++) { double r = S[i][0] * value[0]; r += S[i][1] * value[1]; r += S[i][2] * value[2]; u[i] = r; }
		return;
	}
	double sg[3][3], St[3][3], t1[3][3], sl[3][3], t2[3][3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { sg[i][j] = u[3 + sym_index3(i, j)]; St[i][j] = S[j][i]; }
#define MM(A, B, C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double r = A[i][0] * B[0][j]; r += A[i][1] * B[1][j]; r += A[i][2] * B[2][j]; C[i][j] = r; }
	MM(St, sg, t1)
	MM(t1, S, sl)
	for (int i = 0; i < 3; i++) { sl[i][2] = value[i]; }
	for (int j = 0; j < 3; j++) { sl[2][j] = value[j]; }
	MM(S, sl, t2)
	MM(t2, St, sg)
#undef MM
	/* setSigmaTo writes (i,j) for all i,j into symmetric storage: the lower triangle wins */
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { u[3 + sym_index3(i, j)] = sg[i][j]; }
}

void gcmo_simplex_plain_border(int model, int M, int n_border, const int* border_node, const double* border_normal,
		const int* border_cond, const int* cond_type, const double* cond_b, double* pde) {
	const int outer = model == 0 ? 3 : 1;
	for (int i = 0; i < n_border; i++) {
		const int c = border_cond[i];
		if (c < 0) { continue; }
		plain_border(model, M, cond_type[c], r3(border_normal + 3 * i), cond_b + (size_t) c * outer, pde + (size_t) border_node[i] * M);
	}
}

/* engine/simplex/common.hpp:187-207 */
static int outer_wave_correction(int M, int outer, const double* u, const double* Omega /*[M][outer]*/,
		const double* B /*[outer][M]*/, const double* b, double min_det, double* value, double* det_fabs) {
	double Mx[3][3];
	for (int i = 0; i < outer; i++) for (int j = 0; j < outer; j++) {
		double r = B[i * M] * Omega[j];
		for (int n = 1; n < M; n++) { r += B[i * M + n] * Omega[n * outer + j]; }
		Mx[i][j] = r;
	}
	const double det = outer == 1 ? Mx[0][0]
			: det3(Mx[0][0], Mx[0][1], Mx[0][2], Mx[1][0], Mx[1][1], Mx[1][2], Mx[2][0], Mx[2][1], Mx[2][2]);
	*det_fabs = fabs(det);
	for (int i = 0; i < M; i++) { value[i] = 0; }
	if (!(*det_fabs > min_det)) { return 0; }
	double rhs[3], alpha[3];
	for (int i = 0; i < outer; i++) {
		double r = B[i * M] * u[0];
		for (int n = 1; n < M; n++) { r += B[i * M + n] * u[n]; }
		rhs[i] = b[i] - r;
	}
	if (outer == 1) { if (Mx[0][0] == 0) { return 0; } alpha[0] = rhs[0] / Mx[0][0]; }
	else if (solve3(Mx, rhs, alpha)) { return 0; }
	for (int i = 0; i < M; i++) {
		double r = Omega[i * outer] * alpha[0];
		for (int n = 1; n < outer; n++) { r += Omega[i * outer + n] * alpha[n]; }
		value[i] = r;
	}
	return 1;
}

static void mat_vec(int M, const double* A, const double* x, double* y) {
	for (int i = 0; i < M; i++) {
		double r = A[i * M] * x[0];
		for (int n = 1; n < M; n++) { r += A[i * M + n] * x[n]; }
		y[i] = r;
	}
}

/* columns `mask` (ascending) of U1 into Omega [M][outer] (common.hpp:153-165) */
static void columns(int M, int outer, const double* U1, unsigned mask, double* Omega) {
	int c = 0;
	memset(Omega, 0, (size_t) (M * outer) * sizeof(double));
	for (int k = 0; k < M; k++) {
		if (!((mask >> k) & 1u)) { continue; }
		if (c < outer) { for (int i = 0; i < M; i++) { Omega[i * outer + c] = U1[i * M + k]; } }
		c++;
	}
}


/* [M][2*outer] = RIGHT columns then LEFT columns (ContactCorrector.hpp:186-190) */
static void columns_right_left(int M, int outer, const double* U1, unsigned RIGHT, unsigned LEFT, double* Omega) {
	double part[9 * 3];
	columns(M, outer, U1, RIGHT, part);
	for (int i = 0; i < M; i++) for (int c = 0; c < outer; c++) { Omega[i * 2 * outer + c] = part[i * outer + c]; }
	columns(M, outer, U1, LEFT, part);
	for (int i = 0; i < M; i++) for (int c = 0; c < outer; c++) { Omega[i * 2 * outer + outer + c] = part[i * outer + c]; }
}

/* C[r][c] = A[r][n] * B[n][c], first product assigned, the rest added (linal/operators.hpp:109-123) */
static void mat_mul(int r, int n, int c, const double* A, const double* B, double* C) {
	for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) {
		double x = A[i * n] * B[j];
		for (int k = 1; k < n; k++) { x += A[i * n + k] * B[k * c + j]; }
		C[i * c + j] = x;
	}
}

/* LU with partial pivoting standing in for GSL (util/math/GslUtils.hpp:70-150: gsl_linalg_LU_decomp/_det/_solve);
 * the reference reaches it for N > 3 only (linal/linearSystems.hpp:25-32, determinants.hpp:62-70) */
static double lu_decompose(int n, double* a, int* perm) {
	double sign = 1;
	for (int i = 0; i < n; i++) { perm[i] = i; }
	for (int j = 0; j < n - 1; j++) {
		double best = fabs(a[j * n + j]);
		int piv = j;
		for (int i = j + 1; i < n; i++) { if (fabs(a[i * n + j]) > best) { best = fabs(a[i * n + j]); piv = i; } }
		if (piv != j) {
			for (int k = 0; k < n; k++) { const double x = a[j * n + k]; a[j * n + k] = a[piv * n + k]; a[piv * n + k] = x; }
			const int x = perm[j]; perm[j] = perm[piv]; perm[piv] = x;
			sign = -sign;
		}
		const double ajj = a[j * n + j];
		if (ajj != 0) {
			for (int i = j + 1; i < n; i++) {
				const double aij = a[i * n + j] / ajj;
				a[i * n + j] = aij;
				for (int k = j + 1; k < n; k++) { a[i * n + k] = a[i * n + k] - aij * a[j * n + k]; }
			}
		}
	}
	double det = sign;
	for (int i = 0; i < n; i++) { det = det * a[i * n + i]; }
	return det;
}
static void lu_solve(int n, const double* lu, const int* perm, const double* b, double* x) {
	for (int i = 0; i < n; i++) { x[i] = b[perm[i]]; }
	for (int i = 0; i < n; i++) { double t = x[i]; for (int j = 0; j < i; j++) { t -= lu[i * n + j] * x[j]; } x[i] = t; }
	for (int i = n - 1; i >= 0; i--) { double t = x[i]; for (int j = i + 1; j < n; j++) { t -= lu[i * n + j] * x[j]; } x[i] = t / lu[i * n + i]; }
}

/* determinant of an n x n matrix the way linal dispatches it (determinants.hpp) */
static double det_n(int n, const double* m) {
	if (n == 1) { return m[0]; }
	if (n == 2) { return m[0] * m[3] - m[1] * m[2]; }
	if (n == 3) { return det3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }
	double a[36];
	int perm[6];
	memcpy(a, m, (size_t) (n * n) * sizeof(double));
	return lu_decompose(n, a, perm);
}
/* linal::solveLinearSystem (linearSystems.hpp:35-129); returns 1 when the reference would throw */
static int solve_n(int n, const double* m, const double* b, double* x) {
	if (n == 1) { if (m[0] == 0) { return 1; } x[0] = b[0] / m[0]; return 0; }
	if (n == 2) {
		const double det = m[0] * m[3] - m[1] * m[2];
		if (det == 0) { return 1; }
		const double d1 = b[0] * m[3] - m[1] * b[1];
		const double d2 = m[0] * b[1] - b[0] * m[2];
		x[0] = d1 / det; x[1] = d2 / det;
		return 0;
	}
	if (n == 3) {
		double A[3][3];
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = m[i * 3 + j]; }
		return solve3(A, b, x);
	}
	double a[36];
	int perm[6];
	memcpy(a, m, (size_t) (n * n) * sizeof(double));
	lu_decompose(n, a, perm);
	lu_solve(n, a, perm, b, x);
	return 0;
}

/* border form of calculateOuterWaveCorrection for any outer count n <= 6 (common.hpp:187-207) */
static int outer_wave_correction_n(int M, int n, const double* u, const double* Omega /*[M][n]*/,
		const double* B /*[n][M]*/, const double* b, double min_det, double* value, double* det_fabs, int* err) {
	double Mx[36], rhs[6], alpha[6], Bu[6];
	mat_mul(n, M, n, B, Omega, Mx);
	*det_fabs = fabs(det_n(n, Mx));
	for (int i = 0; i < M; i++) { value[i] = 0; }
	if (!(*det_fabs > min_det)) { return 0; }
	mat_mul(n, M, 1, B, u, Bu);
	for (int i = 0; i < n; i++) { rhs[i] = b[i] - Bu[i]; }
	if (solve_n(n, Mx, rhs, alpha)) { *err += 1; return 0; }
	mat_mul(M, n, 1, Omega, alpha, value);
	return 1;
}

/* linal::invert for 1x1 and 3x3 (linal/functions.hpp:101-134) */
static void invert_n(int n, const double* m, double* r) {
	if (n == 1) { r[0] = 1.0 / m[0]; return; }
	const double det = det3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
	const double adj[9] = {
		m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[4] * m[2],
		m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
		m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]};
	for (int i = 0; i < 9; i++) { r[i] = adj[i] / det; }
}

/* contact form of calculateOuterWaveCorrection (common.hpp:209-260) */
static int contact_wave_correction(int M, int o, const double* uA, const double* OmA, const double* B1A, const double* B2A,
		const double* uB, const double* OmB, const double* B1B, const double* B2B, double min1, double min2,
		double* valueA, double* valueB, double* det1, double* det2, int* err) {
	double R1[9], R[9], t1[3], t2[3], d[3], p[3], BO[9], Q[9], B2Bo[9], B2Ao[9], BQ[9], A[9], f[3], Bp[3], alphaB[3], alphaA[3], Qa[3];
	*det1 = *det2 = 0;
	for (int i = 0; i < M; i++) { valueA[i] = 0; valueB[i] = 0; }
	mat_mul(o, M, o, B1A, OmA, R1);
	*det1 = fabs(det_n(o, R1));
	if (!(*det1 > min1)) { return 0; }
	invert_n(o, R1, R);
	mat_mul(o, M, 1, B1B, uB, t1);
	mat_mul(o, M, 1, B1A, uA, t2);
	for (int i = 0; i < o; i++) { d[i] = t1[i] - t2[i]; }
	mat_mul(o, o, 1, R, d, p);
	mat_mul(o, M, o, B1B, OmB, BO);
	mat_mul(o, o, o, R, BO, Q);
	mat_mul(o, M, o, B2B, OmB, B2Bo);
	mat_mul(o, M, o, B2A, OmA, B2Ao);
	mat_mul(o, o, o, B2Ao, Q, BQ);
	for (int i = 0; i < o * o; i++) { A[i] = B2Bo[i] - BQ[i]; }
	mat_mul(o, o, 1, B2Ao, p, Bp);
	mat_mul(o, M, 1, B2A, uA, t1);
	mat_mul(o, M, 1, B2B, uB, t2);
	for (int i = 0; i < o; i++) { f[i] = (Bp[i] + t1[i]) - t2[i]; }
	*det2 = fabs(det_n(o, A));
	if (!(*det2 > min2)) { return 0; }
	if (solve_n(o, A, f, alphaB)) { *err += 1; return 0; }
	mat_mul(o, o, 1, Q, alphaB, Qa);
	for (int i = 0; i < o; i++) { alphaA[i] = p[i] + Qa[i]; }
	mat_mul(M, o, 1, OmA, alphaA, valueA);
	mat_mul(M, o, 1, OmB, alphaB, valueB);
	return 1;
}

/* contact matrices: elastic ADHESION = fixed velocity / fixed force in the GLOBAL basis (ElasticModel.hpp:156-189,
 * ContactCorrector.hpp:443-462); acoustic SLIDE = normal velocity / pressure (AcousticModel.hpp:95-117,
 * ContactCorrector.hpp:463-481).  which: 1 -> B1, 2 -> B2 */
static void contact_matrix(int model, int M, int which, R3 normal, double* B) {
	const int outer = model == 0 ? 3 : 1;
	memset(B, 0, (size_t) (outer * M) * sizeof(double));
	if (model == 1) {
		if (which == 1) { for (int i = 0; i < 3; i++) { B[i] = normal.v[i]; } } else { B[3] = 1; }
		return;
	}
	for (int i = 0; i < 3; i++) {
		if (which == 1) { B[i * M + i] = 1; }
		else { for (int j = 0; j < 3; j++) { B[i * M + 3 + sym_index3(i, j)] = normal.v[j]; } }
	}
}

#define MM33(A, B, C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double r = A[i][0] * B[0][j]; r += A[i][1] * B[1][j]; r += A[i][2] * B[2][j]; C[i][j] = r; }
/* Model::applyPlainContactCorrectionAsAverage (average != 0) / applyPlainContactCorrection (A takes B's values)
 * (ElasticModel.hpp:243-298, AcousticModel.hpp:158-210) */
static void plain_contact(int model, int average, R3 normal, double* uA, double* uB) {
	double S[3][3], St[3][3];
	local_basis3(normal, S);
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
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { ga[i][j] = uA[3 + sym_index3(i, j)]; gb[i][j] = uB[3 + sym_index3(i, j)]; }
	MM33(St, ga, t)
	MM33(t, S, la)
	MM33(St, gb, t)
	MM33(t, S, lb)
	for (int i = 0; i < 3; i++) { sn[i] = average ? (la[i][2] + lb[i][2]) / 2 : lb[i][2]; }
	for (int i = 0; i < 3; i++) { la[i][2] = sn[i]; }
	for (int j = 0; j < 3; j++) { la[2][j] = sn[j]; }
	MM33(S, la, t)
	MM33(t, St, ga)
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { uA[3 + sym_index3(i, j)] = ga[i][j]; }
	if (average) {
		for (int i = 0; i < 3; i++) { lb[i][2] = sn[i]; }
		for (int j = 0; j < 3; j++) { lb[2][j] = sn[j]; }
		MM33(S, lb, t)
		MM33(t, St, gb)
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { uB[3 + sym_index3(i, j)] = gb[i][j]; }
	}
}
#undef MM33

/* AbstractContactCorrector::applyPlainCorrection over a list of node pairs (ContactCorrector.hpp:256-269) */
void gcmo_simplex_plain_contact(int model, int M, int n, const int* node_a, const int* node_b, const double* normal,
		double* pde_a, double* pde_b) {
	for (int i = 0; i < n; i++) {
		plain_contact(model, 1, r3(normal + 3 * i), pde_a + (size_t) node_a[i] * M, pde_b + (size_t) node_b[i] * M);
	}
}

/* ------------------------------------------------------------------------------------------ */
/* stage, in the four phases simplex::Engine::gcmStage calls (Engine.cpp:118-141)             */
/* ------------------------------------------------------------------------------------------ */
struct gcmo_sstage {
	const gcmo_tri* t;
	int model, M, s, outer;
	unsigned LEFT, RIGHT;   /* Model.cpp:65-82 */
	double tau;
	const double *U, *U1, *L;
	R3 direction;
	double *riem, *grad, *next;
	unsigned* waves;
	int errors;
	int pde_mode;          /* GcmType::ADVECT_PDE_VECTORS: riem holds the PDE vectors themselves */
};

/* beforeStage (…InRiemannInvariants.hpp:44-56) */
gcmo_sstage* gcmo_sx_begin(const gcmo_tri* t, int model, int M, int s, double tau, const double* U_, const double* U1_,
		const double* L_, const double* basis, const double* cur, double* next, int gcm_type) {
	gcmo_sstage* h = (gcmo_sstage*) calloc(1, sizeof(gcmo_sstage));
	const int n = t->n_local;
	h->t = t; h->model = model; h->M = M; h->s = s; h->tau = tau;
	h->outer = model == 0 ? 3 : 1;
	h->LEFT = model == 0 ? 0x15u : 0x1u;
	h->RIGHT = model == 0 ? 0x2au : 0x2u;
	h->U = U_ + (size_t) s * M * M; h->U1 = U1_ + (size_t) s * M * M; h->L = L_ + (size_t) s * M;
	h->riem = (double*) malloc((size_t) n * M * sizeof(double));
	h->grad = (double*) malloc((size_t) n * 3 * M * sizeof(double));
	h->waves = (unsigned*) calloc((size_t) n, sizeof(unsigned));
	h->next = next;
	h->pde_mode = gcm_type == 1;
	if (h->pde_mode) { memcpy(h->riem, cur, (size_t) n * M * sizeof(double)); }   /* …InPdeVectors.hpp:37-44 */
	else { for (int v = 0; v < n; v++) { mat_vec(M, h->U, cur + (size_t) v * M, h->riem + (size_t) v * M); } }
	h->errors += gcmo_simplex_gradient(t, M, h->riem, h->grad);
	const R3 direction = {{basis[0 * 3 + s], basis[1 * 3 + s], basis[2 * 3 + s]}};
	h->direction = direction;
	return h;
}

/* pass 0: contactAndBorderStage (…InRiemannInvariants.hpp:59-96), pass 1: innerStage (:99-113); both through
 * interpolateValuesAround (:146-198) */
static void nodes_pde_vectors(gcmo_sstage* h, int pass);

void gcmo_sx_nodes(gcmo_sstage* h, int pass) {
	if (h->pde_mode) { nodes_pde_vectors(h, pass); return; }
	const gcmo_tri* t = h->t;
	const int M = h->M;
	for (int it = 0; it < t->n_local; it++) {
		const int state = gcmo_simplex_border_state(t, it);
		if ((pass == 0) != (state != 0)) { continue; }
		const int can_st = pass == 1;
		unsigned outers = 0;
		double* ans = h->next + (size_t) it * M;
		const R3 x0 = point(t, t->global_of[it]);
		for (int k = 0; k < M; k++) {
			const double dx = -h->tau * h->L[k];
			if (dx == 0) { ans[k] = h->riem[(size_t) it * M + k]; continue; }
			const R3 shift = scale(h->direction, dx);
			int cell[5];
			h->errors += gcmo_simplex_locate(t, it, shift.v, cell);
			double u = 0;
			if (cell[0] == 4) {
				const R3 q = add(x0, shift);
				int e = 0;
				u = gcmo_simplex_hybrid_interpolate(t, M, h->riem, h->grad, cell + 1, k, q.v, &e);
				h->errors += e;
			} else if (cell[0] == 0) {
				outers |= 1u << k;
			} else if (cell[0] == 3) {
				if (can_st) {
					R3 r[3];
					double vc[3], vn[3];
					for (int i = 0; i < 3; i++) {
						r[i] = point(t, t->global_of[cell[1 + i]]);
						vc[i] = h->riem[(size_t) cell[1 + i] * M + k];
						vn[i] = h->next[(size_t) cell[1 + i] * M + k];
					}
					int e = 0;
					u = interpolate_space_time(shift, x0, r, vc, vn, &e);
					h->errors += e;
				} else { outers |= 1u << k; }
			} else if (cell[0] == 2) {
				if (can_st) { h->errors++; /* THROW_UNSUPPORTED in 3-D */ } else { outers |= 1u << k; }
			}
			ans[k] = u;
		}
		if (pass == 0) {
			/* …InRiemannInvariants.hpp:73-85 */
			if (outers != h->RIGHT && outers != h->LEFT && outers != (h->LEFT | h->RIGHT) && outers != 0) {
				if (outers & h->RIGHT) { outers |= h->RIGHT; }
				if (outers & h->LEFT) { outers |= h->LEFT; }
				for (int k = 0; k < M; k++) { if ((outers >> k) & 1u) { ans[k] = 0; } }
			}
			h->waves[it] = outers;
		}
	}
}

/* BorderCorrectorInRiemannInvariants::applyInGlobalBasis for every condition in order
 * (Engine.cpp:158-167; BorderCorrector.hpp:122-174,241-286) */
void gcmo_sx_border_correct(gcmo_sstage* h, int n_border, const int* border_node, const double* border_normal,
		const int* border_cond, int n_cond, const int* cond_type, const double* cond_b) {
	const int M = h->M, outer = h->outer, model = h->model;
	const double *U = h->U, *U1 = h->U1;
	double* next = h->next;
	for (int c = 0; c < n_cond; c++) {
		int first = -1;
		for (int i = 0; i < n_border; i++) { if (border_cond[i] == c) { first = i; break; } }
		if (first < 0) { continue; }
		const double* b = cond_b + (size_t) c * outer;
		double Omega[9 * 3], B[3 * 9], tmp[9], value[9], det;
		/* getMaximalPossibleDeterminant: aligned case, right invariants */
		columns(M, outer, U1, h->RIGHT, Omega);
		border_matrix(model, M, cond_type[c], h->direction, B);
		memset(tmp, 0, sizeof tmp);
		outer_wave_correction(M, outer, tmp, Omega, B, b, 0, value, &det);
		const double min_det = 1e-3 * det;
		for (int i = 0; i < n_border; i++) {
			if (border_cond[i] != c) { continue; }
			const int node = border_node[i];
			const R3 normal = r3(border_normal + 3 * i);
			double u[9], w[9];
			if (h->pde_mode) { memcpy(u, next + (size_t) node * M, (size_t) M * sizeof(double)); }
			else { mat_vec(M, U1, next + (size_t) node * M, u); }  /* to PDE variables */
			border_matrix(model, M, cond_type[c], normal, B);
			const unsigned outers = h->waves[node];
			if (outers == h->RIGHT || outers == h->LEFT) {
				columns(M, outer, U1, outers, Omega);
				if (outer_wave_correction(M, outer, u, Omega, B, b, min_det, value, &det)) {
					for (int k = 0; k < M; k++) { u[k] += value[k]; }
				} else { plain_border(model, M, cond_type[c], normal, b, u); }
			} else {
				double vr[9], vl[9], d2;
				columns(M, outer, U1, h->RIGHT, Omega);
				const int okr = outer_wave_correction(M, outer, u, Omega, B, b, min_det, vr, &det);
				columns(M, outer, U1, h->LEFT, Omega);
				const int okl = outer_wave_correction(M, outer, u, Omega, B, b, min_det, vl, &d2);
				if (okr && okl) { for (int k = 0; k < M; k++) { u[k] += (vr[k] + vl[k]) / 2; } }
				else { plain_border(model, M, cond_type[c], normal, b, u); }
			}
			if (h->pde_mode) { memcpy(w, u, (size_t) M * sizeof(double)); }
			else { mat_vec(M, U, u, w); }                           /* back to invariants */
			memcpy(next + (size_t) node * M, w, (size_t) M * sizeof(double));
		}
	}
}

static int popcount_u(unsigned x) { int c = 0; while (x) { c += (int) (x & 1u); x >>= 1; } return c; }

/* ContactCorrectorInRiemannInvariants::applyInGlobalBasis (ContactCorrector.hpp:334-356, 381-410) around
 * ContactCorrectorInPdeVectors::applyInGlobalBasis (:133-253).  Both bodies carry the same model (the factory
 * offers elastic-elastic ADHESION and acoustic-acoustic SLIDE only, :484-560). */
void gcmo_sx_contact_correct(gcmo_sstage* a, gcmo_sstage* b, int n, const int* node_a, const int* node_b, const double* normals) {
	if (n == 0) { return; }
	const int M = a->M, o = a->outer, model = a->model;
	const unsigned LEFT = a->LEFT, RIGHT = a->RIGHT;
	double OmA[9 * 6], OmB[9 * 6], B1A[27], B1B[27], B2A[27], B2B[27], zero[9], vA[9], vB[9], det1, det2;
	/* getMaximalPossibleDeterminants (:273-300): calculation direction as the normal */
	memset(zero, 0, sizeof zero);
	columns(M, o, a->U1, LEFT, OmA);
	columns(M, o, b->U1, RIGHT, OmB);
	contact_matrix(model, M, 1, a->direction, B1A); contact_matrix(model, M, 1, a->direction, B1B);
	contact_matrix(model, M, 2, a->direction, B2A); contact_matrix(model, M, 2, a->direction, B2B);
	if (!contact_wave_correction(M, o, zero, OmA, B1A, B2A, zero, OmB, B1B, B2B, 0, 0, vA, vB, &det1, &det2, &a->errors)) { a->errors++; }
	const double min1 = 1e-3 * det1, min2 = 1e-3 * det2;
	for (int i = 0; i < n; i++) {
		double* ra = a->next + (size_t) node_a[i] * M;
		double* rb = b->next + (size_t) node_b[i] * M;
		unsigned wa = a->waves[node_a[i]], wb = b->waves[node_b[i]];
		/* matchInnersAndOuters */
		const int N = (popcount_u(wa) + popcount_u(wb)) / o;
		if (!a->pde_mode && N % 2 != 0) {
			if (N == 3) { wa = wb = LEFT | RIGHT; }
			else if (wa == 0) { if (wb == LEFT) { wa = RIGHT; } else { if (wb != RIGHT) { a->errors++; } wa = LEFT; } }
			else { if (wb != 0) { a->errors++; } if (wa == LEFT) { wb = RIGHT; } else { if (wa != RIGHT) { a->errors++; } wb = LEFT; } }
			for (int k = 0; k < M; k++) { if ((wa >> k) & 1u) { ra[k] = 0; } if ((wb >> k) & 1u) { rb[k] = 0; } }
		}
		a->waves[node_a[i]] = wa; b->waves[node_b[i]] = wb;
		double uA[9], uB[9], w[9];
		if (a->pde_mode) { memcpy(uA, ra, (size_t) M * sizeof(double)); memcpy(uB, rb, (size_t) M * sizeof(double)); }
		else { mat_vec(M, a->U1, ra, uA); mat_vec(M, b->U1, rb, uB); }
		const R3 normal = r3(normals + 3 * i);
		contact_matrix(model, M, 1, normal, B1A); contact_matrix(model, M, 1, normal, B1B);
		contact_matrix(model, M, 2, normal, B2A); contact_matrix(model, M, 2, normal, B2B);
		const int na = popcount_u(wa), nb = popcount_u(wb);
		if (na == o && nb == o) {
			columns(M, o, a->U1, wa, OmA);
			columns(M, o, b->U1, wb, OmB);
			if (contact_wave_correction(M, o, uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA, vB, &det1, &det2, &a->errors)) {
				for (int k = 0; k < M; k++) { uA[k] += vA[k]; uB[k] += vB[k]; }
			} else { plain_contact(model, 1, normal, uA, uB); }
		} else if ((na == 2 * o && nb == 0) || (nb == 2 * o && na == 0)) {
			/* the node with both families outer is a border with two conditions taken from the other node */
			const int first = na == 2 * o;
			double* uX = first ? uA : uB;
			double* uY = first ? uB : uA;
			const double* U1X = first ? a->U1 : b->U1;
			const double *B1X = first ? B1A : B1B, *B2X = first ? B2A : B2B, *B1Y = first ? B1B : B1A, *B2Y = first ? B2B : B2A;
			double Bc[6 * 9], b12[6], value[9], det;
			memcpy(Bc, B1X, (size_t) (o * M) * sizeof(double));
			memcpy(Bc + o * M, B2X, (size_t) (o * M) * sizeof(double));
			mat_mul(o, M, 1, B1Y, uY, b12);
			mat_mul(o, M, 1, B2Y, uY, b12 + o);
			columns_right_left(M, o, U1X, RIGHT, LEFT, OmA);
			if (outer_wave_correction_n(M, 2 * o, uX, OmA, Bc, b12, min1, value, &det, &a->errors)) {
				for (int k = 0; k < M; k++) { uX[k] += value[k]; }
			} else { plain_contact(model, 0, normal, uX, uY); }
		} else {
			double vA2[9], vB2[9], d1, d2;
			columns(M, o, a->U1, RIGHT, OmA);
			columns(M, o, b->U1, LEFT, OmB);
			const int ok1 = contact_wave_correction(M, o, uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA, vB, &d1, &d2, &a->errors);
			columns(M, o, a->U1, LEFT, OmA);
			columns(M, o, b->U1, RIGHT, OmB);
			const int ok2 = contact_wave_correction(M, o, uA, OmA, B1A, B2A, uB, OmB, B1B, B2B, min1, min2, vA2, vB2, &d1, &d2, &a->errors);
			if (ok1 && ok2) { for (int k = 0; k < M; k++) { uA[k] += (vA[k] + vA2[k]) / 2; uB[k] += (vB[k] + vB2[k]) / 2; } }
			else { plain_contact(model, 1, normal, uA, uB); }
		}
		if (a->pde_mode) { memcpy(ra, uA, (size_t) M * sizeof(double)); memcpy(rb, uB, (size_t) M * sizeof(double)); }
		else {
			mat_vec(M, a->U, uA, w); memcpy(ra, w, (size_t) M * sizeof(double));
			mat_vec(M, b->U, uB, w); memcpy(rb, w, (size_t) M * sizeof(double));
		}
	}
}

/* afterStage (…InRiemannInvariants.hpp:116-127); frees the handle and returns its error count */
int gcmo_sx_end(gcmo_sstage* h) {
	const int M = h->M;
	for (int v = 0; v < h->t->n_local && !h->pde_mode; v++) {
		double w[9];
		mat_vec(M, h->U1, h->next + (size_t) v * M, w);
		memcpy(h->next + (size_t) v * M, w, (size_t) M * sizeof(double));
	}
	const int errors = h->errors;
	free(h->riem); free(h->grad); free(h->waves); free(h);
	return errors;
}

int gcmo_simplex_stage(const gcmo_tri* t, int model, int M, int s, double tau,
		const double* U_, const double* U1_, const double* L_, const double* basis,
		int n_border, const int* border_node, const double* border_normal, const int* border_cond,
		int n_cond, const int* cond_type, const double* cond_b,
		const double* cur, double* next) {
	gcmo_sstage* h = gcmo_sx_begin(t, model, M, s, tau, U_, U1_, L_, basis, cur, next, 0);
	gcmo_sx_nodes(h, 0);
	gcmo_sx_border_correct(h, n_border, border_node, border_normal, border_cond, n_cond, cond_type, cond_b);
	gcmo_sx_nodes(h, 1);
	return gcmo_sx_end(h);
}

/* linal/geometry.hpp:275-284 */
static double minimal_height4(R3 a, R3 b, R3 c, R3 d) {
	const double V = volume(a, b, c, d);
	const double A = area3(b, c, d), B = area3(c, d, a), C = area3(d, a, b), D = area3(a, b, c);
	return 3 * V / fmax(A, fmax(B, fmax(C, D)));
}

/* SimplexGrid::getAverageHeight: the MEAN OF A 100-BIN HISTOGRAM of the cells' minimal heights
 * (SimplexGrid.cpp:183-194,266-274; util/math/Histogram.hpp:14-58), which Engine::estimateTimeStep uses
 * (engine/simplex/Engine.hpp:77-92).  out[0] = average, out[1] = minimum */
void gcmo_simplex_heights(const gcmo_tri* t, double out[2]) {
	const size_t bins_n = 100;
	double* h = (double*) malloc((size_t) t->nC * sizeof(double));
	size_t n = 0;
	for (int c = 0; c < t->nC; c++) {
		if (!is_local(t, c)) { continue; }
		const int* v = t->cell_v + 4 * c;
		h[n++] = minimal_height4(point(t, v[0]), point(t, v[1]), point(t, v[2]), point(t, v[3]));
	}
	double lo = h[0], hi = h[0];
	for (size_t i = 1; i < n; i++) { if (h[i] < lo) { lo = h[i]; } if (hi < h[i]) { hi = h[i]; } }
	size_t bins[101];
	memset(bins, 0, sizeof bins);
	size_t used = bins_n;
	if (hi == lo) { bins[0] = n; }
	else {
		const double size = (hi - lo) / (double) bins_n;
		for (size_t i = 0; i < n; i++) { bins[(size_t) ((h[i] - lo) / size)]++; }
		bins[bins_n - 1] += bins[bins_n];
	}
	const double bin_size = (hi - lo) / (double) used;
	double dot_sum = 0, count = 0;
	for (size_t i = 0; i < used; i++) {
		const double center = lo + ((double) i + 0.5) * bin_size;
		dot_sum = dot_sum + (double) bins[i] * center;
		count = count + (double) bins[i];
	}
	out[0] = dot_sum / count;
	out[1] = lo;
	free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* GcmType::ADVECT_PDE_VECTORS (engine/simplex/GridCharacteristicMethodInPdeVectors.hpp:37-169)  */
/* ------------------------------------------------------------------------------------------ */
/* TetrahedronInterpolator<PdeVector>::hybridInterpolate (…Interpolator.hpp:44-104): the quadratic interpolant of
 * the whole vector, replaced by the linear one as soon as ANY component leaves the [min, max] of the 4 values */
static void hybrid_interpolate_vector(const gcmo_tri* t, int M, const double* values, const double* grad,
		const int cell[4], R3 q, double* out, int* err) {
	R3 c[4];
	double l[4], quadratic[9];
	for (int i = 0; i < 4; i++) { c[i] = point(t, t->global_of[cell[i]]); }
	if (barycentric4(c[0], c[1], c[2], c[3], q, l) || !is_interpolation(l)) { *err = 1; }
	int same = 1;
	for (int k = 0; k < M; k++) {
		double v[4];
		R3 g[4];
		for (int i = 0; i < 4; i++) {
			v[i] = values[(size_t) cell[i] * M + k];
			for (int d = 0; d < 3; d++) { g[i].v[d] = grad[((size_t) cell[i] * 3 + d) * M + k]; }
		}
		double x = l[0] * (v[0] + dot(g[0], sub(q, c[0])) / 2.0);
		for (int i = 1; i < 4; i++) { x = x + l[i] * (v[i] + dot(g[i], sub(q, c[i])) / 2.0); }
		quadratic[k] = x;
		const double mn = fmin(fmin(v[0], v[1]), fmin(v[2], v[3]));
		const double mx = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
		if (!(x == fmin(fmax(x, mn), mx))) { same = 0; }
	}
	for (int k = 0; k < M; k++) {
		if (same) { out[k] = quadratic[k]; continue; }
		double v[4];
		for (int i = 0; i < 4; i++) { v[i] = values[(size_t) cell[i] * M + k]; }
		out[k] = l[0] * v[0] + l[1] * v[1] + l[2] * v[2] + l[3] * v[3];
	}
}

static void nodes_pde_vectors(gcmo_sstage* h, int pass) {
	const gcmo_tri* t = h->t;
	const int M = h->M;
	for (int it = 0; it < t->n_local; it++) {
		const int state = gcmo_simplex_border_state(t, it);
		if ((pass == 0) != (state != 0)) { continue; }
		const int can_st = pass == 1;
		unsigned outers = 0;
		double V[9][9];   /* V[j][k]: component j of the vector interpolated at the foot of characteristic k */
		const R3 x0 = point(t, t->global_of[it]);
		for (int k = 0; k < M; k++) {
			double u[9];
			memset(u, 0, sizeof u);
			const double dx = -h->tau * h->L[k];
			if (dx == 0) {
				memcpy(u, h->riem + (size_t) it * M, (size_t) M * sizeof(double));
			} else {
				const R3 shift = scale(h->direction, dx);
				int cell[5];
				h->errors += gcmo_simplex_locate(t, it, shift.v, cell);
				if (cell[0] == 4) {
					int e = 0;
					hybrid_interpolate_vector(t, M, h->riem, h->grad, cell + 1, add(x0, shift), u, &e);
					h->errors += e;
				} else if (cell[0] == 0) {
					outers |= 1u << k;
				} else if (cell[0] == 3) {
					if (can_st) {
						R3 r[3];
						for (int i = 0; i < 3; i++) { r[i] = point(t, t->global_of[cell[1 + i]]); }
						for (int j = 0; j < M; j++) {
							double vc[3], vn[3];
							for (int i = 0; i < 3; i++) {
								vc[i] = h->riem[(size_t) cell[1 + i] * M + j];
								vn[i] = h->next[(size_t) cell[1 + i] * M + j];
							}
							int e = 0;
							u[j] = interpolate_space_time(shift, x0, r, vc, vn, &e);
							h->errors += e;
						}
					} else { outers |= 1u << k; }
				} else if (cell[0] == 2) {
					if (can_st) { h->errors++; } else { outers |= 1u << k; }
				}
			}
			for (int j = 0; j < M; j++) { V[j][k] = u[j]; }
		}
		/* localGcmStep (util/math/GridCharacteristicMethod.hpp:10-17): U1 * diagonalMultiply(U, V) */
		double r[9], w[9];
		for (int k = 0; k < M; k++) {
			double x = h->U[k * M] * V[0][k];
			for (int j = 1; j < M; j++) { x += h->U[k * M + j] * V[j][k]; }
			r[k] = x;
		}
		mat_vec(M, h->U1, r, w);
		memcpy(h->next + (size_t) it * M, w, (size_t) M * sizeof(double));
		if (pass == 0) { h->waves[it] = outers; }
	}
}


/* ---------------------------------------------------------------------------------------------------------------------
 * TriangleInterpolator (util/math/interpolation/TriangleInterpolator.hpp:8-130): the 2-D member of the simplex
 * interpolators.  Pinned bit for bit to the reference's own class (oracle/_ref/gcm_ref_interp, tests/golden/
 * triangle_interpolator.npz) and to the known answers of src/test/sequence/TestInterpolator.cpp:129-189,267-274.
 * ------------------------------------------------------------------------------------------------------------------- */
/* linal::barycentricCoordinates(a, b, c, q) (linal/geometry.hpp:108-116) through solveLinearSystem 2x2
 * (linal/linearSystems.hpp:60-88) and determinant 2x2 (linal/determinants.hpp:20-35); 0 = "SLE determinant is zero" */
static int tri2_barycentric(const double* a, const double* b, const double* c, const double* q, double* l) {
	const double T00 = a[0] - c[0], T01 = b[0] - c[0];
	const double T10 = a[1] - c[1], T11 = b[1] - c[1];
	const double r0 = q[0] - c[0], r1 = q[1] - c[1];
	const double det = T00 * T11 - T01 * T10;
	if (det == 0) { return 0; }
	l[0] = (r0 * T11 - T01 * r1) / det;
	l[1] = (T00 * r1 - r0 * T10) / det;
	l[2] = 1 - l[0] - l[1];
	return 1;
}
static int tri2_is_interpolation(const double* l) { return l[0] > -1e-9 && l[1] > -1e-9 && l[2] > -1e-9; }   /* :14-18 */
static double tri2_min(double a, double b) { return b < a ? b : a; }   /* std::min, folded left (linal/functions.hpp:560-598) */
static double tri2_max(double a, double b) { return a < b ? b : a; }

void gcmo_triangle_interpolate(int mode, int n, const double* points, const double* values, const double* grads,
                               const double* queries, double* out, int* status) {
	for (int i = 0; i < n; i++) {
		const double* q = queries + 2 * (size_t) i;
		out[i] = 0;
		status[i] = 1;
		if (mode == 4) {   /* interpolateInOwner :108-129 */
			static const int T[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
			const double* c = points + 8 * (size_t) i;
			const double* v = values + 4 * (size_t) i;
			for (int k = 0; k < 4; k++) {
				double l[3];
				if (!tri2_barycentric(c + 2 * T[k][0], c + 2 * T[k][1], c + 2 * T[k][2], q, l)) { break; }
				if (tri2_is_interpolation(l)) {
					out[i] = l[0] * v[T[k][0]] + l[1] * v[T[k][1]] + l[2] * v[T[k][2]];
					status[i] = 0;
					break;
				}
			}
			continue;
		}
		const double* c = points + 6 * (size_t) i;
		const double* v = values + 3 * (size_t) i;
		double l[3];
		if (!tri2_barycentric(c, c + 2, c + 4, q, l) || !tri2_is_interpolation(l)) { continue; }
		const double linear = l[0] * v[0] + l[1] * v[1] + l[2] * v[2];   /* :27-36 */
		status[i] = 0;
		if (mode == 0) { out[i] = linear; continue; }
		const double* g = grads + 6 * (size_t) i;
		double quadratic = 0;   /* :47-58 */
		for (int k = 0; k < 3; k++) {
			double dot = g[2 * k] * (q[0] - c[2 * k]);
			dot += g[2 * k + 1] * (q[1] - c[2 * k + 1]);
			const double term = l[k] * (v[k] + dot / 2.0);
			quadratic = k == 0 ? term : quadratic + term;
		}
		if (mode == 1) { out[i] = quadratic; continue; }
		const double lo = tri2_min(tri2_min(v[0], v[1]), v[2]), hi = tri2_max(tri2_max(v[0], v[1]), v[2]);
		const double limited = tri2_min(tri2_max(quadratic, lo), hi);   /* limiterMinMax, linal/functions.hpp:676-679 */
		out[i] = mode == 2 ? limited : (quadratic == limited ? quadratic : linear);   /* :69-78, :91-101 */
	}
}
