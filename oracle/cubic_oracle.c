/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's cubic GCM hot path.
 * See cubic_oracle.h for the parity status (PINNED against the unmodified reference) and the
 * rule that the product never touches this file.  Compiled with -ffp-contract=off so that, like
 * the reference build (CMakeLists.txt:6-7: no -march, no FMA), every operation rounds separately.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/src/libgcm).
 */
#include "cubic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* linal/Symmetry.hpp:41-47: index of (i,j) in the packed upper triangle of a DxD symmetric matrix */
static int sym_index(int D, int i, int j) {
	if (i > j) { int t = i; i = j; j = t; }
	return i * D - ((i - 1) * i) / 2 + j - i;
}

int gcmo_pde_size(int model, int D) {
	/* rheology/variables/VelocitySigmaVariables.hpp:16-18, AcousticVariables.hpp:15 */
	return model == 1 ? D + 1 : D + (D * (D + 1)) / 2;
}

/* linal/geometry.hpp:46-52 and linal/basis.hpp:49-66: local basis whose LAST column is n */
static void local_basis(int D, const double* n, double b[3][3]) {
	memset(b, 0, 9 * sizeof(double));
	if (D == 1) {
		b[0][0] = n[0];
	} else if (D == 2) {
		b[0][0] = n[1];  b[0][1] = n[0];
		b[1][0] = -n[0]; b[1][1] = n[1];
	} else {
		double a[3] = {n[1], -n[0], 0};
		if (n[0] == 0 && n[1] == 0) { a[0] = n[2]; a[1] = 0; a[2] = 0; }
		const double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
		const double la = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
		double t1[3], t2[3];
		for (int i = 0; i < 3; i++) { t1[i] = a[i] * ln / la; }
		t2[0] = n[1] * t1[2] - n[2] * t1[1];
		t2[1] = n[2] * t1[0] - n[0] * t1[2];
		t2[2] = n[0] * t1[1] - n[1] * t1[0];
		for (int i = 0; i < 3; i++) { b[i][0] = t1[i]; b[i][1] = t2[i]; b[i][2] = n[i]; }
	}
}

/* linal/functions.hpp:546-558 */
static void symm_direct_product(int D, const double* v1, const double* v2, double out[3][3]) {
	for (int i = 0; i < D; i++) {
		for (int j = 0; j <= i; j++) {
			out[i][j] = (v1[i] * v2[j] + v2[i] * v1[j]) / 2;
			out[j][i] = out[i][j];
		}
	}
}

/* VelocitySigmaVariables::setVelocity / setSigma into a PDE vector */
static void put_velocity(int D, const double* v, double* vec) {
	for (int i = 0; i < D; i++) { vec[i] = v[i]; }
}
static void put_sigma(int D, double s[3][3], double* vec) {
	for (int i = 0; i < D; i++) {
		for (int j = 0; j <= i; j++) { vec[D + sym_index(D, i, j)] = s[i][j]; }
	}
}
static void negate_sigma(int D, int M, double* vec) {
	for (int i = D; i < M; i++) { vec[i] = -vec[i]; }
}
static void set_row(int M, double* A, int r, const double* vec) {
	for (int j = 0; j < M; j++) { A[r * M + j] = vec[j]; }
}
static void set_col(int M, double* A, int c, const double* vec) {
	for (int i = 0; i < M; i++) { A[i * M + c] = vec[i]; }
}
/* ElasticModel.hpp:157-164: 2*s - Diag(s) */
static void tensor_to_vector(int D, double s[3][3], double out[3][3]) {
	for (int i = 0; i < D; i++) {
		for (int j = 0; j < D; j++) {
			out[i][j] = s[i][j] * 2 - (i == j ? s[i][j] : 0.0);
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* eigen-systems                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* rheology/models/ElasticModel.hpp:362-553 for one direction given by the local basis */
static void elastic_isotropic_direction(int D, double rho, double lambda, double mu,
		double basis[3][3], double* U, double* U1, double* L) {
	const int M = gcmo_pde_size(0, D);
	const double c1 = sqrt((lambda + 2 * mu) / rho);
	const double c2 = sqrt(mu / rho);
	const double alpha = 0.5;
	double n[3][3]; /* n[i] = basis column (i + D - 1) % D : ElasticModel.hpp:421-424 */
	for (int i = 0; i < D; i++) {
		for (int a = 0; a < D; a++) { n[i][a] = basis[a][(i + D - 1) % D]; }
	}
	double N[3][3][3][3];
	for (int i = 0; i < D; i++) {
		for (int j = 0; j <= i; j++) {
			symm_direct_product(D, n[i], n[j], N[i][j]);
			memcpy(N[j][i], N[i][j], sizeof(N[i][j]));
		}
	}
	double I[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	double vec[GCMO_MAX_M], v[3], s[3][3], t[3][3];
	memset(U, 0, (size_t) (M * M) * sizeof(double));
	memset(U1, 0, (size_t) (M * M) * sizeof(double));
	memset(L, 0, (size_t) M * sizeof(double));

	/* eigenvalues: ElasticModel.hpp:399-407 */
	L[0] = c1; L[1] = -c1;
	for (int i = 1; i < D; i++) { L[2 * i] = c2; L[2 * i + 1] = -c2; }

	/* ---- U1, right eigenvectors in columns: ElasticModel.hpp:415-480 ---- */
	memset(vec, 0, sizeof(vec));
	for (int a = 0; a < D; a++) { v[a] = n[0][a] * alpha; }
	put_velocity(D, v, vec);
	{
		const double f = -alpha / c1;
		const double twomu = 2 * mu;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) {
			s[a][b] = (I[a][b] * lambda + N[0][0][a][b] * twomu) * f;
		}
	}
	put_sigma(D, s, vec);
	set_col(M, U1, 0, vec);
	negate_sigma(D, M, vec);
	set_col(M, U1, 1, vec);
	for (int i = 1; i < D; i++) {
		for (int a = 0; a < D; a++) { v[a] = n[i][a] * alpha; }
		put_velocity(D, v, vec);
		const double f = -2 * alpha * mu / c2;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { s[a][b] = N[0][i][a][b] * f; }
		put_sigma(D, s, vec);
		set_col(M, U1, 2 * i, vec);
		negate_sigma(D, M, vec);
		set_col(M, U1, 2 * i + 1, vec);
	}
	for (int a = 0; a < D; a++) { v[a] = 0; }
	put_velocity(D, v, vec);
	if (D == 3) {
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s[a][b] = N[1][2][a][b] * 2; }
		put_sigma(D, s, vec); set_col(M, U1, 6, vec);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s[a][b] = (N[1][1][a][b] - N[2][2][a][b]) / 2; }
		put_sigma(D, s, vec); set_col(M, U1, 7, vec);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s[a][b] = (N[1][1][a][b] + N[2][2][a][b]) / 2; }
		put_sigma(D, s, vec); set_col(M, U1, 8, vec);
	} else if (D == 2) {
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { s[a][b] = I[a][b] - N[0][0][a][b]; }
		put_sigma(D, s, vec); set_col(M, U1, 4, vec);
	}

	/* ---- U, left eigenvectors in rows: ElasticModel.hpp:484-553 ---- */
	memset(vec, 0, sizeof(vec));
	put_velocity(D, n[0], vec);
	{
		const double d = -c1 * rho;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { t[a][b] = N[0][0][a][b] / d; }
		tensor_to_vector(D, t, s);
	}
	put_sigma(D, s, vec);
	set_row(M, U, 0, vec);
	negate_sigma(D, M, vec);
	set_row(M, U, 1, vec);
	for (int i = 1; i < D; i++) {
		put_velocity(D, n[i], vec);
		const double d = -c2 * rho;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { t[a][b] = N[0][i][a][b] / d; }
		tensor_to_vector(D, t, s);
		put_sigma(D, s, vec);
		set_row(M, U, 2 * i, vec);
		negate_sigma(D, M, vec);
		set_row(M, U, 2 * i + 1, vec);
	}
	for (int a = 0; a < D; a++) { v[a] = 0; }
	put_velocity(D, v, vec);
	if (D == 3) {
		tensor_to_vector(D, N[1][2], s);
		put_sigma(D, s, vec); set_row(M, U, 6, vec);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { t[a][b] = N[1][1][a][b] - N[2][2][a][b]; }
		tensor_to_vector(D, t, s);
		put_sigma(D, s, vec); set_row(M, U, 7, vec);
		const double g = 2 * lambda / (lambda + 2 * mu);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
			t[a][b] = (N[1][1][a][b] + N[2][2][a][b]) - N[0][0][a][b] * g;
		}
		tensor_to_vector(D, t, s);
		put_sigma(D, s, vec); set_row(M, U, 8, vec);
	} else if (D == 2) {
		const double g = lambda / (lambda + 2 * mu);
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
			t[a][b] = N[1][1][a][b] - N[0][0][a][b] * g;
		}
		tensor_to_vector(D, t, s);
		put_sigma(D, s, vec); set_row(M, U, 4, vec);
	}
}

/* ElasticModel.hpp:56-65: identity global basis, direction i along the i-th axis */
void gcmo_elastic_isotropic(int D, double rho, double lambda, double mu,
		double* U, double* U1, double* L) {
	const int M = gcmo_pde_size(0, D);
	for (int i = 0; i < D; i++) {
		double n[3] = {0, 0, 0};
		n[i] = 1;
		double b[3][3];
		local_basis(D, n, b);
		elastic_isotropic_direction(D, rho, lambda, mu, b,
				U + (size_t) i * M * M, U1 + (size_t) i * M * M, L + (size_t) i * M);
	}
}

/* rheology/models/AcousticModel.hpp:214-317 */
void gcmo_acoustic(int D, double rho, double lambda, double* U, double* U1, double* L) {
	const int M = D + 1;
	const double c1 = sqrt(lambda / rho);
	const double alpha = 0.5;
	for (int dir = 0; dir < D; dir++) {
		double nn[3] = {0, 0, 0};
		nn[dir] = 1;
		double basis[3][3], n[3][3];
		local_basis(D, nn, basis);
		for (int i = 0; i < D; i++) {
			for (int a = 0; a < D; a++) { n[i][a] = basis[a][(i + D - 1) % D]; }
		}
		double* u = U + (size_t) dir * M * M;
		double* u1 = U1 + (size_t) dir * M * M;
		double* l = L + (size_t) dir * M;
		memset(u, 0, (size_t) (M * M) * sizeof(double));
		memset(u1, 0, (size_t) (M * M) * sizeof(double));
		memset(l, 0, (size_t) M * sizeof(double));
		l[0] = c1; l[1] = -c1;
		double vec[GCMO_MAX_M];
		/* U1: AcousticModel.hpp:247-272 */
		for (int a = 0; a < D; a++) { vec[a] = n[0][a]; }
		vec[D] = c1 * rho;
		set_col(M, u1, 0, vec);
		vec[D] = -vec[D];
		set_col(M, u1, 1, vec);
		vec[D] = 0;
		for (int i = 1; i < D; i++) {
			for (int a = 0; a < D; a++) { vec[a] = n[i][a]; }
			set_col(M, u1, i + 1, vec);
		}
		/* U: AcousticModel.hpp:276-303 */
		for (int a = 0; a < D; a++) { vec[a] = n[0][a] * alpha; }
		vec[D] = alpha / (c1 * rho);
		set_row(M, u, 0, vec);
		vec[D] = -vec[D];
		set_row(M, u, 1, vec);
		vec[D] = 0;
		for (int i = 1; i < D; i++) {
			for (int a = 0; a < D; a++) { vec[a] = n[i][a]; }
			set_row(M, u, i + 1, vec);
		}
	}
}

/* rheology/models/ElasticModel3D.cpp:288-429 (3-D) and ElasticModel2D.cpp:8-76 (2-D), material axes
 * along the coordinate axes.  c = {c11,c12,c13,c22,c23,c33,c44,c55,c66} (OrthotropicMaterial.hpp:13-22) */
void gcmo_elastic_orthotropic(int D, double rho, const double c[9], double* U, double* U1, double* L) {
	const int M = gcmo_pde_size(0, D);
	memset(U, 0, (size_t) (D * M * M) * sizeof(double));
	memset(U1, 0, (size_t) (D * M * M) * sizeof(double));
	memset(L, 0, (size_t) (D * M) * sizeof(double));
	if (D == 3) {
		const double c11 = c[0], c12 = c[1], c13 = c[2], c22 = c[3], c23 = c[4], c33 = c[5],
		             c44 = c[6], c55 = c[7], c66 = c[8];
		/* stiffness tables: diag[s] = c_ss, shear(a,b), coupling(p,s) */
		const double diag[3] = {c11, c22, c33};
		const double shear[3][3] = {{0, c66, c55}, {c66, 0, c44}, {c55, c44, 0}};
		const double coup[3][3] = {{0, c12, c13}, {c12, 0, c23}, {c13, c23, 0}};
		for (int s = 0; s < 3; s++) {
			double* u = U + (size_t) s * 81;
			double* u1 = U1 + (size_t) s * 81;
			double* l = L + (size_t) s * 9;
			int pair = 0;
			/* shear waves first (ascending velocity component), then the P wave */
			int order[3], k = 0;
			for (int a = 0; a < 3; a++) if (a != s) { order[k++] = a; }
			order[2] = s;
			for (pair = 0; pair < 3; pair++) {
				const int a = order[pair];
				const double cc = (a == s) ? diag[s] : shear[a][s];
				const int sg = 3 + sym_index(3, a, s);
				const int r = 2 * pair;
				l[r] = -sqrt(cc / rho);
				l[r + 1] = sqrt(cc / rho);
				u[r * 9 + a] = 1.0;
				u[r * 9 + sg] = 1.0 / (sqrt(cc) * sqrt(rho));
				u[(r + 1) * 9 + a] = 1.0;
				u[(r + 1) * 9 + sg] = -1.0 / (sqrt(cc) * sqrt(rho));
				u1[a * 9 + r] = 0.5;
				u1[a * 9 + r + 1] = 0.5;
				u1[sg * 9 + r] = 0.5 * sqrt(cc) * sqrt(rho);
				u1[sg * 9 + r + 1] = -0.5 * sqrt(cc) * sqrt(rho);
			}
			/* passive stress components (not touching axis s), ascending component index */
			int row = 6;
			for (int comp = 3; comp < 9; comp++) {
				int pa = -1, pb = -1;
				for (int a = 0; a < 3; a++) for (int b = a; b < 3; b++) {
					if (3 + sym_index(3, a, b) == comp) { pa = a; pb = b; }
				}
				if (pa == s || pb == s) { continue; }
				u[row * 9 + comp] = 1.0;
				u1[comp * 9 + row] = 1;
				if (pa == pb) {
					const int ss = 3 + sym_index(3, s, s);
					const double cps = coup[pa][s];
					u[row * 9 + ss] = -cps / diag[s];
					double w;
					if (s == 1 && pa == 0) {
						/* ElasticModel3D.cpp:361-362 writes this one entry in a different form */
						w = (0.5 * cps) / sqrt(diag[s] / rho);
					} else {
						w = (0.5 * cps * sqrt(rho)) / sqrt(diag[s]);
					}
					u1[comp * 9 + 4] = w;
					u1[comp * 9 + 5] = -w;
				}
				row++;
			}
		}
	} else if (D == 2) {
		/* 2-D takes c11, c12, c22, c66 (ElasticModel2D.cpp:84-91) */
		const double c11 = c[0], c12 = c[1], c22 = c[3], c66 = c[8];
		const double cp1 = sqrt(c11 / rho), cp2 = sqrt(c22 / rho), cs = sqrt(c66 / rho);
		double* u = U; double* u1 = U1; double* l = L;
		l[0] = -cs; l[1] = cs; l[2] = -cp1; l[3] = cp1; l[4] = 0;
		u[0 * 5 + 1] = 1.0; u[0 * 5 + 3] = 1.0 / (rho * cs);
		u[1 * 5 + 1] = 1.0; u[1 * 5 + 3] = -1.0 / (rho * cs);
		u[2 * 5 + 0] = 1.0; u[2 * 5 + 2] = 1.0 / (rho * cp1);
		u[3 * 5 + 0] = 1.0; u[3 * 5 + 2] = -1.0 / (rho * cp1);
		u[4 * 5 + 2] = 1.0 / c11; u[4 * 5 + 4] = -1.0 / c12;
		u1[0 * 5 + 2] = 0.5; u1[0 * 5 + 3] = 0.5;
		u1[1 * 5 + 0] = 0.5; u1[1 * 5 + 1] = 0.5;
		u1[2 * 5 + 2] = 0.5 * rho * cp1; u1[2 * 5 + 3] = -0.5 * rho * cp1;
		u1[3 * 5 + 0] = 0.5 * rho * cs; u1[3 * 5 + 1] = -0.5 * rho * cs;
		u1[4 * 5 + 2] = 0.5 * c12 / cp1; u1[4 * 5 + 3] = -0.5 * c12 / cp1; u1[4 * 5 + 4] = -c12;
		u += 25; u1 += 25; l += 5;
		l[0] = -cs; l[1] = cs; l[2] = -cp2; l[3] = cp2; l[4] = 0;
		u[0 * 5 + 0] = 1.0; u[0 * 5 + 3] = 1.0 / (rho * cs);
		u[1 * 5 + 0] = 1.0; u[1 * 5 + 3] = -1.0 / (rho * cs);
		u[2 * 5 + 1] = 1.0; u[2 * 5 + 4] = 1.0 / (rho * cp2);
		u[3 * 5 + 1] = 1.0; u[3 * 5 + 4] = -1.0 / (rho * cp2);
		u[4 * 5 + 2] = 1.0; u[4 * 5 + 4] = -c12 / c22;
		u1[0 * 5 + 0] = 0.5; u1[0 * 5 + 1] = 0.5;
		u1[1 * 5 + 2] = 0.5; u1[1 * 5 + 3] = 0.5;
		u1[2 * 5 + 2] = 0.5 * c12 / cp2; u1[2 * 5 + 3] = -0.5 * c12 / cp2; u1[2 * 5 + 4] = 1.0;
		u1[3 * 5 + 0] = 0.5 * rho * cs; u1[3 * 5 + 1] = -0.5 * rho * cs;
		u1[4 * 5 + 2] = 0.5 * rho * cp2; u1[4 * 5 + 3] = -0.5 * rho * cp2;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* interpolation: util/math/interpolation/EqualDistanceLineInterpolator.hpp                    */
/* ------------------------------------------------------------------------------------------ */


/* ---- rotated orthotropic material, 3-D ---------------------------------------------------------------------------
 * rheology/models/ElasticModel3D.cpp:8-283 (constructRotated and its helpers), rheology/materials/AbstractMaterial.hpp:27-133
 * (rotation of the elastic matrix through the rank-4 tensor), util/math/GslUtils.hpp:163-204 (cubic roots, around
 * GSL's gsl_poly_solve_cubic, poly/solve_cubic.c of GSL, restated here since GSL is absent),
 * linal/linearSystems.hpp:104-129,169-244 (Cramer, degenerate systems).  long double where the reference uses it. */
typedef long double ldbl;
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* <math.h> hides it under -std=c11 */
#endif

static int gsl_cubic(double a, double b, double c, double x[3]) {
	const double q = (a * a - 3 * b), r = (2 * a * a * a - 9 * a * b + 27 * c);
	const double Q = q / 9, R = r / 54;
	const double Q3 = Q * Q * Q, R2 = R * R;
	const double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
	if (R == 0 && Q == 0) { x[0] = x[1] = x[2] = -a / 3; return 3; }
	if (CR2 == CQ3) {
		const double sq = sqrt(Q);
		if (R > 0) { x[0] = -2 * sq - a / 3; x[1] = sq - a / 3; x[2] = sq - a / 3; }
		else { x[0] = -sq - a / 3; x[1] = -sq - a / 3; x[2] = 2 * sq - a / 3; }
		return 3;
	}
	if (R2 < Q3) {
		const double sgn = (R >= 0 ? 1 : -1);
		const double theta = acos(sgn * sqrt(R2 / Q3));
		const double norm = -2 * sqrt(Q);
		double t;
		x[0] = norm * cos(theta / 3) - a / 3;
		x[1] = norm * cos((theta + 2.0 * M_PI) / 3) - a / 3;
		x[2] = norm * cos((theta - 2.0 * M_PI) / 3) - a / 3;
		if (x[0] > x[1]) { t = x[0]; x[0] = x[1]; x[1] = t; }
		if (x[1] > x[2]) {
			t = x[1]; x[1] = x[2]; x[2] = t;
			if (x[0] > x[1]) { t = x[0]; x[0] = x[1]; x[1] = t; }
		}
		return 3;
	}
	return 1; /* one real root: the reference would go on to gsl_poly_complex_solve_cubic */
}

/* GslUtils.hpp:163-204: nearly equal roots (relative 1e-2) are averaged and moved to the end */
static int third_order_roots(const double p[3], double out[3]) {
	const double eps = 1e-2;
	double x[3];
	if (gsl_cubic(p[0], p[1], p[2], x) != 3) { return -1; }
	double x1 = x[0], x2 = x[1], x3 = x[2];
	if (fabs(x1 - x2) < fmax(fabs(x1), fabs(x2)) * eps) {
		if (fabs(x3 - x2) < fmax(fabs(x3), fabs(x2)) * eps) { x1 = x2 = x3 = (x1 + x2 + x3) / 3; }
		else { x2 = (x1 + x2) / 2; x1 = x3; x3 = x2; }
	} else if (fabs(x1 - x3) < fmax(fabs(x1), fabs(x3)) * eps) {
		x3 = (x1 + x3) / 2; x1 = x2; x2 = x3;
	} else if (fabs(x2 - x3) < fmax(fabs(x2), fabs(x3)) * eps) {
		x2 = x3 = (x2 + x3) / 2;
	}
	out[0] = x1; out[1] = x2; out[2] = x3;
	return 0;
}

static const int VOIGT[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}};
static int pair3(int i, int j) { return i <= j ? i * 3 - ((i - 1) * i) / 2 + j - i : pair3(j, i); }

/* AbstractMaterial.hpp:27-133: C (Voigt 6x6) -> c_ijkl -> G G G G c -> Voigt, G = Z(a2) Y(a1) X(a0) */
static void rotate_stiffness(const double c9[9], const double ang[3], double C[6][6]) {
	double C0[6][6] = {{0}};
	C0[0][0] = c9[0]; C0[0][1] = C0[1][0] = c9[1]; C0[0][2] = C0[2][0] = c9[2];
	C0[1][1] = c9[3]; C0[1][2] = C0[2][1] = c9[4]; C0[2][2] = c9[5];
	C0[3][3] = c9[6]; C0[4][4] = c9[7]; C0[5][5] = c9[8];
	double t[6][6], r[6][6] = {{0}};
	for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
		t[pair3(VOIGT[a][0], VOIGT[a][1])][pair3(VOIGT[b][0], VOIGT[b][1])] = C0[a][b];
	}
	const double X[3][3] = {{1.0, 0.0, 0.0}, {0.0, cos(ang[0]), sin(ang[0])}, {0.0, -sin(ang[0]), cos(ang[0])}};
	const double Y[3][3] = {{cos(ang[1]), 0.0, -sin(ang[1])}, {0.0, 1.0, 0.0}, {sin(ang[1]), 0.0, cos(ang[1])}};
	const double Z[3][3] = {{cos(ang[2]), sin(ang[2]), 0.0}, {-sin(ang[2]), cos(ang[2]), 0.0}, {0.0, 0.0, 1.0}};
	double ZY[3][3], G[3][3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		ZY[i][j] = Z[i][0] * Y[0][j]; ZY[i][j] += Z[i][1] * Y[1][j]; ZY[i][j] += Z[i][2] * Y[2][j];
	}
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		G[i][j] = ZY[i][0] * X[0][j]; G[i][j] += ZY[i][1] * X[1][j]; G[i][j] += ZY[i][2] * X[2][j];
	}
	for (int m = 0; m < 3; m++) for (int n = m; n < 3; n++) for (int p = 0; p < 3; p++) for (int q = p; q < 3; q++) {
		double acc = 0;
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
			acc += G[m][i] * G[n][j] * G[p][k] * G[q][l] * t[pair3(i, j)][pair3(k, l)];
		}
		r[pair3(m, n)][pair3(p, q)] = acc;
	}
	/* back to Voigt: the reference reads the upper triangle, and within the shear block takes (later, earlier) */
	for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) {
		const int f = (a >= 3 && b > a) ? b : a, g = (a >= 3 && b > a) ? a : b;
		C[a][b] = C[b][a] = r[pair3(VOIGT[f][0], VOIGT[f][1])][pair3(VOIGT[g][0], VOIGT[g][1])];
	}
}

static const int RHO_COLS[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};   /* getColumnsWithRho, ElasticModel3D.cpp:8-22 */
static const int ZERO_COLS[3][3] = {{6, 7, 8}, {3, 5, 8}, {3, 4, 6}};  /* getZeroColumns, :25-39 */

/* linal/linearSystems.hpp:169-244 for a 3x3 long double matrix of rank 2 (one solution) or 1 (two) */
static void degenerate_solutions(ldbl A[3][3], int count, ldbl x[3], ldbl y[3]) {
	x[0] = x[1] = x[2] = y[0] = y[1] = y[2] = 0;
	if (count == 1) {
		int I = 0, J = 1, P = 0, Q = 1;
		ldbl det = 0;
		for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
			const ldbl d = A[p][i] * A[q][j] - A[q][i] * A[p][j];
			if (fabsl(d) > fabsl(det)) { det = A[p][i] * A[q][j] - A[q][i] * A[p][j]; I = i; J = j; P = p; Q = q; }
		}
		int U = 2;
		for (int k = 0; k < 2; k++) { if (k != I && k != J) { U = k; break; } }
		x[U] = 1;
		x[I] = (-A[P][U] * A[Q][J] + A[Q][U] * A[P][J]) / det;
		x[J] = (-A[P][I] * A[Q][U] + A[Q][I] * A[P][U]) / det;
		return;
	}
	int I = 0, J = 0;
	ldbl det = 0;
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		if (fabsl(det) < fabsl(A[i][j])) { det = A[i][j]; I = i; J = j; }
	}
	const int p = (J == 0) ? 1 : 0, q = (J == 2) ? 1 : 2;
	x[p] = y[q] = 1;
	x[J] = -A[I][p] / det;
	y[J] = -A[I][q] / det;
}

/* findEigenvectors / findEigenstrings, ElasticModel3D.cpp:74-148 */
static void eigen_of(int vectors, ldbl l, double A[9][9], int s, int count, ldbl out[2][9]) {
	const int i = RHO_COLS[s][0], j = RHO_COLS[s][1], k = RHO_COLS[s][2];
	const int rows[3] = {i, j, k};
	const ldbl r = A[0][i];
	ldbl m[3][3], sol[2][3];
	for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
		m[a][b] = vectors ? (ldbl) A[rows[a]][b] : (ldbl) A[rows[b]][a];
		if (a == b) { m[a][b] = A[rows[a]][a] - l * l / r; }
	}
	degenerate_solutions(m, count, sol[0], sol[1]);
	for (int n = 0; n < count; n++) {
		ldbl* e = out[n];
		for (int a = 0; a < 9; a++) { e[a] = 0; }
		if (vectors) {
			for (int a = 0; a < 3; a++) { e[a] = sol[n][a]; }
			for (int a = 0; a < 3; a++) { e[rows[a]] = l / r * e[a]; }
			for (int z = 0; z < 3; z++) {
				const int p = ZERO_COLS[s][z];
				e[p] = (A[p][0] * e[0] + A[p][1] * e[1] + A[p][2] * e[2]) / l;
			}
		} else {
			for (int a = 0; a < 3; a++) { e[rows[a]] = sol[n][a]; }
			for (int a = 0; a < 3; a++) { e[a] = l / r * e[rows[a]]; }
		}
	}
}

static ldbl det3_ld(ldbl a11, ldbl a12, ldbl a13, ldbl a21, ldbl a22, ldbl a23, ldbl a31, ldbl a32, ldbl a33) {
	return a11 * (a22 * a33 - a23 * a32) - a12 * (a21 * a33 - a23 * a31) + a13 * (a21 * a32 - a22 * a31);
}

/* ElasticModel<3>::constructRotated, ElasticModel3D.cpp:151-283.  Returns 0, or -1 when the characteristic cubic has
 * complex roots (the reference's fallback needs gsl_poly_complex_solve_cubic), -2 for a singular system. */
int gcmo_elastic_orthotropic_rotated(double rho, const double c9[9], const double angles[3], double* U, double* U1, double* L) {
	double C[6][6];
	rotate_stiffness(c9, angles, C);
	memset(U, 0, 3 * 81 * sizeof(double));
	memset(U1, 0, 3 * 81 * sizeof(double));
	memset(L, 0, 3 * 9 * sizeof(double));
	/* sigma rows in PDE order xx xy xz yy yz zz = Voigt 0 5 4 1 3 2; velocity-gradient columns of stage s:
	 * d v_x / d x_s, d v_y / d x_s, d v_z / d x_s pair with Voigt entries of (x,s), (y,s), (z,s) */
	static const int ROW_VOIGT[6] = {0, 5, 4, 1, 3, 2};
	static const int COL_VOIGT[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
	for (int s = 0; s < 3; s++) {
		double A[9][9] = {{0}};
		for (int v = 0; v < 3; v++) { A[v][RHO_COLS[s][v]] = -1.0 / rho; }
		for (int row = 0; row < 6; row++) for (int v = 0; v < 3; v++) { A[3 + row][v] = -C[ROW_VOIGT[row]][COL_VOIGT[s][v]]; }
		double* Us = U + s * 81;
		double* U1s = U1 + s * 81;
		double* Ls = L + s * 9;
		const int i = RHO_COLS[s][0], j = RHO_COLS[s][1], k = RHO_COLS[s][2];
		/* constructEigenvaluesPolynomial, :50-71 */
		const ldbl r = A[0][i];
		const ldbl p0 = r * (-A[k][2] - A[j][1] - A[i][0]);
		const ldbl p1 = r * r * ((A[j][1] + A[i][0]) * A[k][2] - A[j][2] * A[k][1] - A[i][2] * A[k][0] + A[i][0] * A[j][1] - A[i][1] * A[j][0]);
		const ldbl p2 = r * r * r * ((-A[i][0] * A[j][1] + A[i][1] * A[j][0]) * A[k][2] + (A[i][0] * A[j][2] - A[i][2] * A[j][0]) * A[k][1] +
		                             (-A[i][1] * A[j][2] + A[i][2] * A[j][1]) * A[k][0]);
		const double pd[3] = {(double) p0, (double) p1, (double) p2};
		double sq[3];
		if (third_order_roots(pd, sq) != 0) { return -1; }
		const double s1 = sqrt(sq[2]), s2 = sqrt(sq[1]), pw = sqrt(sq[0]);
		const double lam[9] = {-s1, s1, -s2, s2, -pw, pw, 0, 0, 0};
		memcpy(Ls, lam, sizeof lam);
		ldbl e[2][9];
#define GCMO_COL(col, v) for (int a = 0; a < 9; a++) { U1s[a * 9 + (col)] = (double) (v)[a]; }
#define GCMO_ROW(row, v) for (int a = 0; a < 9; a++) { Us[(row) * 9 + a] = (double) (v)[a]; }
		const int single_from = (sq[1] != sq[2]) ? 0 : 4;
		for (int n = single_from; n < 6; n++) {
			eigen_of(1, Ls[n], A, s, 1, e); GCMO_COL(n, e[0]);
			eigen_of(0, Ls[n], A, s, 1, e); GCMO_ROW(n, e[0]);
		}
		if (single_from == 4) {
			for (int n = 0; n < 2; n++) {
				eigen_of(1, Ls[n], A, s, 2, e); GCMO_COL(n, e[0]); GCMO_COL(n + 2, e[1]);
				eigen_of(0, Ls[n], A, s, 2, e); GCMO_ROW(n, e[0]); GCMO_ROW(n + 2, e[1]);
			}
		}
#undef GCMO_COL
#undef GCMO_ROW
		for (int z = 0; z < 3; z++) {
			const int zc = ZERO_COLS[s][z];
			U1s[zc * 9 + 6 + z] = 1;
			Us[(6 + z) * 9 + zc] = 1;
		}
		/* rows 6..8: Cramer's rule; determinant(M) of the long double matrix binds to the double overload
		 * (linal/determinants.hpp:56-60), the three numerators stay long double */
		const double Md[3][3] = {{A[i][0], A[j][0], A[k][0]}, {A[i][1], A[j][1], A[k][1]}, {A[i][2], A[j][2], A[k][2]}};
		const double det = Md[0][0] * (Md[1][1] * Md[2][2] - Md[1][2] * Md[2][1]) - Md[0][1] * (Md[1][0] * Md[2][2] - Md[1][2] * Md[2][0]) +
		                   Md[0][2] * (Md[1][0] * Md[2][1] - Md[1][1] * Md[2][0]);
		if (det == 0) { return -2; }
		for (int z = 0; z < 3; z++) {
			const int zc = ZERO_COLS[s][z];
			const double b[3] = {-A[zc][0], -A[zc][1], -A[zc][2]};
			const ldbl d1 = det3_ld(b[0], Md[0][1], Md[0][2], b[1], Md[1][1], Md[1][2], b[2], Md[2][1], Md[2][2]);
			const ldbl d2 = det3_ld(Md[0][0], b[0], Md[0][2], Md[1][0], b[1], Md[1][2], Md[2][0], b[2], Md[2][2]);
			const ldbl d3 = det3_ld(Md[0][0], Md[0][1], b[0], Md[1][0], Md[1][1], b[1], Md[2][0], Md[2][1], b[2]);
			Us[(6 + z) * 9 + i] = (double) (d1 / det);
			Us[(6 + z) * 9 + j] = (double) (d2 / det);
			Us[(6 + z) * 9 + k] = (double) (d3 / det);
		}
		/* U * U1 is diagonal: scale to the identity (:272-279) */
		for (int n = 0; n < 9; n++) {
			double d = Us[n * 9] * U1s[n];
			for (int a = 1; a < 9; a++) { d += Us[n * 9 + a] * U1s[a * 9 + n]; }
			if (d == 0) { return -2; }
			const double nz = sqrt(fabs(d));
			const int sg = d > 0 ? 1 : -1;
			for (int a = 0; a < 9; a++) { U1s[a * 9 + n] = U1s[a * 9 + n] / nz; }
			for (int a = 0; a < 9; a++) { Us[n * 9 + a] = (sg * Us[n * 9 + a]) / nz; }
		}
	}
	return 0;
}

/* :56-71  Newton forward interpolation; src (n vectors of m) is overwritten */
int gcmo_interpolate(int m, int n, double* src, double q, double* out) {
	for (int c = 0; c < m; c++) { out[c] = src[c]; }
	const int p = n - 1;
	for (int i = 1; i <= p; i++) {
		const double f = (q - i + 1) / i;
		for (int j = 0; j < p - i + 1; j++) {
			for (int c = 0; c < m; c++) {
				src[j * m + c] = (src[(j + 1) * m + c] - src[j * m + c]) * f;
			}
		}
		for (int c = 0; c < m; c++) { out[c] += src[c]; }
	}
	return 0;
}

/* :18-43  min-max limited interpolation */
int gcmo_minmax_interpolate(int m, int n, double* src, double q, double* out) {
	if (!(q >= 0)) { return 1; }               /* assert_ge(q, 0) */
	const size_t k = (size_t) q;
	if (k > (size_t) n - 1) { return 2; }       /* assert_le(k, src.size() - 1) */
	/* q == n-1 exactly (k == n-1) passes the reference's asserts but makes it read src[k+1] one past
	 * the end of the std::vector (undefined behaviour; its own TestGridCharacteristicMethod.cpp:15-70
	 * relies on the result being src[k]).  The only defined reading is a bracket made of src[k] alone. */
	const size_t k1 = (k + 1 > (size_t) n - 1) ? k : k + 1;
	double maximum[GCMO_MAX_M], minimum[GCMO_MAX_M];
	for (int c = 0; c < m; c++) {
		maximum[c] = fmax(src[k * m + c], src[k1 * m + c]);
		minimum[c] = fmin(src[k * m + c], src[k1 * m + c]);
	}
	gcmo_interpolate(m, n, src, q, out);
	for (int c = 0; c < m; c++) {
		if (out[c] > maximum[c]) { out[c] = maximum[c]; }
		else if (out[c] < minimum[c]) { out[c] = minimum[c]; }
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* grid: grid/cubic/CubicGrid.hpp:141-147,204-226                                             */
/* ------------------------------------------------------------------------------------------ */

static void index_maker(int D, const int* sizes, int bs, size_t* im) {
	if (D == 1) { im[0] = 1; }
	else if (D == 2) { im[0] = (size_t) (2 * bs + sizes[1]); im[1] = 1; }
	else {
		im[0] = (size_t) (2 * bs + sizes[1]) * (size_t) (2 * bs + sizes[2]);
		im[1] = (size_t) (2 * bs + sizes[2]);
		im[2] = 1;
	}
}

size_t gcmo_all_nodes(int D, const int* sizes, int bs) {
	size_t im[3];
	index_maker(D, sizes, bs, im);
	return im[0] * (size_t) (2 * bs + sizes[0]);
}

size_t gcmo_index(int D, const int* sizes, int bs, const int* it) {
	size_t im[3], ans = 0;
	index_maker(D, sizes, bs, im);
	for (int i = 0; i < D; i++) { ans += im[i] * (size_t) (it[i] + bs); }
	return ans;
}

/* advance a SlowXFastZ multi-index inside [lo, hi) ; returns 0 at the end */
static int next_index(int D, const int* lo, const int* hi, int* it) {
	for (int i = D - 1; i >= 0; i--) {
		if (++it[i] < hi[i]) { return 1; }
		it[i] = lo[i];
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* stage: engine/cubic/GridCharacteristicMethod.hpp:42-87, util/math/GridCharacteristicMethod.hpp:10-17,
 * linal/functions.hpp:254-267 (diagonalMultiply), linal/operators.hpp:109-123 (operator*)     */
/* ------------------------------------------------------------------------------------------ */
int gcmo_stage(int D, int M, const int* sizes, int bs, const double* h, int s, double tau,
		int n_tables, const double* U, const double* U1, const double* L,
		const uint8_t* node_table, const double* cur, double* next) {
	(void) n_tables;
	size_t im[3];
	index_maker(D, sizes, bs, im);
	const int lo[3] = {0, 0, 0};
	int it[3] = {0, 0, 0};
	double* src = (double*) malloc((size_t) (bs + 1) * (size_t) M * sizeof(double));
	double values[GCMO_MAX_M * GCMO_MAX_M]; /* values[j*M + k] = component j interpolated at foot k */
	int rc = 0;
	do {
		const size_t idx = gcmo_index(D, sizes, bs, it);
		const int t = node_table ? node_table[idx] : 0;
		const double* u = U + ((size_t) t * D + s) * M * M;
		const double* u1 = U1 + ((size_t) t * D + s) * M * M;
		const double* l = L + ((size_t) t * D + s) * M;
		for (int k = 0; k < M; k++) {
			const double dx = -tau * l[k];                 /* crossingPoints :56-59 */
			const long shift = (dx > 0) ? 1 : -1;          /* :79 */
			for (int i = 0; i <= bs; i++) {
				const size_t nb = (size_t) ((long) idx + shift * i * (long) im[s]);
				memcpy(src + (size_t) i * M, cur + nb * M, (size_t) M * sizeof(double));
			}
			double col[GCMO_MAX_M];
			const int e = gcmo_minmax_interpolate(M, bs + 1, src, fabs(dx) / h[s], col);
			if (e) { rc = e; }
			for (int j = 0; j < M; j++) { values[j * M + k] = col[j]; }
		}
		double r[GCMO_MAX_M];
		for (int i = 0; i < M; i++) {                      /* diagonalMultiply(U, values) */
			r[i] = u[i * M + 0] * values[0 * M + i];
			for (int j = 1; j < M; j++) { r[i] += u[i * M + j] * values[j * M + i]; }
		}
		double* out = next + idx * M;
		for (int i = 0; i < M; i++) {                      /* U1 * r */
			double acc = u1[i * M + 0] * r[0];
			for (int j = 1; j < M; j++) { acc += u1[i * M + j] * r[j]; }
			out[i] = acc;
		}
	} while (next_index(D, lo, sizes, it));
	free(src);
	return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* quantities: rheology/variables/VelocitySigmaVariables.hpp:100-111, GetSetter maps in *.cpp   */
/* ------------------------------------------------------------------------------------------ */
double gcmo_get_quantity(int D, int M, int code, const double* node) {
	(void) M;
	if (code >= 0) { return node[code]; }
	double trace = 0;
	for (int i = 0; i < D; i++) { trace += node[D + sym_index(D, i, i)]; }
	return -trace / D;
}

void gcmo_set_quantity(int D, int M, int code, double value, double* node) {
	if (code >= 0) { node[code] = value; return; }
	for (int i = 0; i < M; i++) { node[i] = 0; }  /* setPressure clears the whole vector first */
	for (int i = 0; i < D; i++) { node[D + sym_index(D, i, i)] = -value; }
}

/* ------------------------------------------------------------------------------------------ */
/* border ghost fill: engine/cubic/BorderConditions.hpp:81-114                                 */
/* ------------------------------------------------------------------------------------------ */
void gcmo_border_apply(int D, int M, const int* sizes, int bs, int dir,
		const uint8_t* left_mask, const uint8_t* right_mask,
		int nq, const int* q, const double* val, double* pde) {
	for (int side = 0; side < 2; side++) {
		const uint8_t* mask = side == 0 ? left_mask : right_mask;
		if (!mask) { continue; }
		const int inner_sign = side == 0 ? 1 : -1;
		int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, it[3];
		for (int i = 0; i < D; i++) { hi[i] = sizes[i]; }
		lo[dir] = side == 0 ? 0 : sizes[dir] - 1;
		hi[dir] = lo[dir] + 1;
		for (int i = 0; i < 3; i++) { it[i] = lo[i]; }
		size_t n = 0;
		do {
			if (mask[n]) {
				for (int a = 1; a <= bs; a++) {
					int inner[3] = {it[0], it[1], it[2]}, ghost[3] = {it[0], it[1], it[2]};
					inner[dir] += inner_sign * a;
					ghost[dir] -= inner_sign * a;
					double* g = pde + gcmo_index(D, sizes, bs, ghost) * M;
					const double* in = pde + gcmo_index(D, sizes, bs, inner) * M;
					memcpy(g, in, (size_t) M * sizeof(double));
					for (int j = 0; j < nq; j++) {
						const double innerValue = gcmo_get_quantity(D, M, q[j], in);
						const double ghostValue = -innerValue + 2 * val[j];
						gcmo_set_quantity(D, M, q[j], ghostValue, g);
					}
				}
			}
			n++;
		} while (next_index(D, lo, hi, it));
	}
}

/* ------------------------------------------------------------------------------------------ */
/* contact ghost copy: engine/cubic/ContactConditions.hpp:56-68                                */
/* ------------------------------------------------------------------------------------------ */
void gcmo_contact_copy(int D, int M, const int* sizesA, const int* sizesB, int bs,
		const int* boxA_min, const int* boxB_min, const int* extent,
		double* pdeA, const double* pdeB) {
	int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, it[3] = {0, 0, 0};
	for (int i = 0; i < D; i++) { hi[i] = extent[i]; }
	do {
		int a[3], b[3];
		for (int i = 0; i < D; i++) { a[i] = boxA_min[i] + it[i]; b[i] = boxB_min[i] + it[i]; }
		memcpy(pdeA + gcmo_index(D, sizesA, bs, a) * M,
		       pdeB + gcmo_index(D, sizesB, bs, b) * M, (size_t) M * sizeof(double));
	} while (next_index(D, lo, hi, it));
}

/* ------------------------------------------------------------------------------------------ */
/* Maxwell viscosity: rheology/ode/Ode.hpp:28-38 (decay = exp(-tau/tau0) evaluated by the caller) */
/* ------------------------------------------------------------------------------------------ */
void gcmo_ode_maxwell(int D, int M, int model, const int* sizes, int bs,
		const double* decay_per_table, const uint8_t* node_table, double* pde) {
	(void) model;
	const int lo[3] = {0, 0, 0};
	int it[3] = {0, 0, 0};
	do {
		const size_t idx = gcmo_index(D, sizes, bs, it);
		const double f = decay_per_table[node_table ? node_table[idx] : 0];
		/* elastic: all sigma components; acoustic: getSigma() is the pressure (component D) */
		for (int c = D; c < M; c++) { pde[idx * M + c] = pde[idx * M + c] * f; }
	} while (next_index(D, lo, sizes, it));
}
