// TEST INFRASTRUCTURE ONLY.  Shadows <libgcm/util/math/GslUtils.hpp> (same include guard) because GSL is absent here.
// determinant / solveLinearSystem / invert for N > 3 restate what the reference gets from GSL
// (util/math/GslUtils.hpp:70-150 -> gsl_linalg_LU_decomp, _LU_det, _LU_solve, _LU_invert): Gaussian elimination with
// partial pivoting in GSL's loop order, the determinant as signum * product of the diagonal, the solve as permute +
// forward substitution (unit lower) + back substitution.  Only the simplex contact corrector's "both families outer"
// sub-case reaches them (engine/simplex/ContactCorrector.hpp:180-222: a 6x6 system for elastic bodies).
// solveThirdOrderPolynomial (rotated orthotropic materials, rheology/models/ElasticModel3D.cpp:208) keeps the
// reference's own root-sorting wrapper around a restatement of gsl_poly_solve_cubic.
#ifndef LIBGCM_GSLUTILS_HPP
#define LIBGCM_GSLUTILS_HPP
#include <cmath>
#include <libgcm/util/infrastructure/infrastructure.hpp>
#include <libgcm/linal/Matrix.hpp>
namespace gcm {
namespace gsl_utils {
static constexpr real eps = 1e-2;

template<int N>
struct LuStandIn {
	double a[N][N];
	int perm[N];
	double signum = 1;
	template<typename TMatrix>
	explicit LuStandIn(const TMatrix& m) {
		for (int i = 0; i < N; i++) { perm[i] = i; for (int j = 0; j < N; j++) { a[i][j] = m(i, j); } }
		for (int j = 0; j < N - 1; j++) {
			double best = std::fabs(a[j][j]);
			int pivot = j;
			for (int i = j + 1; i < N; i++) { if (std::fabs(a[i][j]) > best) { best = std::fabs(a[i][j]); pivot = i; } }
			if (pivot != j) {
				for (int k = 0; k < N; k++) { std::swap(a[j][k], a[pivot][k]); }
				std::swap(perm[j], perm[pivot]);
				signum = -signum;
			}
			const double ajj = a[j][j];
			if (ajj != 0.0) {
				for (int i = j + 1; i < N; i++) {
					const double aij = a[i][j] / ajj;
					a[i][j] = aij;
					for (int k = j + 1; k < N; k++) { a[i][k] = a[i][k] - aij * a[j][k]; }
				}
			}
		}
	}
	double det() const {
		double d = signum;
		for (int i = 0; i < N; i++) { d = d * a[i][i]; }
		return d;
	}
	void solve(const double* b, double* x) const {
		for (int i = 0; i < N; i++) { x[i] = b[perm[i]]; }
		for (int i = 0; i < N; i++) { double t = x[i]; for (int j = 0; j < i; j++) { t -= a[i][j] * x[j]; } x[i] = t; }
		for (int i = N - 1; i >= 0; i--) { double t = x[i]; for (int j = i + 1; j < N; j++) { t -= a[i][j] * x[j]; } x[i] = t / a[i][i]; }
	}
};

template<int TM, template<int, typename> class C>
linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>
invert(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>& m) {
	const LuStandIn<TM> lu(m);
	linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C> ans;
	for (int col = 0; col < TM; col++) {   // gsl_linalg_LU_invert: solve for every column of the identity
		double e[TM], x[TM];
		for (int i = 0; i < TM; i++) { e[i] = i == col ? 1.0 : 0.0; }
		lu.solve(e, x);
		for (int i = 0; i < TM; i++) { ans(i, col) = x[i]; }
	}
	return ans;
}
template<int TM, template<int, typename> class C>
real determinant(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>& m) {
	return LuStandIn<TM>(m).det();
}
template<int TM, template<int, typename> class C>
linal::MatrixBase<TM, 1, real, linal::NonSymmetric, C>
solveLinearSystem(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>& A,
		const linal::MatrixBase<TM, 1, real, linal::NonSymmetric, C>& b) {
	const LuStandIn<TM> lu(A);
	double rhs[TM], x[TM];
	for (int i = 0; i < TM; i++) { rhs[i] = b(i); }
	lu.solve(rhs, x);
	linal::MatrixBase<TM, 1, real, linal::NonSymmetric, C> ans;
	for (int i = 0; i < TM; i++) { ans(i) = x[i]; }
	return ans;
}
/// gsl_poly_solve_cubic (GSL poly/solve_cubic.c) restated: real roots of x^3 + a x^2 + b x + c, ascending
inline int polySolveCubicStandIn(double a, double b, double c, double* x0, double* x1, double* x2) {
	const double q = (a * a - 3 * b);
	const double r = (2 * a * a * a - 9 * a * b + 27 * c);
	const double Q = q / 9;
	const double R = r / 54;
	const double Q3 = Q * Q * Q;
	const double R2 = R * R;
	const double CR2 = 729 * r * r;
	const double CQ3 = 2916 * q * q * q;
	if (R == 0 && Q == 0) {
		*x0 = -a / 3; *x1 = -a / 3; *x2 = -a / 3;
		return 3;
	} else if (CR2 == CQ3) {
		const double sqrtQ = std::sqrt(Q);
		if (R > 0) { *x0 = -2 * sqrtQ - a / 3; *x1 = sqrtQ - a / 3; *x2 = sqrtQ - a / 3; }
		else { *x0 = -sqrtQ - a / 3; *x1 = -sqrtQ - a / 3; *x2 = 2 * sqrtQ - a / 3; }
		return 3;
	} else if (R2 < Q3) {
		const double sgnR = (R >= 0 ? 1 : -1);
		const double ratio = sgnR * std::sqrt(R2 / Q3);
		const double theta = std::acos(ratio);
		const double norm = -2 * std::sqrt(Q);
		*x0 = norm * std::cos(theta / 3) - a / 3;
		*x1 = norm * std::cos((theta + 2.0 * M_PI) / 3) - a / 3;
		*x2 = norm * std::cos((theta - 2.0 * M_PI) / 3) - a / 3;
		if (*x0 > *x1) { std::swap(*x0, *x1); }
		if (*x1 > *x2) {
			std::swap(*x1, *x2);
			if (*x0 > *x1) { std::swap(*x0, *x1); }
		}
		return 3;
	}
	const double sgnR = (R >= 0 ? 1 : -1);
	const double A = -sgnR * std::pow(std::fabs(R) + std::sqrt(R2 - Q3), 1.0 / 3.0);
	const double B = Q / A;
	*x0 = A + B - a / 3;
	return 1;
}

/// the reference's own wrapper (util/math/GslUtils.hpp:163-204) around the cubic solver: two (nearly) equal roots go
/// to the end.  The complex-root fallback (gsl_poly_complex_solve_cubic) is not provided.
inline linal::Vector<3> solveThirdOrderPolynomial(const linal::Vector<3> p) {
	double x1 = 0, x2 = 0, x3 = 0;
	const int numberOfRoots = polySolveCubicStandIn(p(0), p(1), p(2), &x1, &x2, &x3);
	if (numberOfRoots != 3) { THROW_UNSUPPORTED("complex roots: gsl_poly_complex_solve_cubic is not available in the oracle build"); }
	if (std::fabs(x1 - x2) < std::fmax(std::fabs(x1), std::fabs(x2)) * eps) {
		if (std::fabs(x3 - x2) < std::fmax(std::fabs(x3), std::fabs(x2)) * eps) {
			x1 = x2 = x3 = (x1 + x2 + x3) / 3;
		} else {
			x2 = (x1 + x2) / 2;
			x1 = x3;
			x3 = x2;
		}
	} else if (std::fabs(x1 - x3) < std::fmax(std::fabs(x1), std::fabs(x3)) * eps) {
		x3 = (x1 + x3) / 2;
		x1 = x2;
		x2 = x3;
	} else if (std::fabs(x2 - x3) < std::fmax(std::fabs(x2), std::fabs(x3)) * eps) {
		x2 = x3 = (x2 + x3) / 2;
	}
	return {x1, x2, x3};
}
}
}
#endif
