// TEST INFRASTRUCTURE ONLY.  Shadows <libgcm/util/math/GslUtils.hpp> (same include guard) because GSL is absent here.
// determinant / solveLinearSystem / invert for N > 3 restate what the reference gets from GSL
// (util/math/GslUtils.hpp:70-150 -> gsl_linalg_LU_decomp, _LU_det, _LU_solve, _LU_invert): Gaussian elimination with
// partial pivoting in GSL's loop order, the determinant as signum * product of the diagonal, the solve as permute +
// forward substitution (unit lower) + back substitution.  Only the simplex contact corrector's "both families outer"
// sub-case reaches them (engine/simplex/ContactCorrector.hpp:180-222: a 6x6 system for elastic bodies).
// solveThirdOrderPolynomial (rotated orthotropic materials) is not provided.
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
inline linal::Vector<3> solveThirdOrderPolynomial(const linal::Vector<3>) {
	THROW_UNSUPPORTED("GSL is not available in the oracle build");
}
}
}
#endif
