// Shadows <libgcm/util/math/GslUtils.hpp> (same include guard) because GSL is absent here.
// Only rotated-orthotropic materials and N>3 linear solves reach these; none of the cubic
// configurations used by the oracle do, so they throw.
#ifndef LIBGCM_GSLUTILS_HPP
#define LIBGCM_GSLUTILS_HPP
#include <libgcm/util/infrastructure/infrastructure.hpp>
#include <libgcm/linal/Matrix.hpp>
namespace gcm {
namespace gsl_utils {
static constexpr real eps = 1e-2;
template<int TM, template<int, typename> class C>
linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>
invert(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>&) {
	THROW_UNSUPPORTED("GSL is not available in the oracle build");
}
template<int TM, template<int, typename> class C>
real determinant(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>&) {
	THROW_UNSUPPORTED("GSL is not available in the oracle build");
}
template<int TM, template<int, typename> class C>
linal::MatrixBase<TM, 1, real, linal::NonSymmetric, C>
solveLinearSystem(const linal::MatrixBase<TM, TM, real, linal::NonSymmetric, C>&,
		const linal::MatrixBase<TM, 1, real, linal::NonSymmetric, C>&) {
	THROW_UNSUPPORTED("GSL is not available in the oracle build");
}
inline linal::Vector<3> solveThirdOrderPolynomial(const linal::Vector<3>) {
	THROW_UNSUPPORTED("GSL is not available in the oracle build");
}
}
}
#endif
