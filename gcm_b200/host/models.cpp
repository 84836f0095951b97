// Eigen-systems of the rheology models, materials and areas of the gcm_b200 host layer.
//
// Closed forms follow the reference so that the tables uploaded to the GPU are the ones its CPU
// engine would use (bit for bit, checked in tests/ against matrices dumped from the reference):
//   isotropic elastic   rheology/models/ElasticModel.hpp:362-553 (any orthonormal basis)
//   orthotropic elastic rheology/models/ElasticModel3D.cpp:288-429, ElasticModel2D.cpp:8-76 (material
//                       axes == coordinate axes; rotated materials are not supported yet)
//   acoustic            rheology/models/AcousticModel.hpp:214-317
#include <cmath>
#include <cstring>

#include "gcmb_host.hpp"

namespace gcmb {

// ---- areas ------------------------------------------------------------------------------------
AxisAlignedBoxArea::AxisAlignedBoxArea(const Real3& min_, const Real3& max_) : min(min_), max(max_) {
	for (int i = 0; i < 3; i++) {
		if (!(max[i] - min[i] > 0)) { throw Exception(GCMB_E_INVALID_ARG, "AxisAlignedBoxArea: max must exceed min"); }
	}
}
bool AxisAlignedBoxArea::contains(const Real3& c) const {
	for (int i = 0; i < 3; i++) {
		if (c[i] <= min[i] || c[i] >= max[i]) { return false; }
	}
	return true;
}
void AxisAlignedBoxArea::move(const Real3& s) { for (int i = 0; i < 3; i++) { min[i] += s[i]; max[i] += s[i]; } }
std::vector<double> AxisAlignedBoxArea::deviceParams() const {
	return {min[0], min[1], min[2], max[0], max[1], max[2]};
}

SphereArea::SphereArea(const real& radius_, const Real3& center_) : radius(radius_), center(center_) {
	if (!(radius > 0)) { throw Exception(GCMB_E_INVALID_ARG, "SphereArea: radius must be positive"); }
}
bool SphereArea::contains(const Real3& c) const {
	const real dx = c[0] - center[0], dy = c[1] - center[1], dz = c[2] - center[2];
	return std::sqrt(dx * dx + dy * dy + dz * dz) < radius;
}
void SphereArea::move(const Real3& s) { for (int i = 0; i < 3; i++) { center[i] += s[i]; } }
std::vector<double> SphereArea::deviceParams() const { return {radius, center[0], center[1], center[2]}; }

StraightBoundedCylinderArea::StraightBoundedCylinderArea(const real& radius_, const Real3& begin_,
		const Real3& end_) : radius(radius_), begin(begin_), end(end_) {
	if (!(radius > 0)) { throw Exception(GCMB_E_INVALID_ARG, "cylinder: radius must be positive"); }
	Real3 d = {{end[0] - begin[0], end[1] - begin[1], end[2] - begin[2]}};
	const real len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	for (int i = 0; i < 3; i++) { axis[i] = d[i] / len; }
}
bool StraightBoundedCylinderArea::contains(const Real3& c) const {
	const real pb[3] = {c[0] - begin[0], c[1] - begin[1], c[2] - begin[2]};
	const real pe[3] = {c[0] - end[0], c[1] - end[1], c[2] - end[2]};
	const real d1 = pb[0] * axis[0] + pb[1] * axis[1] + pb[2] * axis[2];
	const real d2 = pe[0] * axis[0] + pe[1] * axis[1] + pe[2] * axis[2];
	if (d1 * d2 >= 0) { return false; }
	return (pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]) - d1 * d1 < radius * radius;
}
void StraightBoundedCylinderArea::move(const Real3& s) {
	for (int i = 0; i < 3; i++) { begin[i] += s[i]; end[i] += s[i]; }
}
std::vector<double> StraightBoundedCylinderArea::deviceParams() const {
	return {radius, begin[0], begin[1], begin[2], end[0], end[1], end[2], axis[0], axis[1], axis[2]};
}

// ---- materials ----------------------------------------------------------------------------------
IsotropicMaterial::IsotropicMaterial(real rho_, real lambda_, real mu_, real yieldStrength_,
		real continualDamageParameter_, int materialNumber_, real tau0_) :
		rho(rho_), lambda(lambda_), mu(mu_), yieldStrength(yieldStrength_),
		continualDamageParameter(continualDamageParameter_), tau0(tau0_) {
	materialNumber = materialNumber_;
}

OrthotropicMaterial::OrthotropicMaterial(real rho_, std::initializer_list<real> c_, real yieldStrength_,
		real continualDamageParameter_, Real3 phi, real tau0_) :
		rho(rho_), yieldStrength(yieldStrength_), continualDamageParameter(continualDamageParameter_),
		tau0(tau0_), anglesOfRotation(phi) {
	if (c_.size() != 9) { throw Exception(GCMB_E_INVALID_ARG, "OrthotropicMaterial needs 9 elastic constants"); }
	int i = 0;
	for (real v : c_) { c[i++] = v; }
}

OrthotropicMaterial::OrthotropicMaterial(const IsotropicMaterial& iso) {
	rho = iso.rho;
	yieldStrength = iso.yieldStrength;
	continualDamageParameter = iso.continualDamageParameter;
	tau0 = iso.tau0;
	const real p = iso.lambda + 2 * iso.mu;
	c[0] = c[3] = c[5] = p;              // c11 c22 c33
	c[1] = c[2] = c[4] = iso.lambda;     // c12 c13 c23
	c[6] = c[7] = c[8] = iso.mu;         // c44 c55 c66
	materialNumber = iso.materialNumber;
}

// ---- PDE vector layout ----------------------------------------------------------------------------
int pdeSize(Models::T model, int D) {
	return model == Models::T::ELASTIC ? D + (D * (D + 1)) / 2 : D + 1;
}

/// position of sigma_ij in the PDE vector (reference linal/Symmetry.hpp:41-47 packing, after D velocities)
static int sigmaIndex(int D, int i, int j) {
	if (i > j) { std::swap(i, j); }
	return D + i * D - ((i - 1) * i) / 2 + j - i;
}

int sigmaComponent(int D, int i, int j) { return sigmaIndex(D, i, j); }

int quantityCode(Models::T model, int D, PhysicalQuantities::T q) {
	typedef PhysicalQuantities::T Q;
	auto bad = [] { throw Exception(GCMB_E_INVALID_ARG, "quantity is not a variable of this model"); return 0; };
	switch (q) {
		case Q::Vx: return D > 0 ? 0 : bad();
		case Q::Vy: return D > 1 ? 1 : bad();
		case Q::Vz: return D > 2 ? 2 : bad();
		case Q::PRESSURE: return model == Models::T::ELASTIC ? GCMB_Q_PRESSURE_TRACE : D;
		default: break;
	}
	if (model != Models::T::ELASTIC) { return bad(); }
	static const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
	const int n = (int) q - (int) Q::Sxx;
	if (n < 0 || n > 5 || ij[n][0] >= D || ij[n][1] >= D) { return bad(); }
	return sigmaIndex(D, ij[n][0], ij[n][1]);
}

int waveColumn(Models::T model, Materials::T material, int D, Waves::T wave) {
	const int w = (int) wave;  // P_F, P_B, S1_F, S1_B, S2_F, S2_B
	if (model != Models::T::ELASTIC || material == Materials::T::ISOTROPIC) {
		const int limit = model == Models::T::ELASTIC ? 2 * D : 2;
		if (w >= limit) { throw Exception(GCMB_E_INVALID_ARG, "no such wave in this model"); }
		return w;
	}
	// orthotropic tables are sorted shear-first with the backward wave first (Model.cpp:22-27,39-46)
	static const int col3[6] = {5, 4, 1, 0, 3, 2};
	static const int col2[4] = {3, 2, 1, 0};
	if (D == 3) { return col3[w]; }
	if (D == 2 && w < 4) { return col2[w]; }
	throw Exception(GCMB_E_INVALID_ARG, "no such wave in this model");
}

// ---- small dense helpers for the isotropic closed form ---------------------------------------------
namespace {

struct Tensor {  // D x D, full storage, used for symmetric products
	real a[3][3];
};

Tensor symmProduct(int D, const real* v, const real* w) {
	Tensor t;
	std::memset(&t, 0, sizeof t);
	for (int i = 0; i < D; i++) for (int j = 0; j <= i; j++) {
		t.a[i][j] = (v[i] * w[j] + w[i] * v[j]) / 2;
		t.a[j][i] = t.a[i][j];
	}
	return t;
}

/// stress as the PDE vector stores it, from the tensor used in dot products (ElasticModel.hpp:157-164)
Tensor tensorToVector(int D, const Tensor& s) {
	Tensor t = s;
	for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) {
		t.a[i][j] = s.a[i][j] * 2 - (i == j ? s.a[i][j] : 0.0);
	}
	return t;
}

struct Column {  // one PDE vector being assembled
	int D, M;
	real v[GCMB_MAX_M];
	Column(int D_, int M_) : D(D_), M(M_) { std::memset(v, 0, sizeof v); }
	void velocity(const real* n, real scale) { for (int i = 0; i < D; i++) { v[i] = n[i] * scale; } }
	void velocity(const real* n) { for (int i = 0; i < D; i++) { v[i] = n[i]; } }
	void noVelocity() { for (int i = 0; i < D; i++) { v[i] = 0; } }
	void sigma(const Tensor& t) {
		for (int i = 0; i < D; i++) for (int j = 0; j <= i; j++) { v[sigmaIndex(D, i, j)] = t.a[i][j]; }
	}
	void flipSigma() { for (int i = D; i < M; i++) { v[i] = -v[i]; } }
	void intoColumn(real* A, int c) const { for (int i = 0; i < M; i++) { A[i * M + c] = v[i]; } }
	void intoRow(real* A, int r) const { for (int j = 0; j < M; j++) { A[r * M + j] = v[j]; } }
};

/// orthonormal basis whose last column is n (reference linal/basis.hpp:49-66, geometry.hpp:35-52)
void localBasis(int D, const real* n, real (&b)[3][3]) {
	std::memset(b, 0, sizeof b);
	if (D == 1) { b[0][0] = n[0]; return; }
	if (D == 2) {
		b[0][0] = n[1]; b[0][1] = n[0];
		b[1][0] = -n[0]; b[1][1] = n[1];
		return;
	}
	real p[3] = {n[1], -n[0], 0};
	if (n[0] == 0 && n[1] == 0) { p[0] = n[2]; p[1] = 0; p[2] = 0; }
	const real ln = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
	const real lp = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
	real t1[3], t2[3];
	for (int i = 0; i < 3; i++) { t1[i] = p[i] * ln / lp; }
	t2[0] = n[1] * t1[2] - n[2] * t1[1];
	t2[1] = n[2] * t1[0] - n[0] * t1[2];
	t2[2] = n[0] * t1[1] - n[1] * t1[0];
	for (int i = 0; i < 3; i++) { b[i][0] = t1[i]; b[i][1] = t2[i]; b[i][2] = n[i]; }
}

void isotropicElasticDirection(int D, const IsotropicMaterial& m, const real (&basis)[3][3],
		real* U, real* U1, real* L) {
	const int M = pdeSize(Models::T::ELASTIC, D);
	const real rho = m.rho, lambda = m.lambda, mu = m.mu;
	const real c1 = std::sqrt((lambda + 2 * mu) / rho);
	const real c2 = std::sqrt(mu / rho);
	const real alpha = 0.5;
	// n[0]: propagation direction, n[1..]: polarisations of the shear waves
	real n[3][3];
	for (int i = 0; i < D; i++) for (int a = 0; a < D; a++) { n[i][a] = basis[a][(i + D - 1) % D]; }
	Tensor N[3][3];
	for (int i = 0; i < D; i++) for (int j = 0; j <= i; j++) { N[i][j] = N[j][i] = symmProduct(D, n[i], n[j]); }

	L[0] = c1; L[1] = -c1;
	for (int i = 1; i < D; i++) { L[2 * i] = c2; L[2 * i + 1] = -c2; }

	Tensor s;
	Column col(D, M);
	// right eigenvectors -> columns of U1
	col.velocity(n[0], alpha);
	{
		const real f = -alpha / c1, twoMu = 2 * mu;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) {
			s.a[a][b] = ((a == b ? 1.0 : 0.0) * lambda + N[0][0].a[a][b] * twoMu) * f;
		}
	}
	col.sigma(s); col.intoColumn(U1, 0);
	col.flipSigma(); col.intoColumn(U1, 1);
	for (int i = 1; i < D; i++) {
		col.velocity(n[i], alpha);
		const real f = -2 * alpha * mu / c2;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { s.a[a][b] = N[0][i].a[a][b] * f; }
		col.sigma(s); col.intoColumn(U1, 2 * i);
		col.flipSigma(); col.intoColumn(U1, 2 * i + 1);
	}
	col.noVelocity();
	if (D == 3) {
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s.a[a][b] = N[1][2].a[a][b] * 2; }
		col.sigma(s); col.intoColumn(U1, 6);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s.a[a][b] = (N[1][1].a[a][b] - N[2][2].a[a][b]) / 2; }
		col.sigma(s); col.intoColumn(U1, 7);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s.a[a][b] = (N[1][1].a[a][b] + N[2][2].a[a][b]) / 2; }
		col.sigma(s); col.intoColumn(U1, 8);
	} else if (D == 2) {
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { s.a[a][b] = (a == b ? 1.0 : 0.0) - N[0][0].a[a][b]; }
		col.sigma(s); col.intoColumn(U1, 4);
	}

	// left eigenvectors -> rows of U
	Column row(D, M);
	row.velocity(n[0]);
	{
		const real d = -c1 * rho;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { s.a[a][b] = N[0][0].a[a][b] / d; }
	}
	row.sigma(tensorToVector(D, s)); row.intoRow(U, 0);
	row.flipSigma(); row.intoRow(U, 1);
	for (int i = 1; i < D; i++) {
		row.velocity(n[i]);
		const real d = -c2 * rho;
		for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) { s.a[a][b] = N[0][i].a[a][b] / d; }
		row.sigma(tensorToVector(D, s)); row.intoRow(U, 2 * i);
		row.flipSigma(); row.intoRow(U, 2 * i + 1);
	}
	row.noVelocity();
	if (D == 3) {
		row.sigma(tensorToVector(D, N[1][2])); row.intoRow(U, 6);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { s.a[a][b] = N[1][1].a[a][b] - N[2][2].a[a][b]; }
		row.sigma(tensorToVector(D, s)); row.intoRow(U, 7);
		const real g = 2 * lambda / (lambda + 2 * mu);
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
			s.a[a][b] = (N[1][1].a[a][b] + N[2][2].a[a][b]) - N[0][0].a[a][b] * g;
		}
		row.sigma(tensorToVector(D, s)); row.intoRow(U, 8);
	} else if (D == 2) {
		const real g = lambda / (lambda + 2 * mu);
		for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { s.a[a][b] = N[1][1].a[a][b] - N[0][0].a[a][b] * g; }
		row.sigma(tensorToVector(D, s)); row.intoRow(U, 4);
	}
}

void acousticDirection(int D, const IsotropicMaterial& m, const real (&basis)[3][3], real* U, real* U1, real* L) {
	const int M = D + 1;
	const real c1 = std::sqrt(m.lambda / m.rho);
	const real alpha = 0.5;
	real n[3][3];
	for (int i = 0; i < D; i++) for (int a = 0; a < D; a++) { n[i][a] = basis[a][(i + D - 1) % D]; }
	L[0] = c1; L[1] = -c1;
	real v[GCMB_MAX_M];
	auto putCol = [&](real* A, int c) { for (int i = 0; i < M; i++) { A[i * M + c] = v[i]; } };
	auto putRow = [&](real* A, int r) { for (int j = 0; j < M; j++) { A[r * M + j] = v[j]; } };
	for (int a = 0; a < D; a++) { v[a] = n[0][a]; }
	v[D] = c1 * m.rho; putCol(U1, 0);
	v[D] = -v[D]; putCol(U1, 1);
	v[D] = 0;
	for (int i = 1; i < D; i++) { for (int a = 0; a < D; a++) { v[a] = n[i][a]; } putCol(U1, i + 1); }
	for (int a = 0; a < D; a++) { v[a] = n[0][a] * alpha; }
	v[D] = alpha / (c1 * m.rho); putRow(U, 0);
	v[D] = -v[D]; putRow(U, 1);
	v[D] = 0;
	for (int i = 1; i < D; i++) { for (int a = 0; a < D; a++) { v[a] = n[i][a]; } putRow(U, i + 1); }
}

/// orthotropic, material axes along the coordinate axes, 3-D (ElasticModel3D.cpp:288-429)
void orthotropic3D(const OrthotropicMaterial& m, GcmMatrices& g) {
	const real rho = m.rho;
	const real c11 = m.c[0], c12 = m.c[1], c13 = m.c[2], c22 = m.c[3], c23 = m.c[4], c33 = m.c[5],
	           c44 = m.c[6], c55 = m.c[7], c66 = m.c[8];
	const real normal[3] = {c11, c22, c33};                                   // c_ss
	const real shear[3][3] = {{0, c66, c55}, {c66, 0, c44}, {c55, c44, 0}};   // modulus of sigma_ab, a != b
	const real cross[3][3] = {{0, c12, c13}, {c12, 0, c23}, {c13, c23, 0}};   // coupling of sigma_pp with e_ss
	for (int s = 0; s < 3; s++) {
		real* U = g.U.data() + s * 81;
		real* U1 = g.U1.data() + s * 81;
		real* L = g.L.data() + s * 9;
		// wave pairs: the two shear waves (ascending velocity component), then the longitudinal one
		int vel[3], k = 0;
		for (int a = 0; a < 3; a++) { if (a != s) { vel[k++] = a; } }
		vel[2] = s;
		for (int w = 0; w < 3; w++) {
			const int a = vel[w], sig = sigmaIndex(3, a, s), r = 2 * w;
			const real modulus = (a == s) ? normal[s] : shear[a][s];
			L[r] = -std::sqrt(modulus / rho);
			L[r + 1] = std::sqrt(modulus / rho);
			U[r * 9 + a] = 1.0;
			U[r * 9 + sig] = 1.0 / (std::sqrt(modulus) * std::sqrt(rho));
			U[(r + 1) * 9 + a] = 1.0;
			U[(r + 1) * 9 + sig] = -1.0 / (std::sqrt(modulus) * std::sqrt(rho));
			U1[a * 9 + r] = 0.5;
			U1[a * 9 + r + 1] = 0.5;
			U1[sig * 9 + r] = 0.5 * std::sqrt(modulus) * std::sqrt(rho);
			U1[sig * 9 + r + 1] = -0.5 * std::sqrt(modulus) * std::sqrt(rho);
		}
		// stresses that do not act on the plane normal to s: carried unchanged, the normal ones also
		// pick up a share of the longitudinal wave
		int row = 6;
		for (int comp = 3; comp < 9; comp++) {
			int p = -1, q = -1;
			for (int a = 0; a < 3; a++) for (int b = a; b < 3; b++) { if (sigmaIndex(3, a, b) == comp) { p = a; q = b; } }
			if (p == s || q == s) { continue; }
			U[row * 9 + comp] = 1.0;
			U1[comp * 9 + row] = 1;
			if (p == q) {
				const real cps = cross[p][s];
				U[row * 9 + sigmaIndex(3, s, s)] = -cps / normal[s];
				// the reference writes U1(Sxx, P) of the y-stage as (0.5*c12)/sqrt(c22/rho)
				// (ElasticModel3D.cpp:361-362) and all the others as (0.5*c*sqrt(rho))/sqrt(c_ss)
				const real w = (s == 1 && p == 0) ? (0.5 * cps) / std::sqrt(normal[s] / rho)
				                                  : (0.5 * cps * std::sqrt(rho)) / std::sqrt(normal[s]);
				U1[comp * 9 + 4] = w;
				U1[comp * 9 + 5] = -w;
			}
			row++;
		}
	}
}

/// orthotropic 2-D (ElasticModel2D.cpp:8-76); uses c11, c12, c22, c66
void orthotropic2D(const OrthotropicMaterial& m, GcmMatrices& g) {
	const real rho = m.rho, c11 = m.c[0], c12 = m.c[1], c22 = m.c[3], c66 = m.c[8];
	const real cp[2] = {std::sqrt(c11 / rho), std::sqrt(c22 / rho)};
	const real cs = std::sqrt(c66 / rho);
	for (int s = 0; s < 2; s++) {
		real* U = g.U.data() + s * 25;
		real* U1 = g.U1.data() + s * 25;
		real* L = g.L.data() + s * 5;
		const int vs = 1 - s;          // velocity component of the shear wave
		const int nn = s == 0 ? 2 : 4; // sigma_ss
		L[0] = -cs; L[1] = cs; L[2] = -cp[s]; L[3] = cp[s]; L[4] = 0;
		U[0 * 5 + vs] = 1.0; U[0 * 5 + 3] = 1.0 / (rho * cs);
		U[1 * 5 + vs] = 1.0; U[1 * 5 + 3] = -1.0 / (rho * cs);
		U[2 * 5 + s] = 1.0; U[2 * 5 + nn] = 1.0 / (rho * cp[s]);
		U[3 * 5 + s] = 1.0; U[3 * 5 + nn] = -1.0 / (rho * cp[s]);
		U1[vs * 5 + 0] = 0.5; U1[vs * 5 + 1] = 0.5;
		U1[s * 5 + 2] = 0.5; U1[s * 5 + 3] = 0.5;
		U1[3 * 5 + 0] = 0.5 * rho * cs; U1[3 * 5 + 1] = -0.5 * rho * cs;
		U1[nn * 5 + 2] = 0.5 * rho * cp[s]; U1[nn * 5 + 3] = -0.5 * rho * cp[s];
		if (s == 0) {
			// the x-stage normalises its static row differently (ElasticModel2D.cpp:33,41)
			U[4 * 5 + 2] = 1.0 / c11; U[4 * 5 + 4] = -1.0 / c12;
			U1[4 * 5 + 2] = 0.5 * c12 / cp[0]; U1[4 * 5 + 3] = -0.5 * c12 / cp[0]; U1[4 * 5 + 4] = -c12;
		} else {
			U[4 * 5 + 2] = 1.0; U[4 * 5 + 4] = -c12 / c22;
			U1[2 * 5 + 2] = 0.5 * c12 / cp[1]; U1[2 * 5 + 3] = -0.5 * c12 / cp[1]; U1[2 * 5 + 4] = 1.0;
		}
	}
}

}  // namespace

GcmMatrices constructGcmMatrices(Models::T model, int D, const IsotropicMaterial& material, const real* calcBasis) {
	if (D < 1 || D > 3) { throw Exception(GCMB_E_INVALID_ARG, "dimensionality must be 1..3"); }
	GcmMatrices g;
	g.D = D;
	g.M = pdeSize(model, D);
	g.U.assign((size_t) D * g.M * g.M, 0.0);
	g.U1.assign((size_t) D * g.M * g.M, 0.0);
	g.L.assign((size_t) D * g.M, 0.0);
	for (int s = 0; s < D; s++) {
		real n[3] = {0, 0, 0};
		for (int i = 0; i < D; i++) { n[i] = calcBasis[i * D + s]; }
		real basis[3][3];
		localBasis(D, n, basis);
		real* U = g.U.data() + (size_t) s * g.M * g.M;
		real* U1 = g.U1.data() + (size_t) s * g.M * g.M;
		real* L = g.L.data() + (size_t) s * g.M;
		if (model == Models::T::ELASTIC) { isotropicElasticDirection(D, material, basis, U, U1, L); }
		else if (model == Models::T::ACOUSTIC) { acousticDirection(D, material, basis, U, U1, L); }
		else { throw Exception(GCMB_E_UNSUPPORTED, "Unknown model type"); }
	}
	g.checkDecomposition(100 * 1e-9 * 1000);
	return g;
}

GcmMatrices constructGcmMatrices(Models::T model, int D, const AbstractMaterial& material) {
	if (D < 1 || D > 3) { throw Exception(GCMB_E_INVALID_ARG, "dimensionality must be 1..3"); }
	GcmMatrices g;
	g.D = D;
	g.M = pdeSize(model, D);
	g.U.assign((size_t) D * g.M * g.M, 0.0);
	g.U1.assign((size_t) D * g.M * g.M, 0.0);
	g.L.assign((size_t) D * g.M, 0.0);
	const auto* iso = dynamic_cast<const IsotropicMaterial*>(&material);
	const auto* ortho = dynamic_cast<const OrthotropicMaterial*>(&material);
	if (iso) {
		for (int s = 0; s < D; s++) {
			real n[3] = {0, 0, 0};
			n[s] = 1;
			real basis[3][3];
			localBasis(D, n, basis);
			real* U = g.U.data() + (size_t) s * g.M * g.M;
			real* U1 = g.U1.data() + (size_t) s * g.M * g.M;
			real* L = g.L.data() + (size_t) s * g.M;
			if (model == Models::T::ELASTIC) { isotropicElasticDirection(D, *iso, basis, U, U1, L); }
			else if (model == Models::T::ACOUSTIC) { acousticDirection(D, *iso, basis, U, U1, L); }
			else { throw Exception(GCMB_E_UNSUPPORTED, "Unknown model type"); }
		}
		g.checkDecomposition(100 * 1e-9 * 1000);
	} else if (ortho) {
		if (model != Models::T::ELASTIC) { throw Exception(GCMB_E_UNSUPPORTED, "Unknown or inappropriate model type"); }
		if (ortho->anglesOfRotation[0] != 0 || ortho->anglesOfRotation[1] != 0 || ortho->anglesOfRotation[2] != 0) {
			throw Exception(GCMB_E_UNSUPPORTED, "rotated orthotropic materials are not supported yet");
		}
		if (D == 3) { orthotropic3D(*ortho, g); }
		else if (D == 2) { orthotropic2D(*ortho, g); g.checkDecomposition(1e-9 * 1000); }
		else { throw Exception(GCMB_E_UNSUPPORTED, "OrthotropicMaterial in 1D is meaningless"); }
	} else {
		throw Exception(GCMB_E_UNSUPPORTED, "Unknown material type");
	}
	return g;
}

real GcmMatrices::getMaximalEigenvalue() const {
	real ans = 0;
	for (int s = 0; s < D; s++) {
		real a = 0;
		for (int i = 0; i < M; i++) { a = std::fmax(a, std::fabs(L[(size_t) s * M + i])); }
		ans = std::fmax(ans, a);
	}
	return ans;
}

void GcmMatrices::checkDecomposition(real eps) const {
	for (int s = 0; s < D; s++) {
		const real* a = u(s);
		const real* b = u1(s);
		for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) {
			real acc = 0;
			for (int k = 0; k < M; k++) { acc += a[i * M + k] * b[k * M + j]; }
			const real want = i == j ? 1.0 : 0.0;
			if (!(std::fabs(acc - want) <= eps * std::fmax(1.0, std::fabs(want)))) {
				throw Exception(GCMB_E_INVALID_OP, "eigen-system check failed: U * U1 != I");
			}
		}
	}
}

}  // namespace gcmb
