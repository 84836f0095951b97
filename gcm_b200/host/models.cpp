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


// ---- rotated orthotropic material, 3-D (ElasticModel3D.cpp:8-283, AbstractMaterial.hpp:27-133) ------------------
namespace rotated {

/// gsl_poly_solve_cubic (GSL poly/solve_cubic.c), which the reference calls through
/// gsl_utils::solveThirdOrderPolynomial (util/math/GslUtils.hpp:163-204): roots of x^3 + a x^2 + b x + c
int solveCubic(double a, double b, double c, double* x0, double* x1, double* x2) {
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

/// gsl_utils::solveThirdOrderPolynomial: three real roots, a pair of (nearly) equal ones moved to the end
void solveThirdOrderPolynomial(const double (&p)[3], double (&x)[3]) {
	const double eps = 1e-2;
	double x1 = 0, x2 = 0, x3 = 0;
	if (solveCubic(p[0], p[1], p[2], &x1, &x2, &x3) != 3) {
		throw Exception(GCMB_E_UNSUPPORTED, "rotated orthotropic material: the characteristic polynomial has complex roots");
	}
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
	x[0] = x1; x[1] = x2; x[2] = x3;
}

inline int sym3(int i, int j) { if (i > j) { std::swap(i, j); } return i * 3 - ((i - 1) * i) / 2 + j - i; }
inline int sym6(int i, int j) { if (i > j) { std::swap(i, j); } return i * 6 - ((i - 1) * i) / 2 + j - i; }

/// AbstractMaterial::rotate(ElasticMatrix, phi): 6x6 -> 3^4 tensor -> rotation by G = Z(phi2) Y(phi1) X(phi0) -> 6x6
/// (AbstractMaterial.hpp:27-133; symmetric storage throughout, like linal::SymmetricMatrix)
void rotatedElasticMatrix(const OrthotropicMaterial& m, double (&c)[21]) {
	double c0[21];
	for (double& x : c0) { x = 0; }
	c0[sym6(0, 0)] = m.c[0]; c0[sym6(0, 1)] = m.c[1]; c0[sym6(0, 2)] = m.c[2];
	c0[sym6(1, 1)] = m.c[3]; c0[sym6(1, 2)] = m.c[4]; c0[sym6(2, 2)] = m.c[5];
	c0[sym6(3, 3)] = m.c[6]; c0[sym6(4, 4)] = m.c[7]; c0[sym6(5, 5)] = m.c[8];
	// convert(ElasticMatrix) -> tensor t[sym3(i,j)][sym3(k,l)]
	static const int voigt[6][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}};
	double t[6][6];
	for (auto& row : t) for (double& x : row) { x = 0; }
	for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) {
		const int ij = sym3(voigt[a][0], voigt[a][1]), kl = sym3(voigt[b][0], voigt[b][1]);
		t[ij][kl] = t[kl][ij] = c0[sym6(a, b)];
	}
	const double phi = m.anglesOfRotation[0], teta = m.anglesOfRotation[1], khi = m.anglesOfRotation[2];
	const double X[9] = {1.0, 0.0, 0.0, 0.0, std::cos(phi), std::sin(phi), 0.0, -std::sin(phi), std::cos(phi)};
	const double Y[9] = {std::cos(teta), 0.0, -std::sin(teta), 0.0, 1.0, 0.0, std::sin(teta), 0.0, std::cos(teta)};
	const double Z[9] = {std::cos(khi), std::sin(khi), 0.0, -std::sin(khi), std::cos(khi), 0.0, 0.0, 0.0, 1.0};
	auto mul = [](const double* a, const double* b, double* out) {
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
			double r = a[i * 3] * b[j];
			for (int k = 1; k < 3; k++) { r += a[i * 3 + k] * b[k * 3 + j]; }
			out[i * 3 + j] = r;
		}
	};
	double zy[9], G[9];
	mul(Z, Y, zy);
	mul(zy, X, G);
	double ans[6][6];
	for (auto& row : ans) for (double& x : row) { x = 0; }
	for (int mm = 0; mm < 3; mm++) for (int n = mm; n < 3; n++) for (int p = 0; p < 3; p++) for (int q = p; q < 3; q++) {
		double& acc = ans[sym3(mm, n)][sym3(p, q)];
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
			acc += G[mm * 3 + i] * G[n * 3 + j] * G[p * 3 + k] * G[q * 3 + l] * t[sym3(i, j)][sym3(k, l)];
		}
	}
	// convert(ElasticTensor) -> 6x6 (the reference reads q(a)(b) with a <= b in Voigt order)
	for (double& x : c) { x = 0; }
	for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) {
		// the reference's table takes (min, max) pairs such that the FIRST index pair is the "earlier" Voigt entry,
		// except for the shear-shear block where it reads q(0,2)(1,2), q(0,1)(1,2), q(0,1)(0,2)
		int first = a, second = b;
		if (a >= 3 && b >= 3 && a != b) { first = b; second = a; }
		c[sym6(a, b)] = ans[sym3(voigt[first][0], voigt[first][1])][sym3(voigt[second][0], voigt[second][1])];
	}
}

/// linal::solveDegenerateLinearSystem for long double 3x3 (linal/linearSystems.hpp:169-244)
void solveDegenerate(const long double (&A)[3][3], int numberOfSolutions, long double (&x)[3], long double (&y)[3]) {
	if (numberOfSolutions == 1) {
		int I = 0, J = 1, P = 0, Q = 1;
		long double det = 0;
		for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) {
			for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
				if (std::fabs(A[p][i] * A[q][j] - A[q][i] * A[p][j]) > std::fabs(det)) {
					det = A[p][i] * A[q][j] - A[q][i] * A[p][j];
					I = i; J = j; P = p; Q = q;
				}
			}
		}
		int U = 2;
		for (int k = 0; k < 2; k++) { if (k != I && k != J) { U = k; break; } }
		x[U] = 1;
		x[I] = (-A[P][U] * A[Q][J] + A[Q][U] * A[P][J]) / det;
		x[J] = (-A[P][I] * A[Q][U] + A[Q][I] * A[P][U]) / det;
		return;
	}
	int I = 0, J = 0;
	long double det = 0;
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		if (std::fabs(det) < std::fabs(A[i][j])) { det = A[i][j]; I = i; J = j; }
	}
	int p, q;
	if (J == 0) { p = 1; q = 2; } else if (J == 1) { p = 0; q = 2; } else { p = 0; q = 1; }
	x[p] = y[q] = 1;
	x[q] = y[p] = 0;
	x[J] = -A[I][p] / det;
	y[J] = -A[I][q] / det;
}

void columnsWithRho(int stage, int& i, int& j, int& k) {
	if (stage == 0) { i = 3; j = 4; k = 5; } else if (stage == 1) { i = 4; j = 6; k = 7; } else { i = 5; j = 7; k = 8; }
}
void zeroColumns(int stage, int& i, int& j, int& k) {
	if (stage == 0) { i = 6; j = 7; k = 8; } else if (stage == 1) { i = 3; j = 5; k = 8; } else { i = 3; j = 4; k = 6; }
}

/// findEigenvectors (vectors = true) / findEigenstrings (false) of ElasticModel3D.cpp:74-148
void eigen(bool vectors, long double l, const double (&A)[9][9], int stage, int count, long double (&out)[2][9]) {
	int i, j, k, p, q, m;
	columnsWithRho(stage, i, j, k);
	zeroColumns(stage, p, q, m);
	const long double r = A[0][i];
	long double M[3][3];
	if (vectors) {
		const long double init[3][3] = {{A[i][0] - l * l / r, A[i][1], A[i][2]}, {A[j][0], A[j][1] - l * l / r, A[j][2]},
		                                {A[k][0], A[k][1], A[k][2] - l * l / r}};
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { M[a][b] = init[a][b]; }
	} else {
		const long double init[3][3] = {{A[i][0] - l * l / r, A[j][0], A[k][0]}, {A[i][1], A[j][1] - l * l / r, A[k][1]},
		                                {A[i][2], A[j][2], A[k][2] - l * l / r}};
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { M[a][b] = init[a][b]; }
	}
	long double sol[2][3] = {{0, 0, 0}, {0, 0, 0}};
	solveDegenerate(M, count, sol[0], sol[1]);
	for (int n = 0; n < count; n++) {
		long double* ev = out[n];
		for (int a = 0; a < 9; a++) { ev[a] = 0; }
		if (vectors) {
			ev[0] = sol[n][0]; ev[1] = sol[n][1]; ev[2] = sol[n][2];
			ev[i] = l / r * ev[0];
			ev[j] = l / r * ev[1];
			ev[k] = l / r * ev[2];
			ev[p] = (A[p][0] * ev[0] + A[p][1] * ev[1] + A[p][2] * ev[2]) / l;
			ev[q] = (A[q][0] * ev[0] + A[q][1] * ev[1] + A[q][2] * ev[2]) / l;
			ev[m] = (A[m][0] * ev[0] + A[m][1] * ev[1] + A[m][2] * ev[2]) / l;
		} else {
			ev[i] = sol[n][0]; ev[j] = sol[n][1]; ev[k] = sol[n][2];
			ev[0] = l / r * ev[i];
			ev[1] = l / r * ev[j];
			ev[2] = l / r * ev[k];
			ev[p] = 0; ev[q] = 0; ev[m] = 0;
		}
	}
}

long double det3l(long double a11, long double a12, long double a13, long double a21, long double a22, long double a23,
                  long double a31, long double a32, long double a33) {
	return a11 * (a22 * a33 - a23 * a32) - a12 * (a21 * a33 - a23 * a31) + a13 * (a21 * a32 - a22 * a31);
}

}  // namespace rotated

/// ElasticModel<3>::constructRotated (ElasticModel3D.cpp:151-283)
void orthotropicRotated3D(const OrthotropicMaterial& material, GcmMatrices& g) {
	using namespace rotated;
	const double rho = material.rho;
	double c21[21];
	rotatedElasticMatrix(material, c21);
	auto c = [&c21](int i, int j) { return c21[sym6(i, j)]; };
	double A3[3][9][9];
	for (auto& a : A3) for (auto& row : a) for (double& x : row) { x = 0; }
	// columns of stress rates per velocity component, rows 3..8 (ElasticModel3D.cpp:163-203)
	static const int rhoCol[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};
	static const int cIdx[3][6][3][2] = {
		{{{0, 0}, {0, 5}, {0, 4}}, {{0, 5}, {5, 5}, {4, 5}}, {{0, 4}, {4, 5}, {4, 4}}, {{0, 1}, {1, 5}, {1, 4}}, {{0, 3}, {3, 5}, {3, 4}}, {{0, 2}, {2, 5}, {2, 4}}},
		{{{0, 5}, {0, 1}, {0, 3}}, {{5, 5}, {1, 5}, {3, 5}}, {{4, 5}, {1, 4}, {3, 4}}, {{1, 5}, {1, 1}, {1, 3}}, {{3, 5}, {1, 3}, {3, 3}}, {{2, 5}, {1, 2}, {2, 3}}},
		{{{0, 4}, {0, 3}, {0, 2}}, {{4, 5}, {3, 5}, {2, 5}}, {{4, 4}, {3, 4}, {2, 4}}, {{1, 4}, {1, 3}, {1, 2}}, {{3, 4}, {3, 3}, {2, 3}}, {{2, 4}, {2, 3}, {2, 2}}}};
	for (int s = 0; s < 3; s++) {
		for (int v = 0; v < 3; v++) { A3[s][v][rhoCol[s][v]] = -1.0 / rho; }
		for (int row = 0; row < 6; row++) for (int v = 0; v < 3; v++) { A3[s][3 + row][v] = -c(cIdx[s][row][v][0], cIdx[s][row][v][1]); }
	}
	for (int stage = 0; stage < 3; stage++) {
		const double (&A)[9][9] = A3[stage];
		double* U = g.U.data() + stage * 81;
		double* U1 = g.U1.data() + stage * 81;
		double* L = g.L.data() + stage * 9;
		int i, j, k;
		columnsWithRho(stage, i, j, k);
		// constructEigenvaluesPolynomial (ElasticModel3D.cpp:50-71)
		const long double r = A[0][i];
		long double pl[3];
		pl[0] = r * (-A[k][2] - A[j][1] - A[i][0]);
		pl[1] = r * r * ((A[j][1] + A[i][0]) * A[k][2] - A[j][2] * A[k][1] - A[i][2] * A[k][0] + A[i][0] * A[j][1] - A[i][1] * A[j][0]);
		pl[2] = r * r * r * ((-A[i][0] * A[j][1] + A[i][1] * A[j][0]) * A[k][2] + (A[i][0] * A[j][2] - A[i][2] * A[j][0]) * A[k][1] +
		                     (-A[i][1] * A[j][2] + A[i][2] * A[j][1]) * A[k][0]);
		const double pd[3] = {(double) pl[0], (double) pl[1], (double) pl[2]};
		double sq[3];
		solveThirdOrderPolynomial(pd, sq);
		const double s1w = (double) std::sqrt(sq[2]), s2w = (double) std::sqrt(sq[1]), pw = (double) std::sqrt(sq[0]);
		const double Ls[9] = {-s1w, s1w, -s2w, s2w, -pw, pw, 0, 0, 0};
		for (int n = 0; n < 9; n++) { L[n] = Ls[n]; }
		long double ev[2][9];
		auto setColumn = [&](int col, const long double* v) { for (int a = 0; a < 9; a++) { U1[a * 9 + col] = (double) v[a]; } };
		auto setRow = [&](int row, const long double* v) { for (int a = 0; a < 9; a++) { U[row * 9 + a] = (double) v[a]; } };
		if (sq[1] != sq[2]) {
			for (int n = 0; n < 6; n++) {
				eigen(true, L[n], A, stage, 1, ev);
				setColumn(n, ev[0]);
				eigen(false, L[n], A, stage, 1, ev);
				setRow(n, ev[0]);
			}
		} else {
			for (int n = 4; n < 6; n++) {
				eigen(true, L[n], A, stage, 1, ev);
				setColumn(n, ev[0]);
				eigen(false, L[n], A, stage, 1, ev);
				setRow(n, ev[0]);
			}
			for (int n = 0; n < 2; n++) {
				eigen(true, L[n], A, stage, 2, ev);
				setColumn(n, ev[0]);
				setColumn(n + 2, ev[1]);
				eigen(false, L[n], A, stage, 2, ev);
				setRow(n, ev[0]);
				setRow(n + 2, ev[1]);
			}
		}
		int p, q, rr;
		zeroColumns(stage, p, q, rr);
		U1[p * 9 + 6] = U1[q * 9 + 7] = U1[rr * 9 + 8] = 1;
		U[6 * 9 + p] = U[7 * 9 + q] = U[8 * 9 + rr] = 1;
		// rows 6..8 of U: solve M x = -A(row, 0..2) by Cramer's rule in long double (ElasticModel3D.cpp:253-270)
		const long double M[3][3] = {{A[i][0], A[j][0], A[k][0]}, {A[i][1], A[j][1], A[k][1]}, {A[i][2], A[j][2], A[k][2]}};
		const int zero[3] = {p, q, rr};
		for (int z = 0; z < 3; z++) {
			const double b[3] = {-A[zero[z]][0], -A[zero[z]][1], -A[zero[z]][2]};
			// determinant(A) of a long double matrix resolves to the Matrix33 (double) overload in the reference
			// (linal/determinants.hpp:56-60): the denominator is computed in double, the numerators in long double
			const double Md[3][3] = {{(double) M[0][0], (double) M[0][1], (double) M[0][2]}, {(double) M[1][0], (double) M[1][1], (double) M[1][2]},
			                         {(double) M[2][0], (double) M[2][1], (double) M[2][2]}};
			const double det = Md[0][0] * (Md[1][1] * Md[2][2] - Md[1][2] * Md[2][1]) - Md[0][1] * (Md[1][0] * Md[2][2] - Md[1][2] * Md[2][0]) +
			                   Md[0][2] * (Md[1][0] * Md[2][1] - Md[1][1] * Md[2][0]);
			if (det == 0) { throw Exception(GCMB_E_INVALID_ARG, "SLE determinant is zero"); }
			const long double d1 = det3l(b[0], M[0][1], M[0][2], b[1], M[1][1], M[1][2], b[2], M[2][1], M[2][2]);
			const long double d2 = det3l(M[0][0], b[0], M[0][2], M[1][0], b[1], M[1][2], M[2][0], b[2], M[2][2]);
			const long double d3 = det3l(M[0][0], M[0][1], b[0], M[1][0], M[1][1], b[1], M[2][0], M[2][1], b[2]);
			U[(6 + z) * 9 + i] = (double) (d1 / det);
			U[(6 + z) * 9 + j] = (double) (d2 / det);
			U[(6 + z) * 9 + k] = (double) (d3 / det);
		}
		// U*U1 is diagonal now: normalise to the identity (ElasticModel3D.cpp:272-279)
		for (int n = 0; n < 9; n++) {
			double diag = U[n * 9] * U1[n];
			for (int a = 1; a < 9; a++) { diag += U[n * 9 + a] * U1[a * 9 + n]; }
			const double normalizer = std::sqrt(std::fabs(diag));
			if (diag == 0) { throw Exception(GCMB_E_INVALID_ARG, "degenerate eigen-system of a rotated orthotropic material"); }
			const int sign = diag > 0 ? 1 : -1;
			for (int a = 0; a < 9; a++) { U1[a * 9 + n] = U1[a * 9 + n] / normalizer; }
			for (int a = 0; a < 9; a++) { U[n * 9 + a] = (sign * U[n * 9 + a]) / normalizer; }
		}
		// GcmMatrix::checkDecomposition(1e-2) (util/math/GridCharacteristicMethod.hpp:60-69): the traces within eps,
		// A*U1 ~ U1*L, U*A ~ L*U and U*U1 ~ I under Utils::approximatelyEqual (util/Utils.hpp:33-41) with
		// tolerance eps*1000 -- loose enough to accept the eigenvectors of an averaged pair of close roots,
		// tight enough to reject a sign error
		const double tol = 1e-2 * 1000;
		auto near = [tol](double f1, double f2) {
			return 4 * (f1 - f2) * (f1 - f2) / ((f1 + f2) * (f1 + f2) + tol) < tol * tol;
		};
		double trL = 0;
		for (int n = 0; n < 9; n++) { trL += L[n]; }
		bool ok = std::fabs(0.0 - trL) <= 1e-2;   // trace(A) == 0: A has no diagonal entry
		for (int a = 0; a < 9 && ok; a++) for (int b = 0; b < 9 && ok; b++) {
			double au1 = A[a][0] * U1[b], ua = U[a * 9] * A[0][b], uu1 = U[a * 9] * U1[b];
			for (int k = 1; k < 9; k++) {
				au1 += A[a][k] * U1[k * 9 + b];
				ua += U[a * 9 + k] * A[k][b];
				uu1 += U[a * 9 + k] * U1[k * 9 + b];
			}
			ok = near(au1, U1[a * 9 + b] * L[b]) && near(ua, L[a] * U[a * 9 + b]) && near(uu1, a == b ? 1.0 : 0.0);
		}
		if (!ok) { throw Exception(GCMB_E_INVALID_OP, "eigen-system check failed for a rotated orthotropic material"); }
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
		const bool isRotated = ortho->anglesOfRotation[0] != 0 || ortho->anglesOfRotation[1] != 0 || ortho->anglesOfRotation[2] != 0;
		if (isRotated && D != 3) { throw Exception(GCMB_E_UNSUPPORTED, "rotated orthotropic materials exist in 3-D only"); }
		if (D == 3 && isRotated) { orthotropicRotated3D(*ortho, g); }
		else if (D == 3) { orthotropic3D(*ortho, g); }
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
