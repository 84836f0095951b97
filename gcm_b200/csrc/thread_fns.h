// Per-thread work functions of the gcm_b200 kernels.
//
// Every function here is the complete work of ONE CUDA thread of the corresponding kernel in
// kernels.cu, written as a __host__ __device__ inline so that the host-logic tests can step through
// exactly the same index math and arithmetic on the build machine (tests/emul/, never shipped, never
// loaded by the product).  All arithmetic follows the reference's operation order and the library is
// compiled with -fmad=false, so results are bit-identical to the reference CPU engine.
#pragma once
#include "internal.cuh"

namespace gcmb {

// read-only load (LDG on the device)
GCMB_HD double ldg_f64(const double* p) {
#ifdef __CUDA_ARCH__
	return __ldg(p);
#else
	return *p;
#endif
}
#define GCMB_LDG(p) ::gcmb::ldg_f64(p)

// ---------------------------------------------------------------------------------------------
// EqualDistanceLineInterpolator::minMaxInterpolate for ONE scalar
// (reference util/math/interpolation/EqualDistanceLineInterpolator.hpp:18-71).
// s[0..BS] are the upwind values (s[0] = the node itself), F[i-1] = (q-i+1)/i, k0 = (size_t) q.
// ---------------------------------------------------------------------------------------------
template<int BS, bool RUNTIME_K0>
GCMB_HD double limited_newton(double (&s)[BS + 1], const double* __restrict__ F, int k0) {
	double a, b;
	if (RUNTIME_K0) {
		// bracket = s[k0], s[k0+1]; k0 == BS (q == BS exactly) makes the reference read one past the
		// end of its vector (undefined); the bracket then degenerates to s[BS]
		a = s[0];
		b = s[BS > 0 ? 1 : 0];
#pragma unroll
		for (int i = 1; i <= BS; i++) {
			if (i == k0) { a = s[i]; b = s[i < BS ? i + 1 : i]; }
		}
	} else {
		a = s[0];
		b = s[1];
	}
	// fmax/fmin of the reference, written as one comparison: identical values for all non-NaN inputs (the
	// C functions' NaN handling costs ~8 instructions each on sm_100, there is no fp64 min/max instruction)
	const bool a_gt_b = a > b;
	const double maximum = a_gt_b ? a : b;
	const double minimum = a_gt_b ? b : a;
	double ans = s[0];
#pragma unroll
	for (int i = 1; i <= BS; i++) {
		const double f = F[i - 1];
#pragma unroll
		for (int j = 0; j < BS - i + 1; j++) {
			s[j] = (s[j + 1] - s[j]) * f;
		}
		ans += s[0];
	}
	if (ans > maximum) {
		ans = maximum;
	} else if (ans < minimum) {
		ans = minimum;
	}
	return ans;
}

// ---------------------------------------------------------------------------------------------
// One node of one stage, sparse eigen-system known at compile time (pattern P):
//   u_new = U1 * diag(U * V),  V(:,k) = limited interpolation of the PDE vector at foot k
// (reference engine/cubic/GridCharacteristicMethod.hpp:42-87, util/math/GridCharacteristicMethod.hpp:10-17,
//  linal/functions.hpp:254-267, linal/operators.hpp:109-123).  Only the structurally non-zero
// products are formed, in the reference's index order: the dropped terms are exact zeros.
// LOAD(j, o) returns component j at offset o (in nodes) along the sweep axis.
// ---------------------------------------------------------------------------------------------
// Coefficient sources for gcm_node_sparse.
// (1) the full StageTable in global memory (one-thread-per-node kernels)
template<int M>
struct TableCoef {
	const StageTable* __restrict__ T;
	GCMB_HD double u(int k, int j) const { return GCMB_LDG(&T->U[k * M + j]); }
	GCMB_HD double u1(int i, int k) const { return GCMB_LDG(&T->U1[i * M + k]); }
	GCMB_HD const double* f(int k) const { return T->F + k * MAXBS; }
};

// (2) the structurally non-zero coefficients only, packed in the order the kernel consumes them:
//     [Newton factors of the interpolated rows][non-zeros of U, row by row][non-zeros of U1, row by row].
//     All positions are compile-time constants of the pattern, so every coefficient is one LDS with an
//     immediate offset from the node's table base in shared memory.
GCMB_HD constexpr int popcount_u(unsigned m) { int n = 0; for (; m; m &= m - 1) { n++; } return n; }
// mask of the rows/columns that store their own coefficients
template<class P>
GCMB_HD constexpr unsigned packed_base_mask() {
	unsigned m = 0;
	for (int k = 0; k < P::M; k++) { if (P::base(k) == k) { m |= 1u << k; } }
	return m;
}
template<class P, int BS>
GCMB_HD constexpr int packed_fpos(int k) {  // k must be a base row (or P::M for the total)
	int n = 0;
	for (int r = 0; r < k; r++) { if (P::sgn(r) != 0 && P::base(r) == r) { n += BS; } }
	return n;
}
template<class P, int BS>
GCMB_HD constexpr int packed_upos(int k, int j) {
	int n = packed_fpos<P, BS>(P::M);
	for (int r = 0; r < k; r++) { if (P::base(r) == r) { n += popcount_u(P::um(r)); } }
	return k < P::M ? n + popcount_u(P::um(k) & ((1u << j) - 1u)) : n;
}
template<class P, int BS>
GCMB_HD constexpr int packed_u1pos(int i, int k) {
	constexpr unsigned B = packed_base_mask<P>();
	int n = packed_upos<P, BS>(P::M, 0);
	for (int r = 0; r < i; r++) { n += popcount_u(P::u1m(r) & B); }
	return i < P::M ? n + popcount_u(P::u1m(i) & B & ((1u << k) - 1u)) : n;
}
template<class P, int BS>
struct Packed {
	static constexpr int SIZE = packed_u1pos<P, BS>(P::M, 0);
};

template<class P, int BS>
struct PackedCoef {
	const double* __restrict__ t;  // packed table of the node's material (shared memory)
	// a row/column that shares reads its base's entry (the compiler merges the two loads) and flips the
	// sign where the pattern says so: (-c) * v == -(c * v) exactly
	GCMB_HD double u(int k, int j) const {
		const double c = t[packed_upos<P, BS>(P::base(k), j)];
		return ((P::uneg(k) >> j) & 1u) ? -c : c;
	}
	GCMB_HD double u1(int i, int k) const {
		const double c = t[packed_u1pos<P, BS>(i, P::base(k))];
		return ((P::u1neg(k) >> i) & 1u) ? -c : c;
	}
	GCMB_HD const double* f(int k) const { return t + packed_fpos<P, BS>(P::base(k)); }
};

template<class P, int BS, class COEF, class LOAD>
GCMB_HD void gcm_node_sparse(const COEF coef, LOAD load, double (&out)[P::M]) {
	constexpr int M = P::M;
	double r[M];
#pragma unroll
	for (int k = 0; k < M; k++) {
		const int sg = P::sgn(k);
		const unsigned mask = P::um(k);
		double acc = 0.0;
		bool first = true;
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((mask >> j) & 1u) {
				double v;
				if (sg == 0) {
					v = load(j, 0);
				} else {
					double s[BS + 1];
#pragma unroll
					for (int i = 0; i <= BS; i++) { s[i] = load(j, sg * i); }
					v = limited_newton<BS, false>(s, coef.f(k), 0);
				}
				const double t = coef.u(k, j) * v;
				if (first) { acc = t; first = false; } else { acc += t; }
			}
		}
		r[k] = acc;
	}
#pragma unroll
	for (int i = 0; i < M; i++) {
		const unsigned mask = P::u1m(i);
		double acc = 0.0;
		bool first = true;
#pragma unroll
		for (int k = 0; k < M; k++) {
			if ((mask >> k) & 1u) {
				const double t = coef.u1(i, k) * r[k];
				if (first) { acc = t; first = false; } else { acc += t; }
			}
		}
		out[i] = acc;
	}
}

// Same, dense eigen-system and runtime border size / foot cell (any material, any Courant number
// up to bs).  A literal restatement of the reference loop.
template<int M, class LOAD>
GCMB_HD void gcm_node_dense(const StageTable* __restrict__ T, int bs, LOAD load, double (&out)[M]) {
	double r[M];
	for (int k = 0; k < M; k++) {
		const int dir = T->dir[k];
		const int k0 = T->k0[k];
		const int k1 = k0 < bs ? k0 + 1 : k0;
		const double* F = T->F + k * MAXBS;
		double acc = 0.0;
		for (int j = 0; j < M; j++) {
			double s[MAXBS + 1];
			for (int i = 0; i <= bs; i++) { s[i] = load(j, dir * i); }
			const double maximum = fmax(s[k0], s[k1]);
			const double minimum = fmin(s[k0], s[k1]);
			double ans = s[0];
			for (int i = 1; i <= bs; i++) {
				const double f = F[i - 1];
				for (int jj = 0; jj < bs - i + 1; jj++) { s[jj] = (s[jj + 1] - s[jj]) * f; }
				ans += s[0];
			}
			if (ans > maximum) { ans = maximum; } else if (ans < minimum) { ans = minimum; }
			const double t = T->U[k * M + j] * ans;
			if (j == 0) { acc = t; } else { acc += t; }
		}
		r[k] = acc;
	}
	for (int i = 0; i < M; i++) {
		double acc = T->U1[i * M] * r[0];
		for (int k = 1; k < M; k++) { acc += T->U1[i * M + k] * r[k]; }
		out[i] = acc;
	}
}

// Dense eigen-system, all feet inside the first cell (k0 == 0 everywhere: Courant number <= 1) and a compile-time
// border size: the kernel of rotated orthotropic materials and of any eigen-system no pattern covers.  The node's
// stencil is read once; per side of the node the quantities that do not depend on the foot (first differences,
// the bracket's minimum and maximum) are formed once per component and shared by every characteristic whose foot
// lies on that side.  Every value is produced by the same operations, in the same order, as gcm_node_dense.
// A foot at the node itself (q == 0: Newton factor F[0] == 0) takes the node's value, as in gcm_node_sparse.
// the limiter's clamp `if (ans > mx) ans = mx; else if (ans < mn) ans = mn;` (EqualDistanceLineInterpolator.hpp:30-35)
// as two compare-and-select pairs: the same value for every input (mn <= mx, or a NaN that fails both comparisons), and
// no divergent branch -- nvcc turns the if/else into BSSY/BRA/BSYNC plus six moves per clamp
GCMB_HD double clamp_select(double ans, double mn, double mx) {
#ifdef __CUDA_ARCH__
	asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %0, %1;\n\tselp.f64 %0, %1, %0, p;\n\tsetp.lt.f64 p, %0, %2;\n\tselp.f64 %0, %2, %0, p;\n\t}"
	    : "+d"(ans) : "d"(mx), "d"(mn));
	return ans;
#else
	if (ans > mx) { ans = mx; } else if (ans < mn) { ans = mn; }
	return ans;
#endif
}

// coefficient sources: (1) the StageTable of the node's material in global memory
template<int M>
struct DenseTableCoef {
	const StageTable* __restrict__ T;
	GCMB_HD double u(int k, int j) const { return GCMB_LDG(&T->U[k * M + j]); }
	GCMB_HD double u1(int i, int k) const { return GCMB_LDG(&T->U1[i * M + k]); }
	GCMB_HD double f(int k, int i) const { return GCMB_LDG(&T->F[k * MAXBS + i]); }
	GCMB_HD int side(int k) const { return f(k, 0) == 0.0 ? 0 : T->dir[k]; }
};
// (2) a body of ONE material: the coefficients travel as a kernel parameter, so that every one of them is a
//     constant-bank operand of the arithmetic instruction that uses it (no load, no register)
template<int M, int BS>
struct DenseParamCoef {
	double U[M * M], U1[M * M], F[M * BS];
	int sd[M];
	GCMB_HD double u(int k, int j) const { return U[k * M + j]; }
	GCMB_HD double u1(int i, int k) const { return U1[i * M + k]; }
	GCMB_HD double f(int k, int i) const { return F[k * BS + i]; }
	GCMB_HD int side(int k) const { return sd[k]; }
};

template<int M, int BS, class COEF, class LOAD>
GCMB_HD void gcm_node_dense_k0(const COEF& co, LOAD load, double (&out)[M]) {
	static_assert(BS == 1 || BS == 2, "border sizes 1 and 2");
	double c[M], r[M];
#pragma unroll
	for (int j = 0; j < M; j++) { c[j] = load(j, 0); }
	int side[M];
#pragma unroll
	for (int k = 0; k < M; k++) { side[k] = co.side(k); }
#pragma unroll
	for (int k = 0; k < M; k++) {
		if (side[k] == 0) {
			double acc = co.u(k, 0) * c[0];
#pragma unroll
			for (int j = 1; j < M; j++) { acc += co.u(k, j) * c[j]; }
			r[k] = acc;
		}
	}
	// both sides unrolled: measured 33.1 ms/step against 34.8 ms with `#pragma unroll 1` (half the code, but 24 B of
	// spills) on 2 x 512x256x512 rotated plies
#pragma unroll
	for (int sd = 1; sd >= -1; sd -= 2) {
		double d0[M], d1[BS == 2 ? M : 1], mx[M], mn[M];
#pragma unroll
		for (int j = 0; j < M; j++) {
			const double s1 = load(j, sd);
			d0[j] = s1 - c[j];
			if (BS == 2) { d1[j] = load(j, 2 * sd) - s1; }
			const bool gt = c[j] > s1;
			mx[j] = gt ? c[j] : s1;
			mn[j] = gt ? s1 : c[j];
		}
#pragma unroll
		for (int k = 0; k < M; k++) {
			if (side[k] == sd) {
				const double f0 = co.f(k, 0);
				const double f1 = BS == 2 ? co.f(k, 1) : 0.0;
				double acc = 0.0;
#pragma unroll
				for (int j = 0; j < M; j++) {
					const double a0 = d0[j] * f0;
					double ans = c[j] + a0;
					if (BS == 2) {
						const double a1 = d1[j] * f0;
						ans += (a1 - a0) * f1;
					}
					ans = clamp_select(ans, mn[j], mx[j]);
					const double t = co.u(k, j) * ans;
					if (j == 0) { acc = t; } else { acc += t; }
				}
				r[k] = acc;
			}
		}
	}
#pragma unroll
	for (int i = 0; i < M; i++) {
		double acc = co.u1(i, 0) * r[0];
#pragma unroll
		for (int k = 1; k < M; k++) { acc += co.u1(i, k) * r[k]; }
		out[i] = acc;
	}
}

// loader over the structure-of-arrays volumes: component j at `o` nodes along the sweep axis
struct SoaLoad {
	const double* __restrict__ base;  // cur + idx
	long long comp, sstride;
	GCMB_HD double operator()(int j, int o) const { return GCMB_LDG(base + j * comp + o * sstride); }
};

template<class P, int BS>
GCMB_HD void stage_thread_sparse(const StageArgs& a, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	const StageTable* T = a.tables + ((int) a.node_table[idx] * a.g.D + a.dir);
	SoaLoad ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	double out[P::M];
	gcm_node_sparse<P, BS>(TableCoef<P::M>{T}, ld, out);
#pragma unroll
	for (int c = 0; c < P::M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

template<int M>
GCMB_HD void stage_thread_dense(const StageArgs& a, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	const StageTable* T = a.tables + ((int) a.node_table[idx] * a.g.D + a.dir);
	SoaLoad ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	double out[M];
	gcm_node_dense<M>(T, a.g.bs, ld, out);
	for (int c = 0; c < M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

template<int M, int BS>
GCMB_HD void stage_thread_dense_k0(const StageArgs& a, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	const DenseTableCoef<M> co{a.tables + ((int) a.node_table[idx] * a.g.D + a.dir)};
	SoaLoad ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	double out[M];
	gcm_node_dense_k0<M, BS>(co, ld, out);
#pragma unroll
	for (int c = 0; c < M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

template<int M, int BS>
GCMB_HD void stage_thread_dense_k0_one(const StageArgs& a, const DenseParamCoef<M, BS>& co, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	SoaLoad ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	double out[M];
	gcm_node_dense_k0<M, BS>(co, ld, out);
#pragma unroll
	for (int c = 0; c < M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

// ---------------------------------------------------------------------------------------------
// quantities (reference rheology/variables/VelocitySigmaVariables.hpp:100-111 + GetSetter maps)
// ---------------------------------------------------------------------------------------------
GCMB_HD int sym_index(int D, int i, int j) {
	if (i > j) { const int t = i; i = j; j = t; }
	return i * D - ((i - 1) * i) / 2 + j - i;
}

GCMB_HD double get_quantity(int D, int code, const double* v) {
	if (code >= 0) { return v[code]; }
	double trace = 0;
	for (int i = 0; i < D; i++) { trace += v[D + sym_index(D, i, i)]; }
	return -trace / D;
}

GCMB_HD void set_quantity(int D, int M, int code, double value, double* v) {
	if (code >= 0) { v[code] = value; return; }
	for (int i = 0; i < M; i++) { v[i] = 0; }
	for (int i = 0; i < D; i++) { v[D + sym_index(D, i, i)] = -value; }
}

// ---------------------------------------------------------------------------------------------
// face enumeration: node f of the face normal to internal axis `axis` (x-slowest order)
// ---------------------------------------------------------------------------------------------
GCMB_HD void face_node(const Geom& g, int axis, long long f, int fixed, int (&it)[3]) {
	const int p = axis == 0 ? 1 : 0;
	const int q = axis == 2 ? 1 : 2;
	it[axis] = fixed;
	it[q] = (int) (f % g.n[q]);
	it[p] = (int) (f / g.n[q]);
}

// ---------------------------------------------------------------------------------------------
// border ghost fill (reference engine/cubic/BorderConditions.hpp:97-114): thread = (face node, a)
// ---------------------------------------------------------------------------------------------
struct BorderArgs {
	double* pde;
	const uint8_t* mask;  // face mask or nullptr
	Geom g;
	int axis;             // internal axis
	int side;             // 0 left (inner sign +1), 1 right (inner sign -1)
	int nq;
	int q[MAXM + 1];
	double val[MAXM + 1];
};

GCMB_HD void border_thread(const BorderArgs& b, long long f, int a /* 1..bs */) {
	if (b.mask && !b.mask[f]) { return; }
	int it[3];
	face_node(b.g, b.axis, f, b.side == 0 ? 0 : b.g.n[b.axis] - 1, it);
	const int sign = b.side == 0 ? 1 : -1;
	int in[3] = {it[0], it[1], it[2]}, gh[3] = {it[0], it[1], it[2]};
	in[b.axis] += sign * a;
	gh[b.axis] -= sign * a;
	const long long ii = b.g.index(in[0], in[1], in[2]);
	const long long gi = b.g.index(gh[0], gh[1], gh[2]);
	double inner[MAXM], ghost[MAXM];
	for (int c = 0; c < b.g.M; c++) { inner[c] = b.pde[c * b.g.comp + ii]; ghost[c] = inner[c]; }
	for (int j = 0; j < b.nq; j++) {
		const double innerValue = get_quantity(b.g.D, b.q[j], inner);
		const double ghostValue = -innerValue + 2 * b.val[j];
		set_quantity(b.g.D, b.g.M, b.q[j], ghostValue, ghost);
	}
	for (int c = 0; c < b.g.M; c++) { b.pde[c * b.g.comp + gi] = ghost[c]; }
}

// ---------------------------------------------------------------------------------------------
// contact ghost copy (reference engine/cubic/ContactConditions.hpp:56-68): thread = node of the box
// ---------------------------------------------------------------------------------------------
struct ContactArgs {
	double* a;
	const double* b;
	Geom ga, gb;
	int amin[3], bmin[3], ext[3];  // internal axes
};

GCMB_HD void contact_thread(const ContactArgs& c, long long t) {
	const int i2 = (int) (t % c.ext[2]);
	const int i1 = (int) ((t / c.ext[2]) % c.ext[1]);
	const int i0 = (int) (t / ((long long) c.ext[2] * c.ext[1]));
	const long long ia = c.ga.index(c.amin[0] + i0, c.amin[1] + i1, c.amin[2] + i2);
	const long long ib = c.gb.index(c.bmin[0] + i0, c.bmin[1] + i1, c.bmin[2] + i2);
	for (int m = 0; m < c.ga.M; m++) { c.a[m * c.ga.comp + ia] = c.b[m * c.gb.comp + ib]; }
}

// ---------------------------------------------------------------------------------------------
// Maxwell viscosity (reference rheology/ode/Ode.hpp:28-38): thread = real node
// ---------------------------------------------------------------------------------------------
GCMB_HD void ode_maxwell_thread(const Geom& g, double* pde, const uint8_t* node_table,
                                const double* decay, int i0, int i1, int i2) {
	const long long idx = g.index(i0, i1, i2);
	const double f = decay[node_table[idx]];
	for (int c = g.D; c < g.M; c++) { pde[c * g.comp + idx] = pde[c * g.comp + idx] * f; }
}

// ---------------------------------------------------------------------------------------------
// areas (reference util/math/Area.hpp:23-123) and node coordinates (grid/cubic/CubicGrid.hpp:113-125)
// ---------------------------------------------------------------------------------------------
struct AreaArgs {
	int kind;        // 0 infinite, 1 box, 2 sphere, 3 cylinder
	double p[10];    // box: min[3],max[3]; sphere: r,c[3]; cylinder: r,begin[3],end[3],axis[3]
};

GCMB_HD void node_coords(const Geom& g, int i0, int i1, int i2, double (&x)[3]) {
	const int it[3] = {i0, i1, i2};
	x[0] = x[1] = x[2] = 0;
	for (int a = g.shift; a < 3; a++) {
		x[a - g.shift] = (g.start[a] * g.h[a]) + (it[a] * g.h[a]);
	}
}

GCMB_HD bool area_contains(const AreaArgs& A, const double (&x)[3]) {
	switch (A.kind) {
		case 0: return true;
		case 1:
			for (int i = 0; i < 3; i++) {
				if (x[i] <= A.p[i] || x[i] >= A.p[3 + i]) { return false; }
			}
			return true;
		case 2: {
			const double dx = x[0] - A.p[1], dy = x[1] - A.p[2], dz = x[2] - A.p[3];
			return sqrt(dx * dx + dy * dy + dz * dz) < A.p[0];
		}
		case 3: {
			const double* b = A.p + 1;
			const double* e = A.p + 4;
			const double* ax = A.p + 7;
			const double pb[3] = {x[0] - b[0], x[1] - b[1], x[2] - b[2]};
			const double pe[3] = {x[0] - e[0], x[1] - e[1], x[2] - e[2]};
			const double d1 = pb[0] * ax[0] + pb[1] * ax[1] + pb[2] * ax[2];
			const double d2 = pe[0] * ax[0] + pe[1] * ax[1] + pe[2] * ax[2];
			if (d1 * d2 >= 0) { return false; }
			return (pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]) - d1 * d1 < A.p[0] * A.p[0];
		}
		default: return false;
	}
}

// ---------------------------------------------------------------------------------------------
// host array (reference layout) <-> device structure of arrays: thread = node of the host array
// ---------------------------------------------------------------------------------------------
struct XferArgs {
	double* soa;
	double* aos;        // device staging buffer holding the slice [x_begin, x_end) of internal axis 0
	Geom g;
	int with_ghosts;
	int x_begin, x_end; // in host-array coordinates of internal axis 0 (ghost-inclusive when with_ghosts)
};

GCMB_HD void xfer_thread(const XferArgs& x, long long t, bool to_device) {
	const Geom& g = x.g;
	const int e1 = x.with_ghosts ? g.n[1] + 2 * g.g[1] : g.n[1];
	const int e2 = x.with_ghosts ? g.n[2] + 2 * g.g[2] : g.n[2];
	const int j2 = (int) (t % e2);
	const int j1 = (int) ((t / e2) % e1);
	const int j0 = (int) (t / ((long long) e2 * e1)) + x.x_begin;
	const int o0 = x.with_ghosts ? g.g[0] : 0, o1 = x.with_ghosts ? g.g[1] : 0, o2 = x.with_ghosts ? g.g[2] : 0;
	const long long idx = g.index(j0 - o0, j1 - o1, j2 - o2);
	for (int c = 0; c < g.M; c++) {
		if (to_device) { x.soa[c * g.comp + idx] = x.aos[t * g.M + c]; }
		else { x.aos[t * g.M + c] = x.soa[c * g.comp + idx]; }
	}
}

}  // namespace gcmb

// ---------------------------------------------------------------------------------------------
// Marching variant for the two strided axes (internal axis 0 or 1): one thread owns one z and walks
// along the sweep axis over [s_begin, s_end), keeping the 2*BS+1 values of every interpolated
// component in registers, so that each value is read from HBM exactly once per stage (plus the 2*BS
// planes re-read at segment starts) and every warp access is a contiguous 256-byte row.
// Loads of the next plane are issued before the arithmetic of the current node.
// ---------------------------------------------------------------------------------------------
namespace gcmb {

template<class P>
struct PatternSets {
	// components that are interpolated (appear in a row with a non-zero eigenvalue)
	GCMB_HD static constexpr unsigned interp() {
		unsigned m = 0;
		for (int k = 0; k < P::M; k++) { if (P::sgn(k) != 0) { m |= P::um(k); } }
		return m;
	}
	// components only read at the node itself
	GCMB_HD static constexpr unsigned center() {
		unsigned m = 0;
		for (int k = 0; k < P::M; k++) { if (P::sgn(k) == 0) { m |= P::um(k); } }
		return m & ~interp();
	}
};

template<class P, int BS>
GCMB_HD void stage_thread_march(const StageArgs& a, int perp, int i2, int s_begin, int s_end) {
	constexpr int M = P::M;
	constexpr int W = 2 * BS + 1;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	const Geom& g = a.g;
	const long long sstride = g.stride(a.axis);
	const long long idx0 = a.axis == 0 ? g.index(s_begin, perp, i2) : g.index(perp, s_begin, i2);
	const double* __restrict__ cur = a.cur;

	double w[M][W];   // w[j][BS + o] = component j at (s + o); only rows in IC are live
	double nv[M];     // plane s + BS + 1 (in flight during the arithmetic)
	double cv[M];     // centre-only components at s
	double cn[M];     // ... and at s + 1 (in flight)
#pragma unroll
	for (int j = 0; j < M; j++) {
		if ((IC >> j) & 1u) {
#pragma unroll
			for (int o = 1; o < W; o++) { w[j][o] = GCMB_LDG(cur + j * g.comp + idx0 + (long long) (o - 1 - BS) * sstride); }
			nv[j] = GCMB_LDG(cur + j * g.comp + idx0 + (long long) BS * sstride);
		}
		if ((CC >> j) & 1u) { cn[j] = GCMB_LDG(cur + j * g.comp + idx0); }
	}
	int tn = a.node_table[idx0];

	for (int s = s_begin; s < s_end; s++) {
		const long long idx = idx0 + (long long) (s - s_begin) * sstride;
		const int t = tn;
		// rotate the window, take over the prefetched plane, issue the next prefetch
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((IC >> j) & 1u) {
#pragma unroll
				for (int o = 0; o < W - 1; o++) { w[j][o] = w[j][o + 1]; }
				w[j][W - 1] = nv[j];
			}
			if ((CC >> j) & 1u) { cv[j] = cn[j]; }
		}
		if (s + 1 < s_end) {
#pragma unroll
			for (int j = 0; j < M; j++) {
				if ((IC >> j) & 1u) { nv[j] = GCMB_LDG(cur + j * g.comp + idx + (long long) (BS + 1) * sstride); }
				if ((CC >> j) & 1u) { cn[j] = GCMB_LDG(cur + j * g.comp + idx + sstride); }
			}
			tn = a.node_table[idx + sstride];
		}
		const StageTable* T = a.tables + (t * g.D + a.dir);
		auto load = [&](int j, int o) -> double { return ((IC >> j) & 1u) ? w[j][BS + o] : cv[j]; };
		double out[M];
		gcm_node_sparse<P, BS>(TableCoef<M>{T}, load, out);
#pragma unroll
		for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
	}
}

}  // namespace gcmb
