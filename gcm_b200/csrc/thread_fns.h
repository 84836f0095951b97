// Per-thread work functions of the gcm_b200 kernels.
//
// Every function here is the complete work of ONE CUDA thread of the corresponding kernel in
// kernels.cu, written as a __host__ __device__ inline so that the host-logic tests can step through
// exactly the same index math and arithmetic on the build machine (tests/emul/, never shipped, never
// loaded by the product).  All arithmetic follows the reference's operation order and the library is
// compiled with -fmad=false, so results are bit-identical to the reference CPU engine.
#pragma once
#include <math.h>
#include <string.h>

#include "internal.cuh"

namespace gcmb {

// read-only load (LDG on the device)
template<class R>
GCMB_HD R ldg_real(const R* p) {
#ifdef __CUDA_ARCH__
	return __ldg(p);
#else
	return *p;
#endif
}
#define GCMB_LDG(p) ::gcmb::ldg_real(p)

// ---------------------------------------------------------------------------------------------
// EqualDistanceLineInterpolator::minMaxInterpolate for ONE scalar
// (reference util/math/interpolation/EqualDistanceLineInterpolator.hpp:18-71).
// s[0..BS] are the upwind values (s[0] = the node itself), F[i-1] = (q-i+1)/i, k0 = (size_t) q.
// ---------------------------------------------------------------------------------------------
template<int BS, bool RUNTIME_K0, class R>
GCMB_HD R limited_newton(R (&s)[BS + 1], const R* __restrict__ F, int k0) {
	R a, b;
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
	const R maximum = a_gt_b ? a : b;
	const R minimum = a_gt_b ? b : a;
	R ans = s[0];
#pragma unroll
	for (int i = 1; i <= BS; i++) {
		const R f = F[i - 1];
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
// The structurally non-zero coefficients only, packed in the order the kernel consumes them:
//     [Newton factors (and, K0RT, the foot cell) of the interpolated rows][non-zeros of U, row by row]
//     [non-zeros of U1, row by row].
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
template<class P, int BS, bool K0RT>
GCMB_HD constexpr int packed_fpos(int k) {  // k must be a base row (or P::M for the total)
	int n = 0;
	for (int r = 0; r < k; r++) { if (P::sgn(r) != 0 && P::base(r) == r) { n += BS + (K0RT ? 1 : 0); } }
	return n;
}
template<class P, int BS, bool K0RT>
GCMB_HD constexpr int packed_upos(int k, int j) {
	int n = packed_fpos<P, BS, K0RT>(P::M);
	for (int r = 0; r < k; r++) { if (P::base(r) == r) { n += popcount_u(P::um(r)); } }
	return k < P::M ? n + popcount_u(P::um(k) & ((1u << j) - 1u)) : n;
}
template<class P, int BS, bool K0RT>
GCMB_HD constexpr int packed_u1pos(int i, int k) {
	constexpr unsigned B = packed_base_mask<P>();
	int n = packed_upos<P, BS, K0RT>(P::M, 0);
	for (int r = 0; r < i; r++) { n += popcount_u(P::u1m(r) & B); }
	return i < P::M ? n + popcount_u(P::u1m(i) & B & ((1u << k) - 1u)) : n;
}
template<class P, int BS, bool K0RT>
struct Packed {
	static constexpr int SIZE = packed_u1pos<P, BS, K0RT>(P::M, 0);
};

template<class R, class P, int BS, bool K0RT>
struct PackedCoef {
	const R* __restrict__ t;  // packed table of the node's material (shared memory)
	// a row/column that shares reads its base's entry (the compiler merges the two loads) and flips the
	// sign where the pattern says so: (-c) * v == -(c * v) exactly
	GCMB_HD R u(int k, int j) const {
		const R c = t[packed_upos<P, BS, K0RT>(P::base(k), j)];
		return ((P::uneg(k) >> j) & 1u) ? -c : c;
	}
	GCMB_HD R u1(int i, int k) const {
		const R c = t[packed_u1pos<P, BS, K0RT>(i, P::base(k))];
		return ((P::u1neg(k) >> i) & 1u) ? -c : c;
	}
	GCMB_HD const R* f(int k) const { return t + packed_fpos<P, BS, K0RT>(P::base(k)); }
	GCMB_HD int k0(int k) const { return K0RT ? (int) t[packed_fpos<P, BS, K0RT>(P::base(k)) + BS] : 0; }
};

template<class R, class P, int BS, bool K0RT, class COEF, class LOAD>
GCMB_HD void gcm_node_sparse(const COEF coef, LOAD load, R (&out)[P::M]) {
	constexpr int M = P::M;
	R r[M];
#pragma unroll
	for (int k = 0; k < M; k++) {
		const int sg = P::sgn(k);
		const unsigned mask = P::um(k);
		R acc = 0;
		bool first = true;
		const int k0 = sg != 0 ? coef.k0(k) : 0;
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((mask >> j) & 1u) {
				R v;
				if (sg == 0) {
					v = load(j, 0);
				} else {
					R s[BS + 1];
#pragma unroll
					for (int i = 0; i <= BS; i++) { s[i] = load(j, sg * i); }
					v = limited_newton<BS, K0RT>(s, coef.f(k), k0);
				}
				const R t = coef.u(k, j) * v;
				if (first) { acc = t; first = false; } else { acc += t; }
			}
		}
		r[k] = acc;
	}
#pragma unroll
	for (int i = 0; i < M; i++) {
		const unsigned mask = P::u1m(i);
		R acc = 0;
		bool first = true;
#pragma unroll
		for (int k = 0; k < M; k++) {
			if ((mask >> k) & 1u) {
				const R t = coef.u1(i, k) * r[k];
				if (first) { acc = t; first = false; } else { acc += t; }
			}
		}
		out[i] = acc;
	}
}

// Same, dense eigen-system and runtime border size / foot cell (any material, any Courant number
// up to bs).  A literal restatement of the reference loop.
template<class R, int M, class LOAD>
GCMB_HD void gcm_node_dense(const StageTableT<R>* __restrict__ T, int bs, LOAD load, R (&out)[M]) {
	R r[M];
	for (int k = 0; k < M; k++) {
		const int dir = T->dir[k];
		const int k0 = T->k0[k];
		const int k1 = k0 < bs ? k0 + 1 : k0;
		const R* F = T->F + k * MAXBS;
		R acc = 0;
		for (int j = 0; j < M; j++) {
			R s[MAXBS + 1];
			for (int i = 0; i <= bs; i++) { s[i] = load(j, dir * i); }
			const R maximum = (R) fmax(s[k0], s[k1]);
			const R minimum = (R) fmin(s[k0], s[k1]);
			R ans = s[0];
			for (int i = 1; i <= bs; i++) {
				const R f = F[i - 1];
				for (int jj = 0; jj < bs - i + 1; jj++) { s[jj] = (s[jj + 1] - s[jj]) * f; }
				ans += s[0];
			}
			if (ans > maximum) { ans = maximum; } else if (ans < minimum) { ans = minimum; }
			const R t = T->U[k * M + j] * ans;
			if (j == 0) { acc = t; } else { acc += t; }
		}
		r[k] = acc;
	}
	for (int i = 0; i < M; i++) {
		R acc = T->U1[i * M] * r[0];
		for (int k = 1; k < M; k++) { acc += T->U1[i * M + k] * r[k]; }
		out[i] = acc;
	}
}

// Dense eigen-system, all feet inside the first cell (k0 == 0 everywhere: Courant number < 1) -- or, K0RT, inside the
// first two cells (Courant number up to 2 at border size 2) -- and a compile-time border size: the kernel of rotated
// orthotropic materials and of any eigen-system no pattern covers.  The node's
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
GCMB_HD float clamp_select(float ans, float mn, float mx) {
#ifdef __CUDA_ARCH__
	asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %0, %1;\n\tselp.f32 %0, %1, %0, p;\n\tsetp.lt.f32 p, %0, %2;\n\tselp.f32 %0, %2, %0, p;\n\t}"
	    : "+f"(ans) : "f"(mx), "f"(mn));
	return ans;
#else
	if (ans > mx) { ans = mx; } else if (ans < mn) { ans = mn; }
	return ans;
#endif
}

// coefficient sources: (1) the StageTable of the node's material in global memory
template<class R, int M>
struct DenseTableCoef {
	const StageTableT<R>* __restrict__ T;
	GCMB_HD R u(int k, int j) const { return GCMB_LDG(&T->U[k * M + j]); }
	GCMB_HD R u1(int i, int k) const { return GCMB_LDG(&T->U1[i * M + k]); }
	GCMB_HD R f(int k, int i) const { return GCMB_LDG(&T->F[k * MAXBS + i]); }
	GCMB_HD int side(int k) const { return f(k, 0) == R(0) ? 0 : T->dir[k]; }
	GCMB_HD int k0(int k) const { return T->k0[k]; }
};
// (2) a body of ONE material: the coefficients travel as a kernel parameter, so that every one of them is a
//     constant-bank operand of the arithmetic instruction that uses it (no load, no register)
template<class R, int M, int BS>
struct DenseParamCoef {
	R U[M * M], U1[M * M], F[M * BS];
	int sd[M];
	int kz[M];
	GCMB_HD R u(int k, int j) const { return U[k * M + j]; }
	GCMB_HD R u1(int i, int k) const { return U1[i * M + k]; }
	GCMB_HD R f(int k, int i) const { return F[k * BS + i]; }
	GCMB_HD int side(int k) const { return sd[k]; }
	GCMB_HD int k0(int k) const { return kz[k]; }
};

template<class R, int M, int BS, bool K0RT, class COEF, class LOAD>
GCMB_HD void gcm_node_dense_k0(const COEF& co, LOAD load, R (&out)[M]) {
	static_assert(BS == 1 || BS == 2, "border sizes 1 and 2");
	R c[M], r[M];
#pragma unroll
	for (int j = 0; j < M; j++) { c[j] = load(j, 0); }
	int side[M];
#pragma unroll
	for (int k = 0; k < M; k++) { side[k] = co.side(k); }
#pragma unroll
	for (int k = 0; k < M; k++) {
		if (side[k] == 0) {
			R acc = co.u(k, 0) * c[0];
#pragma unroll
			for (int j = 1; j < M; j++) { acc += co.u(k, j) * c[j]; }
			r[k] = acc;
		}
	}
	// both sides unrolled: measured 33.1 ms/step against 34.8 ms with `#pragma unroll 1` (half the code, but 24 B of
	// spills) on 2 x 512x256x512 rotated plies
#pragma unroll
	for (int sd = 1; sd >= -1; sd -= 2) {
		R d0[M], d1[BS == 2 ? M : 1], mx[M], mn[M];
		R mx1[K0RT ? M : 1], mn1[K0RT ? M : 1];  // bracket of the second cell (foot cell 1)
#pragma unroll
		for (int j = 0; j < M; j++) {
			const R s1 = load(j, sd);
			d0[j] = s1 - c[j];
			const bool gt = c[j] > s1;
			mx[j] = gt ? c[j] : s1;
			mn[j] = gt ? s1 : c[j];
			if (BS == 2) {
				const R s2 = load(j, 2 * sd);
				d1[j] = s2 - s1;
				if (K0RT) {
					const bool gt1 = s1 > s2;
					mx1[j] = gt1 ? s1 : s2;
					mn1[j] = gt1 ? s2 : s1;
				}
			} else if (K0RT) {
				mx1[j] = s1;  // border size 1, foot cell 1 (q == 1 exactly): the bracket degenerates to s[1]
				mn1[j] = s1;
			}
		}
#pragma unroll
		for (int k = 0; k < M; k++) {
			if (side[k] == sd) {
				const R f0 = co.f(k, 0);
				const R f1 = BS == 2 ? co.f(k, 1) : R(0);
				const bool second = K0RT && co.k0(k) != 0;
				R acc = 0;
#pragma unroll
				for (int j = 0; j < M; j++) {
					const R a0 = d0[j] * f0;
					R ans = c[j] + a0;
					if (BS == 2) {
						const R a1 = d1[j] * f0;
						ans += (a1 - a0) * f1;
					}
					if (K0RT) { ans = clamp_select(ans, second ? mn1[j] : mn[j], second ? mx1[j] : mx[j]); }
					else { ans = clamp_select(ans, mn[j], mx[j]); }
					const R t = co.u(k, j) * ans;
					if (j == 0) { acc = t; } else { acc += t; }
				}
				r[k] = acc;
			}
		}
	}
#pragma unroll
	for (int i = 0; i < M; i++) {
		R acc = co.u1(i, 0) * r[0];
#pragma unroll
		for (int k = 1; k < M; k++) { acc += co.u1(i, k) * r[k]; }
		out[i] = acc;
	}
}

// loader over the structure-of-arrays volumes: component j at `o` nodes along the sweep axis
template<class R>
struct SoaLoad {
	const R* __restrict__ base;  // cur + idx
	long long comp, sstride;
	GCMB_HD R operator()(int j, int o) const { return GCMB_LDG(base + j * comp + o * sstride); }
};

template<class R, int M>
GCMB_HD void stage_thread_dense(const StageArgsT<R>& a, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	const StageTableT<R>* T = a.tables + ((int) a.node_table[idx] * a.g.D + a.dir);
	SoaLoad<R> ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	R out[M];
	gcm_node_dense<R, M>(T, a.g.bs, ld, out);
	for (int c = 0; c < M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

template<class R, int M, int BS, bool K0RT>
GCMB_HD void stage_thread_dense_k0(const StageArgsT<R>& a, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	const DenseTableCoef<R, M> co{a.tables + ((int) a.node_table[idx] * a.g.D + a.dir)};
	SoaLoad<R> ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	R out[M];
	gcm_node_dense_k0<R, M, BS, K0RT>(co, ld, out);
#pragma unroll
	for (int c = 0; c < M; c++) { a.nxt[c * a.g.comp + idx] = out[c]; }
}

template<class R, int M, int BS, bool K0RT>
GCMB_HD void stage_thread_dense_k0_one(const StageArgsT<R>& a, const DenseParamCoef<R, M, BS>& co, int i0, int i1, int i2) {
	const long long idx = a.g.index(i0, i1, i2);
	SoaLoad<R> ld{a.cur + idx, a.g.comp, a.g.stride(a.axis)};
	R out[M];
	gcm_node_dense_k0<R, M, BS, K0RT>(co, ld, out);
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

template<class R>
GCMB_HD R get_quantity(int D, int code, const R* v) {
	if (code >= 0) { return v[code]; }
	R trace = 0;
	for (int i = 0; i < D; i++) { trace += v[D + sym_index(D, i, i)]; }
	return -trace / D;
}

template<class R>
GCMB_HD void set_quantity(int D, int M, int code, R value, R* v) {
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
template<class R>
struct BorderArgs {
	R* pde;
	const uint8_t* mask;  // face mask or nullptr
	Geom g;
	int axis;             // internal axis
	int side;             // 0 left (inner sign +1), 1 right (inner sign -1)
	int nq;
	int q[MAXM + 1];
	R val[MAXM + 1];
};

template<class R>
GCMB_HD void border_thread(const BorderArgs<R>& b, long long f, int a /* 1..bs */) {
	if (b.mask && !b.mask[f]) { return; }
	int it[3];
	face_node(b.g, b.axis, f, b.side == 0 ? 0 : b.g.n[b.axis] - 1, it);
	const int sign = b.side == 0 ? 1 : -1;
	int in[3] = {it[0], it[1], it[2]}, gh[3] = {it[0], it[1], it[2]};
	in[b.axis] += sign * a;
	gh[b.axis] -= sign * a;
	const long long ii = b.g.index(in[0], in[1], in[2]);
	const long long gi = b.g.index(gh[0], gh[1], gh[2]);
	R inner[MAXM], ghost[MAXM];
	for (int c = 0; c < b.g.M; c++) { inner[c] = b.pde[c * b.g.comp + ii]; ghost[c] = inner[c]; }
	for (int j = 0; j < b.nq; j++) {
		const R innerValue = get_quantity(b.g.D, b.q[j], inner);
		const R ghostValue = -innerValue + 2 * b.val[j];
		set_quantity(b.g.D, b.g.M, b.q[j], ghostValue, ghost);
	}
	for (int c = 0; c < b.g.M; c++) { b.pde[c * b.g.comp + gi] = ghost[c]; }
}

// The same for a face across the CONTIGUOUS axis, one thread per face node, all ghost layers: the `bs` ghost nodes of a
// row and the padding next to them make up one 32-byte sector, which the thread writes WHOLE (padding = 0).  A sector that
// is written completely needs no read-modify-write in L2/DRAM; the one-layer-at-a-time fill above moved 13x its useful
// bytes (profiles/r2_launches_traffic_1024.csv: 3.1 GB read + 1.0 GB written per face of a 1024^3 body).
// Requires: border size <= SECT = 32 / sizeof(R) elements; right side: n2 % SECT == 0 (the host checks both).
template<class R>
GCMB_HD void border_thread_zsector(const BorderArgs<R>& b, long long f) {
	constexpr int SECT = 32 / (int) sizeof(R);
	if (b.mask && !b.mask[f]) { return; }
	const Geom& g = b.g;
	int it[3];
	face_node(g, 2, f, b.side == 0 ? 0 : g.n[2] - 1, it);
	const int sign = b.side == 0 ? 1 : -1;
	const long long face = g.index(it[0], it[1], it[2]);
	// first element of the sector that holds the ghosts: [pad .. pad, ghost(-bs) .. ghost(-1)] on the left,
	// [ghost(n2) .. ghost(n2+bs-1), pad .. pad] on the right
	const long long first = b.side == 0 ? face - SECT : face + 1;
	R sector[MAXM][SECT];
	for (int c = 0; c < g.M; c++) { for (int e = 0; e < SECT; e++) { sector[c][e] = R(0); } }
	for (int a = 1; a <= g.bs; a++) {
		R inner[MAXM], ghost[MAXM];
		for (int c = 0; c < g.M; c++) { inner[c] = b.pde[c * g.comp + face + sign * a]; ghost[c] = inner[c]; }
		for (int j = 0; j < b.nq; j++) {
			const R innerValue = get_quantity(g.D, b.q[j], inner);
			const R ghostValue = -innerValue + 2 * b.val[j];
			set_quantity(g.D, g.M, b.q[j], ghostValue, ghost);
		}
		const int e = b.side == 0 ? SECT - a : a - 1;
		for (int c = 0; c < g.M; c++) { sector[c][e] = ghost[c]; }
	}
	for (int c = 0; c < g.M; c++) {
		R* dst = b.pde + c * g.comp + first;
#if defined(__CUDA_ARCH__)
		// two 16-byte stores: the sector is 32-byte aligned (zoff and n2 are multiples of SECT)
		double2* d2 = reinterpret_cast<double2*>(dst);
		const double2* s2 = reinterpret_cast<const double2*>(sector[c]);
		d2[0] = s2[0];
		d2[1] = s2[1];
#else
		for (int e = 0; e < SECT; e++) { dst[e] = sector[c][e]; }
#endif
	}
}

// ---------------------------------------------------------------------------------------------
// contact ghost copy (reference engine/cubic/ContactConditions.hpp:56-68): thread = node of the box
// ---------------------------------------------------------------------------------------------
template<class R>
struct ContactArgs {
	R* a;
	const R* b;
	Geom ga, gb;
	int amin[3], bmin[3], ext[3];  // internal axes
};

template<class R>
GCMB_HD void contact_thread(const ContactArgs<R>& c, long long t) {
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
template<class R>
GCMB_HD void ode_maxwell_thread(const Geom& g, R* pde, const uint8_t* node_table,
                                const R* decay, int i0, int i1, int i2) {
	const long long idx = g.index(i0, i1, i2);
	const R f = decay[node_table[idx]];
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
template<class R>
struct XferArgs {
	R* soa;
	R* aos;             // device staging buffer holding the slice [x_begin, x_end) of internal axis 0
	Geom g;
	int with_ghosts;
	int x_begin, x_end; // in host-array coordinates of internal axis 0 (ghost-inclusive when with_ghosts)
};

template<class R>
GCMB_HD void xfer_thread(const XferArgs<R>& x, long long t, bool to_device) {
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

// which components of a pattern are interpolated / read at the node only
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

}  // namespace gcmb
