// Internal declarations shared by the translation units of libgcm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/gcm_b200.h"

// Kernel-launch plumbing.  The host-logic tests (tests/emul/, never shipped) pre-define these macros to
// step through the same sources on the build machine, where no GPU exists; the product always gets
// the CUDA definitions below.
#ifndef GCMB_LAUNCH
#define GCMB_GLOBAL __global__
#define GCMB_DEV __device__ __forceinline__
#define GCMB_BOUNDS(n) __launch_bounds__(n)
#define GCMB_BOUNDS2(n, b) __launch_bounds__(n, b)
#define GCMB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// kernels with block barriers: written as phases over GCMB_BLOCK_THREADS (a single pass on the device)
#define GCMB_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GCMB_BLOCK_THREADS(tid) for (int tid = threadIdx.x, gcmb_once_ = 1; gcmb_once_; gcmb_once_ = 0)
#define GCMB_DYN_SMEM_RAW(name)                                      \
	extern __shared__ __align__(128) unsigned char gcmb_dyn_smem_[]; \
	unsigned char* name = gcmb_dyn_smem_
#endif
#ifdef __CUDACC__
#define GCMB_HD __host__ __device__ __forceinline__
#else
#define GCMB_HD inline
#endif

namespace gcmb {

constexpr int MAXM = GCMB_MAX_M;
constexpr int MAXBS = GCMB_MAX_BORDER;

// ---------------------------------------------------------------------------------------------
// Device layout of one body (DESIGN.md "Data layout in HBM").
// The D reference axes are mapped onto the LAST D of three internal axes, so that the reference's
// fastest axis is always internal axis 2 (contiguous).  Structure of arrays: M component volumes,
// each (n0+2g0) x (n1+2g1) x pitch doubles; the real node z=0 of every row sits at element `zoff`
// (a multiple of 16 doubles = 128 B, so that warps reading a row start on a cache line).
// ---------------------------------------------------------------------------------------------
struct Geom {
	int D, M, bs;
	int shift;       // internal axis = reference axis + shift  (shift = 3 - D)
	int n[3];        // real nodes per internal axis (1 on inactive axes)
	int g[3];        // ghost width per internal axis (bs or 0)
	int zoff;        // element offset of real z=0 inside a row
	int pitch;       // row pitch in elements
	long long plane; // (n1 + 2 g1) * pitch : stride of internal axis 0
	long long comp;  // elements per component volume (stride between components)
	int start[3];    // global index of the first real node per internal axis
	double h[3];     // spatial step per internal axis (1 on inactive axes)

	GCMB_HD long long stride(int a) const { return a == 0 ? plane : (a == 1 ? (long long) pitch : 1LL); }
	GCMB_HD long long index(int i0, int i1, int i2) const {
		return (long long) (i0 + g[0]) * plane + (long long) (i1 + g[1]) * pitch + (i2 + zoff);
	}
};

// One eigen-system prepared for the stage kernels: per (material table, reference direction), in the
// arithmetic type R of the body (double, or float for gcmb_create(..., 4)).
template<class R>
struct StageTableT {
	R U[MAXM * MAXM];   // row-major, row k = left eigenvector k
	R U1[MAXM * MAXM];  // row-major
	R F[MAXM * MAXBS];  // F[k*MAXBS + i-1] = (q_k - i + 1) / i, i = 1..bs  (Newton factors)
	int k0[MAXM];       // (size_t) q_k : cell containing the characteristic foot
	int dir[MAXM];      // +1 / -1 : side the foot lies on (dx_k > 0 ? +1 : -1)
};
typedef StageTableT<double> StageTable;

// ghost fill of the two faces across the contiguous axis, done by the marching stage kernel that writes the
// layer (reference engine/cubic/BorderConditions.hpp:97-114 for conditions over the whole face with plain
// component quantities): ghost(-a) = inner(+a), components in `flip` negated and shifted by add[c] = 2 b_c(t)
template<class R>
struct ZFaceFill {
	int on[2];        // left / right face
	unsigned set[2];  // bit c: component c is Set by some condition of the face
	R add[2][MAXM];   // 2 * b_c(t_n) of the LAST condition that sets component c
};

template<class R>
struct StageArgsT {
	const R* cur;
	R* nxt;
	const uint8_t* node_table;
	const StageTableT<R>* tables; // indexed [table * D + dir]
	const R* packed;          // this direction's packed coefficient tables [table][Packed::SIZE] or null
	int n_tables;
	Geom g;
	int axis;                 // internal axis of the sweep
	int dir;                  // reference direction
	int x_begin, x_end;       // range of internal axis 0 to process (slab sub-ranges for overlap)
	const StageTableT<R>* host_tables; // HOST copy of `tables` (launchers only: coefficients passed as kernel parameters)
	int zfill;                // != 0: also fill the z ghosts of the written layer (zf)
	ZFaceFill<R> zf;
};
typedef StageArgsT<double> StageArgs;

// type-erased launcher: `args` points to the StageArgsT<R> of the kernel set the launcher belongs to
typedef void (*StageLauncher)(const void* args, cudaStream_t);

// kernel sets: arithmetic type x floating-point contraction
enum { SET_F64_EXACT = 0,  // double, no FMA contraction: bit-identical to the reference CPU engine
       SET_F64_FMA = 1,    // double, contraction allowed (gcmb_set_fma): within 1e-12 of the reference, not bit-identical
       SET_F32 = 2,        // float (gcmb_create(..., 4)), contraction allowed
       N_SETS = 3 };
// variants of the specialised kernels: border size x where the characteristic feet may lie
enum { VAR_BS1 = 0,        // border size 1, foot cell read from the table
       VAR_BS2_K0 = 1,     // border size 2, every foot inside the first cell (Courant number < 1)
       VAR_BS2 = 2,        // border size 2, foot cell read from the table (Courant number up to 2)
       VAR_BS3 = 3,        // border size 3, foot cell read from the table
       N_VARIANTS = 4 };
constexpr int variant_bs(int v) { return v == VAR_BS1 ? 1 : (v == VAR_BS3 ? 3 : 2); }
constexpr bool variant_k0rt(int v) { return v != VAR_BS2_K0; }

struct PatternInfo {
	const char* name;
	int group;             // translation unit the kernels of the pattern are compiled in (build.py)
	int M;
	int axis;              // internal axis of the pattern's direction (2 = contiguous)
	int sgn[MAXM];
	unsigned um[MAXM];
	unsigned u1m[MAXM];
	int base[MAXM];        // row/column whose coefficients row/column k shares (k itself when none)
	unsigned uneg[MAXM];   // bit j: U(k,j) == -U(base,j)
	unsigned u1neg[MAXM];  // bit i: U1(i,k) == -U1(i,base)
};

// host side of thread_fns.h Packed<P,BS,K0RT>: the non-zero coefficients of one table in kernel order, the
// rows/columns that share their coefficients with a base row stored once; with k0rt the foot cell of every
// interpolated base row follows its Newton factors
template<class R>
inline int pack_table(const PatternInfo& p, int bs, bool k0rt, const StageTableT<R>& T, R* out) {
	int n = 0;
	for (int k = 0; k < p.M; k++) {
		if (p.sgn[k] != 0 && p.base[k] == k) {
			for (int i = 0; i < bs; i++) { out[n++] = T.F[k * MAXBS + i]; }
			if (k0rt) { out[n++] = (R) T.k0[k]; }
		}
	}
	for (int k = 0; k < p.M; k++) {
		if (p.base[k] != k) { continue; }
		for (int j = 0; j < p.M; j++) { if ((p.um[k] >> j) & 1u) { out[n++] = T.U[k * p.M + j]; } }
	}
	for (int i = 0; i < p.M; i++) {
		for (int k = 0; k < p.M; k++) { if (((p.u1m[i] >> k) & 1u) && p.base[k] == k) { out[n++] = T.U1[i * p.M + k]; } }
	}
	return n;
}

// true when the table really has the sharing structure the pattern assumes (checked bit for bit)
template<class R>
inline bool table_shares_as_pattern(const PatternInfo& p, int bs, const StageTableT<R>& T) {
	for (int k = 0; k < p.M; k++) {
		const int b = p.base[k];
		if (b == k) { continue; }
		if (T.k0[k] != T.k0[b]) { return false; }
		for (int i = 0; i < bs; i++) { if (T.F[k * MAXBS + i] != T.F[b * MAXBS + i]) { return false; } }
		for (int j = 0; j < p.M; j++) {
			const R want = ((p.uneg[k] >> j) & 1u) ? -T.U[b * p.M + j] : T.U[b * p.M + j];
			if (T.U[k * p.M + j] != want) { return false; }
			const R want1 = ((p.u1neg[k] >> j) & 1u) ? -T.U1[j * p.M + b] : T.U1[j * p.M + b];
			if (T.U1[j * p.M + k] != want1) { return false; }
		}
	}
	return true;
}

// stage_dispatch.cu: registry of the stage kernels compiled in stage_inst.cu (one translation unit per
// kernel set and pattern group)
int pattern_count();
const PatternInfo& pattern(int i);
StageLauncher sparse_launcher(int set, int pattern, int variant);  // null when the set was not built
StageLauncher dense_launcher(int set, int M);                       // literal restatement, any border size / foot cell
StageLauncher dense_k0_launcher(int set, int M, int bs, bool k0rt); // feet within the first two cells, bs 1..2

// error handling ------------------------------------------------------------------------------
void set_error(const std::string& msg);
#define GCMB_CUDA(call)                                                                        \
	do {                                                                                       \
		cudaError_t e__ = (call);                                                              \
		if (e__ != cudaSuccess) {                                                              \
			char buf__[512];                                                                   \
			snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,        \
			         cudaGetErrorString(e__));                                                 \
			gcmb::set_error(buf__);                                                            \
			return GCMB_E_CUDA;                                                                \
		}                                                                                      \
	} while (0)
#define GCMB_FAIL(code, msg)                                                                   \
	do {                                                                                       \
		gcmb::set_error(std::string(__func__) + ": " + (msg));                                 \
		return (code);                                                                         \
	} while (0)

}  // namespace gcmb
