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
#define GCMB_DYN_SMEM(type, name)                                  \
	extern __shared__ __align__(16) unsigned char gcmb_dyn_smem_[]; \
	type& name = *reinterpret_cast<type*>(gcmb_dyn_smem_)
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

// One eigen-system prepared for the stage kernels: per (material table, reference direction).
struct StageTable {
	double U[MAXM * MAXM];   // row-major, row k = left eigenvector k
	double U1[MAXM * MAXM];  // row-major
	double F[MAXM * MAXBS];  // F[k*MAXBS + i-1] = (q_k - i + 1) / i, i = 1..bs  (Newton factors)
	int k0[MAXM];            // (size_t) q_k : cell containing the characteristic foot
	int dir[MAXM];           // +1 / -1 : side the foot lies on (dx_k > 0 ? +1 : -1)
};

struct StageArgs {
	const double* cur;
	double* nxt;
	const uint8_t* node_table;
	const StageTable* tables; // indexed [table * D + dir]
	const double* packed;     // this direction's packed coefficient tables [table][Packed::SIZE] or null
	int n_tables;
	Geom g;
	int axis;                 // internal axis of the sweep
	int dir;                  // reference direction
	int x_begin, x_end;       // range of internal axis 0 to process (slab sub-ranges for overlap)
	const StageTable* host_tables; // HOST copy of `tables` (launchers only: coefficients passed as kernel parameters)
};

typedef void (*StageLauncher)(const StageArgs&, cudaStream_t);

struct PatternInfo {
	const char* name;
	int M;
	int sgn[MAXM];
	unsigned um[MAXM];
	unsigned u1m[MAXM];
	int base[MAXM];        // row/column whose coefficients row/column k shares (k itself when none)
	unsigned uneg[MAXM];   // bit j: U(k,j) == -U(base,j)
	unsigned u1neg[MAXM];  // bit i: U1(i,k) == -U1(i,base)
	StageLauncher launch_bs1;
	StageLauncher launch_bs2;
};

// host side of thread_fns.h Packed<P,BS>: the non-zero coefficients of one table in kernel order, the
// rows/columns that share their coefficients with a base row stored once
inline int pack_table(const PatternInfo& p, int bs, const StageTable& T, double* out) {
	int n = 0;
	for (int k = 0; k < p.M; k++) {
		if (p.sgn[k] != 0 && p.base[k] == k) { for (int i = 0; i < bs; i++) { out[n++] = T.F[k * MAXBS + i]; } }
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
inline bool table_shares_as_pattern(const PatternInfo& p, int bs, const StageTable& T) {
	for (int k = 0; k < p.M; k++) {
		const int b = p.base[k];
		if (b == k) { continue; }
		for (int i = 0; i < bs; i++) { if (T.F[k * MAXBS + i] != T.F[b * MAXBS + i]) { return false; } }
		for (int j = 0; j < p.M; j++) {
			const double want = ((p.uneg[k] >> j) & 1u) ? -T.U[b * p.M + j] : T.U[b * p.M + j];
			if (T.U[k * p.M + j] != want) { return false; }
			const double want1 = ((p.u1neg[k] >> j) & 1u) ? -T.U1[j * p.M + b] : T.U1[j * p.M + b];
			if (T.U1[j * p.M + k] != want1) { return false; }
		}
	}
	return true;
}

// stage_dispatch.cu
int pattern_count();
const PatternInfo& pattern(int i);
StageLauncher dense_launcher(int M);
StageLauncher dense_k0_launcher(int M, int bs);

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
