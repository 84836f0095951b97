// Stage kernels of gcm_b200: one specialised kernel per sparsity class of the eigen-system (patterns.inc)
// x variant {border size, foot cell} (internal.cuh VAR_*), plus dense kernels for everything else.
//
// This file is compiled once per (kernel set, pattern group): build.py passes -DGCMB_SET=<SET_*> and
// -DGCMB_GROUP=<group of patterns.inc | 100 for the dense kernels>; SET_F64_EXACT is compiled with
// -fmad=false, the other sets with FMA contraction.  Every kernel carries SET as a template parameter so
// that the objects of different sets do not collide at link time.  The launchers register themselves in
// the table of stage_dispatch.cu.
#include <cstdlib>

#include "ztile.h"
#ifndef GCMB_EMUL
#include "tma_pipe.h"
#endif

#ifndef GCMB_SET
#error "GCMB_SET is not defined"
#endif
#ifndef GCMB_GROUP
#error "GCMB_GROUP is not defined"
#endif

namespace gcmb {

void register_sparse_launcher(int set, const char* pattern_name, int variant, StageLauncher f);
void register_dense_launcher(int set, int M, StageLauncher f);
void register_dense_k0_launcher(int set, int M, int bs, bool k0rt, StageLauncher f);

namespace {

constexpr int SET = GCMB_SET;
#if GCMB_SET == 2
typedef float Real;
#else
typedef double Real;
#endif
typedef StageArgsT<Real> Args;

constexpr int ZT = 128;  // threads per block, all along the contiguous axis

#if GCMB_GROUP == 100
// block -> node mapping of the one-thread-per-node kernels: blockIdx.x runs along the sweep axis when it is
// strided, so that blocks scheduled together share their halo planes in L2
GCMB_DEV bool block_node(const Args& a, int& i0, int& i1, int& i2) {
	int zc;
	if (a.axis == 0) { i0 = blockIdx.x + a.x_begin; zc = blockIdx.y; i1 = blockIdx.z; }
	else if (a.axis == 1) { i1 = blockIdx.x; zc = blockIdx.y; i0 = blockIdx.z + a.x_begin; }
	else { zc = blockIdx.x; i1 = blockIdx.y; i0 = blockIdx.z + a.x_begin; }
	i2 = zc * ZT + threadIdx.x;
	return i2 < a.g.n[2];
}

dim3 node_blocks(const Args& a) {
	const unsigned zc = (unsigned) ((a.g.n[2] + ZT - 1) / ZT);
	const unsigned nx = (unsigned) (a.x_end - a.x_begin);
	if (a.axis == 0) { return dim3(nx, zc, (unsigned) a.g.n[1]); }
	if (a.axis == 1) { return dim3((unsigned) a.g.n[1], zc, nx); }
	return dim3(zc, (unsigned) a.g.n[1], nx);
}

// ------------------------------------------------------------------------------------------------
// dense eigen-systems
// ------------------------------------------------------------------------------------------------
template<int S, int M>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense(const Args a) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense<Real, M>(a, i0, i1, i2); }
}

template<int S, int M, int BS, bool K0RT>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense_k0(const Args a) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense_k0<Real, M, BS, K0RT>(a, i0, i1, i2); }
}

template<int S, int M, int BS, bool K0RT>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense_k0_one(const Args a, const DenseParamCoef<Real, M, BS> co) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense_k0_one<Real, M, BS, K0RT>(a, co, i0, i1, i2); }
}

template<int M>
void launch_dense(const void* args, cudaStream_t stream) {
	const Args& a = *static_cast<const Args*>(args);
	auto kernel = k_stage_dense<SET, M>;
	GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a);
}

template<int M, int BS, bool K0RT>
void launch_dense_k0(const void* args, cudaStream_t stream) {
	const Args& a = *static_cast<const Args*>(args);
	if (a.n_tables == 1 && a.host_tables) {
		const StageTableT<Real>& T = a.host_tables[a.dir];
		DenseParamCoef<Real, M, BS> co;
		for (int k = 0; k < M; k++) {
			for (int j = 0; j < M; j++) { co.U[k * M + j] = T.U[k * M + j]; co.U1[k * M + j] = T.U1[k * M + j]; }
			for (int i = 0; i < BS; i++) { co.F[k * BS + i] = T.F[k * MAXBS + i]; }
			co.sd[k] = T.F[k * MAXBS] == Real(0) ? 0 : T.dir[k];
			co.kz[k] = T.k0[k];
		}
		auto kernel = k_stage_dense_k0_one<SET, M, BS, K0RT>;
		GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a, co);
		return;
	}
	auto kernel = k_stage_dense_k0<SET, M, BS, K0RT>;
	GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a);
}

template<int M>
void register_dense_M() {
	register_dense_launcher(SET, M, &launch_dense<M>);
	register_dense_k0_launcher(SET, M, 1, false, &launch_dense_k0<M, 1, false>);
	register_dense_k0_launcher(SET, M, 1, true, &launch_dense_k0<M, 1, true>);
	register_dense_k0_launcher(SET, M, 2, false, &launch_dense_k0<M, 2, false>);
	register_dense_k0_launcher(SET, M, 2, true, &launch_dense_k0<M, 2, true>);
}

struct Registrar {
	Registrar() {
		register_dense_M<2>();
		register_dense_M<3>();
		register_dense_M<4>();
		register_dense_M<5>();
		register_dense_M<9>();
	}
};
const Registrar g_registrar;

#else
int env_int(const char* name, int dflt) {
	const char* v = getenv(name);
	return v ? atoi(v) : dflt;
}

// cudaFuncSetAttribute once per (kernel, device)
template<class K>
void func_attr_once(K kernel, cudaFuncAttribute attr, int value, unsigned long long& done_mask) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (!((done_mask >> (dev & 63)) & 1ull)) {
		cudaFuncSetAttribute(kernel, attr, value);
		done_mask |= 1ull << (dev & 63);
	}
}

// ------------------------------------------------------------------------------------------------
// sparse eigen-systems: the patterns of this group
// ------------------------------------------------------------------------------------------------
// the lists stay parenthesised while they are forwarded through GCMB_IN_GROUP_n and become braces at the end
#define GCMB_L(...) (__VA_ARGS__)
#define GCMB_BRACES(...) {__VA_ARGS__}
#define GCMB_PATTERN_DEFINE(NAME, MM, AXIS, SGN, UM, U1M, BASE, UNEG, U1NEG)                            \
	struct Pat_##NAME {                                                                           \
		static constexpr int M = MM;                                                              \
		static constexpr int axis = AXIS;                                                         \
		GCMB_HD static constexpr int sgn(int k) { constexpr int t[9] = GCMB_BRACES SGN; return t[k]; }         \
		GCMB_HD static constexpr unsigned um(int k) { constexpr unsigned t[9] = GCMB_BRACES UM; return t[k]; } \
		GCMB_HD static constexpr unsigned u1m(int k) { constexpr unsigned t[9] = GCMB_BRACES U1M; return t[k]; } \
		GCMB_HD static constexpr int base(int k) { constexpr int t[9] = GCMB_BRACES BASE; return t[k]; }       \
		GCMB_HD static constexpr unsigned uneg(int k) { constexpr unsigned t[9] = GCMB_BRACES UNEG; return t[k]; } \
		GCMB_HD static constexpr unsigned u1neg(int k) { constexpr unsigned t[9] = GCMB_BRACES U1NEG; return t[k]; } \
		static constexpr const char* name() { return #NAME; }                                     \
	};
#define GCMB_CAT2(a, b) a##b
#define GCMB_CAT(a, b) GCMB_CAT2(a, b)
#define GCMB_IN_GROUP_0(...)
#define GCMB_IN_GROUP_1(...)
#define GCMB_IN_GROUP_2(...)
#define GCMB_IN_GROUP_3(...)
#define GCMB_IN_GROUP_4(...)
#define GCMB_IN_GROUP_5(...)
#undef GCMB_CAT_GROUP
#if GCMB_GROUP == 0
#undef GCMB_IN_GROUP_0
#define GCMB_IN_GROUP_0(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#elif GCMB_GROUP == 1
#undef GCMB_IN_GROUP_1
#define GCMB_IN_GROUP_1(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#elif GCMB_GROUP == 2
#undef GCMB_IN_GROUP_2
#define GCMB_IN_GROUP_2(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#elif GCMB_GROUP == 3
#undef GCMB_IN_GROUP_3
#define GCMB_IN_GROUP_3(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#elif GCMB_GROUP == 4
#undef GCMB_IN_GROUP_4
#define GCMB_IN_GROUP_4(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#elif GCMB_GROUP == 5
#undef GCMB_IN_GROUP_5
#define GCMB_IN_GROUP_5(...) GCMB_PATTERN_DEFINE(__VA_ARGS__)
#endif
#define GCMB_PATTERN(GROUP, ...) GCMB_CAT(GCMB_IN_GROUP_, GROUP)(__VA_ARGS__)
#include "patterns.inc"
#undef GCMB_PATTERN
#undef GCMB_PATTERN_DEFINE

// blocks per SM the marching kernel's registers are capped for (see launch_sparse): 6 (80 registers, 24 warps) is
// the best of the variants measured for the fp64 isotropic patterns (profiles/r1_variants.md); the orthotropic ones
// spill 32-40 B at 80 registers and run 8 % faster at 5 (96 registers, no spill): one body 1024^3, 90.9 -> 84.0
// ms/step (isotropic: 82.8 -> 83.6), profiles/r1_carveout.md.  Variants that read the foot cell from the table or
// carry a 7-plane window (border size 3) need the registers too.
template<class P, int BS, bool K0RT>
struct MarchBlocks {
	static constexpr bool ortho3d = P::M == 9 && P::sgn(0) == 1;  // the orthotropic 3-D patterns list (-s, +s, ...)
	static constexpr int value = sizeof(Real) == 4 ? 6 : (BS >= 3 ? 4 : ((ortho3d || (K0RT && P::M == 9)) ? 5 : 6));
};

// marching kernel fed by the cp.async ring (march_async.h): grid = (segments along the sweep axis, z chunks,
// perpendicular axis)
template<int S, class P, int BS, bool K0RT, int LEAD, int MINB, bool ZF>
GCMB_GLOBAL void GCMB_BOUNDS2(MARCH_ZT, MINB) k_stage_march_async(const Args a, int seg) {
	__shared__ Real ring[LEAD + 1][P::M][MARCH_ZT];
	GCMB_DYN_SMEM_RAW(dyn);
	Real* tab = reinterpret_cast<Real*>(dyn);
	GCMB_BLOCK_THREADS(tid) { copy_tables(a, tab, Packed<P, BS, K0RT>::SIZE, tid, MARCH_ZT); }
	__syncthreads();
	const int lo = a.axis == 0 ? a.x_begin : 0;
	const int hi = a.axis == 0 ? a.x_end : a.g.n[1];
	const int s_begin = lo + blockIdx.x * seg;
	const int s_end = min(hi, s_begin + seg);
	const int perp = a.axis == 0 ? (int) blockIdx.z : (int) blockIdx.z + a.x_begin;
	GCMB_BLOCK_THREADS(tid) {
		const int i2 = blockIdx.y * MARCH_ZT + tid;
		if (i2 < a.g.n[2]) { stage_thread_march_async<Real, P, BS, K0RT, LEAD, ZF>(a, ring, tab, tid, perp, i2, s_begin, s_end); }
	}
}

// contiguous-axis kernel: row tiles staged in shared memory by cp.async (ztile.h)
template<int S, class P, int BS, bool K0RT, int ZLEAD>
GCMB_GLOBAL void GCMB_BOUNDS(ZTILE) k_stage_ztile(const Args a, int rows) {
	constexpr int ZRING = ZLEAD + 1;
	typedef ZTileSmem<Real, P::M, ZLEAD> Smem;
	GCMB_DYN_SMEM_RAW(dyn);
	Smem& sm = *reinterpret_cast<Smem*>(dyn);
	Real* tab = reinterpret_cast<Real*>(dyn + sizeof(Smem));
	GCMB_BLOCK_THREADS(tid) { copy_tables(a, tab, Packed<P, BS, K0RT>::SIZE, tid, ZTILE); }
	const int z0 = blockIdx.x * ZTILE;
	const int r0 = blockIdx.y * rows;
	const int r1 = min(a.g.n[1], r0 + rows);
	const int i0 = blockIdx.z + a.x_begin;
	for (int d = 0; d < ZLEAD; d++) {
		GCMB_BLOCK_THREADS(tid) { ztile_issue<Real, P, BS, ZLEAD>(a, sm, d % ZRING, tid, i0, r0 + d, z0, r1); }
	}
	for (int r = r0; r < r1; r++) {
		const int it = r - r0;
		// tile it+ZLEAD goes into the slot read one iteration ago (protected by the barrier below)
		GCMB_BLOCK_THREADS(tid) { ztile_issue<Real, P, BS, ZLEAD>(a, sm, (it + ZLEAD) % ZRING, tid, i0, r + ZLEAD, z0, r1); }
		cp_async_wait<ZLEAD>();
		__syncthreads();
		if (a.zfill && (z0 == 0 || z0 + ZTILE + BS > a.g.n[2])) {  // (the same for all threads of the block)
			GCMB_BLOCK_THREADS(tid) {
				if (tid < BS) { zface_mirror_tile<Real, PatternSets<P>::interp(), P::M, BS, ZROW>(a, sm.v[it % ZRING], ZHALO, z0, ZTILE, tid); }
			}
			__syncthreads();
		}
		GCMB_BLOCK_THREADS(tid) { ztile_compute<Real, P, BS, K0RT, ZLEAD>(a, sm, tab, it % ZRING, tid, i0, r, z0); }
		__syncthreads();
	}
	cp_async_wait<0>();
}

// launch of the marching kernel of a pattern whose direction runs along internal axis 0 or 1
template<class P, int BS, bool K0RT, bool ZF>
void launch_march(const Args& a, cudaStream_t stream, int impl, size_t tab_bytes) {
	// planes per block along the sweep axis.  GCMB_MARCH_SEG: > 0 fixed, 0 the whole axis, unset: 256 (the best at 1024^3,
	// profiles/r1_variants.md), shortened on small grids until the launch has about four waves of blocks -- a 128^3 body
	// with 256-plane segments is 128 blocks for 888 slots and ran at a quarter of the roofline (profiles/r2_small_grids.md);
	// every segment re-reads 2 * border_size planes (from L2 at these sizes), hence not below 16
	static const int seg_env = env_int("GCMB_MARCH_SEG", -1);
	const int len = a.axis == 0 ? a.x_end - a.x_begin : a.g.n[1];
	const int perp = a.axis == 0 ? a.g.n[1] : a.x_end - a.x_begin;
	int seg = seg_env == 0 ? len : seg_env;
	if (seg_env < 0) {
		const long long columns = (long long) ((a.g.n[2] + ZT - 1) / ZT) * perp;   // blocks per segment
		const long long want = 4LL * 148 * 6;
		const long long fit = (long long) len * columns / want;                     // segment length that gives `want` blocks
		seg = (int) (fit > 256 ? 256 : (fit < 16 ? 16 : fit));
	}
	const dim3 grid((unsigned) ((len + seg - 1) / seg), (unsigned) ((a.g.n[2] + ZT - 1) / ZT), (unsigned) perp);
	constexpr int MINB = MarchBlocks<P, BS, K0RT>::value;
#ifndef GCMB_EMUL
	if (impl == 3) {
		// Ring depth and occupancy of the bulk-copy marching kernels (profiles/r2_variants.md, calls 12-31).  A bulk copy lands
		// in shared memory without passing through L1, so -- unlike the cp.async rings, whose requests in flight are bounded by
		// what is left of L1 next to the rings -- the pipeline can be made as deep as shared memory allows.  For the 9-component
		// fp64 patterns with compile-time foot cells the best point measured is 6 slots per warp (5 planes in flight) at 3
		// blocks = 12 warps per SM with the registers uncapped (130, no spills): 27.5 / 26.5 ms per launch at 1024^3 against
		// 28.3 / 27.7 for the cp.async kernel at 24 warps and 29.6 / 29.6 for 3 slots at 5 blocks; 7 slots at 3 blocks and 8 at
		// 2 are within 0.4 ms, 10-12 slots at 2 blocks fall to 30-31 ms.  The kernels with run-time foot cells stay at 3 slots
		// and 5 blocks (deeper was slower at 512^3), the 4- and 5-component and the fp32 patterns on cp.async (launch_sparse).
		// (The wrong values the first 80-register builds gave came from the uninitialised register window, see tma_pipe.h.)
		constexpr bool DEEP = sizeof(Real) == 8 && P::M == 9 && !K0RT;
#if defined(GCMB_TMA_MARCH_MAXB) && defined(GCMB_TMA_MARCH_NST)   // (experiments)
		constexpr int TMINB = MINB > GCMB_TMA_MARCH_MAXB ? GCMB_TMA_MARCH_MAXB : MINB;
		constexpr int NST = GCMB_TMA_MARCH_NST;
#else
		constexpr int TMINB = DEEP ? 3 : (MINB > 5 ? 5 : MINB);
		constexpr int NST = DEEP ? 6 : 3;   // ring slots per warp (planes in flight + the one being consumed)
#endif
		// GCMB_TMA_MARCH: 0 = a pipeline per warp (256-byte copies, no coupling between warps);
		// 1 = one ring per block, refilled by lane 0 of warp 0; 2 = one ring per block, producer warp
		static const int mode = env_int("GCMB_TMA_MARCH", 0);
		static unsigned long long done[3] = {0, 0, 0};
		if (mode == 0) {
			// threads per block of the deep kernels (a pipeline per warp: the block size only sets how many warps share an SM)
#ifndef GCMB_TMA_DEEP_ZT
#define GCMB_TMA_DEEP_ZT 128
#endif
			constexpr int ZTB = DEEP ? GCMB_TMA_DEEP_ZT : MARCH_ZT;
			const dim3 grid_t((unsigned) ((len + seg - 1) / seg), (unsigned) ((a.g.n[2] + ZTB - 1) / ZTB), (unsigned) perp);
			auto kernel = k_stage_march_tma<SET, Real, P, BS, K0RT, NST, 1, false, TMINB, ZF, ZTB>;
			// (ring + the packed tables of all materials: 56 KB + 63 KB for 255 isotropic ones, more for orthotropic tables)
			const size_t smem = sizeof(MarchTmaSmem<Real, P::M, NST, 32, ZTB / 32>) + tab_bytes;
			func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024, done[0]);
			if (smem > 160 * 1024) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); }  // (hundreds of orthotropic materials)
			kernel<<<grid_t, ZTB, smem, stream>>>(a, seg);
#ifdef GCMB_TMA_ALL_MODES
		} else if (mode == 1) {
			auto kernel = k_stage_march_tma<SET, Real, P, BS, K0RT, NST, 4, false, TMINB, ZF>;
			const size_t smem = sizeof(MarchTmaSmem<Real, P::M, NST, 128, 1>) + tab_bytes;
			func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024, done[1]);
			kernel<<<grid, MARCH_ZT, smem, stream>>>(a, seg);
		} else {
			auto kernel = k_stage_march_tma<SET, Real, P, BS, K0RT, NST, 4, true, TMINB, ZF>;
			const size_t smem = sizeof(MarchTmaSmem<Real, P::M, NST, 128, 1>) + tab_bytes;
			func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024, done[2]);
			kernel<<<grid, MARCH_ZT + 32, smem, stream>>>(a, seg);
#endif
		}
		return;
	}
#endif
	auto kernel = k_stage_march_async<SET, P, BS, K0RT, MARCH_LEAD, MINB, ZF>;
	static unsigned long long done = 0;
	if (tab_bytes > 16 * 1024) { func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024, done); }
	static const int carve = env_int("GCMB_MARCH_CARVEOUT", -1);
	static unsigned long long done2 = 0;
	if (carve >= 0) { func_attr_once(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve, done2); }
	GCMB_LAUNCH_COOP(kernel, grid, MARCH_ZT, tab_bytes, stream, a, seg);
}

// launch of the tile kernel of a pattern whose direction runs along the contiguous axis
template<class P, int BS, bool K0RT>
void launch_ztile(const Args& a, cudaStream_t stream, int impl, size_t tab_bytes) {
	static const int rows_env = env_int("GCMB_ZTILE_ROWS", 0);
	// rows per block: 32 for the LDGSTS tiles (round 1), 16 for the bulk-copy tiles (25.0 ms against 25.9 at 32 and 27.4 at 64)
	int rows = rows_env > 0 ? rows_env : (impl == 3 ? 16 : 32);
	if (rows_env <= 0) {
		// small grids: fewer rows per block until the launch has about four waves of blocks (rows are independent: no re-reads)
		const long long tiles = (long long) ((a.g.n[2] + 255) / 256) * (a.x_end - a.x_begin);
		const long long fit = tiles * a.g.n[1] / (4LL * 148 * 3);
		if (fit < rows) { rows = (int) (fit < 4 ? 4 : fit); }
	}
#ifndef GCMB_EMUL
	if (impl == 3) {
		// GCMB_TMA_ZTILE: 0 = a pipeline per warp; 1 = one ring per block of 8 warps
		static const int mode = env_int("GCMB_TMA_ZTILE", 0);
		static unsigned long long done[2] = {0, 0};
		constexpr int NST = 3;
		constexpr int UNIT = 16 / (int) sizeof(Real);
		constexpr int HALO = (BS + UNIT - 1) / UNIT * UNIT;
		constexpr int NW = 8;
		const dim3 grid((unsigned) ((a.g.n[2] + NW * 32 - 1) / (NW * 32)), (unsigned) ((a.g.n[1] + rows - 1) / rows),
		                (unsigned) (a.x_end - a.x_begin));
		static const int carve = env_int("GCMB_ZTILE_CARVEOUT", -1);
		if (mode == 0) {
			auto kernel = k_stage_ztile_tma<SET, Real, P, BS, K0RT, NST, 1, NW, HALO>;
			const size_t smem = sizeof(ZTileTmaSmem<Real, P::M, NST, 32, NW, HALO>) + tab_bytes;
			func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024, done[0]);
			static unsigned long long done2 = 0;
			if (carve >= 0) { func_attr_once(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve, done2); }
			kernel<<<grid, NW * 32, smem, stream>>>(a, rows);
#ifdef GCMB_TMA_ALL_MODES
		} else {
			auto kernel = k_stage_ztile_tma<SET, Real, P, BS, K0RT, NST, NW, NW, HALO>;
			const size_t smem = sizeof(ZTileTmaSmem<Real, P::M, NST, 32 * NW, 1, HALO>) + tab_bytes;
			func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024, done[1]);
			static unsigned long long done2 = 0;
			if (carve >= 0) { func_attr_once(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve, done2); }
			kernel<<<grid, NW * 32, smem, stream>>>(a, rows);
#endif
		}
		return;
	}
#endif
	const dim3 grid((unsigned) ((a.g.n[2] + ZTILE - 1) / ZTILE), (unsigned) ((a.g.n[1] + rows - 1) / rows),
	                (unsigned) (a.x_end - a.x_begin));
	typedef ZTileSmem<Real, P::M, ZLEAD> Smem;
	auto kernel = k_stage_ztile<SET, P, BS, K0RT, ZLEAD>;
	const size_t smem = sizeof(Smem) + tab_bytes;
	static unsigned long long done = 0;
	func_attr_once(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024, done);
	// shared-memory carve-out of the SM in percent of the maximum (the rest is L1); -1 = the driver's choice.
	// The driver sizes it for the most blocks the registers allow (5 x 44 KB for the orthotropic patterns): the
	// tiles arrive by cp.async through L1, and with 28 KB of it left the kernel is 2.5 ms slower per launch at
	// 1024^3.  60 % = 3 tiles' worth: measured 98.2 -> 90.5 ms/step (orthotropic), 83.1 -> 82.8 (isotropic);
	// anything <= 78 % is as good, >= 86 % is the slow mode (profiles/r1_carveout.md)
	static const int carve = env_int("GCMB_ZTILE_CARVEOUT", 60);
	static unsigned long long done2 = 0;
	if (carve >= 0) { func_attr_once(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve, done2); }
	GCMB_LAUNCH_COOP(kernel, grid, ZTILE, smem, stream, a, rows);
}

// A pattern belongs to one direction of one dimensionality, hence to one internal axis: only that axis' kernel
// is compiled for it.
template<class P, int BS, bool K0RT>
void launch_sparse(const void* args, cudaStream_t stream) {
	const Args& a = *static_cast<const Args*>(args);
	// GCMB_STAGE_IMPL: 2 = cp.async (LDGSTS) rings everywhere; 3 = bulk copies (TMA) + mbarriers everywhere; default: the
	// faster of the two per kernel as measured (profiles/r2_variants.md) -- bulk copies for the fp64 tile kernel of the
	// contiguous axis (25.0 against 26.8 ms at 1024^3), cp.async rings for fp32 (17.6 against 18.5 ms)
	static const int impl_env = env_int("GCMB_STAGE_IMPL", 0);
	// ... and, for fp64, bulk copies for the marching kernels that read the foot cell from the table as well (they need
	// 96 registers either way, and there the bulk-copy kernel has no spills: 28.2 / 27.7 against 29.0 / 29.1 ms at Courant 1)
	// and for the marching kernels of the 9-component patterns, with a deep ring at low occupancy (launch_march: 27.5 / 26.5
	// against 28.3 / 27.7 ms).  The 4- and 5-component patterns (acoustic, 2-D) are faster on the cp.async rings
	// (512^3 acoustic 4.53 against 4.72 ms per step, 4096^2 elastic 0.508 against 0.532)
	int impl = impl_env ? impl_env : ((sizeof(Real) == 8 && (P::axis == 2 || K0RT || P::M == 9)) ? 3 : 2);
	// the boundary strips of a decomposed x stage (border_size planes each, launched behind the halo exchange) are too short
	// for a ring that is filled through 2 * border_size warm-up iterations: the cp.async kernel reads its window directly
	if (!impl_env && P::axis == 0 && a.x_end - a.x_begin <= 2 * BS) { impl = 2; }
	const size_t tab_bytes = (size_t) a.n_tables * Packed<P, BS, K0RT>::SIZE * sizeof(Real);
	if (a.axis != P::axis) { return; }  // (the caller matched the pattern by axis)
	if constexpr (P::axis == 2) {
		launch_ztile<P, BS, K0RT>(a, stream, impl, tab_bytes);
	} else if constexpr (P::axis == 1) {
		// the row-writing kernel of the direction before the contiguous one can fill the z ghosts as well
		if (a.zfill) { launch_march<P, BS, K0RT, true>(a, stream, impl, tab_bytes); }
		else { launch_march<P, BS, K0RT, false>(a, stream, impl, tab_bytes); }
	} else {
		launch_march<P, BS, K0RT, false>(a, stream, impl, tab_bytes);
	}
}

template<class P>
void register_pattern() {
	register_sparse_launcher(SET, P::name(), VAR_BS2_K0, &launch_sparse<P, 2, false>);
#ifndef GCMB_QUICK  // (development builds: the headline variant only)
	register_sparse_launcher(SET, P::name(), VAR_BS1, &launch_sparse<P, 1, true>);
	register_sparse_launcher(SET, P::name(), VAR_BS2, &launch_sparse<P, 2, true>);
	register_sparse_launcher(SET, P::name(), VAR_BS3, &launch_sparse<P, 3, true>);
#endif
}

struct Registrar {
	Registrar() {
#define GCMB_REGISTER(NAME, ...) register_pattern<Pat_##NAME>();
#if GCMB_GROUP == 0
#undef GCMB_IN_GROUP_0
#define GCMB_IN_GROUP_0(...) GCMB_REGISTER(__VA_ARGS__)
#elif GCMB_GROUP == 1
#undef GCMB_IN_GROUP_1
#define GCMB_IN_GROUP_1(...) GCMB_REGISTER(__VA_ARGS__)
#elif GCMB_GROUP == 2
#undef GCMB_IN_GROUP_2
#define GCMB_IN_GROUP_2(...) GCMB_REGISTER(__VA_ARGS__)
#elif GCMB_GROUP == 3
#undef GCMB_IN_GROUP_3
#define GCMB_IN_GROUP_3(...) GCMB_REGISTER(__VA_ARGS__)
#elif GCMB_GROUP == 4
#undef GCMB_IN_GROUP_4
#define GCMB_IN_GROUP_4(...) GCMB_REGISTER(__VA_ARGS__)
#elif GCMB_GROUP == 5
#undef GCMB_IN_GROUP_5
#define GCMB_IN_GROUP_5(...) GCMB_REGISTER(__VA_ARGS__)
#endif
#define GCMB_PATTERN(GROUP, ...) GCMB_CAT(GCMB_IN_GROUP_, GROUP)(__VA_ARGS__)
#include "patterns.inc"
#undef GCMB_PATTERN
	}
};
const Registrar g_registrar;
#endif

}  // namespace
}  // namespace gcmb
