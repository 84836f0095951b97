// Stage kernels of gcm_b200 and their registry: one specialised kernel per sparsity class of the
// eigen-system (patterns.inc) x border size {1,2}, plus dense kernels for everything else.
#include <cstdlib>

#include "ztile.h"

namespace gcmb {

#define GCMB_L(...) {__VA_ARGS__}
#define GCMB_PATTERN(NAME, MM, SGN, UM, U1M, BASE, UNEG, U1NEG)                                    \
	struct Pat_##NAME {                                                                           \
		static constexpr int M = MM;                                                              \
		GCMB_HD static constexpr int sgn(int k) { constexpr int t[9] = SGN; return t[k]; }         \
		GCMB_HD static constexpr unsigned um(int k) { constexpr unsigned t[9] = UM; return t[k]; } \
		GCMB_HD static constexpr unsigned u1m(int k) { constexpr unsigned t[9] = U1M; return t[k]; } \
		GCMB_HD static constexpr int base(int k) { constexpr int t[9] = BASE; return t[k]; }       \
		GCMB_HD static constexpr unsigned uneg(int k) { constexpr unsigned t[9] = UNEG; return t[k]; } \
		GCMB_HD static constexpr unsigned u1neg(int k) { constexpr unsigned t[9] = U1NEG; return t[k]; } \
	};
#include "patterns.inc"
#undef GCMB_PATTERN

constexpr int ZT = 128;  // threads per block, all along the contiguous axis

// blocks per SM the marching kernel's registers are capped for (see launch_sparse)
template<class P> struct MarchBlocks { static constexpr int value = 6; };
template<> struct MarchBlocks<Pat_elastic3d_ortho_x> { static constexpr int value = 5; };
template<> struct MarchBlocks<Pat_elastic3d_ortho_y> { static constexpr int value = 5; };
template<> struct MarchBlocks<Pat_elastic3d_ortho_z> { static constexpr int value = 5; };

// block -> node mapping: blockIdx.x runs along the sweep axis when it is strided, so that blocks
// scheduled together share their halo planes in L2
GCMB_DEV bool block_node(const StageArgs& a, int& i0, int& i1, int& i2) {
	int zc;
	if (a.axis == 0) { i0 = blockIdx.x + a.x_begin; zc = blockIdx.y; i1 = blockIdx.z; }
	else if (a.axis == 1) { i1 = blockIdx.x; zc = blockIdx.y; i0 = blockIdx.z + a.x_begin; }
	else { zc = blockIdx.x; i1 = blockIdx.y; i0 = blockIdx.z + a.x_begin; }
	i2 = zc * ZT + threadIdx.x;
	return i2 < a.g.n[2];
}

static dim3 node_blocks(const StageArgs& a) {
	const unsigned zc = (unsigned) ((a.g.n[2] + ZT - 1) / ZT);
	const unsigned nx = (unsigned) (a.x_end - a.x_begin);
	if (a.axis == 0) { return dim3(nx, zc, (unsigned) a.g.n[1]); }
	if (a.axis == 1) { return dim3((unsigned) a.g.n[1], zc, nx); }
	return dim3(zc, (unsigned) a.g.n[1], nx);
}

template<class P, int BS>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_direct(const StageArgs a) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_sparse<P, BS>(a, i0, i1, i2); }
}

template<int M>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense(const StageArgs a) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense<M>(a, i0, i1, i2); }
}

template<int M, int BS>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense_k0(const StageArgs a) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense_k0<M, BS>(a, i0, i1, i2); }
}

template<int M, int BS>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_dense_k0_one(const StageArgs a, const DenseParamCoef<M, BS> co) {
	int i0, i1, i2;
	if (block_node(a, i0, i1, i2)) { stage_thread_dense_k0_one<M, BS>(a, co, i0, i1, i2); }
}

// marching kernel: grid = (segments along the sweep axis, z chunks, perpendicular axis)
template<class P, int BS>
GCMB_GLOBAL void GCMB_BOUNDS(ZT) k_stage_march(const StageArgs a, int seg) {
	const int i2 = blockIdx.y * ZT + threadIdx.x;
	if (i2 >= a.g.n[2]) { return; }
	const int lo = a.axis == 0 ? a.x_begin : 0;
	const int hi = a.axis == 0 ? a.x_end : a.g.n[1];
	const int s_begin = lo + blockIdx.x * seg;
	const int s_end = min(hi, s_begin + seg);
	const int perp = a.axis == 0 ? (int) blockIdx.z : (int) blockIdx.z + a.x_begin;
	stage_thread_march<P, BS>(a, perp, i2, s_begin, s_end);
}

// marching kernel fed by the cp.async ring (march_async.h)
template<class P, int BS, int LEAD, int MINB>
GCMB_GLOBAL void GCMB_BOUNDS2(MARCH_ZT, MINB) k_stage_march_async(const StageArgs a, int seg) {
	__shared__ double ring[LEAD + 1][P::M][MARCH_ZT];
	__shared__ double tab[SMEM_TABLES * Packed<P, BS>::SIZE];
	GCMB_BLOCK_THREADS(tid) { copy_tables(a, tab, Packed<P, BS>::SIZE, tid, MARCH_ZT); }
	__syncthreads();
	const int lo = a.axis == 0 ? a.x_begin : 0;
	const int hi = a.axis == 0 ? a.x_end : a.g.n[1];
	const int s_begin = lo + blockIdx.x * seg;
	const int s_end = min(hi, s_begin + seg);
	const int perp = a.axis == 0 ? (int) blockIdx.z : (int) blockIdx.z + a.x_begin;
	GCMB_BLOCK_THREADS(tid) {
		const int i2 = blockIdx.y * MARCH_ZT + tid;
		if (i2 < a.g.n[2]) { stage_thread_march_async<P, BS, LEAD>(a, ring, tab, tid, perp, i2, s_begin, s_end); }
	}
}

// contiguous-axis kernel: row tiles staged in shared memory by cp.async (ztile.h)
template<class P, int BS, int ZLEAD>
GCMB_GLOBAL void GCMB_BOUNDS(ZTILE) k_stage_ztile(const StageArgs a, int rows) {
	constexpr int ZRING = ZLEAD + 1;
	typedef ZTileSmem<P::M, Packed<P, BS>::SIZE, ZLEAD> Smem;
	GCMB_DYN_SMEM(Smem, sm);
	GCMB_BLOCK_THREADS(tid) { copy_tables(a, sm.tab, Packed<P, BS>::SIZE, tid, ZTILE); }
	const int z0 = blockIdx.x * ZTILE;
	const int r0 = blockIdx.y * rows;
	const int r1 = min(a.g.n[1], r0 + rows);
	const int i0 = blockIdx.z + a.x_begin;
	for (int d = 0; d < ZLEAD; d++) {
		GCMB_BLOCK_THREADS(tid) { ztile_issue<P, BS, ZLEAD>(a, sm, d % ZRING, tid, i0, r0 + d, z0, r1); }
	}
	for (int r = r0; r < r1; r++) {
		const int it = r - r0;
		// tile it+ZLEAD goes into the slot read one iteration ago (protected by the barrier below)
		GCMB_BLOCK_THREADS(tid) { ztile_issue<P, BS, ZLEAD>(a, sm, (it + ZLEAD) % ZRING, tid, i0, r + ZLEAD, z0, r1); }
		cp_async_wait<ZLEAD>();
		__syncthreads();
		GCMB_BLOCK_THREADS(tid) { ztile_compute<P, BS, ZLEAD>(a, sm, it % ZRING, tid, i0, r, z0); }
		__syncthreads();
	}
	cp_async_wait<0>();
}

static int env_int(const char* name, int dflt) {
	const char* v = getenv(name);
	return v ? atoi(v) : dflt;
}

template<class P, int BS>
static void launch_sparse(const StageArgs& a, cudaStream_t stream) {
	// 0 = one thread per node; 1 = marching, register prefetch; 2 = marching, cp.async ring (default)
	static const int impl = env_int("GCMB_STAGE_IMPL", 2);
	static const int seg_env = env_int("GCMB_MARCH_SEG", 256);
	const bool tables_fit = a.packed && a.n_tables <= SMEM_TABLES;
	if (impl >= 1 && a.axis != 2 && (impl == 1 || tables_fit)) {
		const int len = a.axis == 0 ? a.x_end - a.x_begin : a.g.n[1];
		const int seg = seg_env < 1 ? len : seg_env;
		const int perp = a.axis == 0 ? a.g.n[1] : a.x_end - a.x_begin;
		const dim3 grid((unsigned) ((len + seg - 1) / seg), (unsigned) ((a.g.n[2] + ZT - 1) / ZT), (unsigned) perp);
		if (impl == 2) {
			// 2 planes in flight per thread.  Resident blocks per SM the registers are capped for: 6 (80 registers,
			// 24 warps) is the best of the variants measured for the isotropic patterns (profiles/r1_variants.md); the
			// orthotropic ones spill 32-40 B at 80 registers and run 8 % faster at 5 (96 registers, no spill): one body
			// 1024^3, 90.9 -> 84.0 ms/step (isotropic: 82.8 -> 83.6), profiles/r1_carveout.md
			static const int minb_env = env_int("GCMB_MARCH_MINB", -1);
			const int minb = minb_env > 0 ? minb_env : MarchBlocks<P>::value;
			if (P::M == 9 && minb == 5) {
				auto kernel5 = k_stage_march_async<P, BS, MARCH_LEAD, 5>;
				GCMB_LAUNCH_COOP(kernel5, grid, MARCH_ZT, 0, stream, a, seg);
				return;
			}
			auto kernel = k_stage_march_async<P, BS, MARCH_LEAD, 6>;
			static const int carve = env_int("GCMB_MARCH_CARVEOUT", -1);
			static const cudaError_t attr2 = carve >= 0 ? cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve) : cudaSuccess;
			(void) attr2;
			GCMB_LAUNCH_COOP(kernel, grid, MARCH_ZT, 0, stream, a, seg);
		} else {
			auto kernel = k_stage_march<P, BS>;
			GCMB_LAUNCH(kernel, grid, ZT, stream, a, seg);
		}
	} else if (impl == 2 && a.axis == 2 && tables_fit) {
		static const int rows = env_int("GCMB_ZTILE_ROWS", 32);
		const dim3 grid((unsigned) ((a.g.n[2] + ZTILE - 1) / ZTILE), (unsigned) ((a.g.n[1] + rows - 1) / rows),
		                (unsigned) (a.x_end - a.x_begin));
		typedef ZTileSmem<P::M, Packed<P, BS>::SIZE, ZLEAD> Smem;
		auto kernel = k_stage_ztile<P, BS, ZLEAD>;
		// experiment knob: extra dynamic shared memory per block (lowers the number of resident blocks per SM)
		static const int pad = env_int("GCMB_ZTILE_SMEM_PAD", 0);
		static const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem) + pad);
		(void) attr;
		// shared-memory carve-out of the SM in percent of the maximum (the rest is L1); -1 = the driver's choice.
		// The driver sizes it for the most blocks the registers allow (5 x 44 KB for the orthotropic patterns): the
		// tiles arrive by cp.async through L1, and with 28 KB of it left the kernel is 2.5 ms slower per launch at
		// 1024^3.  60 % = 3 tiles' worth: measured 98.2 -> 90.5 ms/step (orthotropic), 83.1 -> 82.8 (isotropic);
		// anything <= 78 % is as good, >= 86 % is the slow mode (profiles/r1_carveout.md)
		static const int carve = env_int("GCMB_ZTILE_CARVEOUT", 60);
		static const cudaError_t attr2 = carve >= 0 ? cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve) : cudaSuccess;
		(void) attr2;
		GCMB_LAUNCH_COOP(kernel, grid, ZTILE, sizeof(Smem) + pad, stream, a, rows);
	} else {
		auto kernel = k_stage_direct<P, BS>;
		GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a);
	}
}

template<int M>
static void launch_dense(const StageArgs& a, cudaStream_t stream) {
	auto kernel = k_stage_dense<M>;
	GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a);
}

template<int M, int BS>
static void launch_dense_k0(const StageArgs& a, cudaStream_t stream) {
	if (a.n_tables == 1 && a.host_tables) {
		const StageTable& T = a.host_tables[a.dir];
		DenseParamCoef<M, BS> co;
		for (int k = 0; k < M; k++) {
			for (int j = 0; j < M; j++) { co.U[k * M + j] = T.U[k * M + j]; co.U1[k * M + j] = T.U1[k * M + j]; }
			for (int i = 0; i < BS; i++) { co.F[k * BS + i] = T.F[k * MAXBS + i]; }
			co.sd[k] = T.F[k * MAXBS] == 0.0 ? 0 : T.dir[k];
		}
		auto kernel = k_stage_dense_k0_one<M, BS>;
		GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a, co);
		return;
	}
	auto kernel = k_stage_dense_k0<M, BS>;
	GCMB_LAUNCH(kernel, node_blocks(a), ZT, stream, a);
}

#define GCMB_PATTERN(NAME, MM, SGN, UM, U1M, BASE, UNEG, U1NEG) \
	{#NAME, MM, SGN, UM, U1M, BASE, UNEG, U1NEG, &launch_sparse<Pat_##NAME, 1>, &launch_sparse<Pat_##NAME, 2>},
static const PatternInfo g_patterns[] = {
#include "patterns.inc"
};
#undef GCMB_PATTERN

int pattern_count() { return (int) (sizeof(g_patterns) / sizeof(g_patterns[0])); }
const PatternInfo& pattern(int i) { return g_patterns[i]; }

StageLauncher dense_launcher(int M) {
	switch (M) {
		case 2: return &launch_dense<2>;
		case 3: return &launch_dense<3>;
		case 4: return &launch_dense<4>;
		case 5: return &launch_dense<5>;
		case 9: return &launch_dense<9>;
		default: return nullptr;
	}
}

// dense eigen-system with every foot in the first cell (the caller checks that): stencil held in registers
StageLauncher dense_k0_launcher(int M, int bs) {
	if (bs != 1 && bs != 2) { return nullptr; }
	switch (M) {
		case 2: return bs == 1 ? &launch_dense_k0<2, 1> : &launch_dense_k0<2, 2>;
		case 3: return bs == 1 ? &launch_dense_k0<3, 1> : &launch_dense_k0<3, 2>;
		case 4: return bs == 1 ? &launch_dense_k0<4, 1> : &launch_dense_k0<4, 2>;
		case 5: return bs == 1 ? &launch_dense_k0<5, 1> : &launch_dense_k0<5, 2>;
		case 9: return bs == 1 ? &launch_dense_k0<9, 1> : &launch_dense_k0<9, 2>;
		default: return nullptr;
	}
}

}  // namespace gcmb
