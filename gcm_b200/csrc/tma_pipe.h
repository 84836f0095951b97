// Stage kernels fed by 1-D bulk asynchronous copies (TMA: cp.async.bulk, SASS UBLKCP) completed on
// shared-memory mbarriers (SASS SYNCS) -- the Blackwell replacement for the per-thread 8-byte LDGSTS
// pipelines of march_async.h / ztile.h.  One elected lane issues ONE copy per (component, row segment)
// instead of 32 lanes issuing 32 addresses each; consumers sleep on the mbarrier's phase instead of stalling
// on cp.async.wait_group; there is no __syncthreads in any loop.
//
// A "group" of GW warps shares a ring of NSTAGE slots with one full/empty mbarrier pair per slot:
//   GW == 1   every warp runs a pipeline of its own over its 32 nodes (no coupling between warps at all:
//             the lane that issues the refill belongs to the warp that has just read the slot);
//   GW == all the block shares the ring (fewer, larger copies); the refill is issued by lane 0 of the
//             block's extra producer warp (PRODUCER) or by lane 0 of warp 0 after the empty barrier.
// Device-only: the stepping harness (tests/emul) keeps using the LDGSTS kernels for the index math, the
// results of both kernel families are compared bit for bit on the GPU (tests/test_gpu_parity.py).
#pragma once
#include "march_async.h"

#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
namespace gcmb {

GCMB_DEV unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
GCMB_DEV void mbar_init(unsigned long long* bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
GCMB_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
GCMB_DEV void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
GCMB_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GCMB_DEV void mbar_arrive(unsigned long long* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
GCMB_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
	unsigned done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
		             : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while (!done);
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completing on `bar`
GCMB_DEV void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
	             :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// marching kernel (strided axes): ring slot of iteration s' holds the interpolated components at plane
// s'+BS and the centre-only components and material ids at plane s'.  The window's left part is streamed
// through the same ring by 2*BS warm-up iterations, so every value comes through TMA.
// ---------------------------------------------------------------------------------------------
template<class R, int M, int NSTAGE, int ZW /* nodes per group */, int NG /* groups per block */>
struct MarchTmaSmem {
	alignas(128) R v[NG][NSTAGE][M][ZW];
	alignas(16) uint8_t id[NG][NSTAGE][ZW];
	alignas(8) unsigned long long full[NG][NSTAGE];
	alignas(8) unsigned long long empty[NG][NSTAGE];
};

template<int S, class R, class P, int BS, bool K0RT, int NSTAGE, int GW, bool PRODUCER, int MINB, bool ZF, int ZTB = MARCH_ZT>
__global__ void __launch_bounds__(ZTB + (PRODUCER ? 32 : 0), MINB) k_stage_march_tma(const StageArgsT<R> a, int seg) {
	constexpr int M = P::M;
	constexpr int W = 2 * BS + 1;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	constexpr int NWARP = ZTB / 32;   // ZTB = threads per block (with a pipeline per warp it only decides which warps share an SM)
	constexpr int NG = NWARP / GW;
	constexpr int ZW = 32 * GW;
	static_assert(!PRODUCER || GW == NWARP, "the producer warp serves one block-wide ring");
	typedef MarchTmaSmem<R, M, NSTAGE, ZW, NG> Smem;
	extern __shared__ __align__(128) unsigned char gcmb_dyn_smem_[];
	Smem& sm = *reinterpret_cast<Smem*>(gcmb_dyn_smem_);
	R* tab = reinterpret_cast<R*>(gcmb_dyn_smem_ + sizeof(Smem));
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	copy_tables(a, tab, Packed<P, BS, K0RT>::SIZE, tid, (int) blockDim.x);
	if (tid == 0) {
		for (int gi = 0; gi < NG; gi++) {
			for (int i = 0; i < NSTAGE; i++) { mbar_init(&sm.full[gi][i], 1); mbar_init(&sm.empty[gi][i], GW); }
		}
		mbar_init_fence();
	}
	__syncthreads();

	const Geom& g = a.g;
	const int lo = a.axis == 0 ? a.x_begin : 0;
	const int hi = a.axis == 0 ? a.x_end : g.n[1];
	const int s_begin = lo + blockIdx.x * seg;
	const int s_end = min(hi, s_begin + seg);
	const int perp = a.axis == 0 ? (int) blockIdx.z : (int) blockIdx.z + a.x_begin;
	const int n_it = s_end - s_begin + 2 * BS;   // iteration `it` works on s = s_begin - 2*BS + it
	const long long sstride = g.stride(a.axis);
	const int zb = blockIdx.y * ZTB;

	// all copies of iteration `it` of group `gi` (the caller has made sure the slot is free)
	auto issue = [&](int gi, int it) {
		const int s = s_begin - 2 * BS + it;
		const int slot = it % NSTAGE;
		const long long idx = a.axis == 0 ? g.index(s, perp, zb + gi * ZW) : g.index(perp, s, zb + gi * ZW);
		const bool centre = s >= s_begin;
		constexpr unsigned row = ZW * (unsigned) sizeof(R);
		const unsigned bytes = (unsigned) popcount_u(IC) * row + (centre ? (unsigned) popcount_u(CC) * row + ZW : 0u);
		unsigned long long* bar = &sm.full[gi][slot];
		mbar_expect_tx(bar, bytes);
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((IC >> j) & 1u) { bulk_g2s(sm.v[gi][slot][j], a.cur + j * g.comp + idx + (long long) BS * sstride, row, bar); }
			else if (((CC >> j) & 1u) && centre) { bulk_g2s(sm.v[gi][slot][j], a.cur + j * g.comp + idx, row, bar); }
		}
		if (centre) { bulk_g2s(sm.id[gi][slot], a.node_table + idx, ZW, bar); }
	};

	if (PRODUCER && warp == NWARP) {
		if (lane == 0) {
			for (int it = 0; it < n_it; it++) {
				if (it >= NSTAGE) { mbar_wait(&sm.empty[0][it % NSTAGE], (unsigned) ((it / NSTAGE - 1) & 1)); }
				issue(0, it);
			}
		}
		return;
	}

	const int gi = warp / GW;
	const int zl = tid - gi * ZW;                      // position inside the group's row segment
	const int i2 = zb + tid;
	const bool live = i2 < g.n[2];
	// a warp whose 32 nodes all lie past the end of the row has nothing to copy or compute (rows shorter than the block:
	// small grids); with a pipeline per warp nobody waits for it
	if (GW == 1 && !PRODUCER && zb + gi * ZW >= g.n[2]) { return; }
	const bool issuer = !PRODUCER && zl == 0;          // lane 0 of the group's first warp
	if (issuer) {
		for (int it = 0; it < NSTAGE - 1 && it < n_it; it++) { issue(gi, it); }
	}
	R w[M][W];
	R cv[M];
	// The warm-up iterations rotate the window before it is full: it must hold DEFINED values.  With the window left
	// uninitialised ptxas kept its oldest plane in UNIFORM registers whenever the register cap was tight (R2UR of a
	// per-lane value, then DADD/FSEL with the UR operand: every lane computed with lane 0's plane) -- the wrong results
	// of the 80-register builds in profiles/r2_debug_tma.log; found in the SASS, profiles/r2_variants.md "Call 12-15".
#pragma unroll
	for (int j = 0; j < M; j++) {
#pragma unroll
		for (int o = 0; o < W; o++) { w[j][o] = R(0); }
	}
	const long long idx_first = a.axis == 0 ? g.index(s_begin, perp, i2) : g.index(perp, s_begin, i2);
	for (int it = 0; it < n_it; it++) {
		const int slot = it % NSTAGE;
		const int s = s_begin - 2 * BS + it;
		if (issuer && it + NSTAGE - 1 < n_it) {
			// refill the slot read in the previous iteration: free once every warp of the group has passed it
			const int nit = it + NSTAGE - 1;
			if (GW > 1 && nit >= NSTAGE) { mbar_wait(&sm.empty[gi][nit % NSTAGE], (unsigned) ((nit / NSTAGE - 1) & 1)); }
			issue(gi, nit);
		}
		mbar_wait(&sm.full[gi][slot], (unsigned) ((it / NSTAGE) & 1));
		int t = 0;
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((IC >> j) & 1u) {
#pragma unroll
				for (int o = 0; o < W - 1; o++) { w[j][o] = w[j][o + 1]; }
				w[j][W - 1] = sm.v[gi][slot][j][zl];
			} else if ((CC >> j) & 1u) {
				cv[j] = sm.v[gi][slot][j][zl];
			}
		}
		if (s >= s_begin) { t = sm.id[gi][slot][zl]; }
		__syncwarp();
		if ((GW > 1 || PRODUCER) && lane == 0) { mbar_arrive(&sm.empty[gi][slot]); }
		if (s < s_begin || !live) { continue; }
		const long long idx = idx_first + (long long) (s - s_begin) * sstride;
		auto load = [&](int j, int o) -> R { return ((IC >> j) & 1u) ? w[j][BS + o] : cv[j]; };
		R out[M];
		gcm_node_sparse<R, P, BS, K0RT>(PackedCoef<R, P, BS, K0RT>{tab + t * Packed<P, BS, K0RT>::SIZE}, load, out);
#pragma unroll
		for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
		if (ZF && a.zfill) { zface_fill_warp<R, M>(a, out, idx - i2, i2, lane); }
	}
}

// ---------------------------------------------------------------------------------------------
// contiguous-axis kernel: a group walks over consecutive rows; every row segment (its nodes plus HALO nodes
// on both sides) arrives by one bulk copy per component
// ---------------------------------------------------------------------------------------------
template<class R> struct ZHaloOf { static constexpr int value = 16 / (int) sizeof(R) >= 4 ? 4 : 2; };  // >= 16 bytes; covers BS <= 2 (double) / 4

template<class R, int M, int NSTAGE, int ZW, int NG, int HALO>
struct ZTileTmaSmem {
	alignas(128) R v[NG][NSTAGE][M][ZW + 2 * HALO];
	alignas(16) uint8_t id[NG][NSTAGE][ZW];
	alignas(8) unsigned long long full[NG][NSTAGE];
	alignas(8) unsigned long long empty[NG][NSTAGE];
};

template<int S, class R, class P, int BS, bool K0RT, int NSTAGE, int GW, int NWARP, int HALO>
__global__ void __launch_bounds__(NWARP * 32) k_stage_ztile_tma(const StageArgsT<R> a, int rows) {
	constexpr int M = P::M;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	constexpr int NG = NWARP / GW;
	constexpr int ZW = 32 * GW;
	static_assert(HALO >= BS && (HALO * sizeof(R)) % 16 == 0, "halo: whole 16-byte units");
	typedef ZTileTmaSmem<R, M, NSTAGE, ZW, NG, HALO> Smem;
	extern __shared__ __align__(128) unsigned char gcmb_dyn_smem_[];
	Smem& sm = *reinterpret_cast<Smem*>(gcmb_dyn_smem_);
	R* tab = reinterpret_cast<R*>(gcmb_dyn_smem_ + sizeof(Smem));
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	copy_tables(a, tab, Packed<P, BS, K0RT>::SIZE, tid, (int) blockDim.x);
	if (tid == 0) {
		for (int gi = 0; gi < NG; gi++) {
			for (int i = 0; i < NSTAGE; i++) { mbar_init(&sm.full[gi][i], 1); mbar_init(&sm.empty[gi][i], GW); }
		}
		mbar_init_fence();
	}
	__syncthreads();

	const Geom& g = a.g;
	const int zb = blockIdx.x * (NWARP * 32);
	const int r0 = blockIdx.y * rows;
	const int r1 = min(g.n[1], r0 + rows);
	const int i0 = blockIdx.z + a.x_begin;
	const int n_it = r1 - r0;
	const int gi = warp / GW;
	const int zl = tid - gi * ZW;
	const int i2 = zb + tid;
	const bool live = i2 < g.n[2];
	const bool issuer = zl == 0;
	const int seg = zb + gi * ZW;  // first node of the group's part of the row
	static_assert(GW == 1 || NG == 1, "the in-tile ghost fill synchronises a warp or the whole block");
	if (GW == 1 && seg >= g.n[2]) { return; }   // (a warp past the end of the row: nothing to copy or compute)

	auto issue = [&](int it) {
		const int slot = it % NSTAGE;
		const long long row = g.index(i0, r0 + it, seg);
		constexpr unsigned wide = (ZW + 2 * HALO) * (unsigned) sizeof(R);
		constexpr unsigned bytes = (unsigned) popcount_u(IC | CC) * wide + ZW;
		unsigned long long* bar = &sm.full[gi][slot];
		mbar_expect_tx(bar, bytes);
#pragma unroll
		for (int j = 0; j < M; j++) {
			if (((IC | CC) >> j) & 1u) { bulk_g2s(sm.v[gi][slot][j], a.cur + j * g.comp + row - HALO, wide, bar); }
		}
		bulk_g2s(sm.id[gi][slot], a.node_table + row, ZW, bar);
	};

	if (issuer) {
		for (int it = 0; it < NSTAGE - 1 && it < n_it; it++) { issue(it); }
	}
	for (int it = 0; it < n_it; it++) {
		const int slot = it % NSTAGE;
		if (issuer && it + NSTAGE - 1 < n_it) {
			const int nit = it + NSTAGE - 1;
			if (GW > 1 && nit >= NSTAGE) { mbar_wait(&sm.empty[gi][nit % NSTAGE], (unsigned) ((nit / NSTAGE - 1) & 1)); }
			issue(nit);
		}
		mbar_wait(&sm.full[gi][slot], (unsigned) ((it / NSTAGE) & 1));
		if (a.zfill && (seg == 0 || seg + ZW + BS > g.n[2])) {  // (the same for all threads of the group)
			// border condition of the z faces: ghost nodes of the staged row from its own inner nodes (ztile.h)
			if (zl < BS) {
				zface_mirror_tile<R, IC, M, BS, ZW + 2 * HALO>(a, sm.v[gi][slot], HALO, seg, ZW, zl);
				// generic-proxy writes into a buffer that the next bulk copy (async proxy) of this slot overwrites
				fence_proxy_async_smem();
			}
			if (GW == 1) { __syncwarp(); } else { __syncthreads(); }
		}
		if (live) {
			const long long idx = g.index(i0, r0 + it, i2);
			const R* tb = tab + (int) sm.id[gi][slot][zl] * Packed<P, BS, K0RT>::SIZE;
			const R (*v)[ZW + 2 * HALO] = sm.v[gi][slot];
			auto load = [&](int j, int o) -> R { return v[j][zl + HALO + o]; };
			R out[M];
			gcm_node_sparse<R, P, BS, K0RT>(PackedCoef<R, P, BS, K0RT>{tb}, load, out);
#pragma unroll
			for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
		}
		__syncwarp();
		if (GW > 1 && lane == 0) { mbar_arrive(&sm.empty[gi][slot]); }
	}
}

}  // namespace gcmb
#endif
