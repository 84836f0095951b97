// Marching stage kernel with an asynchronous-copy pipeline (strided axes: internal axis 0 or 1).
//
// One thread owns one z and marches along the sweep axis, keeping the 2*BS+1 values of every interpolated
// component in registers, so that each value is read from HBM exactly once per stage (plus the 2*BS planes
// re-read at segment starts) and every warp access is a contiguous row segment.  Every thread keeps LEAD
// planes in flight with cp.async (LDGSTS) into a small shared-memory ring: each thread copies and later
// reads only ITS OWN slots, so no block barrier is needed, only cp.async.wait_group.  With LEAD planes of
// 9 components in flight per thread the bytes in flight per SM cover the HBM latency-bandwidth product at
// 16 warps/SM, which a one-plane register prefetch could not (profiles/r1_run1: 61 % of the stall samples
// sat on the first use of the prefetched plane).
//
// The kernel can also fill the ghost nodes of the two z faces of the layer it writes (StageArgsT::zfill): the
// edge lanes of a row hold the values the ghosts mirror, a warp shuffle brings them to the lanes that write
// the 32-byte sector next to the row, and the whole sector (ghosts + row padding) goes out in one store, so
// that no partially written sector ever reaches HBM -- the separate ghost-fill kernel this replaces read 10x
// its useful bytes (profiles/r1_*: k_border, 2.7 ms of an 86 ms step).
#pragma once
#include "thread_fns.h"

namespace gcmb {

#if defined(__CUDA_ARCH__)
template<int BYTES>
GCMB_DEV void cp_async_bytes(void* smem_dst, const void* gmem_src) {
	const unsigned dst = (unsigned) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" :: "r"(dst), "l"(gmem_src), "n"(BYTES) : "memory");
}
GCMB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template<int N>
GCMB_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
#else
// stepping harness: the copy is immediate
template<int BYTES>
inline void cp_async_bytes(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, BYTES); }
inline void cp_async_commit() { }
template<int N>
inline void cp_async_wait() { }
#endif
template<class R>
GCMB_HD void cp_async_real(R* smem_dst, const R* gmem_src) { cp_async_bytes<(int) sizeof(R)>(smem_dst, gmem_src); }

constexpr int MARCH_ZT = 128;  // threads per block = z extent of a block
constexpr int MARCH_LEAD = 2;  // planes in flight ahead of the one being consumed

// phase 0 of the pipelined kernels: copy the packed coefficient tables of all materials to shared memory
template<class R>
GCMB_HD void copy_tables(const StageArgsT<R>& a, R* tab, int table_size, int tid, int n_threads) {
	const int n = a.n_tables * table_size;
	for (int i = tid; i < n; i += n_threads) { tab[i] = a.packed[i]; }
}

// value of lane `src` (device: warp shuffle; stepping harness: the caller passes the row's values)
template<class R>
GCMB_DEV R lane_value(R v, int src) {
#if defined(__CUDA_ARCH__)
	return __shfl_sync(0xffffffffu, v, src);
#else
	(void) src;
	return v;
#endif
}

// Ghost fill of the z faces of the row a warp has just computed (see the header).  `out` = the thread's new
// values at z = i2, `row0` = element of z = 0 of the row in the written layer.  Called by ALL lanes of a full
// warp (the launcher guarantees n2 % 32 == 0).
template<class R, int M>
GCMB_DEV void zface_fill_warp(const StageArgsT<R>& a, const R (&out)[M], long long row0, int i2, int lane) {
#if defined(__CUDA_ARCH__)
	constexpr int SECT = 32 / (int) sizeof(R);  // elements per 32-byte sector
	const Geom& g = a.g;
	const int bs = g.bs;
	if (a.zf.on[0] && i2 - lane == 0) {          // the warp that holds z = 0..31
		const int e = lane;                      // element zoff - SECT + e of the row
		const int dist = SECT - e;               // ghost(-dist) mirrors inner(+dist)
#pragma unroll
		for (int c = 0; c < M; c++) {
			R v = lane_value(out[c], dist & 31);
			if ((a.zf.set[0] >> c) & 1u) { v = -v + a.zf.add[0][c]; }
			if (dist > bs) { v = R(0); }         // row padding: written so that the sector is complete
			if (e < SECT) { a.nxt[c * g.comp + row0 - SECT + e] = v; }
		}
	}
	if (a.zf.on[1] && i2 - lane == g.n[2] - 32) {  // the warp that holds the last 32 nodes
		const int e = lane - (32 - SECT);          // element n2 + e
		const int dist = e + 1;                    // ghost(n2-1+dist) mirrors inner(n2-1-dist)
#pragma unroll
		for (int c = 0; c < M; c++) {
			R v = lane_value(out[c], (30 - e) & 31);
			if ((a.zf.set[1] >> c) & 1u) { v = -v + a.zf.add[1][c]; }
			if (dist > bs) { v = R(0); }
			if (e >= 0) { a.nxt[c * g.comp + row0 + g.n[2] + e] = v; }
		}
	}
#endif
}

// ring[slot][component][thread]; slot of iteration s' holds: interpolated components at plane s'+BS,
// centre-only components at plane s'
template<class R, class P, int BS, bool K0RT, int LEAD, bool ZF>
GCMB_HD void stage_thread_march_async(const StageArgsT<R>& a, R (*ring)[P::M][MARCH_ZT], const R* tab,
                                      int tid, int perp, int i2, int s_begin, int s_end) {
	constexpr int M = P::M;
	constexpr int W = 2 * BS + 1;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	const Geom& g = a.g;
	const long long sstride = g.stride(a.axis);
	const long long idx0 = a.axis == 0 ? g.index(s_begin, perp, i2) : g.index(perp, s_begin, i2);
	const R* __restrict__ cur = a.cur;
	constexpr int MARCH_LEAD = LEAD;
	constexpr int MARCH_RING = LEAD + 1;

	auto issue = [&](int s) {  // all copies of iteration s (s may run past the segment: then nothing)
		if (s < s_end) {
			const long long idx = idx0 + (long long) (s - s_begin) * sstride;
			R (*slot)[MARCH_ZT] = ring[(s - s_begin) % MARCH_RING];
#pragma unroll
			for (int j = 0; j < M; j++) {
				if ((IC >> j) & 1u) { cp_async_real(&slot[j][tid], cur + j * g.comp + idx + (long long) BS * sstride); }
				else if ((CC >> j) & 1u) { cp_async_real(&slot[j][tid], cur + j * g.comp + idx); }
			}
		}
		cp_async_commit();
	};

	// prologue: LEAD iterations in flight, the left part of the window straight into registers
#pragma unroll
	for (int d = 0; d < MARCH_LEAD; d++) { issue(s_begin + d); }
	R w[M][W];  // w[j][BS + o] = component j at (s + o); only rows in IC are live
	R cv[M];
#pragma unroll
	for (int j = 0; j < M; j++) {
		if ((IC >> j) & 1u) {
#pragma unroll
			for (int o = 1; o < W; o++) { w[j][o] = GCMB_LDG(cur + j * g.comp + idx0 + (long long) (o - 1 - BS) * sstride); }
		}
	}
	int tn = a.node_table[idx0];

	for (int s = s_begin; s < s_end; s++) {
		const long long idx = idx0 + (long long) (s - s_begin) * sstride;
		const int t = tn;
		if (s + 1 < s_end) { tn = a.node_table[idx + sstride]; }
		cp_async_wait<MARCH_LEAD - 1>();  // the copies of iteration s have landed
		R (*slot)[MARCH_ZT] = ring[(s - s_begin) % MARCH_RING];
#pragma unroll
		for (int j = 0; j < M; j++) {
			if ((IC >> j) & 1u) {
#pragma unroll
				for (int o = 0; o < W - 1; o++) { w[j][o] = w[j][o + 1]; }
				w[j][W - 1] = slot[j][tid];
			} else if ((CC >> j) & 1u) {
				cv[j] = slot[j][tid];
			}
		}
		// refill: iteration s+LEAD goes into the slot consumed one iteration ago (RING = LEAD + 1)
		issue(s + MARCH_LEAD);
		auto load = [&](int j, int o) -> R { return ((IC >> j) & 1u) ? w[j][BS + o] : cv[j]; };
		R out[M];
		gcm_node_sparse<R, P, BS, K0RT>(PackedCoef<R, P, BS, K0RT>{tab + t * Packed<P, BS, K0RT>::SIZE}, load, out);
#pragma unroll
		for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
		if (ZF && a.zfill) { zface_fill_warp<R, M>(a, out, idx - i2, i2, tid & 31); }
	}
	cp_async_wait<0>();
}

}  // namespace gcmb
