// Marching stage kernel with an asynchronous-copy pipeline (strided axes: internal axis 0 or 1).
//
// Same arithmetic and the same register window as stage_thread_march (thread_fns.h); what changes is how
// the next planes reach the thread.  Every thread owns one z and keeps LEAD planes in flight with
// cp.async (LDGSTS) into a small shared-memory ring: each thread copies and later reads only ITS OWN
// 8-byte slots, so no block barrier is needed, only cp.async.wait_group.  With LEAD planes of 9 components
// in flight per thread the bytes in flight per SM cover the HBM latency-bandwidth product at 16 warps/SM,
// which the one-plane register prefetch could not (profiles/r1_run1: 61 % of the stall samples sat on the
// first use of the prefetched plane).
#pragma once
#include "thread_fns.h"

namespace gcmb {

#if defined(__CUDA_ARCH__)
GCMB_DEV void cp_async_f64(double* smem_dst, const double* gmem_src) {
	const unsigned dst = (unsigned) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(dst), "l"(gmem_src) : "memory");
}
GCMB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template<int N>
GCMB_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
#else
// stepping harness: the copy is immediate
inline void cp_async_f64(double* smem_dst, const double* gmem_src) { *smem_dst = *gmem_src; }
inline void cp_async_commit() { }
template<int N>
inline void cp_async_wait() { }
#endif

constexpr int MARCH_ZT = 128;  // threads per block = z extent of a block
constexpr int MARCH_LEAD = 2;  // planes in flight ahead of the one being consumed
constexpr int SMEM_TABLES = 16; // material tables kept in shared memory by the pipelined kernels

// phase 0 of the pipelined kernels, one thread: copy the packed coefficient tables to shared memory
GCMB_HD void copy_tables(const StageArgs& a, double* tab, int table_size, int tid, int n_threads) {
	const int n = a.n_tables * table_size;
	for (int i = tid; i < n; i += n_threads) { tab[i] = a.packed[i]; }
}

// ring[slot][component][thread]; slot of iteration s' holds: interpolated components at plane s'+BS,
// centre-only components at plane s'
template<class P, int BS, int LEAD>
GCMB_HD void stage_thread_march_async(const StageArgs& a, double (*ring)[P::M][MARCH_ZT], const double* tab,
                                      int tid, int perp, int i2, int s_begin, int s_end) {
	constexpr int M = P::M;
	constexpr int W = 2 * BS + 1;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	const Geom& g = a.g;
	const long long sstride = g.stride(a.axis);
	const long long idx0 = a.axis == 0 ? g.index(s_begin, perp, i2) : g.index(perp, s_begin, i2);
	const double* __restrict__ cur = a.cur;
	constexpr int MARCH_LEAD = LEAD;
	constexpr int MARCH_RING = LEAD + 1;

	auto issue = [&](int s) {  // all copies of iteration s (s may run past the segment: then nothing)
		if (s < s_end) {
			const long long idx = idx0 + (long long) (s - s_begin) * sstride;
			double (*slot)[MARCH_ZT] = ring[(s - s_begin) % MARCH_RING];
#pragma unroll
			for (int j = 0; j < M; j++) {
				if ((IC >> j) & 1u) { cp_async_f64(&slot[j][tid], cur + j * g.comp + idx + (long long) BS * sstride); }
				else if ((CC >> j) & 1u) { cp_async_f64(&slot[j][tid], cur + j * g.comp + idx); }
			}
		}
		cp_async_commit();
	};

	// prologue: LEAD iterations in flight, the left part of the window straight into registers
#pragma unroll
	for (int d = 0; d < MARCH_LEAD; d++) { issue(s_begin + d); }
	double w[M][W];  // w[j][BS + o] = component j at (s + o); only rows in IC are live
	double cv[M];
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
		double (*slot)[MARCH_ZT] = ring[(s - s_begin) % MARCH_RING];
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
		auto load = [&](int j, int o) -> double { return ((IC >> j) & 1u) ? w[j][BS + o] : cv[j]; };
		double out[M];
		gcm_node_sparse<P, BS>(PackedCoef<P, BS>{tab + t * Packed<P, BS>::SIZE}, load, out);
#pragma unroll
		for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
	}
	cp_async_wait<0>();
}

}  // namespace gcmb
