// Stage kernel for the contiguous axis (internal axis 2): row tiles staged in shared memory.
//
// A block owns one z-chunk of ZTILE nodes and walks over consecutive rows (internal axis 1).  The
// components of the next rows are copied global -> shared with cp.async (LDGSTS) ZLEAD tiles ahead, so
// the HBM latency is covered by copies in flight instead of by resident warps; the stencil along z is
// then read from shared memory (each value is fetched from HBM once, the 2*BS halo values per chunk
// twice).  The arithmetic is the shared gcm_node_sparse: results are bit-identical to every other
// variant.  profiles/r1_run1: the one-thread-per-node kernel it replaces spent 14 stalled warps per
// issue slot on demand loads.
//
// The kernel body is written as barrier-separated phases over GCMB_BLOCK_THREADS so that the host-logic
// tests can step through it (one pass per phase over all threads of the block).
#pragma once
#include "march_async.h"

namespace gcmb {

constexpr int ZTILE = 256;             // threads per block = nodes per tile
constexpr int ZLEAD = 1;               // tiles in flight ahead of the one being computed
constexpr int ZHALO = 4;               // largest border size served by this kernel
constexpr int ZROW = ZTILE + 2 * ZHALO;

// followed in dynamic shared memory by the packed coefficient tables of all materials
template<class R, int M, int LEAD>
struct ZTileSmem {
	R v[LEAD + 1][M][ZROW];
	uint8_t id[LEAD + 1][ZTILE];
};

// phase A, one thread: start the copies of row `i1` (tile slot `slot`); rows past the end copy nothing
template<class R, class P, int BS, int LEAD>
GCMB_HD void ztile_issue(const StageArgsT<R>& a, ZTileSmem<R, P::M, LEAD>& sm, int slot, int tid, int i0, int i1, int z0, int i1_end) {
	constexpr int M = P::M;
	constexpr unsigned IC = PatternSets<P>::interp();
	constexpr unsigned CC = PatternSets<P>::center();
	const Geom& g = a.g;
	if (i1 < i1_end) {
		const long long row = g.index(i0, i1, 0);  // element of z = 0 of this row
		const int z = z0 + tid;
		if (z < g.n[2] + BS) {
#pragma unroll
			for (int j = 0; j < M; j++) {
				if (((IC | CC) >> j) & 1u) { cp_async_real(&sm.v[slot][j][tid + ZHALO], a.cur + j * g.comp + row + z); }
			}
		}
		// halo: BS values on each side of the chunk, copied by the first 2*BS threads
		if (tid < 2 * BS) {
			const int e = tid < BS ? tid - BS : ZTILE + (tid - BS);  // offset relative to z0
			const int zh = z0 + e;
			if (zh < g.n[2] + BS) {
#pragma unroll
				for (int j = 0; j < M; j++) {
					if ((IC >> j) & 1u) { cp_async_real(&sm.v[slot][j][e + ZHALO], a.cur + j * g.comp + row + zh); }
				}
			}
		}
		if ((tid & 3) == 0 && z < g.n[2]) { cp_async_bytes<4>(&sm.id[slot][tid], a.node_table + row + z); }
	}
	cp_async_commit();
}

// phase B, one thread: one node of the tile from shared memory
template<class R, class P, int BS, bool K0RT, int LEAD>
GCMB_HD void ztile_compute(const StageArgsT<R>& a, const ZTileSmem<R, P::M, LEAD>& sm, const R* tabs, int slot, int tid, int i0, int i1, int z0) {
	constexpr int M = P::M;
	const Geom& g = a.g;
	const int z = z0 + tid;
	if (z >= g.n[2]) { return; }
	const long long idx = g.index(i0, i1, z);
	const R* tab = tabs + (int) sm.id[slot][tid] * Packed<P, BS, K0RT>::SIZE;
	const R (*v)[ZROW] = sm.v[slot];
	auto load = [&](int j, int o) -> R { return v[j][tid + ZHALO + o]; };
	R out[M];
	gcm_node_sparse<R, P, BS, K0RT>(PackedCoef<R, P, BS, K0RT>{tab}, load, out);
#pragma unroll
	for (int c = 0; c < M; c++) { a.nxt[c * g.comp + idx] = out[c]; }
}

}  // namespace gcmb
