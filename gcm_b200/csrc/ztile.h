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

// In-tile ghost fill of the faces across the contiguous axis (a.zfill on a contiguous-axis stage): the ghost nodes of the
// staged row are overwritten IN SHARED MEMORY by the mirrored inner values of the same row, so the border condition
// costs no pass over the faces in HBM at all (reference engine/cubic/BorderConditions.hpp:97-114 for conditions over
// the whole face with plain component quantities; the ghost nodes of these faces in HBM are then never read).
// `row` = the staged components, `base` = index of node z = `seg` in a staged row, `width` = nodes of the row that
// this tile owns, t = 0..BS-1 (one caller per ghost layer).  Reads real nodes only, writes ghost nodes only.  The
// right face's ghost nodes can also lie in the halo of the tile BEFORE the one that holds the last node (when that
// one holds fewer than BS nodes): every tile whose nodes read them fills them.
template<class R, unsigned IC, int M, int BS, int ROWW>
GCMB_HD void zface_mirror_tile(const StageArgsT<R>& a, R (*row)[ROWW], int base, int seg, int width, int t) {
	const int d = t + 1;
	if (a.zf.on[0] && seg == 0) {
#pragma unroll
		for (int j = 0; j < M; j++) {
			if (!((IC >> j) & 1u)) { continue; }
			R x = row[j][base + d];
			if ((a.zf.set[0] >> j) & 1u) { x = -x + a.zf.add[0][j]; }
			row[j][base - d] = x;
		}
	}
	const int last = a.g.n[2] - 1 - seg;
	if (a.zf.on[1] && last >= 0 && last + d < width + BS) {
#pragma unroll
		for (int j = 0; j < M; j++) {
			if (!((IC >> j) & 1u)) { continue; }
			R x = row[j][base + last - d];
			if ((a.zf.set[1] >> j) & 1u) { x = -x + a.zf.add[1][j]; }
			row[j][base + last + d] = x;
		}
	}
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
