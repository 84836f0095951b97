// Private declarations shared by the translation units that implement the C ABI (gcmb_capi.cu: contexts and
// cubic bodies; simplex_capi.cu: tetrahedral bodies).
#pragma once
#include <nccl.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "internal.cuh"

namespace gcmb {

constexpr int N_CLASSES = 8;

struct ProfileSpan {
	int cls;
	cudaEvent_t a, b;
};

}  // namespace gcmb

struct gcmb_ctx {
	int device = 0;
	int real_bytes = 8;          // 8: double (the reference's `real`), 4: float
	bool fma = false;            // fp64 kernels compiled with FMA contraction (gcmb_set_fma): not bit-identical
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t timer_a = nullptr, timer_b = nullptr;
	bool profiling = false;
	std::vector<gcmb::ProfileSpan> spans;
	double class_ms[gcmb::N_CLASSES] = {0};
	long long class_launches[gcmb::N_CLASSES] = {0};
	long long launches = 0;
	size_t bytes = 0;
	std::vector<gcmb_body*> bodies;
	ncclComm_t comm = nullptr;
	int n_ranks = 1, rank = 0;
	double* scratch = nullptr;  // small device scratch (reductions)
	// Halo exchange overlapped with the x stage.  NCCL runs on comm_stream between ev_ready (state complete on
	// `stream`) and ev_halo (ghost planes received).  The x stage of a body whose exchange is in flight launches
	// its interior (which reads no ghost plane) on `stream` at once and its two boundary strips on edge_stream
	// behind ev_halo; everything else on `stream` first waits for ev_halo and for the strips (ev_edge).
	cudaStream_t comm_stream = nullptr, edge_stream = nullptr;
	cudaEvent_t ev_ready = nullptr, ev_halo = nullptr, ev_edge = nullptr;
	bool halo_pending = false, edge_pending = false, halo_defer = false, edge_waits_halo = false;
	// asynchronous read-back (snapshots): copies run on copy_stream behind ev_copy_ready
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_copy_ready = nullptr, ev_copy_done = nullptr;
	bool copy_pending = false;
};

namespace gcmb {

// make `stream` wait for a halo exchange / boundary strips in flight (no-op when there is none)
inline void wait_halo(gcmb_ctx* ctx) {
	if (ctx->halo_pending) {
		cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0);
		ctx->halo_pending = false;
	}
	if (ctx->edge_pending) {
		cudaStreamWaitEvent(ctx->stream, ctx->ev_edge, 0);
		ctx->edge_pending = false;
	}
}

// kernels that overwrite state an asynchronous read-back may still be reading wait for the copy
inline void wait_copy(gcmb_ctx* ctx) {
	if (ctx->copy_pending) {
		cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0);
		ctx->copy_pending = false;
	}
}

struct Launch {
	gcmb_ctx* ctx;
	int cls;
	ProfileSpan span;
	cudaStream_t on;
	Launch(gcmb_ctx* c, int cls_, cudaStream_t other = nullptr) : ctx(c), cls(cls_), on(other ? other : c->stream) {
		if (!ctx->halo_defer) { wait_halo(ctx); }
		wait_copy(ctx);
		if (ctx->profiling) {
			span.cls = cls;
			cudaEventCreate(&span.a);
			cudaEventCreate(&span.b);
			cudaEventRecord(span.a, on);
		}
	}
	~Launch() {
		ctx->launches++;
		if (ctx->profiling) {
			cudaEventRecord(span.b, on);
			ctx->spans.push_back(span);
		}
	}
};

}  // namespace gcmb
