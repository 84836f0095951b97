// libgcm_b200.so — implementation of the C ABI in include/gcm_b200.h on CUDA (sm_100a).
//
// All grid data lives in HBM as a structure of arrays (internal.cuh: Geom); every entry point below
// enqueues hand-written kernels on the context's stream.  There is no host implementation of any
// of the operations: without a device gcmb_create fails.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include "thread_fns.h"

using namespace gcmb;

// =============================================================================================
// kernels (thin wrappers: one thread = one call of a thread function from thread_fns.h)
// =============================================================================================
namespace {

GCMB_GLOBAL void k_border(BorderArgs b, long long n_face) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (b.axis == 2) {
		// a face across the contiguous axis: the ghost layers of one face node share memory sectors with each
		// other (and so do the inner layers they mirror), so one thread fills all of them
		if (t >= n_face) { return; }
		for (int a = 1; a <= b.g.bs; a++) { border_thread(b, t, a); }
		return;
	}
	if (t >= n_face * b.g.bs) { return; }
	// consecutive threads walk the face (contiguous along z when the face contains z)
	const long long f = t % n_face;
	const int a = (int) (t / n_face) + 1;
	border_thread(b, f, a);
}

GCMB_GLOBAL void k_contact(ContactArgs c, long long n) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) { contact_thread(c, t); }
}

GCMB_GLOBAL void k_ode_maxwell(Geom g, double* pde, const uint8_t* node_table, const double* decay) {
	const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
	if (i2 < g.n[2]) { ode_maxwell_thread(g, pde, node_table, decay, blockIdx.z, blockIdx.y, i2); }
}

GCMB_GLOBAL void k_assign_table(Geom g, uint8_t* node_table, int table, AreaArgs area) {
	const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
	if (i2 >= g.n[2]) { return; }
	double x[3];
	node_coords(g, blockIdx.z, blockIdx.y, i2, x);
	if (area_contains(area, x)) { node_table[g.index(blockIdx.z, blockIdx.y, i2)] = (uint8_t) table; }
}

struct VecArg { double v[MAXM]; };

GCMB_GLOBAL void k_add_vector(Geom g, double* pde, VecArg vec, AreaArgs area) {
	const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
	if (i2 >= g.n[2]) { return; }
	double x[3];
	node_coords(g, blockIdx.z, blockIdx.y, i2, x);
	if (area_contains(area, x)) {
		const long long idx = g.index(blockIdx.z, blockIdx.y, i2);
		for (int c = 0; c < g.M; c++) { pde[c * g.comp + idx] += vec.v[c]; }
	}
}

GCMB_GLOBAL void k_face_mask(Geom g, int axis, int side, AreaArgs area, uint8_t* mask, long long n_face) {
	const long long f = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_face) { return; }
	int it[3];
	face_node(g, axis, f, side == 0 ? 0 : g.n[axis] - 1, it);
	double x[3];
	node_coords(g, it[0], it[1], it[2], x);
	mask[f] = area_contains(area, x) ? 1 : 0;
}

GCMB_GLOBAL void k_set_table_real_nodes(Geom g, uint8_t* node_table, const uint8_t* ids /* real nodes, x slowest */) {
	const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
	if (i2 >= g.n[2]) { return; }
	const long long r = ((long long) blockIdx.z * g.n[1] + blockIdx.y) * g.n[2] + i2;
	node_table[g.index(blockIdx.z, blockIdx.y, i2)] = ids ? ids[r] : 0;
}

GCMB_GLOBAL void k_xfer(XferArgs x, long long n, int to_device) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) { xfer_thread(x, t, to_device != 0); }
}

// deterministic block reduction: fixed tree over shared memory
template<int THREADS>
GCMB_DEV void block_sum(double& v, long long& c) {
#ifdef GCMB_EMUL_BLOCK_SUM
	GCMB_EMUL_BLOCK_SUM(v, c);
	return;
#endif
	__shared__ double sv[THREADS];
	__shared__ long long sc[THREADS];
	sv[threadIdx.x] = v;
	sc[threadIdx.x] = c;
	__syncthreads();
	for (int s = THREADS / 2; s > 0; s >>= 1) {
		if ((int) threadIdx.x < s) {
			sv[threadIdx.x] += sv[threadIdx.x + s];
			sc[threadIdx.x] += sc[threadIdx.x + s];
		}
		__syncthreads();
	}
	v = sv[0];
	c = sc[0];
}

// detector: quantity summed over masked nodes of the right face of the last axis; one partial (sum, count)
// per block in a fixed order => deterministic
GCMB_GLOBAL void k_detector(Geom g, const double* pde, const uint8_t* mask, int code, double* out_sum,
                           long long* out_count) {
	const long long n_face = (long long) g.n[0] * g.n[1];
	double sum = 0;
	long long count = 0;
	for (long long f = (long long) blockIdx.x * blockDim.x + threadIdx.x; f < n_face; f += (long long) gridDim.x * blockDim.x) {
		if (!mask[f]) { continue; }
		const long long idx = g.index((int) (f / g.n[1]), (int) (f % g.n[1]), g.n[2] - 1);
		double v[MAXM];
		for (int c = 0; c < g.M; c++) { v[c] = pde[c * g.comp + idx]; }
		sum += get_quantity(g.D, code, v);
		count++;
	}
	block_sum<256>(sum, count);
	if (threadIdx.x == 0) { out_sum[blockIdx.x] = sum; out_count[blockIdx.x] = count; }
}

// checksum partials: sum_i (i+1) * u_i over real nodes, fixed grid => deterministic
GCMB_GLOBAL void k_checksum(Geom g, const double* pde, double* partial) {
	const long long n = (long long) g.n[0] * g.n[1] * g.n[2];
	double sum = 0;
	long long dummy = 0;
	for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long) gridDim.x * blockDim.x) {
		const int i2 = (int) (t % g.n[2]);
		const int i1 = (int) ((t / g.n[2]) % g.n[1]);
		const int i0 = (int) (t / ((long long) g.n[2] * g.n[1]));
		const long long idx = g.index(i0, i1, i2);
		double acc = 0;
		for (int c = 0; c < g.M; c++) { acc += (c + 1) * pde[c * g.comp + idx]; }
		sum += acc;
	}
	block_sum<256>(sum, dummy);
	if (threadIdx.x == 0) { partial[blockIdx.x] = sum; }
}

}  // namespace

// =============================================================================================
// host-side objects
// =============================================================================================
namespace gcmb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool load() {
		if (lib) { return true; }
		lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) { lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
		if (!lib) { return false; }
#define L(name) *(void**) (&name) = dlsym(lib, "nccl" #name)
		L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(Send); L(Recv); L(AllReduce);
		L(GroupStart); L(GroupEnd); L(GetErrorString);
#undef L
		return GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllReduce;
	}
};
static NcclApi g_nccl;

constexpr int N_CLASSES = 8;

struct ProfileSpan {
	int cls;
	cudaEvent_t a, b;
};

}  // namespace gcmb

struct gcmb_ctx {
	int device = 0;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t timer_a = nullptr, timer_b = nullptr;
	bool profiling = false;
	std::vector<ProfileSpan> spans;
	double class_ms[N_CLASSES] = {0};
	long long class_launches[N_CLASSES] = {0};
	long long launches = 0;
	size_t bytes = 0;
	std::vector<gcmb_body*> bodies;
	ncclComm_t comm = nullptr;
	int n_ranks = 1, rank = 0;
	double* scratch = nullptr;  // small device scratch (reductions)
	// halo exchange overlapped with the interior of the x stage: NCCL runs on comm_stream between ev_ready (state
	// complete on `stream`) and ev_halo (ghost planes received); the next launch on `stream` waits for ev_halo
	// unless it is the interior part of the x stage (halo_defer)
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
	bool halo_pending = false, halo_defer = false;
};

// make `stream` wait for a halo exchange in flight (no-op when there is none)
static void wait_halo(gcmb_ctx* ctx) {
	if (ctx->halo_pending) {
		cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0);
		ctx->halo_pending = false;
	}
}

struct BorderCond {
	int cond = 0, dir = 0;
	uint8_t* mask[2] = {nullptr, nullptr};  // device, may be null (= side not handled)
	bool side_on[2] = {false, false};
	std::vector<int> q;
};

struct gcmb_body {
	gcmb_ctx* ctx = nullptr;
	Geom g;
	double* buf[2] = {nullptr, nullptr};  // buf[cur], buf[1-cur]
	int cur = 0;
	uint8_t* node_table = nullptr;
	int n_tables = 0;
	std::vector<double> U, U1, L;         // host copies [n][D][M*M] / [n][D][M]
	StageTable* tables = nullptr;         // device [n*D]
	std::vector<StageTable> host_tables;  // the same tables on the host
	double tables_tau = NAN;
	bool any_k0 = false;                  // some characteristic foot beyond the first cell (Courant > 1)
	int pattern_of_dir[3] = {-1, -1, -1};
	double* packed[3] = {nullptr, nullptr, nullptr};  // device: packed non-zero coefficients per direction
	std::string kernel_name[3];
	std::map<int, BorderCond> borders;    // ordered by condition number
	uint8_t* detector_mask = nullptr;
	int detector_code = 0;
	double* decay_dev = nullptr;
};

namespace {

struct Launch {
	gcmb_ctx* ctx;
	int cls;
	ProfileSpan span;
	Launch(gcmb_ctx* c, int cls_) : ctx(c), cls(cls_) {
		if (ctx->halo_pending && !ctx->halo_defer) { wait_halo(ctx); }
		if (ctx->profiling) {
			span.cls = cls;
			cudaEventCreate(&span.a);
			cudaEventCreate(&span.b);
			cudaEventRecord(span.a, ctx->stream);
		}
	}
	~Launch() {
		ctx->launches++;
		if (ctx->profiling) {
			cudaEventRecord(span.b, ctx->stream);
			ctx->spans.push_back(span);
		}
	}
};

int flush_profile(gcmb_ctx* ctx) {
	if (ctx->spans.empty()) { return GCMB_OK; }
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	for (ProfileSpan& s : ctx->spans) {
		float ms = 0;
		cudaEventElapsedTime(&ms, s.a, s.b);
		ctx->class_ms[s.cls] += ms;
		ctx->class_launches[s.cls]++;
		cudaEventDestroy(s.a);
		cudaEventDestroy(s.b);
	}
	ctx->spans.clear();
	return GCMB_OK;
}

AreaArgs make_area(int kind, const double* p) {
	AreaArgs a;
	std::memset(&a, 0, sizeof a);
	a.kind = kind;
	const int n = kind == 1 ? 6 : (kind == 2 ? 4 : (kind == 3 ? 7 : 0));
	for (int i = 0; i < n; i++) { a.p[i] = p[i]; }
	if (kind == 3) {  // axis = normalize(end - begin) (reference util/math/Area.hpp:93)
		const double d[3] = {p[4] - p[1], p[5] - p[2], p[6] - p[3]};
		const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		for (int i = 0; i < 3; i++) { a.p[7 + i] = d[i] / len; }
	}
	return a;
}

dim3 node_grid(const Geom& g, int threads) {
	return dim3((unsigned) ((g.n[2] + threads - 1) / threads), (unsigned) g.n[1], (unsigned) g.n[0]);
}

long long face_size(const Geom& g, int axis) {
	long long n = 1;
	for (int a = 0; a < 3; a++) { if (a != axis) { n *= g.n[a]; } }
	return n;
}

// Newton factors, foot cells and sides of every (table, direction) for time step tau
// (reference engine/cubic/GridCharacteristicMethod.hpp:56-59,78-84 + EqualDistanceLineInterpolator.hpp:20,63)
int build_tables(gcmb_body* b, double tau) {
	const Geom& g = b->g;
	const int M = g.M, D = g.D;
	std::vector<StageTable>& host = b->host_tables;
	host.assign((size_t) b->n_tables * D, StageTable());
	b->any_k0 = false;
	for (int t = 0; t < b->n_tables; t++) {
		for (int s = 0; s < D; s++) {
			StageTable& T = host[(size_t) t * D + s];
			std::memset(&T, 0, sizeof T);
			const double* u = b->U.data() + ((size_t) t * D + s) * M * M;
			const double* u1 = b->U1.data() + ((size_t) t * D + s) * M * M;
			const double* l = b->L.data() + ((size_t) t * D + s) * M;
			for (int i = 0; i < M * M; i++) { T.U[i] = u[i]; T.U1[i] = u1[i]; }
			const double h = g.h[s + g.shift];
			for (int k = 0; k < M; k++) {
				const double dx = -tau * l[k];
				T.dir[k] = dx > 0 ? 1 : -1;
				const double q = std::fabs(dx) / h;
				if (!(q >= 0)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "characteristic foot is not a number"); }
				const size_t k0 = (size_t) q;
				if (k0 > (size_t) g.bs) {
					GCMB_FAIL(GCMB_E_INVALID_ARG, "characteristic foot lies beyond the ghost layer (Courant number exceeds the border size)");
				}
				T.k0[k] = (int) k0;
				if (k0 != 0) { b->any_k0 = true; }
				for (int i = 1; i <= g.bs; i++) { T.F[k * MAXBS + i - 1] = (q - i + 1) / i; }
			}
		}
	}
	GCMB_CUDA(cudaMemcpyAsync(b->tables, host.data(), host.size() * sizeof(StageTable), cudaMemcpyHostToDevice, b->ctx->stream));
	// packed copies of the structurally non-zero coefficients for the specialised kernels
	std::vector<double> pk;
	for (int s = 0; s < D; s++) {
		cudaFree(b->packed[s]);
		b->packed[s] = nullptr;
		if (b->pattern_of_dir[s] < 0) { continue; }
		const PatternInfo& P = pattern(b->pattern_of_dir[s]);
		pk.assign((size_t) b->n_tables * (MAXM * MAXM * 2 + MAXM * MAXBS), 0.0);
		int size = 0;
		bool shares = true;
		for (int t = 0; t < b->n_tables; t++) {
			double one[MAXM * MAXM * 2 + MAXM * MAXBS];
			size = pack_table(P, g.bs, host[(size_t) t * D + s], one);
			std::memcpy(pk.data() + (size_t) t * size, one, (size_t) size * sizeof(double));
			shares = shares && table_shares_as_pattern(P, g.bs, host[(size_t) t * D + s]);
		}
		if (!shares) { continue; }  // no packed table: the stage falls back to the full-table kernels
		GCMB_CUDA(cudaMalloc(&b->packed[s], (size_t) b->n_tables * size * sizeof(double)));
		GCMB_CUDA(cudaMemcpyAsync(b->packed[s], pk.data(), (size_t) b->n_tables * size * sizeof(double), cudaMemcpyHostToDevice, b->ctx->stream));
		GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	}
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	b->tables_tau = tau;
	return GCMB_OK;
}

// choose the stage kernel of every direction from the sparsity of the tables
void choose_patterns(gcmb_body* b) {
	const Geom& g = b->g;
	const int M = g.M, D = g.D;
	for (int s = 0; s < D; s++) {
		unsigned um[MAXM] = {0}, u1m[MAXM] = {0};
		int sgn[MAXM];
		bool sgn_ok = true;
		for (int t = 0; t < b->n_tables; t++) {
			const double* u = b->U.data() + ((size_t) t * D + s) * M * M;
			const double* u1 = b->U1.data() + ((size_t) t * D + s) * M * M;
			const double* l = b->L.data() + ((size_t) t * D + s) * M;
			for (int k = 0; k < M; k++) {
				for (int j = 0; j < M; j++) {
					if (u[k * M + j] != 0) { um[k] |= 1u << j; }
					if (u1[k * M + j] != 0) { u1m[k] |= 1u << j; }
				}
				// side of the foot: dx = -tau*l > 0 ? +1 : -1 ; l == 0 needs no interpolation
				const int sg = l[k] == 0 ? 0 : (l[k] < 0 ? 1 : -1);
				if (t == 0) { sgn[k] = sg; } else if (sgn[k] != sg) { sgn_ok = false; }
			}
		}
		int chosen = -1;
		if (sgn_ok && (g.bs == 1 || g.bs == 2) && !getenv("GCMB_FORCE_DENSE")) {
			for (int p = 0; p < pattern_count() && chosen < 0; p++) {
				const PatternInfo& P = pattern(p);
				if (P.M != M) { continue; }
				bool ok = true;
				for (int k = 0; k < M && ok; k++) {
					ok = P.sgn[k] == sgn[k] && (um[k] & ~P.um[k]) == 0 && (u1m[k] & ~P.u1m[k]) == 0;
				}
				if (ok) { chosen = p; }
			}
		}
		b->pattern_of_dir[s] = chosen;
		b->kernel_name[s] = chosen >= 0 ? std::string("sparse:") + pattern(chosen).name + "/bs" + std::to_string(g.bs)
		                                : std::string("dense:M") + std::to_string(M);
	}
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* gcmb_last_error(void) { return g_error.c_str(); }
const char* gcmb_version(void) { return "gcm_b200 0.1 (sm_100a, fp64)"; }

int gcmb_create(int device, int real_bytes, gcmb_ctx** out) {
	if (!out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "out is null"); }
	if (real_bytes != 8) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "only real_bytes == 8 (fp64) is implemented"); }
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		GCMB_FAIL(GCMB_E_NO_DEVICE, "no CUDA device: gcm_b200 has no CPU path");
	}
	if (device < 0 || device >= n) { GCMB_FAIL(GCMB_E_INVALID_ARG, "device index out of range"); }
	GCMB_CUDA(cudaSetDevice(device));
	std::unique_ptr<gcmb_ctx> ctx(new gcmb_ctx);
	ctx->device = device;
	GCMB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
	ctx->stream = ctx->own_stream;
	GCMB_CUDA(cudaEventCreate(&ctx->timer_a));
	GCMB_CUDA(cudaEventCreate(&ctx->timer_b));
	GCMB_CUDA(cudaMalloc(&ctx->scratch, 4096 * sizeof(double)));
	*out = ctx.release();
	return GCMB_OK;
}

void gcmb_destroy(gcmb_ctx* ctx) {
	if (!ctx) { return; }
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	while (!ctx->bodies.empty()) { gcmb_cubic_body_destroy(ctx->bodies.back()); }
	for (ProfileSpan& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
	if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); }
	if (ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->comm); }
	if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_halo); }
	cudaFree(ctx->scratch);
	cudaEventDestroy(ctx->timer_a);
	cudaEventDestroy(ctx->timer_b);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

int gcmb_set_stream(gcmb_ctx* ctx, void* cuda_stream) {
	if (!ctx) { GCMB_FAIL(GCMB_E_INVALID_ARG, "ctx is null"); }
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
	return GCMB_OK;
}

int gcmb_sync(gcmb_ctx* ctx) {
	GCMB_CUDA(cudaSetDevice(ctx->device));
	wait_halo(ctx);
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	return GCMB_OK;
}

int gcmb_timer_start(gcmb_ctx* ctx) {
	GCMB_CUDA(cudaEventRecord(ctx->timer_a, ctx->stream));
	return GCMB_OK;
}

int gcmb_timer_stop(gcmb_ctx* ctx, float* ms) {
	wait_halo(ctx);
	GCMB_CUDA(cudaEventRecord(ctx->timer_b, ctx->stream));
	GCMB_CUDA(cudaEventSynchronize(ctx->timer_b));
	GCMB_CUDA(cudaEventElapsedTime(ms, ctx->timer_a, ctx->timer_b));
	return GCMB_OK;
}

int gcmb_profile_enable(gcmb_ctx* ctx, int on) {
	const int rc = flush_profile(ctx);
	ctx->profiling = on != 0;
	return rc;
}

int gcmb_profile_get(gcmb_ctx* ctx, int n_classes, double* ms, long long* launches) {
	const int rc = flush_profile(ctx);
	if (rc) { return rc; }
	for (int i = 0; i < n_classes && i < N_CLASSES; i++) {
		if (ms) { ms[i] = ctx->class_ms[i]; }
		if (launches) { launches[i] = ctx->class_launches[i]; }
	}
	return GCMB_OK;
}

long long gcmb_launch_count(gcmb_ctx* ctx) { return ctx->launches; }
size_t gcmb_device_bytes(gcmb_ctx* ctx) { return ctx->bytes; }

// ---- body ------------------------------------------------------------------------------------
int gcmb_cubic_body_create(gcmb_ctx* ctx, int D, int M, const int* sizes, const int* start,
                           const double* h, int border_size, gcmb_body** out) {
	if (!ctx || !sizes || !start || !h || !out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (D < 1 || D > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "D must be 1..3"); }
	if (M < 1 || M > MAXM) { GCMB_FAIL(GCMB_E_INVALID_ARG, "M must be 1..9"); }
	if (border_size < 1 || border_size > MAXBS) { GCMB_FAIL(GCMB_E_INVALID_ARG, "border_size must be 1..8"); }
	for (int i = 0; i < D; i++) {
		if (sizes[i] < border_size) { GCMB_FAIL(GCMB_E_INVALID_ARG, "sizes[i] must be >= border_size"); }
		if (!(h[i] > 0)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "h[i] must be positive"); }
	}
	GCMB_CUDA(cudaSetDevice(ctx->device));
	std::unique_ptr<gcmb_body> b(new gcmb_body);
	b->ctx = ctx;
	Geom& g = b->g;
	std::memset(&g, 0, sizeof g);
	g.D = D; g.M = M; g.bs = border_size; g.shift = 3 - D;
	for (int a = 0; a < 3; a++) { g.n[a] = 1; g.g[a] = 0; g.start[a] = 0; g.h[a] = 1; }
	for (int i = 0; i < D; i++) {
		const int a = i + g.shift;
		g.n[a] = sizes[i]; g.g[a] = border_size; g.start[a] = start[i]; g.h[a] = h[i];
	}
	g.zoff = 16;  // >= MAXBS, multiple of 16 doubles (128 B)
	g.pitch = ((g.zoff + g.n[2] + g.g[2]) + 15) / 16 * 16;
	g.plane = (long long) (g.n[1] + 2 * g.g[1]) * g.pitch;
	g.comp = ((long long) (g.n[0] + 2 * g.g[0]) * g.plane + 31) / 32 * 32;
	const size_t bytes = (size_t) g.comp * M * sizeof(double);
	for (int i = 0; i < 2; i++) {
		if (cudaMalloc(&b->buf[i], bytes) != cudaSuccess) {
			cudaGetLastError();
			cudaFree(b->buf[0]);
			GCMB_FAIL(GCMB_E_CUDA, "out of device memory for the PDE time layers");
		}
		GCMB_CUDA(cudaMemsetAsync(b->buf[i], 0, bytes, ctx->stream));
	}
	GCMB_CUDA(cudaMalloc(&b->node_table, (size_t) g.comp + 64));  // slack: 4-byte id copies of the z-tile kernel
	GCMB_CUDA(cudaMemsetAsync(b->node_table, 0, (size_t) g.comp, ctx->stream));
	GCMB_CUDA(cudaMalloc(&b->decay_dev, 256 * sizeof(double)));
	ctx->bytes += 2 * bytes + (size_t) g.comp;
	ctx->bodies.push_back(b.get());
	*out = b.release();
	return GCMB_OK;
}

void gcmb_cubic_body_destroy(gcmb_body* b) {
	if (!b) { return; }
	gcmb_ctx* ctx = b->ctx;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	ctx->bytes -= 2 * (size_t) b->g.comp * b->g.M * sizeof(double) + (size_t) b->g.comp;
	cudaFree(b->buf[0]);
	cudaFree(b->buf[1]);
	cudaFree(b->node_table);
	cudaFree(b->tables);
	for (int s = 0; s < 3; s++) { cudaFree(b->packed[s]); }
	cudaFree(b->decay_dev);
	cudaFree(b->detector_mask);
	for (auto& kv : b->borders) { cudaFree(kv.second.mask[0]); cudaFree(kv.second.mask[1]); }
	ctx->bodies.erase(std::remove(ctx->bodies.begin(), ctx->bodies.end(), b), ctx->bodies.end());
	delete b;
}

static int transfer(gcmb_body* b, void* host, int with_ghosts, bool to_device) {
	if (!b || !host) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	const int e0 = with_ghosts ? g.n[0] + 2 * g.g[0] : g.n[0];
	const long long e12 = (long long) (with_ghosts ? g.n[1] + 2 * g.g[1] : g.n[1]) * (with_ghosts ? g.n[2] + 2 * g.g[2] : g.n[2]);
	const long long per_plane = e12 * g.M;  // doubles per slice of internal axis 0
	int chunk = (int) std::max<long long>(1, (64LL << 20) / std::max<long long>(1, per_plane));
	chunk = std::min(chunk, e0);
	double* stage = nullptr;
	GCMB_CUDA(cudaMalloc(&stage, (size_t) chunk * per_plane * sizeof(double)));
	int rc = GCMB_OK;
	for (int x0 = 0; x0 < e0 && rc == GCMB_OK; x0 += chunk) {
		const int x1 = std::min(e0, x0 + chunk);
		const long long n = (long long) (x1 - x0) * e12;
		double* hp = static_cast<double*>(host) + (long long) x0 * per_plane;
		XferArgs x;
		x.soa = b->buf[b->cur]; x.aos = stage; x.g = g; x.with_ghosts = with_ghosts; x.x_begin = x0; x.x_end = x1;
		cudaError_t e = cudaSuccess;
		if (to_device) {
			e = cudaMemcpyAsync(stage, hp, (size_t) n * g.M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
		}
		if (e == cudaSuccess) {
			Launch l(ctx, 6);
			GCMB_LAUNCH(k_xfer, (unsigned) ((n + 255) / 256), 256, ctx->stream, x, n, to_device ? 1 : 0);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess && !to_device) {
			e = cudaMemcpyAsync(hp, stage, (size_t) n * g.M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
		}
		if (e == cudaSuccess) { e = cudaStreamSynchronize(ctx->stream); }
		if (e != cudaSuccess) { set_error(std::string("state transfer: ") + cudaGetErrorString(e)); rc = GCMB_E_CUDA; }
	}
	cudaFree(stage);
	return rc;
}

int gcmb_cubic_upload_state(gcmb_body* body, const void* aos_pde, int with_ghosts) {
	return transfer(body, const_cast<void*>(aos_pde), with_ghosts, true);
}

int gcmb_cubic_download_state(gcmb_body* body, void* aos_pde, int with_ghosts) {
	return transfer(body, aos_pde, with_ghosts, false);
}

int gcmb_cubic_download_tables(gcmb_body* b, uint8_t* node_table_id) {
	if (!b || !node_table_id) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	const Geom& g = b->g;
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	std::vector<uint8_t> all((size_t) g.comp);
	GCMB_CUDA(cudaMemcpyAsync(all.data(), b->node_table, all.size(), cudaMemcpyDeviceToHost, b->ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	size_t r = 0;
	for (int i0 = 0; i0 < g.n[0]; i0++) for (int i1 = 0; i1 < g.n[1]; i1++) for (int i2 = 0; i2 < g.n[2]; i2++) {
		node_table_id[r++] = all[(size_t) g.index(i0, i1, i2)];
	}
	return GCMB_OK;
}

int gcmb_cubic_set_materials(gcmb_body* b, int n_tables, const double* U, const double* U1,
                             const double* L, const uint8_t* node_table_id) {
	if (!b || !U || !U1 || !L) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (n_tables < 1 || n_tables > GCMB_MAX_TABLES) { GCMB_FAIL(GCMB_E_INVALID_ARG, "n_tables must be 1..255"); }
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	const size_t mm = (size_t) n_tables * g.D * g.M * g.M, lm = (size_t) n_tables * g.D * g.M;
	b->n_tables = n_tables;
	b->U.assign(U, U + mm);
	b->U1.assign(U1, U1 + mm);
	b->L.assign(L, L + lm);
	cudaFree(b->tables);
	b->tables = nullptr;
	GCMB_CUDA(cudaMalloc(&b->tables, (size_t) n_tables * g.D * sizeof(StageTable)));
	b->tables_tau = NAN;
	choose_patterns(b);
	uint8_t* ids = nullptr;
	if (node_table_id) {
		const size_t n = (size_t) g.n[0] * g.n[1] * g.n[2];
		for (size_t i = 0; i < n; i++) {
			if (node_table_id[i] >= n_tables) { GCMB_FAIL(GCMB_E_INVALID_ARG, "node_table_id out of range"); }
		}
		GCMB_CUDA(cudaMalloc(&ids, n));
		GCMB_CUDA(cudaMemcpyAsync(ids, node_table_id, n, cudaMemcpyHostToDevice, ctx->stream));
	}
	{
		Launch l(ctx, 6);
		GCMB_LAUNCH(k_set_table_real_nodes, node_grid(g, 128), 128, ctx->stream, g, b->node_table, ids);
	}
	GCMB_CUDA(cudaGetLastError());
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	cudaFree(ids);
	return GCMB_OK;
}

int gcmb_cubic_assign_table_in_area(gcmb_body* b, int table_id, int area_kind, const double* params) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (table_id < 0 || table_id >= b->n_tables) { GCMB_FAIL(GCMB_E_INVALID_ARG, "table_id out of range"); }
	if (area_kind < 0 || area_kind > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "unknown area kind"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	{
		Launch l(b->ctx, 6);
		GCMB_LAUNCH(k_assign_table, node_grid(b->g, 128), 128, b->ctx->stream, b->g, b->node_table, table_id, make_area(area_kind, params));
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

int gcmb_cubic_add_vector_in_area(gcmb_body* b, const double* vector_M, int area_kind, const double* params) {
	if (!b || !vector_M) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (area_kind < 0 || area_kind > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "unknown area kind"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	VecArg v;
	std::memset(&v, 0, sizeof v);
	for (int c = 0; c < b->g.M; c++) { v.v[c] = vector_M[c]; }
	{
		Launch l(b->ctx, 6);
		GCMB_LAUNCH(k_add_vector, node_grid(b->g, 128), 128, b->ctx->stream, b->g, b->buf[b->cur], v, make_area(area_kind, params));
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

// ---- borders ---------------------------------------------------------------------------------
static int border_register(gcmb_body* b, int cond, int dir, int n_q, const int* q_codes, BorderCond** out) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (dir < 0 || dir >= b->g.D) { GCMB_FAIL(GCMB_E_INVALID_ARG, "direction out of range"); }
	if (n_q < 0 || n_q > MAXM + 1 || (n_q > 0 && !q_codes)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bad quantity list"); }
	for (int i = 0; i < n_q; i++) {
		if (q_codes[i] != GCMB_Q_PRESSURE_TRACE && (q_codes[i] < 0 || q_codes[i] >= b->g.M)) {
			GCMB_FAIL(GCMB_E_INVALID_ARG, "quantity code out of range");
		}
	}
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	BorderCond& c = b->borders[cond];
	cudaFree(c.mask[0]);
	cudaFree(c.mask[1]);
	c = BorderCond();
	c.cond = cond;
	c.dir = dir;
	c.q.assign(q_codes, q_codes + n_q);
	*out = &c;
	return GCMB_OK;
}

int gcmb_cubic_border_set(gcmb_body* b, int cond, int dir, const uint8_t* left_mask,
                          const uint8_t* right_mask, int n_q, const int* q_codes) {
	BorderCond* c = nullptr;
	const int rc = border_register(b, cond, dir, n_q, q_codes, &c);
	if (rc) { return rc; }
	const long long nf = face_size(b->g, dir + b->g.shift);
	const uint8_t* src[2] = {left_mask, right_mask};
	for (int s = 0; s < 2; s++) {
		if (!src[s]) { continue; }
		c->side_on[s] = true;
		GCMB_CUDA(cudaMalloc(&c->mask[s], (size_t) nf));
		GCMB_CUDA(cudaMemcpy(c->mask[s], src[s], (size_t) nf, cudaMemcpyHostToDevice));
	}
	return GCMB_OK;
}

int gcmb_cubic_border_set_area(gcmb_body* b, int cond, int dir, int sides, int area_kind,
                               const double* params, int n_q, const int* q_codes) {
	BorderCond* c = nullptr;
	const int rc = border_register(b, cond, dir, n_q, q_codes, &c);
	if (rc) { return rc; }
	if (area_kind < 0 || area_kind > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "unknown area kind"); }
	const int axis = dir + b->g.shift;
	const long long nf = face_size(b->g, axis);
	for (int s = 0; s < 2; s++) {
		if (!((sides >> s) & 1)) { continue; }
		c->side_on[s] = true;
		if (area_kind == 0) { continue; }  // infinite area: no mask needed
		GCMB_CUDA(cudaMalloc(&c->mask[s], (size_t) nf));
		{
			Launch l(b->ctx, 3);
			GCMB_LAUNCH(k_face_mask, (unsigned) ((nf + 255) / 256), 256, b->ctx->stream, b->g, axis, s, make_area(area_kind, params), c->mask[s], nf);
		}
		GCMB_CUDA(cudaGetLastError());
	}
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	return GCMB_OK;
}

int gcmb_cubic_border_apply(gcmb_body* b, int dir, int n_values, const double* values) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	int need = 0;
	for (auto& kv : b->borders) { if (kv.second.dir == dir) { need += (int) kv.second.q.size(); } }
	if (need != n_values || (need > 0 && !values)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "wrong number of border values"); }
	const Geom& g = b->g;
	const int axis = dir + g.shift;
	const long long nf = face_size(g, axis);
	int used = 0;
	for (auto& kv : b->borders) {
		BorderCond& c = kv.second;
		if (c.dir != dir) { continue; }
		for (int s = 0; s < 2; s++) {
			if (!c.side_on[s]) { continue; }
			BorderArgs a;
			std::memset(&a, 0, sizeof a);
			a.pde = b->buf[b->cur]; a.mask = c.mask[s]; a.g = g; a.axis = axis; a.side = s;
			a.nq = (int) c.q.size();
			for (int i = 0; i < a.nq; i++) { a.q[i] = c.q[(size_t) i]; a.val[i] = values[used + i]; }
			const long long n = axis == 2 ? nf : nf * g.bs;
			Launch l(b->ctx, 3);
			GCMB_LAUNCH(k_border, (unsigned) ((n + 127) / 128), 128, b->ctx->stream, a, nf);
		}
		used += (int) c.q.size();
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

// ---- contacts --------------------------------------------------------------------------------
int gcmb_cubic_contact_apply(gcmb_body* a, const gcmb_body* b, const int* boxA_min,
                             const int* boxB_min, const int* extent) {
	if (!a || !b || !boxA_min || !boxB_min || !extent) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (a->g.D != b->g.D || a->g.M != b->g.M || a->g.bs != b->g.bs) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bodies are not compatible"); }
	GCMB_CUDA(cudaSetDevice(a->ctx->device));
	ContactArgs c;
	std::memset(&c, 0, sizeof c);
	c.a = a->buf[a->cur]; c.b = b->buf[b->cur]; c.ga = a->g; c.gb = b->g;
	long long n = 1;
	for (int ax = 0; ax < 3; ax++) { c.amin[ax] = c.bmin[ax] = 0; c.ext[ax] = 1; }
	for (int i = 0; i < a->g.D; i++) {
		const int ax = i + a->g.shift;
		c.amin[ax] = boxA_min[i]; c.bmin[ax] = boxB_min[i]; c.ext[ax] = extent[i];
		if (extent[i] < 1) { GCMB_FAIL(GCMB_E_INVALID_ARG, "empty contact box"); }
		if (boxA_min[i] < -a->g.bs || boxA_min[i] + extent[i] > a->g.n[ax] + a->g.bs ||
		    boxB_min[i] < -b->g.bs || boxB_min[i] + extent[i] > b->g.n[ax] + b->g.bs) {
			GCMB_FAIL(GCMB_E_INVALID_ARG, "contact box leaves the grid");
		}
		n *= extent[i];
	}
	{
		Launch l(a->ctx, 4);
		GCMB_LAUNCH(k_contact, (unsigned) ((n + 127) / 128), 128, a->ctx->stream, c, n);
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

// ---- stage -----------------------------------------------------------------------------------
int gcmb_cubic_stage(gcmb_body* b, int dir, double tau) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (dir < 0 || dir >= b->g.D) { GCMB_FAIL(GCMB_E_INVALID_ARG, "direction out of range"); }
	if (!b->tables) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	if (!(b->tables_tau == tau)) {
		const int rc = build_tables(b, tau);
		if (rc) { return rc; }
	}
	StageArgs a;
	a.cur = b->buf[b->cur];
	a.nxt = b->buf[1 - b->cur];
	a.node_table = b->node_table;
	a.tables = b->tables;
	a.packed = b->packed[dir];
	a.n_tables = b->n_tables;
	a.g = b->g;
	a.axis = dir + b->g.shift;
	a.dir = dir;
	a.x_begin = 0;
	a.x_end = b->g.n[0];
	a.host_tables = b->host_tables.data();
	StageLauncher launch = nullptr;
	const int p = b->pattern_of_dir[dir];
	if (p >= 0 && !b->any_k0) {
		launch = b->g.bs == 1 ? pattern(p).launch_bs1 : pattern(p).launch_bs2;
	}
	static const bool literal_dense = std::getenv("GCMB_DENSE_LITERAL") != nullptr;
	if (!launch && !b->any_k0 && !literal_dense) {
		launch = dense_k0_launcher(b->g.M, b->g.bs);
		if (launch && p < 0) {
			b->kernel_name[dir] = std::string(b->n_tables == 1 ? "dense_k0_one:M" : "dense_k0:M") + std::to_string(b->g.M) + "/bs" + std::to_string(b->g.bs);
		}
	}
	if (!launch) { launch = dense_launcher(b->g.M); }
	if (!launch) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "no stage kernel for this PDE size"); }
	gcmb_ctx* ctx = b->ctx;
	const int bs = b->g.g[0];
	if (ctx->halo_pending && a.axis == 0 && bs > 0 && b->g.n[0] > 2 * bs) {
		// the nodes at least `bs` planes away from both slab faces read no ghost plane: they go first, while the
		// halo exchange is still in flight; the two boundary strips follow once the ghost planes have arrived
		StageArgs part = a;
		part.x_begin = bs; part.x_end = b->g.n[0] - bs;
		ctx->halo_defer = true;
		{ Launch l(ctx, a.axis); launch(part, ctx->stream); }
		ctx->halo_defer = false;
		part.x_begin = 0; part.x_end = bs;
		{ Launch l(ctx, a.axis); launch(part, ctx->stream); }
		part.x_begin = b->g.n[0] - bs; part.x_end = b->g.n[0];
		{ Launch l(ctx, a.axis); launch(part, ctx->stream); }
	} else {
		Launch l(ctx, a.axis);
		launch(a, ctx->stream);
	}
	GCMB_CUDA(cudaGetLastError());
	b->cur = 1 - b->cur;  // swapCurrAndNextPdeTimeLayer
	return GCMB_OK;
}

const char* gcmb_cubic_stage_kernel_name(gcmb_body* b, int dir) {
	if (!b || dir < 0 || dir >= b->g.D) { return ""; }
	return b->kernel_name[dir].c_str();
}

// ---- ode -------------------------------------------------------------------------------------
int gcmb_cubic_ode_maxwell(gcmb_body* b, const double* decay_per_table) {
	if (!b || !decay_per_table) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (b->n_tables < 1) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	GCMB_CUDA(cudaMemcpyAsync(b->decay_dev, decay_per_table, (size_t) b->n_tables * sizeof(double), cudaMemcpyHostToDevice, b->ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));  // the host array may be a temporary
	{
		Launch l(b->ctx, 5);
		GCMB_LAUNCH(k_ode_maxwell, node_grid(b->g, 128), 128, b->ctx->stream, b->g, b->buf[b->cur], b->node_table, b->decay_dev);
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

// ---- seismogram taps ---------------------------------------------------------------------------
int gcmb_cubic_detector_set_mask(gcmb_body* b, int q_code, const uint8_t* face_mask) {
	if (!b || !face_mask) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (q_code != GCMB_Q_PRESSURE_TRACE && (q_code < 0 || q_code >= b->g.M)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "quantity code out of range"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	const long long nf = face_size(b->g, 2);
	cudaFree(b->detector_mask);
	b->detector_mask = nullptr;
	GCMB_CUDA(cudaMalloc(&b->detector_mask, (size_t) nf));
	GCMB_CUDA(cudaMemcpy(b->detector_mask, face_mask, (size_t) nf, cudaMemcpyHostToDevice));
	b->detector_code = q_code;
	return GCMB_OK;
}

int gcmb_cubic_detector_set_area(gcmb_body* b, int q_code, int area_kind, const double* params) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (q_code != GCMB_Q_PRESSURE_TRACE && (q_code < 0 || q_code >= b->g.M)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "quantity code out of range"); }
	if (area_kind < 0 || area_kind > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "unknown area kind"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	const long long nf = face_size(b->g, 2);
	cudaFree(b->detector_mask);
	b->detector_mask = nullptr;
	GCMB_CUDA(cudaMalloc(&b->detector_mask, (size_t) nf));
	{
		Launch l(b->ctx, 7);
		GCMB_LAUNCH(k_face_mask, (unsigned) ((nf + 255) / 256), 256, b->ctx->stream, b->g, 2, 1, make_area(area_kind, params), b->detector_mask, nf);
	}
	GCMB_CUDA(cudaGetLastError());
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	b->detector_code = q_code;
	return GCMB_OK;
}

int gcmb_cubic_seismo(gcmb_body* b, double* sum, long long* count, int line_comp, double* line, int n_line) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	if (sum || count) {
		if (!b->detector_mask) { GCMB_FAIL(GCMB_E_INVALID_OP, "detector is not set"); }
		const long long nf = (long long) g.n[0] * g.n[1];
		const int blocks = (int) std::max<long long>(1, std::min<long long>(1024, (nf + 255) / 256));
		double* d_sum = ctx->scratch;
		long long* d_count = reinterpret_cast<long long*>(ctx->scratch + 1024);
		{
			Launch l(ctx, 7);
			GCMB_LAUNCH(k_detector, blocks, 256, ctx->stream, g, b->buf[b->cur], b->detector_mask, b->detector_code, d_sum, d_count);
		}
		GCMB_CUDA(cudaGetLastError());
		std::vector<double> ps((size_t) blocks);
		std::vector<long long> pc((size_t) blocks);
		GCMB_CUDA(cudaMemcpyAsync(ps.data(), d_sum, (size_t) blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		GCMB_CUDA(cudaMemcpyAsync(pc.data(), d_count, (size_t) blocks * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
		GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
		double hs = 0;
		long long hc = 0;
		for (int i = 0; i < blocks; i++) { hs += ps[(size_t) i]; hc += pc[(size_t) i]; }
		if (sum) { *sum = hs; }
		if (count) { *count = hc; }
	}
	if (line) {
		if (line_comp < 0 || line_comp >= g.M) { GCMB_FAIL(GCMB_E_INVALID_ARG, "line component out of range"); }
		if (n_line != g.n[2]) { GCMB_FAIL(GCMB_E_INVALID_ARG, "n_line must equal the size of the last axis"); }
		const double* src = b->buf[b->cur] + (long long) line_comp * g.comp + g.index(g.n[0] / 2, g.n[1] / 2, 0);
		GCMB_CUDA(cudaMemcpyAsync(line, src, (size_t) n_line * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	return GCMB_OK;
}

int gcmb_cubic_checksum(gcmb_body* b, double* out) {
	if (!b || !out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	gcmb_ctx* ctx = b->ctx;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	const int blocks = 1024;
	{
		Launch l(ctx, 7);
		GCMB_LAUNCH(k_checksum, blocks, 256, ctx->stream, b->g, b->buf[b->cur], ctx->scratch);
	}
	GCMB_CUDA(cudaGetLastError());
	std::vector<double> partial((size_t) blocks);
	GCMB_CUDA(cudaMemcpyAsync(partial.data(), ctx->scratch, (size_t) blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	double s = 0;
	for (double p : partial) { s += p; }
	*out = s;
	return GCMB_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------
#define GCMB_NCCL(call)                                                                          \
	do {                                                                                         \
		ncclResult_t r__ = (call);                                                               \
		if (r__ != ncclSuccess) {                                                                \
			set_error(std::string(#call) + " -> " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); \
			return GCMB_E_NCCL;                                                                  \
		}                                                                                        \
	} while (0)

int gcmb_comm_unique_id(void* id128) {
	if (!id128) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null id"); }
	if (!g_nccl.load()) { GCMB_FAIL(GCMB_E_NCCL, "libnccl.so.2 cannot be loaded"); }
	ncclUniqueId id;
	GCMB_NCCL(g_nccl.GetUniqueId(&id));
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	std::memcpy(id128, &id, 128);
	return GCMB_OK;
}

int gcmb_comm_init(gcmb_ctx* ctx, int n_ranks, int rank, const void* id128) {
	if (!ctx || !id128) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bad rank"); }
	if (!g_nccl.load()) { GCMB_FAIL(GCMB_E_NCCL, "libnccl.so.2 cannot be loaded"); }
	GCMB_CUDA(cudaSetDevice(ctx->device));
	ncclUniqueId id;
	std::memcpy(&id, id128, 128);
	GCMB_NCCL(g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank));
	ctx->n_ranks = n_ranks;
	ctx->rank = rank;
	return GCMB_OK;
}

int gcmb_cubic_halo_exchange(gcmb_body* b) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	gcmb_ctx* ctx = b->ctx;
	if (!ctx->comm || ctx->n_ranks == 1) { return GCMB_OK; }
	const Geom& g = b->g;
	if (g.D != 3 && g.g[0] == 0) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "slab decomposition needs the x axis to be the slowest internal axis (3-D grids)"); }
	GCMB_CUDA(cudaSetDevice(ctx->device));
	// x-planes are contiguous inside every component volume: ghost planes [0,bs) and [n0+bs, n0+2bs),
	// outermost real planes [bs, 2bs) and [n0, n0+bs)
	const size_t count = (size_t) g.g[0] * (size_t) g.plane;
	double* base = b->buf[b->cur];
	// on a stream of its own, so that the interior of the following x stage (which reads no ghost plane) overlaps it
	static const bool overlap = !std::getenv("GCMB_NO_HALO_OVERLAP");
	if (overlap && !ctx->comm_stream) {
		int least = 0, greatest = 0;
		GCMB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		GCMB_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, greatest));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
	}
	wait_halo(ctx);
	cudaStream_t comm_stream = overlap ? ctx->comm_stream : ctx->stream;
	if (overlap) {
		GCMB_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
		GCMB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
	}
	GCMB_NCCL(g_nccl.GroupStart());
	for (int c = 0; c < g.M; c++) {
		double* v = base + (long long) c * g.comp;
		if (ctx->rank > 0) {
			GCMB_NCCL(g_nccl.Send(v + (long long) g.g[0] * g.plane, count, ncclDouble, ctx->rank - 1, ctx->comm, comm_stream));
			GCMB_NCCL(g_nccl.Recv(v, count, ncclDouble, ctx->rank - 1, ctx->comm, comm_stream));
		}
		if (ctx->rank < ctx->n_ranks - 1) {
			GCMB_NCCL(g_nccl.Send(v + (long long) g.n[0] * g.plane, count, ncclDouble, ctx->rank + 1, ctx->comm, comm_stream));
			GCMB_NCCL(g_nccl.Recv(v + (long long) (g.n[0] + g.g[0]) * g.plane, count, ncclDouble, ctx->rank + 1, ctx->comm, comm_stream));
		}
	}
	GCMB_NCCL(g_nccl.GroupEnd());
	if (overlap) {
		GCMB_CUDA(cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
		ctx->halo_pending = true;
	}
	ctx->launches++;
	return GCMB_OK;
}

size_t gcmb_cubic_halo_bytes(gcmb_body* b) {
	return b ? (size_t) b->g.M * (size_t) b->g.g[0] * (size_t) b->g.plane * sizeof(double) : 0;
}

static int halo_host(gcmb_body* b, int side, void* host, bool get) {
	if (!b || !host) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (side != 0 && side != 1) { GCMB_FAIL(GCMB_E_INVALID_ARG, "side must be 0 or 1"); }
	const Geom& g = b->g;
	if (g.g[0] == 0) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "slab decomposition needs the x axis to be the slowest internal axis (3-D grids)"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	wait_halo(b->ctx);
	const size_t count = (size_t) g.g[0] * (size_t) g.plane;
	// real planes next to the face: [bs, 2bs) left, [n0, n0+bs) right; ghost planes: [0, bs) / [n0+bs, n0+2bs)
	const long long first = get ? (side == 0 ? g.g[0] : g.n[0]) : (side == 0 ? 0 : g.n[0] + g.g[0]);
	for (int c = 0; c < g.M; c++) {
		double* dev = b->buf[b->cur] + (long long) c * g.comp + first * g.plane;
		double* h = static_cast<double*>(host) + (size_t) c * count;
		if (get) { GCMB_CUDA(cudaMemcpyAsync(h, dev, count * sizeof(double), cudaMemcpyDeviceToHost, b->ctx->stream)); }
		else { GCMB_CUDA(cudaMemcpyAsync(dev, h, count * sizeof(double), cudaMemcpyHostToDevice, b->ctx->stream)); }
	}
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	return GCMB_OK;
}

int gcmb_cubic_halo_get(gcmb_body* b, int side, void* host_buffer) { return halo_host(b, side, host_buffer, true); }
int gcmb_cubic_halo_put(gcmb_body* b, int side, const void* host_buffer) {
	return halo_host(b, side, const_cast<void*>(host_buffer), false);
}

int gcmb_comm_allreduce_sum(gcmb_ctx* ctx, double* host_values, int n) {
	if (!ctx || !host_values || n < 1 || n > 2048) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bad argument"); }
	if (!ctx->comm || ctx->n_ranks == 1) { return GCMB_OK; }
	GCMB_CUDA(cudaSetDevice(ctx->device));
	double* d = ctx->scratch + 2048;
	GCMB_CUDA(cudaMemcpyAsync(d, host_values, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	GCMB_NCCL(g_nccl.AllReduce(d, d, (size_t) n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
	GCMB_CUDA(cudaMemcpyAsync(host_values, d, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	return GCMB_OK;
}

}  // extern "C"

#include "simplex_capi.inc"
