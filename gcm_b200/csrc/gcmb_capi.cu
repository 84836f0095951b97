// libgcm_b200.so — implementation of the C ABI in include/gcm_b200.h on CUDA (sm_100a).
//
// All grid data lives in HBM as a structure of arrays (internal.cuh: Geom); every entry point below
// enqueues hand-written kernels on the context's stream.  There is no host implementation of any
// of the operations: without a device gcmb_create fails.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <memory>

#include "capi_internal.cuh"
#include "thread_fns.h"
#include "triangle_fns.h"

using namespace gcmb;

// =============================================================================================
// kernels (thin wrappers: one thread = one call of a thread function from thread_fns.h)
// =============================================================================================
namespace {

template<class R>
GCMB_GLOBAL void k_border_zsector(BorderArgs<R> b, long long n_face) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n_face) { border_thread_zsector(b, t); }
}

template<class R>
GCMB_GLOBAL void k_border(BorderArgs<R> b, long long n_face) {
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

template<class R>
GCMB_GLOBAL void k_contact(ContactArgs<R> c, long long n) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) { contact_thread(c, t); }
}

template<class R>
GCMB_GLOBAL void k_ode_maxwell(Geom g, R* pde, const uint8_t* node_table, const R* decay) {
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

template<class R> struct VecArg { R v[MAXM]; };

template<class R>
GCMB_GLOBAL void k_add_vector(Geom g, R* pde, VecArg<R> vec, AreaArgs area) {
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

template<class R>
GCMB_GLOBAL void k_xfer(XferArgs<R> x, long long n, int to_device) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) { xfer_thread(x, t, to_device != 0); }
}

// deterministic block reduction: warp shuffles in a fixed butterfly, then a fixed tree over the warps' partials
template<int THREADS>
GCMB_DEV void block_sum(double& v, long long& c) {
#ifdef GCMB_EMUL_BLOCK_SUM
	GCMB_EMUL_BLOCK_SUM(v, c);
	return;
#else
	__shared__ double sv[THREADS / 32];
	__shared__ long long sc[THREADS / 32];
	for (int o = 16; o > 0; o >>= 1) {
		v += __shfl_down_sync(0xffffffffu, v, o);
		c += __shfl_down_sync(0xffffffffu, c, o);
	}
	if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; sc[threadIdx.x >> 5] = c; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < THREADS / 32; w++) { v += sv[w]; c += sc[w]; }
	}
#endif
}

// detector: quantity summed over masked nodes of the right face of the last axis; one partial (sum, count)
// per block in a fixed order => deterministic
template<class R>
GCMB_GLOBAL void k_detector(Geom g, const R* pde, const uint8_t* mask, int code, double* out_sum,
                           long long* out_count) {
	const long long n_face = (long long) g.n[0] * g.n[1];
	double sum = 0;
	long long count = 0;
	for (long long f = (long long) blockIdx.x * blockDim.x + threadIdx.x; f < n_face; f += (long long) gridDim.x * blockDim.x) {
		if (!mask[f]) { continue; }
		const long long idx = g.index((int) (f / g.n[1]), (int) (f % g.n[1]), g.n[2] - 1);
		R v[MAXM];
		for (int c = 0; c < g.M; c++) { v[c] = pde[c * g.comp + idx]; }
		sum += (double) get_quantity(g.D, code, v);
		count++;
	}
	block_sum<256>(sum, count);
	if (threadIdx.x == 0) { out_sum[blockIdx.x] = sum; out_count[blockIdx.x] = count; }
}

// checksum partials: sum_i (i+1) * u_i over real nodes, fixed grid => deterministic
template<class R>
GCMB_GLOBAL void k_checksum(Geom g, const R* pde, double* partial) {
	const long long n = (long long) g.n[0] * g.n[1] * g.n[2];
	double sum = 0;
	long long dummy = 0;
	for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long) gridDim.x * blockDim.x) {
		const int i2 = (int) (t % g.n[2]);
		const int i1 = (int) ((t / g.n[2]) % g.n[1]);
		const int i0 = (int) (t / ((long long) g.n[2] * g.n[1]));
		const long long idx = g.index(i0, i1, i2);
		double acc = 0;
		for (int c = 0; c < g.M; c++) { acc += (c + 1) * (double) pde[c * g.comp + idx]; }
		sum += acc;
	}
	block_sum<256>(sum, dummy);
	if (threadIdx.x == 0) { partial[blockIdx.x] = sum; }
}

// final reduction of the detector partials in index order (the order the host used to sum them in), and the z-axis line
// as doubles: results of gcmb_cubic_seismo_begin, fetched by one asynchronous copy
GCMB_GLOBAL void k_detector_final(const double* part_sum, const long long* part_count, int blocks, double* out2) {
	double s = 0;
	long long c = 0;
	for (int i = 0; i < blocks; i++) { s += part_sum[i]; c += part_count[i]; }
	out2[0] = s;
	out2[1] = (double) c;
}

template<class R>
GCMB_GLOBAL void k_line_to_double(const R* src, double* dst, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { dst[i] = (double) src[i]; }
}

GCMB_GLOBAL void k_triangle_interpolate(int mode, long long n, const double* points, const double* values, const double* grads,
                                        const double* queries, double* out, int* status) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { tri2::query_thread(mode, i, points, values, grads, queries, out, status); }
}

// gather of a box of nodes into a dense array [component][box node] (asynchronous snapshots, thin-column checks)
template<class R>
GCMB_GLOBAL void k_gather_box(Geom g, const R* pde, R* out, int b0, int b1, int b2, int e0, int e1, int e2) {
	const long long n = (long long) e0 * e1 * e2;
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) { return; }
	const int i2 = (int) (t % e2);
	const int i1 = (int) ((t / e2) % e1);
	const int i0 = (int) (t / ((long long) e2 * e1));
	const long long idx = g.index(b0 + i0, b1 + i1, b2 + i2);
	for (int c = 0; c < g.M; c++) { out[t * g.M + c] = pde[c * g.comp + idx]; }
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

}  // namespace gcmb

struct BorderCond {
	int cond = 0, dir = 0;
	uint8_t* mask[2] = {nullptr, nullptr};  // device, may be null (= side not handled)
	bool side_on[2] = {false, false};
	std::vector<int> q;
};

struct gcmb_body {
	gcmb_ctx* ctx = nullptr;
	Geom g;
	void* buf[2] = {nullptr, nullptr};    // buf[cur], buf[1-cur]: R = double or float (ctx->real_bytes)
	int cur = 0;
	uint8_t* node_table = nullptr;
	int n_tables = 0;
	std::vector<double> U, U1, L;         // host copies [n][D][M*M] / [n][D][M]
	void* tables = nullptr;               // device StageTableT<R> [n*D]
	std::vector<StageTable> host_tables;  // the tables in double
	std::vector<unsigned char> host_tables_r;  // the tables as StageTableT<R> (what the device holds)
	double tables_tau = NAN;
	bool any_k0 = false;                  // some characteristic foot beyond the first cell (Courant >= 1)
	int max_k0 = 0;
	int pattern_of_dir[3] = {-1, -1, -1};
	void* packed[3] = {nullptr, nullptr, nullptr};  // device: packed non-zero coefficients per direction
	int packed_variant[3] = {-1, -1, -1};
	std::string kernel_name[3];
	std::map<int, BorderCond> borders;    // ordered by condition number
	uint8_t* detector_mask = nullptr;
	int detector_code = 0;
	void* decay_dev = nullptr;
	bool halo_inflight = false;           // my ghost x-planes are being received (ctx->halo_pending)
	// ghost fill of the z faces handed to the next marching stage (gcmb_cubic_stage_fill_next_border)
	bool zfill_armed = false;
	ZFaceFill<double> zfill;
	// border condition of the contiguous axis' faces that gcmb_cubic_border_apply has DEFERRED to the stage of that
	// direction (the tile kernel produces the ghost nodes in shared memory); anything else that could see or overwrite
	// these ghost nodes first makes the fill kernel run after all (flush_deferred_border)
	bool border_deferred = false;
	int border_deferred_dir = 0;
	std::vector<double> border_deferred_values;
	bool border_in_tile = false;          // the last gcmb_cubic_stage consumed a deferred condition inside its kernel
	void* gather_dev = nullptr;           // staging of asynchronous box read-backs
	size_t gather_bytes = 0;
	// asynchronous seismogram taps (gcmb_cubic_seismo_begin/end): {sum, count, line...} on the device and in pinned memory
	double* seis_dev = nullptr;
	double* seis_host = nullptr;
	cudaEvent_t ev_seis = nullptr;
	bool seis_pending = false, seis_detector = false;
	int seis_line = 0;
};

namespace {

int flush_profile(gcmb_ctx* ctx) {
	if (ctx->spans.empty()) { return GCMB_OK; }
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ctx->edge_stream) { GCMB_CUDA(cudaStreamSynchronize(ctx->edge_stream)); }
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

int kernel_set(const gcmb_ctx* ctx) { return ctx->real_bytes == 4 ? SET_F32 : (ctx->fma ? SET_F64_FMA : SET_F64_EXACT); }

// variant of the specialised kernels that serves this body at the current time step (-1: none)
int sparse_variant(const gcmb_body* b) {
	switch (b->g.bs) {
		case 1: return VAR_BS1;
		case 2: return b->any_k0 ? VAR_BS2 : VAR_BS2_K0;
		case 3: return VAR_BS3;
		default: return -1;
	}
}

template<class R>
void convert_table(const StageTable& s, StageTableT<R>& d) {
	for (int i = 0; i < MAXM * MAXM; i++) { d.U[i] = (R) s.U[i]; d.U1[i] = (R) s.U1[i]; }
	for (int i = 0; i < MAXM * MAXBS; i++) { d.F[i] = (R) s.F[i]; }
	for (int i = 0; i < MAXM; i++) { d.k0[i] = s.k0[i]; d.dir[i] = s.dir[i]; }
}

// Newton factors, foot cells and sides of every (table, direction) for time step tau
// (reference engine/cubic/GridCharacteristicMethod.hpp:56-59,78-84 + EqualDistanceLineInterpolator.hpp:20,63)
template<class R>
int build_tables(gcmb_body* b, double tau) {
	const Geom& g = b->g;
	const int M = g.M, D = g.D;
	b->tables_tau = NAN;  // a failure below must not leave half-built tables marked as current
	std::vector<StageTable>& host = b->host_tables;
	host.assign((size_t) b->n_tables * D, StageTable());
	b->any_k0 = false;
	b->max_k0 = 0;
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
				b->max_k0 = std::max(b->max_k0, (int) k0);
				for (int i = 1; i <= g.bs; i++) { T.F[k * MAXBS + i - 1] = (q - i + 1) / i; }
			}
		}
	}
	b->host_tables_r.resize(host.size() * sizeof(StageTableT<R>));
	StageTableT<R>* hr = reinterpret_cast<StageTableT<R>*>(b->host_tables_r.data());
	for (size_t i = 0; i < host.size(); i++) { convert_table(host[i], hr[i]); }
	GCMB_CUDA(cudaMemcpyAsync(b->tables, hr, host.size() * sizeof(StageTableT<R>), cudaMemcpyHostToDevice, b->ctx->stream));
	// packed copies of the structurally non-zero coefficients for the specialised kernels
	const int variant = sparse_variant(b);
	std::vector<R> pk;
	for (int s = 0; s < D; s++) {
		cudaFree(b->packed[s]);
		b->packed[s] = nullptr;
		b->packed_variant[s] = -1;
		if (b->pattern_of_dir[s] < 0 || variant < 0) { continue; }
		const PatternInfo& P = pattern(b->pattern_of_dir[s]);
		const bool k0rt = variant_k0rt(variant);
		pk.assign((size_t) b->n_tables * (MAXM * MAXM * 2 + MAXM * (MAXBS + 1)), R(0));
		int size = 0;
		bool shares = true;
		for (int t = 0; t < b->n_tables; t++) {
			R one[MAXM * MAXM * 2 + MAXM * (MAXBS + 1)];
			size = pack_table(P, g.bs, k0rt, hr[(size_t) t * D + s], one);
			std::memcpy(pk.data() + (size_t) t * size, one, (size_t) size * sizeof(R));
			shares = shares && table_shares_as_pattern(P, g.bs, hr[(size_t) t * D + s]);
		}
		if (!shares) { continue; }  // no packed table: the stage falls back to the dense kernels
		GCMB_CUDA(cudaMalloc(&b->packed[s], (size_t) b->n_tables * size * sizeof(R)));
		GCMB_CUDA(cudaMemcpyAsync(b->packed[s], pk.data(), (size_t) b->n_tables * size * sizeof(R), cudaMemcpyHostToDevice, b->ctx->stream));
		GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));
		b->packed_variant[s] = variant;
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
		if (sgn_ok && g.bs <= 3 && !getenv("GCMB_FORCE_DENSE")) {
			for (int p = 0; p < pattern_count() && chosen < 0; p++) {
				const PatternInfo& P = pattern(p);
				if (P.M != M || P.axis != s + g.shift) { continue; }
				bool ok = true;
				for (int k = 0; k < M && ok; k++) {
					ok = P.sgn[k] == sgn[k] && (um[k] & ~P.um[k]) == 0 && (u1m[k] & ~P.u1m[k]) == 0;
				}
				if (ok) { chosen = p; }
			}
		}
		b->pattern_of_dir[s] = chosen;
		b->kernel_name[s] = "unset";
	}
}

template<class R> R* layer(gcmb_body* b, int which) { return static_cast<R*>(b->buf[which]); }

// what gcmb_cubic_stage launches for `dir` at the current tables; fills kernel_name
StageLauncher pick_launcher(gcmb_body* b, int dir) {
	const int set = kernel_set(b->ctx);
	const int M = b->g.M, bs = b->g.bs;
	const int p = b->pattern_of_dir[dir];
	const int variant = sparse_variant(b);
	static const char* const set_name[N_SETS] = {"", "+fma", "/f32"};
	if (p >= 0 && variant >= 0 && b->packed[dir] && b->packed_variant[dir] == variant) {
		StageLauncher f = sparse_launcher(set, p, variant);
		if (f) {
			static const char* const var_name[N_VARIANTS] = {"/bs1", "/bs2", "/bs2+k0", "/bs3+k0"};
			b->kernel_name[dir] = std::string("sparse:") + pattern(p).name + var_name[variant] + set_name[set];
			return f;
		}
	}
	static const bool literal_dense = std::getenv("GCMB_DENSE_LITERAL") != nullptr;
	if (!literal_dense && bs <= 2 && b->max_k0 <= 1) {
		StageLauncher f = dense_k0_launcher(set, M, bs, b->any_k0);
		if (f) {
			b->kernel_name[dir] = std::string(b->n_tables == 1 ? "dense_k0_one:M" : "dense_k0:M") + std::to_string(M) + "/bs" + std::to_string(bs) +
			                      (b->any_k0 ? "+k0" : "") + set_name[set];
			return f;
		}
	}
	StageLauncher f = dense_launcher(set, M);
	b->kernel_name[dir] = std::string("dense:M") + std::to_string(M) + set_name[set];
	return f;
}

}  // namespace

#define GCMB_BY_REAL(ctx, call_double, call_float) ((ctx)->real_bytes == 4 ? (call_float) : (call_double))

// =============================================================================================
// C ABI
// =============================================================================================
// (every entry point below is declared extern "C" by include/gcm_b200.h)

const char* gcmb_last_error(void) { return g_error.c_str(); }
const char* gcmb_version(void) { return "gcm_b200 0.2 (sm_100a, fp64 / fp32)"; }

int gcmb_create(int device, int real_bytes, gcmb_ctx** out) {
	if (!out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "out is null"); }
	if (real_bytes != 8 && real_bytes != 4) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "real_bytes must be 8 (fp64) or 4 (fp32)"); }
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		GCMB_FAIL(GCMB_E_NO_DEVICE, "no CUDA device: gcm_b200 has no CPU path");
	}
	if (device < 0 || device >= n) { GCMB_FAIL(GCMB_E_INVALID_ARG, "device index out of range"); }
	GCMB_CUDA(cudaSetDevice(device));
	std::unique_ptr<gcmb_ctx> ctx(new gcmb_ctx);
	ctx->device = device;
	ctx->real_bytes = real_bytes;
	const char* fma = std::getenv("GCMB_FMA");
	ctx->fma = fma && fma[0] == '1';
	GCMB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
	ctx->stream = ctx->own_stream;
	GCMB_CUDA(cudaEventCreate(&ctx->timer_a));
	GCMB_CUDA(cudaEventCreate(&ctx->timer_b));
	GCMB_CUDA(cudaMalloc(&ctx->scratch, 4096 * sizeof(double)));
	*out = ctx.release();
	return GCMB_OK;
}

int gcmb_real_bytes(gcmb_ctx* ctx) { return ctx ? ctx->real_bytes : 0; }

int gcmb_set_fma(gcmb_ctx* ctx, int on) {
	if (!ctx) { GCMB_FAIL(GCMB_E_INVALID_ARG, "ctx is null"); }
	if (on && ctx->real_bytes == 8 && !dense_launcher(SET_F64_FMA, 9)) {
		GCMB_FAIL(GCMB_E_UNSUPPORTED, "this build of libgcm_b200.so holds no FMA-contracted fp64 kernels");
	}
	ctx->fma = on != 0;
	for (gcmb_body* b : ctx->bodies) { b->tables_tau = NAN; }
	return GCMB_OK;
}

void gcmb_destroy(gcmb_ctx* ctx) {
	if (!ctx) { return; }
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	while (!ctx->bodies.empty()) { gcmb_cubic_body_destroy(ctx->bodies.back()); }
	for (ProfileSpan& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
	if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); }
	if (ctx->edge_stream) { cudaStreamSynchronize(ctx->edge_stream); }
	if (ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->comm); }
	if (ctx->comm_stream) {
		cudaStreamDestroy(ctx->comm_stream); cudaStreamDestroy(ctx->edge_stream);
		cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_halo); cudaEventDestroy(ctx->ev_edge);
	}
	if (ctx->copy_stream) {
		cudaStreamSynchronize(ctx->copy_stream);
		cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_copy_ready); cudaEventDestroy(ctx->ev_copy_done);
	}
	cudaFree(ctx->scratch);
	cudaEventDestroy(ctx->timer_a);
	cudaEventDestroy(ctx->timer_b);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

int gcmb_set_stream(gcmb_ctx* ctx, void* cuda_stream) {
	if (!ctx) { GCMB_FAIL(GCMB_E_INVALID_ARG, "ctx is null"); }
	wait_halo(ctx);
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
	return GCMB_OK;
}

int gcmb_sync(gcmb_ctx* ctx) {
	GCMB_CUDA(cudaSetDevice(ctx->device));
	wait_halo(ctx);
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ctx->copy_stream) { GCMB_CUDA(cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }
	return GCMB_OK;
}

int gcmb_timer_start(gcmb_ctx* ctx) {
	wait_halo(ctx);
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
static int flush_deferred_border(gcmb_body* b);
#define GCMB_FLUSH_BORDER(body) do { const int rc_flush_ = flush_deferred_border(body); if (rc_flush_) { return rc_flush_; } } while (0)

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
	g.zoff = 16;  // >= MAXBS, multiple of 16 elements (128 B of doubles)
	g.pitch = ((g.zoff + g.n[2] + g.g[2]) + 15) / 16 * 16;
	g.plane = (long long) (g.n[1] + 2 * g.g[1]) * g.pitch;
	g.comp = ((long long) (g.n[0] + 2 * g.g[0]) * g.plane + 31) / 32 * 32;
	// slack behind the last component: the bulk-copy kernels fetch whole 128/256-node row segments
	const size_t slack = 512;
	const size_t bytes = ((size_t) g.comp * M + slack) * (size_t) ctx->real_bytes;
	for (int i = 0; i < 2; i++) {
		if (cudaMalloc(&b->buf[i], bytes) != cudaSuccess) {
			cudaGetLastError();
			cudaFree(b->buf[0]);
			GCMB_FAIL(GCMB_E_CUDA, "out of device memory for the PDE time layers");
		}
		GCMB_CUDA(cudaMemsetAsync(b->buf[i], 0, bytes, ctx->stream));
	}
	GCMB_CUDA(cudaMalloc(&b->node_table, (size_t) g.comp + slack));  // slack: id copies of the tile kernels
	GCMB_CUDA(cudaMemsetAsync(b->node_table, 0, (size_t) g.comp + slack, ctx->stream));
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
	wait_halo(ctx);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); }
	ctx->bytes -= 2 * ((size_t) b->g.comp * b->g.M + 512) * (size_t) ctx->real_bytes + (size_t) b->g.comp;
	cudaFree(b->buf[0]);
	cudaFree(b->buf[1]);
	cudaFree(b->node_table);
	cudaFree(b->tables);
	for (int s = 0; s < 3; s++) { cudaFree(b->packed[s]); }
	cudaFree(b->decay_dev);
	cudaFree(b->detector_mask);
	cudaFree(b->gather_dev);
	cudaFree(b->seis_dev);
	if (b->seis_host) { cudaFreeHost(b->seis_host); }
	if (b->ev_seis) { cudaEventDestroy(b->ev_seis); }
	for (auto& kv : b->borders) { cudaFree(kv.second.mask[0]); cudaFree(kv.second.mask[1]); }
	ctx->bodies.erase(std::remove(ctx->bodies.begin(), ctx->bodies.end(), b), ctx->bodies.end());
	delete b;
}

template<class R>
static int transfer(gcmb_body* b, void* host, int with_ghosts, bool to_device) {
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	GCMB_FLUSH_BORDER(b);
	const int e0 = with_ghosts ? g.n[0] + 2 * g.g[0] : g.n[0];
	const long long e12 = (long long) (with_ghosts ? g.n[1] + 2 * g.g[1] : g.n[1]) * (with_ghosts ? g.n[2] + 2 * g.g[2] : g.n[2]);
	const long long per_plane = e12 * g.M;  // reals per slice of internal axis 0
	int chunk = (int) std::max<long long>(1, (64LL << 20) / std::max<long long>(1, per_plane));
	chunk = std::min(chunk, e0);
	R* stage = nullptr;
	GCMB_CUDA(cudaMalloc(&stage, (size_t) chunk * per_plane * sizeof(R)));
	int rc = GCMB_OK;
	for (int x0 = 0; x0 < e0 && rc == GCMB_OK; x0 += chunk) {
		const int x1 = std::min(e0, x0 + chunk);
		const long long n = (long long) (x1 - x0) * e12;
		R* hp = static_cast<R*>(host) + (long long) x0 * per_plane;
		XferArgs<R> x;
		x.soa = layer<R>(b, b->cur); x.aos = stage; x.g = g; x.with_ghosts = with_ghosts; x.x_begin = x0; x.x_end = x1;
		cudaError_t e = cudaSuccess;
		if (to_device) {
			e = cudaMemcpyAsync(stage, hp, (size_t) n * g.M * sizeof(R), cudaMemcpyHostToDevice, ctx->stream);
		}
		if (e == cudaSuccess) {
			Launch l(ctx, 6);
			GCMB_LAUNCH(k_xfer<R>, (unsigned) ((n + 255) / 256), 256, ctx->stream, x, n, to_device ? 1 : 0);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess && !to_device) {
			e = cudaMemcpyAsync(hp, stage, (size_t) n * g.M * sizeof(R), cudaMemcpyDeviceToHost, ctx->stream);
		}
		if (e == cudaSuccess) { e = cudaStreamSynchronize(ctx->stream); }
		if (e != cudaSuccess) { set_error(std::string("state transfer: ") + cudaGetErrorString(e)); rc = GCMB_E_CUDA; }
	}
	cudaFree(stage);
	return rc;
}

int gcmb_cubic_upload_state(gcmb_body* body, const void* aos_pde, int with_ghosts) {
	if (!body || !aos_pde) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	GCMB_FLUSH_BORDER(body);
	return GCMB_BY_REAL(body->ctx, transfer<double>(body, const_cast<void*>(aos_pde), with_ghosts, true),
	                    transfer<float>(body, const_cast<void*>(aos_pde), with_ghosts, true));
}

int gcmb_cubic_download_state(gcmb_body* body, void* aos_pde, int with_ghosts) {
	if (!body || !aos_pde) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (with_ghosts) { GCMB_FLUSH_BORDER(body); }
	return GCMB_BY_REAL(body->ctx, transfer<double>(body, aos_pde, with_ghosts, false), transfer<float>(body, aos_pde, with_ghosts, false));
}

// ---- asynchronous read-back of a box of real nodes (snapshots that do not stall the time loop) ----
template<class R>
static int box_begin(gcmb_body* b, const int* box_min, const int* extent, void* pinned_host) {
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	int lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
	long long n = 1;
	for (int i = 0; i < g.D; i++) {
		const int a = i + g.shift;
		lo[a] = box_min[i]; ext[a] = extent[i];
		if (extent[i] < 1 || box_min[i] < -g.g[a] || box_min[i] + extent[i] > g.n[a] + g.g[a]) { GCMB_FAIL(GCMB_E_INVALID_ARG, "box leaves the grid"); }
		n *= extent[i];
	}
	GCMB_CUDA(cudaSetDevice(ctx->device));
	if (!ctx->copy_stream) {
		GCMB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy_ready, cudaEventDisableTiming));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming));
	}
	const size_t bytes = (size_t) n * g.M * sizeof(R);
	if (b->gather_bytes < bytes) {
		GCMB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
		cudaFree(b->gather_dev);
		b->gather_dev = nullptr;
		b->gather_bytes = 0;
		GCMB_CUDA(cudaMalloc(&b->gather_dev, bytes));
		b->gather_bytes = bytes;
	}
	// the gather runs on the copy stream behind everything enqueued so far; later kernels that overwrite the
	// layer wait for it (Launch -> wait_copy), the PCIe transfer itself overlaps them
	wait_halo(ctx);
	GCMB_CUDA(cudaEventRecord(ctx->ev_copy_ready, ctx->stream));
	GCMB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_ready, 0));
	GCMB_LAUNCH(k_gather_box<R>, (unsigned) ((n + 255) / 256), 256, ctx->copy_stream, g, (const R*) layer<R>(b, b->cur),
	            static_cast<R*>(b->gather_dev), lo[0], lo[1], lo[2], ext[0], ext[1], ext[2]);
	ctx->launches++;
	GCMB_CUDA(cudaGetLastError());
	GCMB_CUDA(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
	ctx->copy_pending = true;
	GCMB_CUDA(cudaMemcpyAsync(pinned_host, b->gather_dev, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
	return GCMB_OK;
}

int gcmb_cubic_download_box_begin(gcmb_body* body, const int* box_min, const int* extent, void* host) {
	if (!body || !box_min || !extent || !host) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	GCMB_FLUSH_BORDER(body);
	return GCMB_BY_REAL(body->ctx, box_begin<double>(body, box_min, extent, host), box_begin<float>(body, box_min, extent, host));
}

int gcmb_cubic_download_box_end(gcmb_body* body) {
	if (!body) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	gcmb_ctx* ctx = body->ctx;
	GCMB_CUDA(cudaSetDevice(ctx->device));
	if (ctx->copy_stream) { GCMB_CUDA(cudaStreamSynchronize(ctx->copy_stream)); }
	return GCMB_OK;
}

int gcmb_triangle_interpolate(gcmb_ctx* ctx, int mode, int n, const double* points, const double* values, const double* gradients,
                              const double* queries, double* out, int* status) {
	if (!ctx || !points || !values || !queries || !out || !status) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (mode < 0 || mode > 4 || n < 1) { GCMB_FAIL(GCMB_E_INVALID_ARG, "mode must be 0..4 and n positive"); }
	if (mode >= 1 && mode <= 3 && !gradients) { GCMB_FAIL(GCMB_E_INVALID_ARG, "the quadratic interpolants need the gradients"); }
	GCMB_CUDA(cudaSetDevice(ctx->device));
	const int np = mode == 4 ? 4 : 3;
	const size_t nn = (size_t) n;
	double *d_p = nullptr, *d_v = nullptr, *d_g = nullptr, *d_q = nullptr, *d_o = nullptr;
	int* d_s = nullptr;
	int rc = GCMB_OK;
	auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == GCMB_OK) { set_error(std::string("gcmb_triangle_interpolate: ") + cudaGetErrorString(e)); rc = GCMB_E_CUDA; } };
	fail(cudaMalloc(&d_p, nn * np * 2 * sizeof(double)));
	fail(cudaMalloc(&d_v, nn * np * sizeof(double)));
	fail(cudaMalloc(&d_g, nn * 3 * 2 * sizeof(double)));
	fail(cudaMalloc(&d_q, nn * 2 * sizeof(double)));
	fail(cudaMalloc(&d_o, nn * sizeof(double)));
	fail(cudaMalloc(&d_s, nn * sizeof(int)));
	if (rc == GCMB_OK) {
		fail(cudaMemcpyAsync(d_p, points, nn * np * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		fail(cudaMemcpyAsync(d_v, values, nn * np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		if (gradients && mode != 4) { fail(cudaMemcpyAsync(d_g, gradients, nn * 3 * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); }
		fail(cudaMemcpyAsync(d_q, queries, nn * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		{
			Launch l(ctx, 7);
			GCMB_LAUNCH(k_triangle_interpolate, (unsigned) ((n + 127) / 128), 128, ctx->stream, mode, (long long) n, (const double*) d_p, (const double*) d_v,
			            (const double*) d_g, (const double*) d_q, d_o, d_s);
		}
		fail(cudaGetLastError());
		fail(cudaMemcpyAsync(out, d_o, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		fail(cudaMemcpyAsync(status, d_s, nn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		fail(cudaStreamSynchronize(ctx->stream));
	}
	cudaFree(d_p); cudaFree(d_v); cudaFree(d_g); cudaFree(d_q); cudaFree(d_o); cudaFree(d_s);
	return rc;
}

int gcmb_host_alloc_pinned(size_t bytes, void** out) {
	if (!out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "out is null"); }
	GCMB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
	return GCMB_OK;
}

void gcmb_host_free_pinned(void* p) {
	if (p) { cudaFreeHost(p); }
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
	wait_halo(ctx);
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));  // stages in flight may still read the old tables
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

template<class R>
static int add_vector(gcmb_body* b, const double* vector_M, int area_kind, const double* params) {
	VecArg<R> v;
	std::memset(&v, 0, sizeof v);
	for (int c = 0; c < b->g.M; c++) { v.v[c] = (R) vector_M[c]; }
	{
		Launch l(b->ctx, 6);
		GCMB_LAUNCH(k_add_vector<R>, node_grid(b->g, 128), 128, b->ctx->stream, b->g, layer<R>(b, b->cur), v, make_area(area_kind, params));
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

int gcmb_cubic_add_vector_in_area(gcmb_body* b, const double* vector_M, int area_kind, const double* params) {
	if (!b || !vector_M) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (area_kind < 0 || area_kind > 3) { GCMB_FAIL(GCMB_E_INVALID_ARG, "unknown area kind"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	GCMB_FLUSH_BORDER(b);
	return GCMB_BY_REAL(b->ctx, add_vector<double>(b, vector_M, area_kind, params), add_vector<float>(b, vector_M, area_kind, params));
}

// ---- borders ---------------------------------------------------------------------------------
static int border_register(gcmb_body* b, int cond, int dir, int n_q, const int* q_codes, BorderCond** out) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_FLUSH_BORDER(b);  // (a deferred fill belongs to the conditions registered when it was requested)
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
		// a mask that selects every node of the face is no mask (an infinite area turned into a mask by the caller, as the
		// reference-side binding does for every condition): the whole-face paths -- the fill inside the stage kernel among
		// them -- apply
		bool all = true;
		for (long long i = 0; i < nf && all; i++) { all = src[s][i] != 0; }
		if (all) { continue; }
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

static int border_values_check(gcmb_body* b, int dir, int n_values, const double* values) {
	int need = 0;
	for (auto& kv : b->borders) { if (kv.second.dir == dir) { need += (int) kv.second.q.size(); } }
	if (need != n_values || (need > 0 && !values)) { GCMB_FAIL(GCMB_E_INVALID_ARG, "wrong number of border values"); }
	return GCMB_OK;
}

template<class R>
static int border_apply(gcmb_body* b, int dir, const double* values) {
	const Geom& g = b->g;
	const int axis = dir + g.shift;
	const long long nf = face_size(g, axis);
	int used = 0;
	for (auto& kv : b->borders) {
		BorderCond& c = kv.second;
		if (c.dir != dir) { continue; }
		for (int s = 0; s < 2; s++) {
			if (!c.side_on[s]) { continue; }
			BorderArgs<R> a;
			std::memset(&a, 0, sizeof a);
			a.pde = layer<R>(b, b->cur); a.mask = c.mask[s]; a.g = g; a.axis = axis; a.side = s;
			a.nq = (int) c.q.size();
			for (int i = 0; i < a.nq; i++) { a.q[i] = c.q[(size_t) i]; a.val[i] = (R) values[used + i]; }
			const long long n = axis == 2 ? nf : nf * g.bs;
			Launch l(b->ctx, 3);
			// z faces: whole-sector ghost writes where the row layout allows (see border_thread_zsector)
			constexpr int SECT = 32 / (int) sizeof(R);
			static const bool no_sector = std::getenv("GCMB_BORDER_LAYERWISE") != nullptr;
			if (axis == 2 && !no_sector && g.bs <= SECT && g.zoff % SECT == 0 && g.zoff >= SECT && g.n[2] > g.bs && (s == 0 || g.n[2] % SECT == 0)) {
				GCMB_LAUNCH(k_border_zsector<R>, (unsigned) ((nf + 127) / 128), 128, b->ctx->stream, a, nf);
				continue;
			}
			GCMB_LAUNCH(k_border<R>, (unsigned) ((n + 127) / 128), 128, b->ctx->stream, a, nf);
		}
		used += (int) c.q.size();
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

// The deferred border condition, if any, applied by the fill kernel now.  Called by every entry point that reads or
// writes ghost nodes of the current layer, or that computes a stage the deferred fill does not ride on.
static int flush_deferred_border(gcmb_body* b) {
	if (!b->border_deferred) { return GCMB_OK; }
	b->border_deferred = false;
	const double* values = b->border_deferred_values.data();
	return GCMB_BY_REAL(b->ctx, border_apply<double>(b, b->border_deferred_dir, values), border_apply<float>(b, b->border_deferred_dir, values));
}

static bool zface_conditions(gcmb_body* b, int dir, const double* values, ZFaceFill<double>& zf);

int gcmb_cubic_border_apply(gcmb_body* b, int dir, int n_values, const double* values) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	int rc = border_values_check(b, dir, n_values, values);
	if (rc) { return rc; }
	rc = flush_deferred_border(b);
	if (rc) { return rc; }
	// Faces across the contiguous axis, whole-face conditions on plain components: nothing is launched here.  The stage of
	// this direction -- the next call on this body in the reference's time step (cubic/Engine.cpp:94-111), unless a
	// contact copy comes between -- mirrors the ghost nodes inside the shared-memory rows it stages anyway, so the
	// condition costs no pass over the faces in HBM (1.8 ms of an 83 ms step at 1024^3, profiles/r2_variants.md call 11).
	// GCMB_ZTILE_BORDER=0: always the fill kernel (measurements).
	static const bool off = std::getenv("GCMB_ZTILE_BORDER") != nullptr && std::getenv("GCMB_ZTILE_BORDER")[0] == '0';
	ZFaceFill<double> zf;
	if (!off && dir == b->g.D - 1 && b->g.n[2] > b->g.bs && b->tables && zface_conditions(b, dir, values, zf)) {
		b->border_deferred = true;
		b->border_deferred_dir = dir;
		b->border_deferred_values.assign(values, values + n_values);
		return GCMB_OK;
	}
	return GCMB_BY_REAL(b->ctx, border_apply<double>(b, dir, values), border_apply<float>(b, dir, values));
}

// The ghost fill of the faces across direction `dir` as data for a stage kernel: possible when, on every face, the LAST
// registered condition (the one whose values survive, BorderConditions.hpp:81-95) covers the whole face with plain
// components.
static bool zface_conditions(gcmb_body* b, int dir, const double* values, ZFaceFill<double>& zf) {
	std::memset(&zf, 0, sizeof zf);
	int used = 0;
	for (auto& kv : b->borders) {
		const BorderCond& c = kv.second;
		if (c.dir != dir) { continue; }
		for (int s = 0; s < 2; s++) {
			if (!c.side_on[s]) { continue; }
			if (c.mask[s]) { return false; }
			zf.on[s] = 1;
			zf.set[s] = 0;
			for (size_t i = 0; i < c.q.size(); i++) {
				if (c.q[i] < 0) { return false; }
				// a later Set of the same component wins, as in the reference's loop over the condition's values
				zf.set[s] |= 1u << c.q[i];
				zf.add[s][c.q[i]] = 2 * values[used + (int) i];
			}
		}
		used += (int) c.q.size();
	}
	return zf.on[0] || zf.on[1];
}

// Can the ghost fill of direction `dir` ride on the marching stage that WRITES the layer (whole-warp sector stores)?
static bool zfill_possible(gcmb_body* b, int dir, const double* values, ZFaceFill<double>& zf) {
#ifdef GCMB_EMUL
	return false;
#else
	// Off unless GCMB_FUSED_BORDER=1.  Measured at 1024^3 (profiles/r2_variants.md): the 2 x 9 sector stores per row cost the
	// marching kernel as much as the separate fill kernel takes (bulk-copy variant: y stage 30.3 -> 32.6 ms against 2.7 ms
	// of k_border; the LDGSTS variant collapses to 80 ms), because the ghost sector lies in a 128-byte line of its own:
	// the partial-line write is what costs, whoever issues it.  The contiguous-axis stage fills the ghosts of the rows it
	// has staged in shared memory instead (gcmb_cubic_stage_with_border), which costs no HBM traffic at all.
	static const bool on = std::getenv("GCMB_FUSED_BORDER") != nullptr && std::getenv("GCMB_FUSED_BORDER")[0] == '1';
	const Geom& g = b->g;
	if (!on || dir != g.D - 1 || g.D < 2 || g.n[2] % 32 != 0 || g.n[2] < 64 || g.bs > 4) { return false; }
	return zface_conditions(b, dir, values, zf);
#endif
}

// ---- contacts --------------------------------------------------------------------------------
template<class R>
static int contact_apply(gcmb_body* a, const gcmb_body* b, const int* boxA_min, const int* boxB_min, const int* extent) {
	ContactArgs<R> c;
	std::memset(&c, 0, sizeof c);
	c.a = layer<R>(a, a->cur); c.b = static_cast<const R*>(b->buf[b->cur]); c.ga = a->g; c.gb = b->g;
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
		GCMB_LAUNCH(k_contact<R>, (unsigned) ((n + 127) / 128), 128, a->ctx->stream, c, n);
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

int gcmb_cubic_contact_apply(gcmb_body* a, const gcmb_body* b, const int* boxA_min,
                             const int* boxB_min, const int* extent) {
	if (!a || !b || !boxA_min || !boxB_min || !extent) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (a->ctx != b->ctx) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bodies belong to different contexts"); }
	if (a->g.D != b->g.D || a->g.M != b->g.M || a->g.bs != b->g.bs) { GCMB_FAIL(GCMB_E_INVALID_ARG, "bodies are not compatible"); }
	GCMB_CUDA(cudaSetDevice(a->ctx->device));
	// the reference's order is border conditions, contact copies, stage: a deferred border fill must not come after the copy
	GCMB_FLUSH_BORDER(a);
	GCMB_FLUSH_BORDER(const_cast<gcmb_body*>(b));
	return GCMB_BY_REAL(a->ctx, contact_apply<double>(a, b, boxA_min, boxB_min, extent), contact_apply<float>(a, b, boxA_min, boxB_min, extent));
}

// ---- stage -----------------------------------------------------------------------------------
template<class R>
static int stage_impl(gcmb_body* b, int dir, double tau) {
	if (!(b->tables_tau == tau)) {
		const int rc = build_tables<R>(b, tau);
		if (rc) { return rc; }
	}
	StageArgsT<R> a;
	std::memset(&a, 0, sizeof a);
	a.cur = layer<R>(b, b->cur);
	a.nxt = layer<R>(b, 1 - b->cur);
	a.node_table = b->node_table;
	a.tables = static_cast<const StageTableT<R>*>(b->tables);
	a.packed = static_cast<const R*>(b->packed[dir]);
	a.n_tables = b->n_tables;
	a.g = b->g;
	a.axis = dir + b->g.shift;
	a.dir = dir;
	a.x_begin = 0;
	a.x_end = b->g.n[0];
	a.host_tables = reinterpret_cast<const StageTableT<R>*>(b->host_tables_r.data());
	StageLauncher launch = pick_launcher(b, dir);
	if (!launch) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "no stage kernel for this PDE size / arithmetic type in this build"); }
	if (b->zfill_armed) {
		b->zfill_armed = false;
		// only the kernels of the specialised patterns carry a ghost fill: the row-writing marching kernel (axis 1) for
		// the NEXT stage's faces, the tile kernel of the contiguous axis (axis 2) for its own
		if (a.axis == 0 || b->kernel_name[dir].compare(0, 7, "sparse:") != 0) { GCMB_FAIL(GCMB_E_INVALID_OP, "fused border fill is not available for this stage"); }
		a.zfill = 1;
		for (int s = 0; s < 2; s++) {
			a.zf.on[s] = b->zfill.on[s];
			a.zf.set[s] = b->zfill.set[s];
			for (int c = 0; c < MAXM; c++) { a.zf.add[s][c] = (R) b->zfill.add[s][c]; }
		}
	}
	gcmb_ctx* ctx = b->ctx;
	const int bs = b->g.g[0];
	if (b->halo_inflight && ctx->halo_pending && a.axis == 0 && bs > 0 && b->g.n[0] > 2 * bs) {
		// the nodes at least `bs` planes away from both slab faces read no ghost plane: they go first, while the
		// halo exchange is still in flight; the two boundary strips run on a stream of their own behind the
		// exchange, so that the interiors of the context's other bodies are not held back either
		StageArgsT<R> part = a;
		part.x_begin = bs; part.x_end = b->g.n[0] - bs;
		ctx->halo_defer = true;
		{ Launch l(ctx, a.axis); launch(&part, ctx->stream); }
		if (!ctx->edge_waits_halo) {
			GCMB_CUDA(cudaStreamWaitEvent(ctx->edge_stream, ctx->ev_halo, 0));
			ctx->edge_waits_halo = true;
		}
		part.x_begin = 0; part.x_end = bs;
		{ Launch l(ctx, a.axis, ctx->edge_stream); launch(&part, ctx->edge_stream); }
		part.x_begin = b->g.n[0] - bs; part.x_end = b->g.n[0];
		{ Launch l(ctx, a.axis, ctx->edge_stream); launch(&part, ctx->edge_stream); }
		ctx->halo_defer = false;
		GCMB_CUDA(cudaEventRecord(ctx->ev_edge, ctx->edge_stream));
		ctx->edge_pending = true;
	} else {
		Launch l(ctx, a.axis);
		launch(&a, ctx->stream);
	}
	b->halo_inflight = false;
	GCMB_CUDA(cudaGetLastError());
	b->cur = 1 - b->cur;  // swapCurrAndNextPdeTimeLayer
	return GCMB_OK;
}

int gcmb_cubic_stage(gcmb_body* b, int dir, double tau) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (dir < 0 || dir >= b->g.D) { GCMB_FAIL(GCMB_E_INVALID_ARG, "direction out of range"); }
	if (!b->tables) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	b->border_in_tile = false;
	if (b->border_deferred) {
		bool in_tile = false;
		ZFaceFill<double> zf;
		if (dir == b->border_deferred_dir && !b->zfill_armed && zface_conditions(b, dir, b->border_deferred_values.data(), zf)) {
			if (!(b->tables_tau == tau)) {  // the tables must be current before the launcher is known
				const int rc = GCMB_BY_REAL(b->ctx, build_tables<double>(b, tau), build_tables<float>(b, tau));
				if (rc) { return rc; }
			}
			pick_launcher(b, dir);
			in_tile = b->kernel_name[dir].compare(0, 7, "sparse:") == 0;
		}
		if (in_tile) {
			b->border_deferred = false;
			b->border_in_tile = true;
			b->zfill = zf;
			b->zfill_armed = true;
		} else {
			const int rc = flush_deferred_border(b);
			if (rc) { return rc; }
		}
	}
	return GCMB_BY_REAL(b->ctx, stage_impl<double>(b, dir, tau), stage_impl<float>(b, dir, tau));
}

int gcmb_cubic_stage_fill_next_border(gcmb_body* b, int dir, double tau, int next_dir, int n_values, const double* values, int* fused) {
	if (!b || !fused) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (dir < 0 || dir >= b->g.D || next_dir < 0 || next_dir >= b->g.D) { GCMB_FAIL(GCMB_E_INVALID_ARG, "direction out of range"); }
	if (!b->tables) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	int rc = border_values_check(b, next_dir, n_values, values);
	if (rc) { return rc; }
	*fused = 0;
	ZFaceFill<double> zf;
	if (dir != next_dir && dir + b->g.shift == 1 && zfill_possible(b, next_dir, values, zf)) {
		// the tables must be current before the launcher is known
		if (!(b->tables_tau == tau)) {
			rc = GCMB_BY_REAL(b->ctx, build_tables<double>(b, tau), build_tables<float>(b, tau));
			if (rc) { return rc; }
		}
		pick_launcher(b, dir);
		if (b->kernel_name[dir].compare(0, 7, "sparse:") == 0) {
			b->zfill = zf;
			b->zfill_armed = true;
			*fused = 1;
		}
	}
	return gcmb_cubic_stage(b, dir, tau);
}

int gcmb_cubic_stage_with_border(gcmb_body* b, int dir, double tau, int n_values, const double* values, int* fused) {
	if (!b || !fused) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (dir < 0 || dir >= b->g.D) { GCMB_FAIL(GCMB_E_INVALID_ARG, "direction out of range"); }
	if (!b->tables) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	*fused = 0;
	int rc = gcmb_cubic_border_apply(b, dir, n_values, values);
	if (rc) { return rc; }
	rc = gcmb_cubic_stage(b, dir, tau);
	*fused = (rc == GCMB_OK && b->border_in_tile) ? 1 : 0;
	return rc;
}

const char* gcmb_cubic_stage_kernel_name(gcmb_body* b, int dir) {
	if (!b || dir < 0 || dir >= b->g.D) { return ""; }
	return b->kernel_name[dir].c_str();
}

// ---- ode -------------------------------------------------------------------------------------
template<class R>
static int ode_maxwell(gcmb_body* b, const double* decay_per_table) {
	std::vector<R> d((size_t) b->n_tables);
	for (int i = 0; i < b->n_tables; i++) { d[(size_t) i] = (R) decay_per_table[i]; }
	GCMB_CUDA(cudaMemcpyAsync(b->decay_dev, d.data(), d.size() * sizeof(R), cudaMemcpyHostToDevice, b->ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(b->ctx->stream));  // the host array is a temporary
	{
		Launch l(b->ctx, 5);
		GCMB_LAUNCH(k_ode_maxwell<R>, node_grid(b->g, 128), 128, b->ctx->stream, b->g, layer<R>(b, b->cur), b->node_table, static_cast<const R*>(b->decay_dev));
	}
	GCMB_CUDA(cudaGetLastError());
	return GCMB_OK;
}

int gcmb_cubic_ode_maxwell(gcmb_body* b, const double* decay_per_table) {
	if (!b || !decay_per_table) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (b->n_tables < 1) { GCMB_FAIL(GCMB_E_INVALID_OP, "materials are not set"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	GCMB_FLUSH_BORDER(b);
	return GCMB_BY_REAL(b->ctx, ode_maxwell<double>(b, decay_per_table), ode_maxwell<float>(b, decay_per_table));
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

template<class R>
static int seismo(gcmb_body* b, double* sum, long long* count, int line_comp, double* line, int n_line, int line_i0, int line_i1) {
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	if (sum || count) {
		if (!b->detector_mask) { GCMB_FAIL(GCMB_E_INVALID_OP, "detector is not set"); }
		const long long nf = (long long) g.n[0] * g.n[1];
		const int blocks = (int) std::max<long long>(1, std::min<long long>(1024, (nf + 255) / 256));
		double* d_sum = ctx->scratch;
		long long* d_count = reinterpret_cast<long long*>(ctx->scratch + 1024);
		{
			Launch l(ctx, 7);
			GCMB_LAUNCH(k_detector<R>, blocks, 256, ctx->stream, g, (const R*) layer<R>(b, b->cur), b->detector_mask, b->detector_code, d_sum, d_count);
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
		if (line_i0 < 0 || line_i0 >= g.n[0] || line_i1 < 0 || line_i1 >= g.n[1]) { GCMB_FAIL(GCMB_E_INVALID_ARG, "line position is outside the body"); }
		wait_halo(ctx);
		const R* src = layer<R>(b, b->cur) + (long long) line_comp * g.comp + g.index(line_i0, line_i1, 0);
		std::vector<R> tmp((size_t) n_line);
		GCMB_CUDA(cudaMemcpyAsync(tmp.data(), src, (size_t) n_line * sizeof(R), cudaMemcpyDeviceToHost, ctx->stream));
		GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
		for (int i = 0; i < n_line; i++) { line[i] = (double) tmp[(size_t) i]; }
	}
	return GCMB_OK;
}

int gcmb_cubic_seismo(gcmb_body* b, double* sum, long long* count, int line_comp, double* line, int n_line) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	const int i0 = b->g.n[0] / 2, i1 = b->g.n[1] / 2;
	return GCMB_BY_REAL(b->ctx, seismo<double>(b, sum, count, line_comp, line, n_line, i0, i1), seismo<float>(b, sum, count, line_comp, line, n_line, i0, i1));
}

int gcmb_cubic_seismo_at(gcmb_body* b, double* sum, long long* count, int line_comp, double* line, int n_line, const int* line_node) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	int it[3] = {0, 0, 0};
	if (line) {
		if (!line_node) { GCMB_FAIL(GCMB_E_INVALID_ARG, "line_node is null"); }
		for (int i = 0; i < b->g.D - 1; i++) { it[i + b->g.shift] = line_node[i]; }
	}
	return GCMB_BY_REAL(b->ctx, seismo<double>(b, sum, count, line_comp, line, n_line, it[0], it[1]), seismo<float>(b, sum, count, line_comp, line, n_line, it[0], it[1]));
}

template<class R>
static int seismo_begin(gcmb_body* b, int with_detector, int line_comp, const int* line_node) {
	gcmb_ctx* ctx = b->ctx;
	const Geom& g = b->g;
	const int n_line = line_node ? g.n[2] : 0;
	if (b->seis_pending) { GCMB_FAIL(GCMB_E_INVALID_OP, "a seismogram read-back is already in flight: call gcmb_cubic_seismo_end first"); }
	if (!b->seis_dev) {
		GCMB_CUDA(cudaMalloc(&b->seis_dev, (size_t) (2 + g.n[2]) * sizeof(double)));
		GCMB_CUDA(cudaHostAlloc(&b->seis_host, (size_t) (2 + g.n[2]) * sizeof(double), cudaHostAllocDefault));
		GCMB_CUDA(cudaEventCreateWithFlags(&b->ev_seis, cudaEventDisableTiming));
	}
	if (with_detector) {
		if (!b->detector_mask) { GCMB_FAIL(GCMB_E_INVALID_OP, "detector is not set"); }
		const long long nf = (long long) g.n[0] * g.n[1];
		const int blocks = (int) std::max<long long>(1, std::min<long long>(1024, (nf + 255) / 256));
		double* d_sum = ctx->scratch;
		long long* d_count = reinterpret_cast<long long*>(ctx->scratch + 1024);
		{
			Launch l(ctx, 7);
			GCMB_LAUNCH(k_detector<R>, blocks, 256, ctx->stream, g, (const R*) layer<R>(b, b->cur), b->detector_mask, b->detector_code, d_sum, d_count);
		}
		{
			Launch l(ctx, 7);
			GCMB_LAUNCH(k_detector_final, 1, 1, ctx->stream, (const double*) d_sum, (const long long*) d_count, blocks, b->seis_dev);
		}
		GCMB_CUDA(cudaGetLastError());
		if (ctx->comm && ctx->n_ranks > 1) {
			// slabs of a decomposed body: the detector sums of all ranks, still on the stream (no host round trip)
			const ncclResult_t r = g_nccl.AllReduce(b->seis_dev, b->seis_dev, 2, ncclDouble, ncclSum, ctx->comm, ctx->stream);
			if (r != ncclSuccess) { GCMB_FAIL(GCMB_E_NCCL, std::string("ncclAllReduce -> ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error")); }
		}
	}
	if (n_line) {
		if (line_comp < 0 || line_comp >= g.M) { GCMB_FAIL(GCMB_E_INVALID_ARG, "line component out of range"); }
		int it[3] = {0, 0, 0};
		for (int i = 0; i < g.D - 1; i++) { it[i + g.shift] = line_node[i]; }
		if (it[0] < 0 || it[0] >= g.n[0] || it[1] < 0 || it[1] >= g.n[1]) { GCMB_FAIL(GCMB_E_INVALID_ARG, "line position is outside the body"); }
		const R* src = layer<R>(b, b->cur) + (long long) line_comp * g.comp + g.index(it[0], it[1], 0);
		Launch l(ctx, 7);
		GCMB_LAUNCH(k_line_to_double<R>, (unsigned) ((n_line + 255) / 256), 256, ctx->stream, src, b->seis_dev + 2, n_line);
		GCMB_CUDA(cudaGetLastError());
	}
	wait_halo(ctx);
	GCMB_CUDA(cudaMemcpyAsync(b->seis_host, b->seis_dev, (size_t) (2 + n_line) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GCMB_CUDA(cudaEventRecord(b->ev_seis, ctx->stream));
	b->seis_pending = true;
	b->seis_detector = with_detector != 0;
	b->seis_line = n_line;
	return GCMB_OK;
}

int gcmb_cubic_seismo_begin(gcmb_body* b, int with_detector, int line_comp, const int* line_node) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	return GCMB_BY_REAL(b->ctx, seismo_begin<double>(b, with_detector, line_comp, line_node), seismo_begin<float>(b, with_detector, line_comp, line_node));
}

int gcmb_cubic_seismo_end(gcmb_body* b, double* sum, long long* count, double* line, int n_line) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	if (!b->seis_pending) { GCMB_FAIL(GCMB_E_INVALID_OP, "no seismogram read-back in flight"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	GCMB_CUDA(cudaEventSynchronize(b->ev_seis));
	b->seis_pending = false;
	if (b->seis_detector) {
		if (sum) { *sum = b->seis_host[0]; }
		if (count) { *count = (long long) b->seis_host[1]; }
	}
	if (line) {
		if (n_line != b->seis_line) { GCMB_FAIL(GCMB_E_INVALID_ARG, "n_line does not match the line requested by gcmb_cubic_seismo_begin"); }
		std::memcpy(line, b->seis_host + 2, (size_t) n_line * sizeof(double));
	}
	return GCMB_OK;
}

template<class R>
static int checksum(gcmb_body* b, double* out) {
	gcmb_ctx* ctx = b->ctx;
	const int blocks = 1024;
	{
		Launch l(ctx, 7);
		GCMB_LAUNCH(k_checksum<R>, blocks, 256, ctx->stream, b->g, (const R*) layer<R>(b, b->cur), ctx->scratch);
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

int gcmb_cubic_checksum(gcmb_body* b, double* out) {
	if (!b || !out) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	return GCMB_BY_REAL(b->ctx, checksum<double>(b, out), checksum<float>(b, out));
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

int gcmb_halo_exchange_bodies(gcmb_body* const* bodies, int n) {
	if (!bodies || n < 1) { GCMB_FAIL(GCMB_E_INVALID_ARG, "no bodies"); }
	for (int i = 0; i < n; i++) {
		if (!bodies[i] || bodies[i]->ctx != bodies[0]->ctx) { GCMB_FAIL(GCMB_E_INVALID_ARG, "the bodies must belong to one context"); }
	}
	gcmb_ctx* ctx = bodies[0]->ctx;
	if (!ctx->comm || ctx->n_ranks == 1) { return GCMB_OK; }
	for (int i = 0; i < n; i++) {
		if (bodies[i]->g.g[0] == 0) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "slab decomposition needs the x axis to be the slowest internal axis (3-D grids)"); }
	}
	GCMB_CUDA(cudaSetDevice(ctx->device));
	for (int i = 0; i < n; i++) { GCMB_FLUSH_BORDER(bodies[i]); }
	// on a stream of its own, so that the interior of the following x stage (which reads no ghost plane) overlaps it
	static const bool overlap = !std::getenv("GCMB_NO_HALO_OVERLAP");
	if (overlap && !ctx->comm_stream) {
		int least = 0, greatest = 0;
		GCMB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		GCMB_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, greatest));
		GCMB_CUDA(cudaStreamCreateWithPriority(&ctx->edge_stream, cudaStreamNonBlocking, greatest));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
		GCMB_CUDA(cudaEventCreateWithFlags(&ctx->ev_edge, cudaEventDisableTiming));
	}
	wait_halo(ctx);
	cudaStream_t comm_stream = overlap ? ctx->comm_stream : ctx->stream;
	if (overlap) {
		GCMB_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
		GCMB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
	}
	// x-planes are contiguous inside every component volume: ghost planes [0,bs) and [n0+bs, n0+2bs),
	// outermost real planes [bs, 2bs) and [n0, n0+bs).  One group for all bodies: one NCCL launch per step.
	const ncclDataType_t type = ctx->real_bytes == 4 ? ncclFloat : ncclDouble;
	const size_t rb = (size_t) ctx->real_bytes;
	ncclResult_t failed = ncclSuccess;
	const char* what = "";
#define GCMB_NCCL_IN_GROUP(call) do { if (failed == ncclSuccess) { failed = (call); if (failed != ncclSuccess) { what = #call; } } } while (0)
	GCMB_NCCL(g_nccl.GroupStart());
	for (int i = 0; i < n; i++) {
		gcmb_body* b = bodies[i];
		const Geom& g = b->g;
		const size_t count = (size_t) g.g[0] * (size_t) g.plane;
		char* base = static_cast<char*>(b->buf[b->cur]);
		for (int c = 0; c < g.M; c++) {
			char* v = base + (size_t) c * (size_t) g.comp * rb;
			if (ctx->rank > 0) {
				GCMB_NCCL_IN_GROUP(g_nccl.Send(v + (size_t) g.g[0] * g.plane * rb, count, type, ctx->rank - 1, ctx->comm, comm_stream));
				GCMB_NCCL_IN_GROUP(g_nccl.Recv(v, count, type, ctx->rank - 1, ctx->comm, comm_stream));
			}
			if (ctx->rank < ctx->n_ranks - 1) {
				GCMB_NCCL_IN_GROUP(g_nccl.Send(v + (size_t) g.n[0] * g.plane * rb, count, type, ctx->rank + 1, ctx->comm, comm_stream));
				GCMB_NCCL_IN_GROUP(g_nccl.Recv(v + (size_t) (g.n[0] + g.g[0]) * g.plane * rb, count, type, ctx->rank + 1, ctx->comm, comm_stream));
			}
		}
	}
	const ncclResult_t ended = g_nccl.GroupEnd();  // always closed, also after a failed call inside
#undef GCMB_NCCL_IN_GROUP
	if (failed != ncclSuccess || ended != ncclSuccess) {
		set_error(std::string(failed != ncclSuccess ? what : "ncclGroupEnd") + " -> " +
		          (g_nccl.GetErrorString ? g_nccl.GetErrorString(failed != ncclSuccess ? failed : ended) : "nccl error"));
		return GCMB_E_NCCL;
	}
	if (overlap) {
		GCMB_CUDA(cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
		ctx->halo_pending = true;
		ctx->edge_waits_halo = false;
		for (int i = 0; i < n; i++) { bodies[i]->halo_inflight = true; }
	}
	ctx->launches++;
	return GCMB_OK;
}

int gcmb_cubic_halo_exchange(gcmb_body* b) {
	if (!b) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null body"); }
	return gcmb_halo_exchange_bodies(&b, 1);
}

size_t gcmb_cubic_halo_bytes(gcmb_body* b) {
	return b ? (size_t) b->g.M * (size_t) b->g.g[0] * (size_t) b->g.plane * (size_t) b->ctx->real_bytes : 0;
}

static int halo_host(gcmb_body* b, int side, void* host, bool get) {
	if (!b || !host) { GCMB_FAIL(GCMB_E_INVALID_ARG, "null argument"); }
	if (side != 0 && side != 1) { GCMB_FAIL(GCMB_E_INVALID_ARG, "side must be 0 or 1"); }
	const Geom& g = b->g;
	if (g.g[0] == 0) { GCMB_FAIL(GCMB_E_UNSUPPORTED, "slab decomposition needs the x axis to be the slowest internal axis (3-D grids)"); }
	GCMB_CUDA(cudaSetDevice(b->ctx->device));
	GCMB_FLUSH_BORDER(b);
	wait_halo(b->ctx);
	const size_t rb = (size_t) b->ctx->real_bytes;
	const size_t count = (size_t) g.g[0] * (size_t) g.plane;
	// real planes next to the face: [bs, 2bs) left, [n0, n0+bs) right; ghost planes: [0, bs) / [n0+bs, n0+2bs)
	const long long first = get ? (side == 0 ? g.g[0] : g.n[0]) : (side == 0 ? 0 : g.n[0] + g.g[0]);
	for (int c = 0; c < g.M; c++) {
		char* dev = static_cast<char*>(b->buf[b->cur]) + ((size_t) c * (size_t) g.comp + (size_t) first * g.plane) * rb;
		char* h = static_cast<char*>(host) + (size_t) c * count * rb;
		if (get) { GCMB_CUDA(cudaMemcpyAsync(h, dev, count * rb, cudaMemcpyDeviceToHost, b->ctx->stream)); }
		else { GCMB_CUDA(cudaMemcpyAsync(dev, h, count * rb, cudaMemcpyHostToDevice, b->ctx->stream)); }
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
	wait_halo(ctx);  // one communicator: the reduction is ordered behind a halo exchange in flight on every rank
	double* d = ctx->scratch + 2048;
	GCMB_CUDA(cudaMemcpyAsync(d, host_values, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	GCMB_NCCL(g_nccl.AllReduce(d, d, (size_t) n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
	GCMB_CUDA(cudaMemcpyAsync(host_values, d, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	GCMB_CUDA(cudaStreamSynchronize(ctx->stream));
	return GCMB_OK;
}

