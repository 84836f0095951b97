// C entry points of the host layer, for bindings (Python ctypes in gcm_b200/capi.py, tests, bench).
#include <cstring>
#include <memory>

#include "gcmb_host.hpp"

using namespace gcmb;

namespace {
thread_local std::string hostError;
struct Handle {
	std::shared_ptr<AbstractEngine> engine;
	cubic::EngineBase* base = nullptr;
	simplex::Engine* sx = nullptr;
};
cubic::EngineBase& cubicOf(void* handle) {
	cubic::EngineBase* e = static_cast<Handle*>(handle)->base;
	if (!e) { throw Exception(GCMB_E_INVALID_OP, "not a cubic engine"); }
	return *e;
}
simplex::Engine& simplexOf(void* handle) {
	simplex::Engine* e = static_cast<Handle*>(handle)->sx;
	if (!e) { throw Exception(GCMB_E_INVALID_OP, "not a simplex engine"); }
	return *e;
}
template<typename F> int guarded(F f) {
	try { f(); return GCMB_OK; }
	catch (const Exception& e) { hostError = e.what(); return e.code() == 0 ? GCMB_E_INVALID_OP : e.code(); }
	catch (const std::exception& e) { hostError = e.what(); return GCMB_E_INVALID_OP; }
}
}  // namespace

extern "C" {

const char* gcmb_host_last_error(void) { return hostError.c_str(); }

/// createEngine(parseTaskText(text)); slab_count > 1 decomposes every body along x across processes;
/// real_bytes 8 | 4 = the arithmetic type of the cubic engines (the reference's compile-time `real`);
/// flags: bit 0 = fp64 stage kernels with FMA contraction
int gcmb_host_engine_create2(const char* task_text, int device, int slab_rank, int slab_count,
		const void* nccl_id128, int real_bytes, int flags, void** out) {
	return guarded([&] {
		Task task = parseTaskText(task_text);
		task.device.device = device;
		task.device.slabRank = slab_rank;
		task.device.slabCount = slab_count;
		task.device.ncclUniqueId = nccl_id128;
		task.device.realBytes = real_bytes;
		task.device.fma = (flags & 1) != 0;
		std::unique_ptr<Handle> h(new Handle);
		h->engine = createEngine(task);
		h->base = dynamic_cast<cubic::EngineBase*>(h->engine.get());
		h->sx = dynamic_cast<simplex::Engine*>(h->engine.get());
		*out = h.release();
	});
}

int gcmb_host_engine_create(const char* task_text, int device, int slab_rank, int slab_count,
		const void* nccl_id128, void** out) {
	return gcmb_host_engine_create2(task_text, device, slab_rank, slab_count, nccl_id128, 8, 0, out);
}

void gcmb_host_engine_destroy(void* handle) { delete static_cast<Handle*>(handle); }

int gcmb_host_engine_run(void* handle) {
	return guarded([&] { static_cast<Handle*>(handle)->engine->run(); });
}

int gcmb_host_engine_advance(void* handle, int n) {
	return guarded([&] { static_cast<Handle*>(handle)->engine->advance(n); });
}

int gcmb_host_engine_info(void* handle, int* steps_done, double* time, double* tau) {
	return guarded([&] {
		*steps_done = static_cast<Handle*>(handle)->engine->stepsDone();
		*time = static_cast<Handle*>(handle)->engine->currentTime();
		*tau = static_cast<Handle*>(handle)->engine->currentTimeStep();
	});
}

/// D, M, sizes[3] of a body (sizes of THIS process' slab when decomposed)
int gcmb_host_engine_body_info(void* handle, size_t id, int* D, int* M, int* sizes, int* start) {
	return guarded([&] {
		auto mesh = cubicOf(handle).getMesh(id);
		*D = mesh->D;
		*M = mesh->M;
		for (int i = 0; i < 3; i++) { sizes[i] = mesh->sizes[(size_t) i]; start[i] = mesh->start[(size_t) i]; }
	});
}

/// real nodes of the current time layer, x slowest, M per node
int gcmb_host_engine_body_pde(void* handle, size_t id, double* out) {
	return guarded([&] {
		auto mesh = cubicOf(handle).getMesh(id);
		const auto& v = mesh->pdeRealNodes();
		std::memcpy(out, v.data(), v.size() * sizeof(double));
	});
}

void* gcmb_host_engine_body_handle(void* handle, size_t id) {
	void* ans = nullptr;
	guarded([&] { ans = cubicOf(handle).getMesh(id)->handle(); });
	return ans;
}

void* gcmb_host_engine_context(void* handle) {
	Handle* h = static_cast<Handle*>(handle);
	return h->base ? h->base->context() : (h->sx ? h->sx->context() : nullptr);
}

// ---- simplex engines ---------------------------------------------------------------------------
/// sizes_out = {nV, nC, nIncident}; arrays may be null to query the sizes only
int gcmb_host_simplex_triangulation(void* handle, int* sizes_out, double* xyz, int* cell_v, int* cell_n, int* cell_grid,
		int* inc_off, int* inc_cell) {
	return guarded([&] {
		const simplex::FlatTriangulation& t = simplexOf(handle).getTriangulation();
		sizes_out[0] = t.nV; sizes_out[1] = t.nC; sizes_out[2] = (int) t.incCell.size();
		if (!xyz) { return; }
		std::memcpy(xyz, t.xyz.data(), t.xyz.size() * sizeof(double));
		std::memcpy(cell_v, t.cellV.data(), t.cellV.size() * sizeof(int));
		std::memcpy(cell_n, t.cellN.data(), t.cellN.size() * sizeof(int));
		std::memcpy(cell_grid, t.cellGrid.data(), t.cellGrid.size() * sizeof(int));
		std::memcpy(inc_off, t.incOff.data(), t.incOff.size() * sizeof(int));
		std::memcpy(inc_cell, t.incCell.data(), t.incCell.size() * sizeof(int));
	});
}

/// info = {n_local, M, n_border_conditions}; reals = {average height, minimal height, maximal eigenvalue};
/// basis = the current calculation basis (9)
int gcmb_host_simplex_body_info(void* handle, size_t id, int* info, double* reals, double* basis) {
	return guarded([&] {
		simplex::Engine& e = simplexOf(handle);
		auto mesh = e.getMesh(id);
		info[0] = (int) mesh->sizeOfRealNodes(); info[1] = mesh->M; info[2] = (int) e.numberOfBorderConditions(id);
		reals[0] = mesh->getAverageHeight(); reals[1] = mesh->getMinimalHeight(); reals[2] = mesh->getMaximalEigenvalue();
		std::memcpy(basis, e.calculationBasis(), 9 * sizeof(double));
	});
}

/// PDE vectors [n_local][M] of the current layer and, when not null, the U/U1 [3][M][M] and L [3][M] in use
int gcmb_host_simplex_body_pde(void* handle, size_t id, double* pde, double* U, double* U1, double* L) {
	return guarded([&] {
		auto mesh = simplexOf(handle).getMesh(id);
		if (pde) { const auto& v = mesh->pdeAll(); std::memcpy(pde, v.data(), v.size() * sizeof(double)); }
		const GcmMatrices& g = mesh->matrices;
		if (U) { std::memcpy(U, g.U.data(), g.U.size() * sizeof(double)); }
		if (U1) { std::memcpy(U1, g.U1.data(), g.U1.size() * sizeof(double)); }
		if (L) { std::memcpy(L, g.L.data(), g.L.size() * sizeof(double)); }
	});
}

/// border nodes of one condition of a body; returns the count through *n (copies at most `capacity`)
int gcmb_host_simplex_border_nodes(void* handle, size_t id, int condition, int capacity, int* n, int* nodes, double* normals) {
	return guarded([&] {
		std::vector<int> nd;
		std::vector<real> nr;
		simplexOf(handle).borderNodes(id, (size_t) condition, nd, nr);
		*n = (int) nd.size();
		for (int i = 0; i < *n && i < capacity; i++) { nodes[i] = nd[(size_t) i]; std::memcpy(normals + 3 * i, &nr[(size_t) 3 * i], 3 * sizeof(double)); }
	});
}

int gcmb_host_simplex_contact_nodes(void* handle, size_t a, size_t b, int capacity, int* n, int* first, int* second, double* normals) {
	return guarded([&] {
		std::vector<int> f, s;
		std::vector<real> nr;
		simplexOf(handle).contactNodes({a, b}, f, s, nr);
		*n = (int) f.size();
		for (int i = 0; i < *n && i < capacity; i++) {
			first[i] = f[(size_t) i]; second[i] = s[(size_t) i];
			std::memcpy(normals + 3 * i, &nr[(size_t) 3 * i], 3 * sizeof(double));
		}
	});
}

/// the engine's triangulation as an INM mesh file (grid/simplex/mesh_loaders/InmMeshLoader.hpp format)
int gcmb_host_simplex_save_inm(void* handle, const char* file_name) {
	return guarded([&] { simplex::saveInmMesh(simplexOf(handle).getTriangulation(), file_name); });
}

/// InmMeshLoader::readFromFile (reference grid/simplex/mesh_loaders/InmMeshLoader.hpp:96-168): points [nV][3], cells
/// [nC][4] (0-based vertex ids) and the material number of every cell.  sizes_out = {nV, nC}; arrays may be null to query
/// the sizes only
int gcmb_host_inm_read(const char* file_name, double scale, int* sizes_out, double* xyz, int* cell_v, int* cell_material) {
	return guarded([&] {
		const simplex::FlatTriangulation t = simplex::loadInmMesh(file_name, scale);
		sizes_out[0] = t.nV; sizes_out[1] = t.nC;
		if (xyz) { std::memcpy(xyz, t.xyz.data(), t.xyz.size() * sizeof(double)); }
		if (cell_v) { std::memcpy(cell_v, t.cellV.data(), t.cellV.size() * sizeof(int)); }
		if (cell_material) { std::memcpy(cell_material, t.cellGrid.data(), t.cellGrid.size() * sizeof(int)); }
	});
}

int gcmb_host_simplex_errors(void* handle, int* count) {
	return guarded([&] { *count = simplexOf(handle).errorCount(); });
}

/// eigen-systems of every material table of a body: U,U1 [tables][D][M][M], L [tables][D][M]
int gcmb_host_engine_body_matrices(void* handle, size_t id, int* n_tables, double* U, double* U1, double* L) {
	return guarded([&] {
		auto mesh = cubicOf(handle).getMesh(id);
		const auto& ms = mesh->tableMatrices();
		*n_tables = (int) ms.size();
		size_t a = 0, b = 0;
		for (const auto& g : ms) {
			if (U) { std::memcpy(U + a, g.U.data(), g.U.size() * sizeof(double)); }
			if (U1) { std::memcpy(U1 + a, g.U1.data(), g.U1.size() * sizeof(double)); }
			if (L) { std::memcpy(L + b, g.L.data(), g.L.size() * sizeof(double)); }
			a += g.U.size();
			b += g.L.size();
		}
	});
}

/// seismogram recorded so far; returns the number of samples (copies at most `capacity`)
int gcmb_host_engine_seismogram(void* handle, double* times, float* values, int capacity) {
	int n = -1;
	guarded([&] {
		const auto& s = cubicOf(handle).seismogram();
		for (int i = 0; i < (int) s.size() && i < capacity; i++) { times[i] = s[(size_t) i].first; values[i] = s[(size_t) i].second; }
		n = (int) s.size();
	});
	return n;
}

/// Model::constructGcmMatrices for one material.  model: 0 elastic, 1 acoustic.  material_kind:
/// 0 isotropic {rho, lambda, mu}, 1 orthotropic {rho, c11, c12, c13, c22, c23, c33, c44, c55, c66}
int gcmb_host_matrices(int model, int D, int material_kind, const double* p, double* U, double* U1, double* L) {
	return guarded([&] {
		const Models::T m = model == 1 ? Models::T::ACOUSTIC : Models::T::ELASTIC;
		GcmMatrices g;
		if (material_kind == 0) {
			g = constructGcmMatrices(m, D, IsotropicMaterial(p[0], p[1], p[2]));
		} else {
			g = constructGcmMatrices(m, D, OrthotropicMaterial(p[0],
					std::initializer_list<real>{p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9]}));
		}
		std::memcpy(U, g.U.data(), g.U.size() * sizeof(double));
		std::memcpy(U1, g.U1.data(), g.U1.size() * sizeof(double));
		std::memcpy(L, g.L.data(), g.L.size() * sizeof(double));
	});
}

}  // extern "C"
