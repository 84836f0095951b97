// C entry points of the host layer, for bindings (Python ctypes in gcm_b200/capi.py, tests, bench).
#include <cstring>

#include "gcmb_host.hpp"

using namespace gcmb;

namespace {
thread_local std::string hostError;
struct Handle {
	std::shared_ptr<AbstractEngine> engine;
	cubic::EngineBase* base = nullptr;
};
template<typename F> int guarded(F f) {
	try { f(); return GCMB_OK; }
	catch (const Exception& e) { hostError = e.what(); return e.code() == 0 ? GCMB_E_INVALID_OP : e.code(); }
	catch (const std::exception& e) { hostError = e.what(); return GCMB_E_INVALID_OP; }
}
}  // namespace

extern "C" {

const char* gcmb_host_last_error(void) { return hostError.c_str(); }

/// createEngine(parseTaskText(text)); slab_count > 1 decomposes every body along x across processes
int gcmb_host_engine_create(const char* task_text, int device, int slab_rank, int slab_count,
		const void* nccl_id128, void** out) {
	return guarded([&] {
		Task task = parseTaskText(task_text);
		task.device.device = device;
		task.device.slabRank = slab_rank;
		task.device.slabCount = slab_count;
		task.device.ncclUniqueId = nccl_id128;
		auto* h = new Handle;
		h->engine = createEngine(task);
		h->base = dynamic_cast<cubic::EngineBase*>(h->engine.get());
		*out = h;
	});
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
		*time = Clock::Time();
		*tau = Clock::TimeStep();
	});
}

/// D, M, sizes[3] of a body (sizes of THIS process' slab when decomposed)
int gcmb_host_engine_body_info(void* handle, size_t id, int* D, int* M, int* sizes, int* start) {
	return guarded([&] {
		auto mesh = static_cast<Handle*>(handle)->base->getMesh(id);
		*D = mesh->D;
		*M = mesh->M;
		for (int i = 0; i < 3; i++) { sizes[i] = mesh->sizes[(size_t) i]; start[i] = mesh->start[(size_t) i]; }
	});
}

/// real nodes of the current time layer, x slowest, M per node
int gcmb_host_engine_body_pde(void* handle, size_t id, double* out) {
	return guarded([&] {
		auto mesh = static_cast<Handle*>(handle)->base->getMesh(id);
		const auto& v = mesh->pdeRealNodes();
		std::memcpy(out, v.data(), v.size() * sizeof(double));
	});
}

void* gcmb_host_engine_body_handle(void* handle, size_t id) {
	try { return static_cast<Handle*>(handle)->base->getMesh(id)->handle(); }
	catch (...) { return nullptr; }
}

void* gcmb_host_engine_context(void* handle) { return static_cast<Handle*>(handle)->base->context(); }

/// eigen-systems of every material table of a body: U,U1 [tables][D][M][M], L [tables][D][M]
int gcmb_host_engine_body_matrices(void* handle, size_t id, int* n_tables, double* U, double* U1, double* L) {
	return guarded([&] {
		auto mesh = static_cast<Handle*>(handle)->base->getMesh(id);
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
	const auto& s = static_cast<Handle*>(handle)->base->seismogram();
	for (int i = 0; i < (int) s.size() && i < capacity; i++) { times[i] = s[(size_t) i].first; values[i] = s[(size_t) i].second; }
	return (int) s.size();
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
