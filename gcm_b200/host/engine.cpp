// gcm_b200 host layer: engines.  The control flow is the reference's
// (engine/AbstractEngine.cpp:9-46, engine/cubic/Engine.cpp:13-151); every operation on grid data is a
// call into the C ABI (include/gcm_b200.h), i.e. a CUDA kernel on device-resident state.
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "gcmb_host.hpp"

namespace gcmb {

real Clock::time = 0;
real Clock::timeStep = 0;

static void check(int rc) {
	if (rc != GCMB_OK) { throw Exception(rc, gcmb_last_error()); }
}

// ---------------------------------------------------------------------------------------------
// AbstractEngine (reference engine/AbstractEngine.cpp)
// ---------------------------------------------------------------------------------------------
AbstractEngine::AbstractEngine(const Task& task) :
		CourantNumber(task.globalSettings.CourantNumber),
		verboseTimeSteps(task.globalSettings.verboseTimeSteps) {
	Clock::setZero();
}

void AbstractEngine::activateClock() const {
	Clock::time = time;
	Clock::timeStep = timeStep;
}

void AbstractEngine::storeClock() {
	time = Clock::time;
	timeStep = Clock::timeStep;
}

void AbstractEngine::afterConstruction(const Task& task) {
	Clock::timeStep = estimateTimeStep();
	storeClock();
	requiredTime = Clock::TimeStep() * task.globalSettings.numberOfSnaps * task.globalSettings.stepsPerSnap;
	if (task.globalSettings.numberOfSnaps <= 0) { requiredTime = task.globalSettings.requiredTime; }
	if (!(requiredTime > 0)) { throw Exception(GCMB_E_INVALID_ARG, "required time must be positive"); }
}

void AbstractEngine::run() {
	activateClock();
	step = 0;
	writeSnapshots(step);
	// the step count is decided by the same floating-point accumulation as in the reference
	while (Clock::Time() < requiredTime) {
		Clock::timeStep = estimateTimeStep();
		if (verboseTimeSteps) {
			printf("Start step %d. Time = %g. TimeStep = %g\n", step, Clock::Time(), Clock::TimeStep());
		}
		nextTimeStep();
		step++;
		Clock::tickTack();
		storeClock();
		writeSnapshots(step);
	}
	finishRun();
}

void AbstractEngine::advance(int n) {
	activateClock();
	for (int i = 0; i < n; i++) {
		Clock::timeStep = estimateTimeStep();
		nextTimeStep();
		step++;
		Clock::tickTack();
		storeClock();
		writeSnapshots(step);
	}
	finishRun();
}

// ---------------------------------------------------------------------------------------------
// Mesh
// ---------------------------------------------------------------------------------------------
namespace cubic {

Mesh::~Mesh() {
	if (body) { gcmb_cubic_body_destroy(body); }
}

size_t Mesh::sizeOfRealNodes() const {
	size_t n = 1;
	for (int i = 0; i < D; i++) { n *= (size_t) sizes[i]; }
	return n;
}

real Mesh::getMinimalSpatialStep() const {
	real ans = h[0];
	for (int i = 1; i < D; i++) { if (ans > h[i]) { ans = h[i]; } }
	return ans;
}

const std::vector<real>& Mesh::pdeRealNodes() const {
	if (!hostValid) {
		host.resize(sizeOfRealNodes() * (size_t) M);
		if (realBytes == 4) {
			// an fp32 context keeps floats (the reference's `real` with LIBGCM_DOUBLE_PRECISION off)
			std::vector<float> tmp(host.size());
			check(gcmb_cubic_download_state(body, tmp.data(), 0));
			for (size_t i = 0; i < tmp.size(); i++) { host[i] = tmp[i]; }
		} else {
			check(gcmb_cubic_download_state(body, host.data(), 0));
		}
		hostValid = true;
	}
	return host;
}

const real* Mesh::pde(const Iterator& it) const {
	size_t idx = 0;
	for (int i = 0; i < D; i++) {
		if (it[i] < 0 || it[i] >= sizes[i]) { throw Exception(GCMB_E_INVALID_ARG, "Mesh::pde: only real nodes are accessible"); }
		idx = idx * (size_t) sizes[i] + (size_t) it[i];
	}
	return pdeRealNodes().data() + idx * (size_t) M;
}

Real3 Mesh::coords(const Iterator& it) const {
	Real3 ans = {{0, 0, 0}};
	for (int i = 0; i < D; i++) { ans[i] = (start[i] * h[i]) + (it[i] * h[i]); }
	return ans;
}

// ---------------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------------
EngineBase::EngineBase(const Task& task, int dimensionality) :
		AbstractEngine(task), D(dimensionality), taskCopy(task) {
	if (task.globalSettings.dimensionality != D) {
		throw Exception(GCMB_E_INVALID_ARG, "task dimensionality does not match the engine");
	}
	slabRank = task.device.slabRank;
	slabCount = task.device.slabCount;
	if (slabCount > 1) {
		// Every body is cut along x by slab, so a contact NORMAL to x between two bodies would join planes that live on
		// different processes: the local contact boxes would not find it and the facing ghost planes would silently stay
		// empty.  Such tasks are detected on the undecomposed cubes and refused before anything is created.
		for (const auto& a : task.cubicGrid.cubics) {
			for (const auto& b : task.cubicGrid.cubics) {
				if (a.first == b.first || (int) a.second.sizes.size() != D || (int) b.second.sizes.size() != D) { continue; }
				int width[3], axis = 0;
				for (int i = 0; i < D; i++) {
					width[i] = std::min(a.second.start[i] + a.second.sizes[i] - 1, b.second.start[i] + b.second.sizes[i] - 1) -
					           std::max(a.second.start[i], b.second.start[i]);
				}
				for (int i = 1; i < D; i++) { if (width[i] < width[axis]) { axis = i; } }
				if (width[axis] == -1 && axis == 0) {
					throw Exception(GCMB_E_UNSUPPORTED, "slab decomposition along x cannot hold a contact normal to x (bodies " +
							std::to_string(a.first) + " and " + std::to_string(b.first) + "): stack the bodies along y or z");
				}
			}
		}
	}
	check(gcmb_create(task.device.device, task.device.realBytes, &ctx));
	if (task.device.fma) { check(gcmb_set_fma(ctx, 1)); }
	if (slabCount > 1) {
		if (!task.device.ncclUniqueId) { throw Exception(GCMB_E_INVALID_ARG, "slab decomposition needs the NCCL id"); }
		check(gcmb_comm_init(ctx, slabCount, slabRank, task.device.ncclUniqueId));
	}
	createGridsAndContacts(task);
	for (const auto& taskBody : task.bodies) {
		Body& body = getBody(taskBody.first);
		setUpPde(task, body);
		setUpBorders(task, body);
		for (const Odes::T ode : taskBody.second.odes) {
			if (ode != Odes::T::MAXWELL_VISCOSITY) { throw Exception(GCMB_E_UNSUPPORTED, "only the Maxwell viscosity ODE exists"); }
			body.odes.push_back(ode);
		}
	}
	for (const Snapshotters::T s : task.globalSettings.snapshottersId) {
		if (s == Snapshotters::T::SLICESNAP) {
			if (task.detector.quantities.size() != 1) { throw Exception(GCMB_E_INVALID_ARG, "exactly one detector quantity is supported"); }
			Body& b = getBody(task.detector.gridId);
			const int code = quantityCode(b.mesh->modelType, D, task.detector.quantities[0]);
			const auto p = task.detector.area->deviceParams();
			check(gcmb_cubic_detector_set_area(b.mesh->body, code, task.detector.area->deviceKind(), p.data()));
		}
	}
	afterConstruction(task);
}

EngineBase::~EngineBase() {
	try { finishPendingSeismo(); finishPendingSnapshots(); } catch (...) { }
	for (void* p : pinned) { gcmb_host_free_pinned(p); }
	bodies.clear();
	if (ctx) { gcmb_destroy(ctx); }
}

EngineBase::Body& EngineBase::getBody(const GridId id) {
	for (Body& b : bodies) { if (b.mesh->id == id) { return b; } }
	throw Exception(GCMB_E_INVALID_ARG, "There isn't a body with given id");
}
const EngineBase::Body& EngineBase::getBody(const GridId id) const {
	for (const Body& b : bodies) { if (b.mesh->id == id) { return b; } }
	throw Exception(GCMB_E_INVALID_ARG, "There isn't a body with given id");
}

std::shared_ptr<const Mesh> EngineBase::getMesh(const GridId gridId) const {
	return getBody(gridId).mesh;
}

/// reference engine/cubic/Engine.cpp:40-87
void EngineBase::createGridsAndContacts(const Task& task) {
	if (task.bodies.size() != task.cubicGrid.cubics.size()) {
		throw Exception(GCMB_E_INVALID_ARG, "every body needs a cube");
	}
	if ((int) task.cubicGrid.h.size() != D) { throw Exception(GCMB_E_INVALID_ARG, "h must have D entries"); }
	for (const auto& taskBody : task.bodies) {
		const Task::CubicGrid::Cube& cube = task.cubicGrid.cubics.at(taskBody.first);
		if ((int) cube.sizes.size() != D || (int) cube.start.size() != D) {
			throw Exception(GCMB_E_INVALID_ARG, "cube sizes/start must have D entries");
		}
		Body body;
		body.mesh = std::make_shared<Mesh>();
		Mesh& m = *body.mesh;
		m.id = taskBody.first;
		m.D = D;
		m.borderSize = task.cubicGrid.borderSize;
		m.modelType = taskBody.second.modelId;
		m.materialType = taskBody.second.materialId;
		m.M = pdeSize(m.modelType, D);
		m.realBytes = task.device.realBytes;
		for (int i = 0; i < D; i++) {
			m.sizes[i] = cube.sizes[i];
			m.start[i] = cube.start[i];
			m.h[i] = task.cubicGrid.h[i];
		}
		m.globalSizes = m.sizes;
		m.globalStart = m.start;
		if (slabCount > 1) {
			// slab decomposition along x: this process keeps planes [lo, hi) of the body
			const int nx = m.sizes[0];
			const int lo = (int) ((long long) nx * slabRank / slabCount);
			const int hi = (int) ((long long) nx * (slabRank + 1) / slabCount);
			m.start[0] += lo;
			m.sizes[0] = hi - lo;
		}
		check(gcmb_cubic_body_create(ctx, D, m.M, m.sizes.data(), m.start.data(), m.h.data(), m.borderSize, &m.body));
		bodies.push_back(body);
	}

	const int bs = task.cubicGrid.borderSize;
	for (Body& body : bodies) {
		for (const Body& other : bodies) {
			if (other.mesh->id == body.mesh->id) { continue; }
			const Mesh& a = *body.mesh;
			const Mesh& b = *other.mesh;
			// intersection of the two index-space boxes (util/math/AABB.hpp:82-91)
			int lo[3], hi[3], width[3];
			bool valid = true;
			for (int i = 0; i < D; i++) {
				lo[i] = std::max(a.start[i], b.start[i]);
				hi[i] = std::min(a.start[i] + a.sizes[i] - 1, b.start[i] + b.sizes[i] - 1);
				width[i] = hi[i] - lo[i];
				if (width[i] < 0) { valid = false; }
			}
			if (valid) { throw Exception(GCMB_E_BAD_MESH, "Bodies must not intersect"); }
			int axis = 0;
			for (int i = 1; i < D; i++) { if (width[i] < width[axis]) { axis = i; } }
			if (width[axis] != -1) { continue; }  // no contact

			Contact c;
			c.neighborId = b.id;
			c.direction = axis;
			if (a.start[axis] > b.start[axis]) { lo[axis] -= bs; } else { hi[axis] += bs; }
			for (int i = 0; i < 3; i++) { c.boxA[i] = c.boxB[i] = 0; c.extent[i] = 1; }
			for (int i = 0; i < D; i++) {
				c.boxA[i] = lo[i] - a.start[i];
				c.boxB[i] = lo[i] - b.start[i];
				c.extent[i] = hi[i] - lo[i] + 1;
			}
			if (task.contactCondition.defaultCondition != ContactConditions::T::ADHESION) {
				throw Exception(GCMB_E_UNSUPPORTED, "only ADHESION contacts exist for cubic grids");
			}
			if (a.modelType != b.modelType) { throw Exception(GCMB_E_UNSUPPORTED, "contacting bodies must share the model"); }
			body.contacts.push_back(c);
		}
	}
}

/// reference engine/cubic/DefaultMesh.hpp:60-66 -> MaterialsCondition.hpp:23-96, InitialCondition.hpp:23-88
void EngineBase::setUpPde(const Task& task, Body& body) {
	Mesh& m = *body.mesh;
	struct Cond { std::shared_ptr<Area> area; Task::MaterialCondition::Material material; };
	std::vector<Cond> conds;
	switch (task.materialConditions.type) {
		case Task::MaterialCondition::Type::BY_AREAS:
			conds.push_back({std::make_shared<InfiniteArea>(), task.materialConditions.byAreas.defaultMaterial});
			for (const auto& inh : task.materialConditions.byAreas.materials) { conds.push_back({inh.area, inh.material}); }
			break;
		case Task::MaterialCondition::Type::BY_BODIES:
			conds.push_back({std::make_shared<InfiniteArea>(), task.materialConditions.byBodies.bodyMaterialMap.at(m.id)});
			break;
		default:
			throw Exception(GCMB_E_UNSUPPORTED, "Unknown type of material condition");
	}
	if (conds.size() > GCMB_MAX_TABLES) { throw Exception(GCMB_E_UNSUPPORTED, "too many material conditions"); }
	std::vector<real> U, U1, L;
	m.maximalEigenvalue = 0;
	for (const Cond& c : conds) {
		if (!c.material) { throw Exception(GCMB_E_INVALID_ARG, "material is not set"); }
		const bool iso = dynamic_cast<const IsotropicMaterial*>(c.material.get()) != nullptr;
		if (iso != (m.materialType == Materials::T::ISOTROPIC)) {
			throw Exception(GCMB_E_INVALID_ARG, "material type does not match the body");
		}
		GcmMatrices g = constructGcmMatrices(m.modelType, D, *c.material);
		U.insert(U.end(), g.U.begin(), g.U.end());
		U1.insert(U1.end(), g.U1.begin(), g.U1.end());
		L.insert(L.end(), g.L.begin(), g.L.end());
		m.maximalEigenvalue = std::fmax(m.maximalEigenvalue, g.getMaximalEigenvalue());
		m.materials.push_back(c.material);
		m.matrices.push_back(g);
	}
	check(gcmb_cubic_set_materials(m.body, (int) conds.size(), U.data(), U1.data(), L.data(), nullptr));
	for (size_t i = 1; i < conds.size(); i++) {
		const auto p = conds[i].area->deviceParams();
		check(gcmb_cubic_assign_table_in_area(m.body, (int) i, conds[i].area->deviceKind(), p.data()));
	}

	// initial conditions, summed in the reference's order: vectors, waves, quantities
	auto add = [&](const std::shared_ptr<Area>& area, const std::vector<real>& v) {
		const auto p = area->deviceParams();
		check(gcmb_cubic_add_vector_in_area(m.body, v.data(), area->deviceKind(), p.data()));
	};
	for (const auto& v : task.initialCondition.vectors) {
		if ((int) v.list.size() != m.M) { throw Exception(GCMB_E_INVALID_ARG, "initial vector has a wrong size"); }
		add(v.area, v.list);
	}
	for (const auto& w : task.initialCondition.waves) {
		if (w.direction >= D) { throw Exception(GCMB_E_INVALID_ARG, "wave direction out of range"); }
		const GcmMatrices& g = m.matrices.front();  // the first condition's material, like the reference
		const int col = waveColumn(m.modelType, m.materialType, D, w.waveType);
		std::vector<real> tmp((size_t) m.M);
		for (int i = 0; i < m.M; i++) { tmp[(size_t) i] = g.u1(w.direction)[i * m.M + col]; }
		const int code = quantityCode(m.modelType, D, w.quantity);
		real current;
		if (code >= 0) { current = tmp[(size_t) code]; }
		else {
			real trace = 0;
			for (int i = 0; i < D; i++) { trace += tmp[(size_t) sigmaComponent(D, i, i)]; }
			current = -trace / D;
		}
		if (current == 0) { throw Exception(GCMB_E_INVALID_ARG, "wave has zero calibration quantity"); }
		const real scale = w.quantityValue / current;
		for (real& x : tmp) { x *= scale; }
		add(w.area, tmp);
	}
	for (const auto& q : task.initialCondition.quantities) {
		std::vector<real> tmp((size_t) m.M, 0.0);
		const int code = quantityCode(m.modelType, D, q.physicalQuantity);
		if (code >= 0) { tmp[(size_t) code] = q.value; }
		else {
			for (int i = 0; i < D; i++) { tmp[(size_t) sigmaComponent(D, i, i)] = -q.value; }
		}
		add(q.area, tmp);
	}
}

/// reference engine/cubic/BorderConditions.hpp:46-78
void EngineBase::setUpBorders(const Task& task, Body& body) {
	Mesh& m = *body.mesh;
	const auto found = task.cubicBorderConditions.find(m.id);
	if (found == task.cubicBorderConditions.end()) { return; }
	int cond = 0;
	for (const Task::CubicBorderCondition& bc : found->second) {
		if (bc.direction < 0 || bc.direction >= D) { throw Exception(GCMB_E_INVALID_ARG, "border direction out of range"); }
		Border b;
		b.direction = bc.direction;
		std::vector<int> codes;
		for (const auto& q : bc.values) {  // std::map order == enum order
			codes.push_back(quantityCode(m.modelType, D, q.first));
			b.values.push_back(q.second);
		}
		int sides = 3;
		if (slabCount > 1 && bc.direction == 0) {
			sides = (slabRank == 0 ? 1 : 0) | (slabRank == slabCount - 1 ? 2 : 0);
		}
		const auto p = bc.area->deviceParams();
		check(gcmb_cubic_border_set_area(m.body, cond++, bc.direction, sides, bc.area->deviceKind(), p.data(),
				(int) codes.size(), codes.data()));
		body.borders.push_back(b);
	}
}

/// reference engine/cubic/Engine.cpp:92-121
void EngineBase::nextTimeStep() {
	auto borderValues = [&](const Body& body, int direction) {
		std::vector<double> values;
		for (const Border& b : body.borders) {
			if (b.direction != direction) { continue; }
			for (const auto& f : b.values) { values.push_back(f(Clock::Time())); }
		}
		return values;
	};
	for (int stage = 0; stage < D; stage++) {
		for (Body& body : bodies) {
			if (body.borderFilledByStage) {  // the previous stage's kernel has written these ghost nodes already
				body.borderFilledByStage = false;
				continue;
			}
			// (for the faces across the last direction the library defers the fill to that direction's stage kernel, which
			// mirrors the ghost nodes in its shared-memory rows -- unless a contact copy below touches the body first)
			const std::vector<double> values = borderValues(body, stage);
			if (!values.empty() || !body.borders.empty()) {
				check(gcmb_cubic_border_apply(body.mesh->body, stage, (int) values.size(), values.data()));
			}
		}
		for (Body& body : bodies) {
			for (const Contact& c : body.contacts) {
				if (c.direction != stage) { continue; }
				// internal axes of the C ABI are the reference axes: boxes are passed with D entries
				check(gcmb_cubic_contact_apply(body.mesh->body, getBody(c.neighborId).mesh->body,
						c.boxA, c.boxB, c.extent));
			}
		}
		if (stage == 0 && slabCount > 1) {
			// one exchange for all bodies; the interior of every body's x stage overlaps it
			std::vector<gcmb_body*> all;
			for (Body& body : bodies) { all.push_back(body.mesh->body); }
			check(gcmb_halo_exchange_bodies(all.data(), (int) all.size()));
		}
		for (Body& body : bodies) {
			if (stage + 1 == D - 1 && !body.borders.empty()) {
				// the stage before the last one can write the ghost nodes of the last direction's faces with the rows
				// it produces (the border values are those of the same time, BorderConditions.hpp:81-95)
				const std::vector<double> values = borderValues(body, D - 1);
				int fused = 0;
				check(gcmb_cubic_stage_fill_next_border(body.mesh->body, stage, Clock::TimeStep(), D - 1,
						(int) values.size(), values.data(), &fused));
				body.borderFilledByStage = fused != 0;
			} else {
				check(gcmb_cubic_stage(body.mesh->body, stage, Clock::TimeStep()));
			}
		}
	}
	for (Body& body : bodies) {
		for (size_t o = 0; o < body.odes.size(); o++) {
			std::vector<double> decay;
			for (const auto& mat : body.mesh->materials) {
				const auto* iso = dynamic_cast<const IsotropicMaterial*>(mat.get());
				const auto* ort = dynamic_cast<const OrthotropicMaterial*>(mat.get());
				const real tau0 = iso ? iso->tau0 : ort->tau0;
				decay.push_back(std::exp(-Clock::TimeStep() / tau0));
			}
			check(gcmb_cubic_ode_maxwell(body.mesh->body, decay.data()));
		}
		body.mesh->invalidateHostCopy();
	}
}

/// reference engine/cubic/Engine.cpp:126-140
real EngineBase::estimateTimeStep() {
	real maxEigenvalue = 0;
	for (const Body& body : bodies) {
		for (int i = 0; i < D; i++) {
			if (!(body.mesh->h[i] == bodies.front().mesh->h[i])) { throw Exception(GCMB_E_BAD_MESH, "bodies must share h"); }
		}
		const real e = body.mesh->getMaximalEigenvalue();
		if (e > maxEigenvalue) { maxEigenvalue = e; }
	}
	return CourantNumber * bodies.front().mesh->getMinimalSpatialStep() / maxEigenvalue;
}

static std::string padded(int v, int digits) {
	char buf[32];
	snprintf(buf, sizeof buf, "%0*d", digits, v);
	return buf;
}

/// reference util/snapshot/SliceSnapshotter.hpp:36-82 (file names: Snapshotter.hpp:53-68).  The taps are read back
/// asynchronously: the detector reduction and the z-axis line of this step are enqueued behind the step's kernels and the
/// time loop goes on; the values are collected, and the step's files written, at the next snapshot or at the end of
/// run() / advance() -- the per-step host synchronisation of a synchronous read-back cost 1 % of a 1024^3 step.
void EngineBase::sliceSnapshot(const int step_) {
	finishPendingSeismo();
	const int last = D - 1;
	pendingSeismoStep = step_;
	pendingSeismoTime = Clock::Time();
	for (Body& body : bodies) {
		const Mesh& m = *body.mesh;
		const bool isDetectorBody = m.id == taskCopy.detector.gridId;
		// the line runs through the centre of the WHOLE body (SliceSnapshotter.hpp:44-58: sizes / 2): of a decomposed
		// body only the slab that holds that node has it
		int lineNode[3] = {0, 0, 0};
		bool holdsLine = true;
		for (int i = 0; i < last; i++) {
			lineNode[i] = m.globalStart[(size_t) i] + m.globalSizes[(size_t) i] / 2 - m.start[(size_t) i];
			if (lineNode[i] < 0 || lineNode[i] >= m.sizes[(size_t) i]) { holdsLine = false; }
		}
		// detector sums of a decomposed body are all-reduced on the stream
		check(gcmb_cubic_seismo_begin(m.body, isDetectorBody ? 1 : 0, last /* velocity along the last axis */, holdsLine ? lineNode : nullptr));
	}
	seismoPending = true;
}

void EngineBase::finishPendingSeismo() {
	if (!seismoPending) { return; }
	seismoPending = false;
	const Task& task = taskCopy;
	const int step_ = pendingSeismoStep;
	// the reference insists on an odd number of processes and lets the middle one write; with slabs the
	// detector value is reduced over all of them and the slab that holds the centre line writes, for any count
	std::string dir = "snapshots";
	if (!task.globalSettings.outputDirectory.empty()) { dir += "/" + task.globalSettings.outputDirectory; }
	if (!seismoDirsMade) {
		mkdir("snapshots", 0777);
		mkdir(dir.c_str(), 0777);
		mkdir((dir + "/zaxis").c_str(), 0777);
		mkdir((dir + "/detector").c_str(), 0777);
		seismoDirsMade = true;
	}
	for (Body& body : bodies) {
		const Mesh& m = *body.mesh;
		const int last = D - 1;
		const bool isDetectorBody = m.id == task.detector.gridId;
		std::vector<double> line((size_t) m.sizes[last]);
		double sum = 0;
		long long count = 0;
		int lineNode[3] = {0, 0, 0};
		bool holdsLine = true;
		for (int i = 0; i < last; i++) {
			lineNode[i] = m.globalStart[(size_t) i] + m.globalSizes[(size_t) i] / 2 - m.start[(size_t) i];
			if (lineNode[i] < 0 || lineNode[i] >= m.sizes[(size_t) i]) { holdsLine = false; }
		}
		check(gcmb_cubic_seismo_end(m.body, isDetectorBody ? &sum : nullptr, isDetectorBody ? &count : nullptr,
				holdsLine ? line.data() : nullptr, holdsLine ? m.sizes[last] : 0));
		const bool writer = holdsLine;  // every slab records the seismogram, the one with the centre line writes the files
		const std::string name = "mesh" + std::to_string(m.id) + "core" + padded(slabRank, 2) + "snap" + padded(step_, 4) + ".txt";
		if (writer) {
			std::ofstream f(dir + "/zaxis/" + name);
			Mesh::Iterator it = {{lineNode[0], lineNode[1], lineNode[2]}};
			for (int k = 0; k < m.sizes[last]; k++) {
				it[(size_t) last] = k;
				f << m.coords(it)[(size_t) last] << "\t" << line[(size_t) k] << "\t\n";   // (one write per file, not per line)
			}
		}
		if (!isDetectorBody) { continue; }
		if (count < 1) { throw Exception(GCMB_E_INVALID_ARG, "the detector area holds no node"); }
		const real value = sum / (real) count;
		seismo.push_back({pendingSeismoTime, (float) value});
		if (!writer) { continue; }
		std::ofstream f(dir + "/detector/" + name);
		for (const auto& s : seismo) { f << s.first << "\t" << (real) s.second << "\t\n"; }
	}
}

void EngineBase::writeSnapshots(const int step_) {
	const Task& task = taskCopy;
	if (step_ % task.globalSettings.stepsPerSnap != 0) { return; }
	for (const Snapshotters::T s : task.globalSettings.snapshottersId) {
		if (s == Snapshotters::T::SLICESNAP) { sliceSnapshot(step_); }
		if (s == Snapshotters::T::VTK) { vtkSnapshot(step_); }
	}
}

/// VtkSnapshotter::snapshotImpl (util/snapshot/VtkSnapshotter.hpp:20-61) for every body of this process.  The state is read
/// back ASYNCHRONOUSLY: a gather + device-to-host copy is enqueued on a side stream into page-locked memory and the time
/// loop goes on; the file is written when the copy has landed -- at the next snapshot or at the end of run()/advance().
/// Bodies whose state does not fit the device-side staging (ASYNC_LIMIT) are read back in chunks, synchronously.
void EngineBase::vtkSnapshot(const int step_) {
	finishPendingSnapshots();
	static const size_t ASYNC_LIMIT = (size_t) 1 << 30;
	for (size_t i = 0; i < bodies.size(); i++) {
		const Mesh& m = *bodies[i].mesh;
		const size_t bytes = m.sizeOfRealNodes() * (size_t) m.M * (size_t) m.realBytes;
		PendingSnapshot p;
		p.step = step_;
		p.body = i;
		if (bytes <= ASYNC_LIMIT) {
			if (pinned.size() <= i) { pinned.resize(bodies.size(), nullptr); }
			if (!pinned[i]) { check(gcmb_host_alloc_pinned(bytes, &pinned[i])); }
			const int lo[3] = {0, 0, 0};
			check(gcmb_cubic_download_box_begin(m.body, lo, m.sizes.data(), pinned[i]));
			p.async = true;
		} else {
			p.values = m.pdeRealNodes();  // chunked, synchronous
			p.async = false;
		}
		pending.push_back(std::move(p));
	}
}

void EngineBase::finishPendingSnapshots() {
	const Task& task = taskCopy;
	for (PendingSnapshot& p : pending) {
		const Mesh& m = *bodies[p.body].mesh;
		const size_t total = m.sizeOfRealNodes();
		if (p.async) {
			check(gcmb_cubic_download_box_end(m.body));
			p.values.resize(total * (size_t) m.M);
			if (m.realBytes == 4) {
				const float* src = static_cast<const float*>(pinned[p.body]);
				for (size_t i = 0; i < p.values.size(); i++) { p.values[i] = src[i]; }
			} else {
				std::memcpy(p.values.data(), pinned[p.body], p.values.size() * sizeof(real));
			}
		}
		const int n[3] = {m.sizes[0], m.sizes[1], m.sizes[2]};   // unused axes have size 1
		std::vector<uint8_t> tables(total);
		check(gcmb_cubic_download_tables(m.body, tables.data()));
		// VTK point order: x fastest (linal::SlowZFastX); ours: x slowest
		std::vector<size_t> order(total);
		std::vector<float> points(3 * total), material(total);
		size_t i = 0;
		for (int z = 0; z < n[2]; z++) for (int y = 0; y < n[1]; y++) for (int x = 0; x < n[0]; x++, i++) {
			const Mesh::Iterator it = {{x, y, z}};
			order[i] = ((size_t) x * (size_t) n[1] + (size_t) y) * (size_t) n[2] + (size_t) z;
			const Real3 c = m.coords(it);
			for (size_t d = 0; d < 3; d++) { points[3 * i + d] = (float) c[d]; }
			material[i] = (float) m.materials[tables[order[i]]]->materialNumber;
		}
		const auto fields = vtk::snapshotFields(m.modelType, D, m.M, p.values, order, task.vtkSnapshotter.quantitiesToSnap, material);
		vtk::writeStructuredGrid(vtk::snapshotFileName(task.globalSettings.outputDirectory, "vtk", m.id, slabRank, p.step, "vts"),
				n, points, fields);
	}
	pending.clear();
}

void EngineBase::finishRun() {
	finishPendingSeismo();
	finishPendingSnapshots();
}

}  // namespace cubic

std::shared_ptr<AbstractEngine> createEngine(const Task& task) {
	if (task.globalSettings.gridId == Grids::T::SIMPLEX) { return std::make_shared<simplex::Engine>(task); }
	switch (task.globalSettings.dimensionality) {
		case 1: return std::make_shared<cubic::Engine<1>>(task);
		case 2: return std::make_shared<cubic::Engine<2>>(task);
		case 3: return std::make_shared<cubic::Engine<3>>(task);
		default: throw Exception(GCMB_E_INVALID_ARG, "Invalid space dimensionality");
	}
}

}  // namespace gcmb
