// simplex::Engine — the reference's engine/simplex/Engine.{hpp,cpp} for 3-D bodies on the flat triangulation,
// driving the gcmb_simplex_* device calls.  Supported: both GCM types (Riemann invariants -- the reference's
// default -- and PDE vectors), GLOBAL_BASIS borders and contacts, PRODUCT splitting, isotropic elastic/acoustic
// bodies, fixed or per-step random calculation basis, PRODUCT and SUMM splitting by directions, borders and contacts
// in the GLOBAL_BASIS or in the LOCAL_BASIS of every border vertex.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>

#include "gcmb_host.hpp"

namespace gcmb {

namespace {
void check(int rc) {
	if (rc != GCMB_OK) { throw Exception(rc, gcmb_last_error()); }
}
}  // namespace

namespace simplex {

// ---------------------------------------------------------------------------------------------
// Mesh
// ---------------------------------------------------------------------------------------------
Mesh::~Mesh() {
	if (body) { gcmb_simplex_body_destroy(body); }
}

Real3 Mesh::coords(const Iterator it) const {
	const double* p = &triangulation->xyz[(size_t) 3 * (size_t) globalOf.at((size_t) it)];
	return {{p[0], p[1], p[2]}};
}

const std::vector<real>& Mesh::pdeAll() const {
	if (!hostValid) {
		host.resize(globalOf.size() * (size_t) M);
		check(gcmb_simplex_download_state(body, host.data()));
		hostValid = true;
	}
	return host;
}

const real* Mesh::pde(const Iterator it) const {
	return pdeAll().data() + (size_t) it * (size_t) M;
}

// ---------------------------------------------------------------------------------------------
// geometry of the set-up (host, once): linal/geometry.hpp:238-284, util/math/Histogram.hpp:14-58
// ---------------------------------------------------------------------------------------------
namespace {

Real3 sub(const double* a, const double* b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
Real3 cross(const Real3& a, const Real3& b) {
	return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}
real length(const Real3& a) { real r = a[0] * a[0]; r += a[1] * a[1]; r += a[2] * a[2]; return std::sqrt(r); }
real area(const double* a, const double* b, const double* c) { return length(cross(sub(b, a), sub(c, a))) / 2; }
real volume(const double* a, const double* b, const double* c, const double* d) {
	const Real3 ba = sub(b, a), ca = sub(c, a), da = sub(d, a);
	const real det = ba[0] * (ca[1] * da[2] - ca[2] * da[1]) - ba[1] * (ca[0] * da[2] - ca[2] * da[0]) +
	                 ba[2] * (ca[0] * da[1] - ca[1] * da[0]);
	return std::fabs(det / 6);
}
real minimalHeight(const double* a, const double* b, const double* c, const double* d) {
	const real V = volume(a, b, c, d);
	const real A = area(b, c, d), B = area(c, d, a), C = area(d, a, b), D = area(a, b, c);
	return 3 * V / std::fmax(A, std::fmax(B, std::fmax(C, D)));
}

/// Histogram(begin, end, 100).mean() and .min() — what SimplexGrid::collectCellHeightsStatistics keeps
/// (grid/simplex/SimplexGrid.cpp:266-274)
void heightStatistics(const std::vector<real>& h, real& mean, real& minimum) {
	const size_t binsNumber = 100;
	real lo = h.front(), hi = h.front();
	for (const real x : h) { if (x < lo) { lo = x; } if (hi < x) { hi = x; } }
	std::vector<size_t> bins;
	if (hi == lo) {
		bins.assign(binsNumber, 0);
		bins[0] = h.size();
	} else {
		const real binSize = (hi - lo) / real(binsNumber);
		bins.assign(binsNumber + 1, 0);
		for (const real x : h) { ++bins[(size_t) ((x - lo) / binSize)]; }
		bins[binsNumber - 1] += bins.back();
		bins.pop_back();
	}
	const real size = (hi - lo) / (real) bins.size();
	real weighted = 0, count = 0;
	for (size_t i = 0; i < bins.size(); i++) {
		const real center = lo + (real(i) + 0.5) * size;
		weighted = weighted + (real) bins[i] * center;
		count = count + (real) bins[i];
	}
	mean = weighted / count;
	minimum = lo;
}

int outerNumber(Models::T model) { return model == Models::T::ACOUSTIC ? 1 : 3; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------------
Engine::Engine(const Task& task) : AbstractEngine(task) {
	if (task.globalSettings.dimensionality != 3) {
		throw Exception(GCMB_E_UNSUPPORTED, "the simplex engine of this build is three-dimensional");
	}
	localBasis = task.simplexGrid.borderCalcMode == BorderCalcMode::LOCAL_BASIS;
	summSplitting = task.globalSettings.splittingType == SplittingType::SUMM;
	if (task.simplexGrid.movable) { throw Exception(GCMB_E_UNSUPPORTED, "movable grids are not built"); }
	vtkSettings = task.vtkSnapshotter;
	snapshotters = task.globalSettings.snapshottersId;
	outputDirectory = task.globalSettings.outputDirectory;
	stepsPerSnap = task.globalSettings.stepsPerSnap;
	check(gcmb_create(task.device.device, 8, &ctx));
	try {
		createTriangulation(task);
		// initializeCalculationBasis (engine/simplex/Engine.hpp:194-206)
		createNewRandomAtEachTimeStep = task.calculationBasis.empty();
		randomState = 0x9E3779B97F4A7C15ull ^ task.randomBasisSeed;
		if (createNewRandomAtEachTimeStep) { changeCalculationBasis(); }
		else {
			if (task.calculationBasis.size() != 9) { throw Exception(GCMB_E_INVALID_ARG, "calculationBasis must have 9 entries"); }
			std::copy(task.calculationBasis.begin(), task.calculationBasis.end(), basis);
		}
		createMeshes(task);
		createContacts(task);
		for (int v = 0; v < triangulation.nV; v++) { addBorderOrContact(v); }
		for (Body& body : bodies) {
			std::vector<int> types, nodes, conds;
			std::vector<real> normals;
			for (size_t c = 0; c < body.borders.size(); c++) {
				const Border& b = body.borders[c];
				types.push_back(b.type == BorderConditions::T::FIXED_FORCE ? 0 : 1);
				nodes.insert(nodes.end(), b.nodes.begin(), b.nodes.end());
				normals.insert(normals.end(), b.normals.begin(), b.normals.end());
				conds.insert(conds.end(), b.nodes.size(), (int) c);
			}
			check(gcmb_simplex_border_set(body.mesh->body, (int) types.size(), types.data(), (int) nodes.size(),
					nodes.data(), normals.data(), conds.data()));
		}
		for (auto& contact : contacts) {
			Contact& c = contact.second;
			check(gcmb_simplex_contact_create(getBody(contact.first.first).mesh->body, getBody(contact.first.second).mesh->body,
					(int) c.first.size(), c.first.data(), c.second.data(), c.normals.data(), &c.handle));
		}
		applyPlainBorderContactCorrection(Clock::Time());
		afterConstruction(task);
	} catch (...) {
		for (auto& contact : contacts) { gcmb_simplex_contact_destroy(contact.second.handle); }
		contacts.clear();
		bodies.clear();
		gcmb_destroy(ctx);
		throw;
	}
}

Engine::~Engine() {
	for (auto& contact : contacts) { gcmb_simplex_contact_destroy(contact.second.handle); }
	bodies.clear();
	if (ctx) { gcmb_destroy(ctx); }
}

Engine::Body& Engine::getBody(const GridId id) {
	for (Body& b : bodies) { if (b.mesh->id == id) { return b; } }
	throw Exception(GCMB_E_INVALID_ARG, "There isn't a body with given id");
}
const Engine::Body& Engine::getBody(const GridId id) const {
	for (const Body& b : bodies) { if (b.mesh->id == id) { return b; } }
	throw Exception(GCMB_E_INVALID_ARG, "There isn't a body with given id");
}

/// the triangulation of the whole calculation space with a body id in every cell (the reference's
/// CgalTriangulation built by the mesher, grid/simplex/cgal/Cgal3DTriangulation.hpp:56-76)
void Engine::createTriangulation(const Task& task) {
	const Task::SimplexGrid& g = task.simplexGrid;
	if (g.mesher == Task::SimplexGrid::Mesher::INM_MESHER) {
		// the cells of the file carry the body ids (InmMeshLoader.hpp:60-83); ids without a Task::Body are an error
		triangulation = loadInmMesh(g.fileName, g.scale);
		for (const int id : triangulation.cellGrid) {
			if (id != EmptySpaceFlag && !task.bodies.count((size_t) id)) {
				throw Exception(GCMB_E_INVALID_ARG, "the mesh file has cells of body " + std::to_string(id) + " which the task does not describe");
			}
		}
		triangulation.cleanBodyIds();
		return;
	}
	if (g.mesher != Task::SimplexGrid::Mesher::BOX_MESHER) {
		throw Exception(GCMB_E_UNSUPPORTED, "CGAL meshing is not part of this build: use the box mesher or an INM mesh file");
	}
	if (g.boxCubes[0] < 1 || g.boxCubes[1] < 1 || g.boxCubes[2] < 1 || !(g.spatialStep > 0)) {
		throw Exception(GCMB_E_INVALID_ARG, "box mesher needs positive sizes and spatial step");
	}
	if (task.bodies.empty()) { throw Exception(GCMB_E_INVALID_ARG, "no bodies"); }
	const int firstId = (int) task.bodies.begin()->first;
	triangulation = makeBoxMesh(g.boxCubes[0], g.boxCubes[1], g.boxCubes[2], g.boxOrigin, g.spatialStep, g.jitter, g.seed,
			nullptr, nullptr, firstId);
	if (g.bodies.empty() && task.bodies.size() != 1) {
		throw Exception(GCMB_E_INVALID_ARG, "several bodies need regions in simplexGrid.bodies");
	}
	for (int c = 0; c < triangulation.nC; c++) {
		Real3 center = {{0, 0, 0}};
		for (int m = 0; m < 4; m++) for (size_t a = 0; a < 3; a++) {
			center[a] += triangulation.xyz[(size_t) 3 * (size_t) triangulation.cellV[(size_t) 4 * c + m] + a] / 4;
		}
		int id = g.bodies.empty() ? firstId : EmptySpaceFlag;
		for (const auto& region : g.bodies) {
			if (!task.bodies.count(region.id)) { throw Exception(GCMB_E_INVALID_ARG, "a region names an unknown body"); }
			if (region.area->contains(center)) { id = (int) region.id; }
		}
		for (const auto& cavity : g.cavities) { if (cavity->contains(center)) { id = EmptySpaceFlag; } }
		triangulation.cellGrid[(size_t) c] = id;
	}
	triangulation.cleanBodyIds();   // CgalTriangulation's constructor (grid/simplex/cgal/CgalTriangulation.cpp:8-39)
}

/// engine/simplex/Engine.cpp:52-91 + DefaultMesh::setUpPde (DefaultMesh.hpp:63-70,227-267)
void Engine::createMeshes(const Task& task) {
	for (const auto& taskBody : task.bodies) {
		if (taskBody.second.materialId != Materials::T::ISOTROPIC) { throw Exception(GCMB_E_UNSUPPORTED, "Unsupported material type"); }
		if (taskBody.second.modelId != Models::T::ELASTIC && taskBody.second.modelId != Models::T::ACOUSTIC) {
			throw Exception(GCMB_E_UNSUPPORTED, "Unknown model type");
		}
		Body body;
		for (const Odes::T ode : taskBody.second.odes) {
			if (ode != Odes::T::MAXWELL_VISCOSITY) { throw Exception(GCMB_E_UNSUPPORTED, "only the Maxwell viscosity ODE exists"); }
			body.odes.push_back(ode);
		}
		body.mesh = std::make_shared<Mesh>();
		Mesh& m = *body.mesh;
		m.id = taskBody.first;
		m.modelType = taskBody.second.modelId;
		m.materialType = taskBody.second.materialId;
		m.M = pdeSize(m.modelType, 3);
		m.triangulation = &triangulation;
		const FlatTriangulation& t = triangulation;
		check(gcmb_simplex_body_create(ctx, m.modelType == Models::T::ACOUSTIC ? 1 : 0, t.nV, t.nC, t.xyz.data(), t.cellV.data(),
				t.cellN.data(), t.cellGrid.data(), t.incOff.data(), t.incCell.data(), (int) m.id, &m.body));
		check(gcmb_simplex_set_gcm_type(m.body, task.globalSettings.gcmType == GcmType::ADVECT_PDE_VECTORS ? 1 : 0));
		check(gcmb_simplex_set_splitting(m.body, summSplitting ? 1 : 0));
		int nLocal = 0, M = 0, nBorder = 0;
		check(gcmb_simplex_info(m.body, &nLocal, &M, &nBorder));
		m.globalOf.resize((size_t) nLocal);
		m.state.resize((size_t) nLocal);
		m.borderNormals.resize((size_t) nLocal * 3);
		m.commonNormals.resize((size_t) nLocal * 3);
		check(gcmb_simplex_vertices(m.body, m.globalOf.data(), m.state.data(), m.borderNormals.data(), m.commonNormals.data()));
		m.localOf.assign((size_t) t.nV, -1);
		for (int l = 0; l < nLocal; l++) { m.localOf[(size_t) m.globalOf[(size_t) l]] = l; }
		// cell heights (SimplexGrid.cpp:183-194,266-274)
		std::vector<real> heights;
		for (int c = 0; c < t.nC; c++) {
			if (t.cellGrid[(size_t) c] != (int) m.id) { continue; }
			const int* v = &t.cellV[(size_t) 4 * c];
			heights.push_back(minimalHeight(&t.xyz[(size_t) 3 * v[0]], &t.xyz[(size_t) 3 * v[1]], &t.xyz[(size_t) 3 * v[2]], &t.xyz[(size_t) 3 * v[3]]));
		}
		heightStatistics(heights, m.averageHeight, m.minimalHeight);

		// applyMaterialsCondition (DefaultMesh.hpp:227-243): BY_BODIES only
		if (task.materialConditions.type != Task::MaterialCondition::Type::BY_BODIES) {
			throw Exception(GCMB_E_UNSUPPORTED, "simplex grids take materials BY_BODIES");
		}
		m.material = task.materialConditions.byBodies.bodyMaterialMap.at(m.id);
		if (!dynamic_cast<const IsotropicMaterial*>(m.material.get())) { throw Exception(GCMB_E_INVALID_ARG, "material type does not match the body"); }
		setMaterial(body);
		applyInitialConditions(task, body);

		for (const Task::BorderCondition& condition : task.borderConditions) {
			if ((int) condition.values.size() != outerNumber(m.modelType)) {
				throw Exception(GCMB_E_INVALID_ARG, "a border condition needs as many values as the model has outer waves");
			}
			Border border;
			border.correctionArea = condition.area;
			border.useForMulticontactNodes = condition.useForMulticontactNodes;
			border.type = condition.type;
			border.values = condition.values;
			body.borders.push_back(border);
		}
		bodies.push_back(body);
	}
}

void Engine::setMaterial(Body& body) {
	Mesh& m = *body.mesh;
	const auto& iso = dynamic_cast<const IsotropicMaterial&>(*m.material);
	m.matrices = constructGcmMatrices(m.modelType, 3, iso, basis);
	m.maximalEigenvalue = m.matrices.getMaximalEigenvalue();
	check(gcmb_simplex_set_material(m.body, m.matrices.U.data(), m.matrices.U1.data(), m.matrices.L.data(), basis));
	if (!localBasis) { return; }
	// DefaultMesh::applyMaterialsCondition, LOCAL_BASIS (engine/simplex/DefaultMesh.hpp:245-266): border (and
	// multicontact) vertices in a basis whose first axis is their border normal, contact vertices in one along the
	// normal towards the body they touch; linal::createLocalBasisWithX (linal/basis.hpp:104-110)
	const FlatTriangulation& t = triangulation;
	std::map<int, std::vector<real>> contactNormals;   // towards a neighbouring body, for every local vertex
	std::vector<int> nodes;
	std::vector<real> U, U1, L, bases;
	for (size_t l = 0; l < m.globalOf.size(); l++) {
		if (m.state[l] == 0) { continue; }
		real n[3];
		if (m.state[l] == 2) {
			const int g = m.globalOf[l];
			int other = EmptySpaceFlag;
			for (int i = t.incOff[(size_t) g]; i < t.incOff[(size_t) g + 1]; i++) {
				const int id = t.cellGrid[(size_t) t.incCell[(size_t) i]];
				if (id != (int) m.id) { other = id; }
			}
			std::vector<real>& all = contactNormals[other];
			if (all.empty()) {
				all.resize(m.globalOf.size() * 3);
				check(gcmb_simplex_contact_normals(m.body, other, all.data()));
			}
			std::copy(&all[3 * l], &all[3 * l] + 3, n);
		} else {
			std::copy(&m.borderNormals[3 * l], &m.borderNormals[3 * l] + 3, n);
		}
		if (n[0] == 0 && n[1] == 0 && n[2] == 0) {
			throw Exception(GCMB_E_UNSUPPORTED, "LOCAL_BASIS: a multicontact vertex without a border normal has no local basis");
		}
		// tau_1 = -perpendicularClockwise(n) (linal/geometry.hpp:46-52), tau_2 = n x tau_1
		real p[3] = {n[1], -n[0], 0};
		if (n[0] == 0 && n[1] == 0) { p[0] = n[2]; p[1] = 0; p[2] = 0; }
		const real lenN = length({{n[0], n[1], n[2]}}), lenP = length({{p[0], p[1], p[2]}});
		real t1[3];
		for (int i = 0; i < 3; i++) { t1[i] = -(p[i] * lenN / lenP); }
		const Real3 t2 = cross({{n[0], n[1], n[2]}}, {{t1[0], t1[1], t1[2]}});
		const real local[9] = {n[0], t1[0], t2[0], n[1], t1[1], t2[1], n[2], t1[2], t2[2]};
		const GcmMatrices g = constructGcmMatrices(m.modelType, 3, iso, local);
		if (g.getMaximalEigenvalue() > m.maximalEigenvalue) { m.maximalEigenvalue = g.getMaximalEigenvalue(); }
		nodes.push_back((int) l);
		U.insert(U.end(), g.U.begin(), g.U.end());
		U1.insert(U1.end(), g.U1.begin(), g.U1.end());
		L.insert(L.end(), g.L.begin(), g.L.end());
		bases.insert(bases.end(), local, local + 9);
	}
	check(gcmb_simplex_set_local_bases(m.body, (int) nodes.size(), nodes.data(), U.data(), U1.data(), L.data(), bases.data()));
}

/// the reference's applyInitialConditions on vertex coordinates (DefaultMesh.hpp:63-70 ->
/// rheology/models/Model.hpp InitialCondition: vectors, then waves, then quantities, summed)
void Engine::applyInitialConditions(const Task& task, Body& body) {
	Mesh& m = *body.mesh;
	const int M = m.M;
	std::vector<std::pair<std::shared_ptr<Area>, std::vector<real>>> terms;
	for (const auto& v : task.initialCondition.vectors) {
		if ((int) v.list.size() != M) { throw Exception(GCMB_E_INVALID_ARG, "initial vector has a wrong size"); }
		terms.push_back({v.area, v.list});
	}
	for (const auto& w : task.initialCondition.waves) {
		if (w.direction >= 3) { throw Exception(GCMB_E_INVALID_ARG, "wave direction out of range"); }
		// the wave is built in the GLOBAL basis like the reference's InitialCondition does
		const GcmMatrices g = constructGcmMatrices(m.modelType, 3, *m.material);
		const int col = waveColumn(m.modelType, m.materialType, 3, w.waveType);
		std::vector<real> tmp((size_t) M);
		for (int i = 0; i < M; i++) { tmp[(size_t) i] = g.u1(w.direction)[i * M + col]; }
		const int code = quantityCode(m.modelType, 3, w.quantity);
		real current;
		if (code >= 0) { current = tmp[(size_t) code]; }
		else {
			real trace = 0;
			for (int i = 0; i < 3; i++) { trace += tmp[(size_t) sigmaComponent(3, i, i)]; }
			current = -trace / 3;
		}
		if (current == 0) { throw Exception(GCMB_E_INVALID_ARG, "wave has zero calibration quantity"); }
		const real scale = w.quantityValue / current;
		for (real& x : tmp) { x *= scale; }
		terms.push_back({w.area, tmp});
	}
	for (const auto& q : task.initialCondition.quantities) {
		std::vector<real> tmp((size_t) M, 0.0);
		const int code = quantityCode(m.modelType, 3, q.physicalQuantity);
		if (code >= 0) { tmp[(size_t) code] = q.value; }
		else { for (int i = 0; i < 3; i++) { tmp[(size_t) sigmaComponent(3, i, i)] = -q.value; } }
		terms.push_back({q.area, tmp});
	}
	std::vector<real> pde(m.globalOf.size() * (size_t) M, 0.0);
	for (size_t l = 0; l < m.globalOf.size(); l++) {
		const Real3 x = m.coords((int) l);
		for (const auto& term : terms) {
			if (!term.first->contains(x)) { continue; }
			for (int i = 0; i < M; i++) { pde[l * (size_t) M + (size_t) i] += term.second[(size_t) i]; }
		}
	}
	check(gcmb_simplex_upload_state(m.body, pde.data()));
	m.invalidateHostCopy();
}

/// engine/simplex/Engine.cpp:218-246; the corrector factory's table (ContactCorrector.hpp:484-560)
void Engine::createContacts(const Task& task) {
	std::set<GridId> ids;
	for (const Body& body : bodies) { ids.insert(body.mesh->id); }
	for (auto i = ids.begin(); i != ids.end(); ++i) {
		for (auto j = std::next(i); j != ids.end(); ++j) {
			const GridsPair pair = {*i, *j};
			Contact contact;
			contact.condition = task.contactCondition.defaultCondition;
			const auto it = task.contactCondition.gridToGridConditions.find(pair);
			if (it != task.contactCondition.gridToGridConditions.end()) { contact.condition = it->second; }
			const Models::T a = getBody(pair.first).mesh->modelType, b = getBody(pair.second).mesh->modelType;
			const bool ok = (contact.condition == ContactConditions::T::ADHESION && a == Models::T::ELASTIC && b == Models::T::ELASTIC) ||
			                (contact.condition == ContactConditions::T::SLIDE && a == Models::T::ACOUSTIC && b == Models::T::ACOUSTIC);
			if (!ok) {
				throw Exception(GCMB_E_UNSUPPORTED, "Incompatible or unsupported contact conditions, models and materials combination");
			}
			contacts.insert({pair, contact});
		}
	}
}

/// engine/simplex/Engine.cpp:250-309
void Engine::addBorderOrContact(const int vertex) {
	const FlatTriangulation& t = triangulation;
	std::set<int> incidentGrids;
	for (int i = t.incOff[(size_t) vertex]; i < t.incOff[(size_t) vertex + 1]; i++) {
		const int c = t.incCell[(size_t) i];
		incidentGrids.insert(t.cellGrid[(size_t) c]);
		// the hull's outside is empty space (CGAL's infinite cells carry EmptySpaceFlag)
		for (int k = 0; k < 4; k++) {
			if (t.cellN[(size_t) 4 * c + k] < 0 && t.cellV[(size_t) 4 * c + k] != vertex) { incidentGrids.insert(EmptySpaceFlag); }
		}
	}
	if (incidentGrids.size() == 1) { return; }

	auto addBorderNode = [&](const GridId gridId) {
		Body& body = getBody(gridId);
		const Mesh& mesh = *body.mesh;
		const int it = mesh.localOf[(size_t) vertex];
		const real* bn = &mesh.borderNormals[(size_t) 3 * it];
		const bool isMulticontact = bn[0] == 0 && bn[1] == 0 && bn[2] == 0;
		Border* chosen = nullptr;
		for (Border& border : body.borders) {
			if (border.correctionArea->contains(mesh.coords(it)) && (!isMulticontact || border.useForMulticontactNodes)) { chosen = &border; }
		}
		if (!chosen) { return; }
		const real* n = &mesh.commonNormals[(size_t) 3 * it];
		if (n[0] == 0 && n[1] == 0 && n[2] == 0) { throw Exception(GCMB_E_BAD_MESH, "border vertex without a normal"); }
		chosen->nodes.push_back(it);
		chosen->normals.insert(chosen->normals.end(), n, n + 3);
	};

	if (incidentGrids.erase(EmptySpaceFlag)) {
		for (const int id : incidentGrids) { addBorderNode((GridId) id); }
	} else if (incidentGrids.size() == 2) {
		const GridsPair pair = {(GridId) *incidentGrids.begin(), (GridId) *incidentGrids.rbegin()};
		const Mesh& first = *getBody(pair.first).mesh;
		const Mesh& second = *getBody(pair.second).mesh;
		// contact normals of the first body towards the second, computed on the device once per pair
		std::vector<real>& normals = contactNormalsCache[pair];
		if (normals.size() != first.globalOf.size() * 3) {
			normals.resize(first.globalOf.size() * 3);
			check(gcmb_simplex_contact_normals(first.body, (int) pair.second, normals.data()));
		}
		const int a = first.localOf[(size_t) vertex], b = second.localOf[(size_t) vertex];
		const real* n = &normals[(size_t) 3 * a];
		if (n[0] != 0 || n[1] != 0 || n[2] != 0) {
			Contact& c = contacts.at(pair);
			c.first.push_back(a);
			c.second.push_back(b);
			c.normals.insert(c.normals.end(), n, n + 3);
		}
	} else {
		for (const int id : incidentGrids) { addBorderNode((GridId) id); }
	}
}

real Engine::randomReal(real lo, real hi) {
	randomState = randomState * 6364136223846793005ull + 1442695040888963407ull;
	const real u = (real) ((randomState >> 11) & ((1ull << 53) - 1)) / (real) (1ull << 53);
	return (hi - lo) * u + lo;
}

/// linal::randomBasis (linal/basis.hpp:125-137, special/RotationMatrix.hpp:15-45) with a reproducible generator;
/// Engine::changeCalculationBasis (engine/simplex/Engine.hpp:208-216)
void Engine::changeCalculationBasis() {
	if (!createNewRandomAtEachTimeStep) { return; }
	const real phi = randomReal(-M_PI, M_PI), teta = randomReal(-M_PI, M_PI), khi = randomReal(-M_PI, M_PI);
	const real X[9] = {1.0, 0.0, 0.0, 0.0, std::cos(phi), std::sin(phi), 0.0, -std::sin(phi), std::cos(phi)};
	const real Y[9] = {std::cos(teta), 0.0, -std::sin(teta), 0.0, 1.0, 0.0, std::sin(teta), 0.0, std::cos(teta)};
	const real Z[9] = {std::cos(khi), std::sin(khi), 0.0, -std::sin(khi), std::cos(khi), 0.0, 0.0, 0.0, 1.0};
	auto mul = [](const real* a, const real* b, real* c) {
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
			real r = a[i * 3] * b[j];
			for (int k = 1; k < 3; k++) { r += a[i * 3 + k] * b[k * 3 + j]; }
			c[i * 3 + j] = r;
		}
	};
	real zy[9], ans[9];
	mul(Z, Y, zy);
	mul(zy, X, ans);
	for (int j = 0; j < 3; j++) {
		const Real3 col = {{ans[j], ans[3 + j], ans[6 + j]}};
		const real len = length(col);
		for (int i = 0; i < 3; i++) { ans[i * 3 + j] = col[(size_t) i] / len; }
	}
	std::copy(ans, ans + 9, basis);
	for (Body& body : bodies) { setMaterial(body); }
}

std::vector<real> Engine::borderValues(const Body& body, const real time) const {
	const int outer = outerNumber(body.mesh->modelType);
	std::vector<real> b(body.borders.size() * (size_t) outer + 1, 0.0);
	for (size_t c = 0; c < body.borders.size(); c++) {
		for (int i = 0; i < outer; i++) { b[c * (size_t) outer + (size_t) i] = body.borders[c].values[(size_t) i](time); }
	}
	return b;
}

/// engine/simplex/Engine.cpp:197-214
void Engine::applyPlainBorderContactCorrection(const real timeForBorderCondition) {
	for (auto& contact : contacts) { check(gcmb_simplex_contact_plain(contact.second.handle)); }
	for (const Body& body : bodies) {
		const std::vector<real> b = borderValues(body, timeForBorderCondition);
		check(gcmb_simplex_plain_border(body.mesh->body, b.data()));
		body.mesh->invalidateHostCopy();
	}
}

/// engine/simplex/Engine.cpp:97-115
void Engine::nextTimeStep() {
	changeCalculationBasis();
	applyPlainBorderContactCorrection(Clock::Time() + Clock::TimeStep());
	for (int stage = 0; stage < 3; stage++) { gcmStage(stage, Clock::Time(), Clock::TimeStep()); }
	if (summSplitting) { for (const Body& body : bodies) { check(gcmb_simplex_average_layers(body.mesh->body)); } }
	for (const Body& body : bodies) {   // engine/simplex/Engine.cpp:110-114
		for (size_t o = 0; o < body.odes.size(); o++) {
			const auto& iso = dynamic_cast<const IsotropicMaterial&>(*body.mesh->material);
			check(gcmb_simplex_ode_maxwell(body.mesh->body, std::exp(-Clock::TimeStep() / iso.tau0)));
		}
	}
	for (const Body& body : bodies) { body.mesh->invalidateHostCopy(); }
}

/// engine/simplex/Engine.cpp:118-191 (GLOBAL_BASIS: contacts first, then every body's borders in task order)
void Engine::gcmStage(const int stage, const real currentTime, const real timeStep) {
	for (const Body& body : bodies) { check(gcmb_simplex_before_stage(body.mesh->body, stage, timeStep)); }
	for (const Body& body : bodies) { check(gcmb_simplex_border_contact_stage(body.mesh->body)); }
	for (auto& contact : contacts) { check(gcmb_simplex_contact_correct(contact.second.handle)); }
	for (const Body& body : bodies) {
		const std::vector<real> b = borderValues(body, currentTime + timeStep);
		check(gcmb_simplex_border_correct(body.mesh->body, b.data()));
	}
	for (const Body& body : bodies) { check(gcmb_simplex_inner_stage(body.mesh->body)); }
	for (const Body& body : bodies) { check(gcmb_simplex_after_stage(body.mesh->body)); }
}

/// engine/simplex/Engine.hpp:77-92
real Engine::estimateTimeStep() {
	real minimalTimeStep = std::numeric_limits<real>::max();
	for (const Body& body : bodies) {
		const real h = body.mesh->getAverageHeight();
		const real bodyTimeStep = CourantNumber * h / body.mesh->getMaximalEigenvalue();
		if (bodyTimeStep < minimalTimeStep) { minimalTimeStep = bodyTimeStep; }
	}
	return minimalTimeStep;
}

/// Engine::writeSnapshots (engine/simplex/Engine.cpp:313-320) -> VtkSnapshotter::snapshotImpl
/// (util/snapshot/VtkSnapshotter.hpp:20-61): the body's vertices and its own cells as a .vtu
void Engine::writeSnapshots(const int step_) {
	if (step_ % stepsPerSnap != 0) { return; }
	for (const Snapshotters::T s : snapshotters) {
		if (s != Snapshotters::T::VTK) { continue; }
		checkNodeErrors();  // never write a snapshot of a state the reference would not have reached
		for (const Body& body : bodies) {
			const Mesh& m = *body.mesh;
			const size_t n = m.sizeOfRealNodes();
			std::vector<size_t> order(n);
			std::vector<float> points(3 * n), material(n, (float) m.material->materialNumber);
			for (size_t l = 0; l < n; l++) {
				order[l] = l;
				const Real3 c = m.coords((int) l);
				for (size_t d = 0; d < 3; d++) { points[3 * l + d] = (float) c[d]; }
			}
			std::vector<int32_t> cells;
			for (int c = 0; c < triangulation.nC; c++) {
				if (triangulation.cellGrid[(size_t) c] != (int) m.id) { continue; }
				for (int k = 0; k < 4; k++) { cells.push_back(m.localOf[(size_t) triangulation.cellV[(size_t) 4 * c + k]]); }
			}
			const auto fields = vtk::snapshotFields(m.modelType, 3, m.M, m.pdeAll(), order, vtkSettings.quantitiesToSnap, material);
			vtk::writeUnstructuredGrid(vtk::snapshotFileName(outputDirectory, "vtk", m.id, 0, step_, "vtu"), points, cells, fields);
		}
	}
}

void Engine::checkNodeErrors() const {
	if (!throwOnNodeErrors) { return; }
	const int n = errorCount();
	if (n > 0) {
		throw Exception(GCMB_E_BAD_MESH, std::to_string(n) + " node computations hit a case on which the reference engine throws "
				"(degenerate cell location, interpolation outside the cell or failed least squares): the results are not valid");
	}
}

void Engine::finishRun() { checkNodeErrors(); }

int Engine::errorCount() const {
	int total = 0;
	for (const Body& body : bodies) {
		int c = 0;
		check(gcmb_simplex_errors(body.mesh->body, &c));
		total += c;
	}
	return total;
}

void Engine::borderNodes(const GridId id, const size_t condition, std::vector<int>& nodes, std::vector<real>& normals) const {
	const Border& b = getBody(id).borders.at(condition);
	nodes = b.nodes;
	normals = b.normals;
}
void Engine::contactNodes(const GridsPair& pair, std::vector<int>& first, std::vector<int>& second, std::vector<real>& normals) const {
	const Contact& c = contacts.at(pair);
	first = c.first; second = c.second; normals = c.normals;
}

size_t Engine::numberOfContactNodes(const GridsPair& pair) const { return contacts.at(pair).first.size(); }
size_t Engine::numberOfBorderNodes(const GridId id, const size_t condition) const { return getBody(id).borders.at(condition).nodes.size(); }

}  // namespace simplex
}  // namespace gcmb
