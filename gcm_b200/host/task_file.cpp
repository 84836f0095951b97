// Plain-text task files -> gcmb::Task.  One statement per line, '#' starts a comment:
//   dimensionality D | courant C | border_size B | h h0 [h1 [h2]] | steps N | required_time T
//   body ID (elastic|acoustic) (isotropic|orthotropic) sizes n.. start s.. [ode maxwell]
//   material default MATERIAL | material area AREA MATERIAL | material body ID MATERIAL
//   initial quantity Q value AREA | initial wave W direction Q value AREA
//   border ID direction AREA {Q (const c | sin amp omega)}...
//   detector ID Q AREA [output-directory] | vtk [every N] Q... | output DIRECTORY
//   AREA = infinite | box x0 y0 z0 x1 y1 z1 | sphere r cx cy cz | cylinder r bx by bz ex ey ez
//   MATERIAL = isotropic rho lambda mu [tau0 t] | orthotropic rho c11 c12 c13 c22 c23 c33 c44 c55 c66 [angles a b c] [tau0 t]
// Simplex grids (box mesher):
//   grid simplex | simplex_box nx ny nz ox oy oz h [jitter j] [seed s] | region ID AREA | cavity AREA
//   simplex_mesh FILE [scale S]   (INM mesh file instead of the box mesher)
//   body ID (elastic|acoustic) isotropic | material body ID MATERIAL | basis b00 b01 .. b22 | basis random [seed]
//   border_condition AREA (fixed_force|fixed_velocity) [no_multicontact] (const c | sin amp omega)...
//   contact (adhesion|slide) [ID ID] | gcm_type (riemann_invariants|pde_vectors) | splitting (product|summ) | border_calc_mode (global|local)
// The test suite feeds the same files to the unmodified reference.
#include <cmath>
#include <sstream>

#include "gcmb_host.hpp"

namespace gcmb {
namespace {

struct Words {
	std::vector<std::string> w;
	size_t pos = 0;
	bool done() const { return pos >= w.size(); }
	const std::string& next() {
		if (done()) { throw Exception(GCMB_E_INVALID_ARG, "task text: unexpected end of line"); }
		return w[pos++];
	}
	std::string peek() const { return done() ? std::string() : w[pos]; }
	real num() { return std::stod(next()); }
	int inum() { return std::stoi(next()); }
	Real3 vec() { Real3 v; for (int i = 0; i < 3; i++) { v[(size_t) i] = num(); } return v; }
};

std::shared_ptr<Area> area(Words& t) {
	const std::string kind = t.next();
	if (kind == "infinite") { return std::make_shared<InfiniteArea>(); }
	if (kind == "box") { const Real3 a = t.vec(), b = t.vec(); return std::make_shared<AxisAlignedBoxArea>(a, b); }
	if (kind == "sphere") { const real r = t.num(); return std::make_shared<SphereArea>(r, t.vec()); }
	if (kind == "cylinder") {
		const real r = t.num();
		const Real3 a = t.vec(), b = t.vec();
		return std::make_shared<StraightBoundedCylinderArea>(r, a, b);
	}
	throw Exception(GCMB_E_INVALID_ARG, "task text: unknown area " + kind);
}

Task::MaterialCondition::Material material(Words& t) {
	const std::string kind = t.next();
	Task::MaterialCondition::Material ans;
	if (kind == "isotropic") {
		const real rho = t.num(), la = t.num(), mu = t.num();
		real tau0 = 0;
		if (t.peek() == "tau0") { t.next(); tau0 = t.num(); }
		ans = std::make_shared<IsotropicMaterial>(rho, la, mu, 0, 0, 0, tau0);
	} else if (kind == "orthotropic") {
		const real rho = t.num();
		real c[9];
		for (real& x : c) { x = t.num(); }
		real tau0 = 0;
		Real3 angles = {{0, 0, 0}};
		if (t.peek() == "angles") { t.next(); angles = t.vec(); }   // rotation of the material axes (radians)
		if (t.peek() == "tau0") { t.next(); tau0 = t.num(); }
		ans = std::make_shared<OrthotropicMaterial>(rho,
				std::initializer_list<real>{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]},
				0, 0, angles, tau0);
	} else {
		throw Exception(GCMB_E_INVALID_ARG, "task text: unknown material " + kind);
	}
	return ans;
}

PhysicalQuantities::T quantity(const std::string& s) {
	typedef PhysicalQuantities::T Q;
	static const std::map<std::string, Q> names = {
			{"Vx", Q::Vx}, {"Vy", Q::Vy}, {"Vz", Q::Vz}, {"Sxx", Q::Sxx}, {"Sxy", Q::Sxy}, {"Sxz", Q::Sxz},
			{"Syy", Q::Syy}, {"Syz", Q::Syz}, {"Szz", Q::Szz}, {"PRESSURE", Q::PRESSURE}};
	const auto it = names.find(s);
	if (it == names.end()) { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown quantity " + s); }
	return it->second;
}

Waves::T wave(const std::string& s) {
	typedef Waves::T W;
	static const std::map<std::string, W> names = {
			{"P_FORWARD", W::P_FORWARD}, {"P_BACKWARD", W::P_BACKWARD}, {"S1_FORWARD", W::S1_FORWARD},
			{"S1_BACKWARD", W::S1_BACKWARD}, {"S2_FORWARD", W::S2_FORWARD}, {"S2_BACKWARD", W::S2_BACKWARD}};
	const auto it = names.find(s);
	if (it == names.end()) { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown wave " + s); }
	return it->second;
}

Task::TimeDependency timeDependency(Words& t) {
	const std::string kind = t.next();
	if (kind == "const") { const real c = t.num(); return [c](real) { return c; }; }
	if (kind == "sin") {
		const real amp = t.num(), omega = t.num();
		return [amp, omega](real time) { return amp * std::sin(omega * time); };
	}
	if (kind == "gauss") {  // amp * exp(-(t - t0)^2 / (2 tau^2)): the pulse of the reference's ndi tasks (src/launcher/ndi.hpp:200-203)
		const real amp = t.num(), t0 = t.num(), tau = t.num();
		return [amp, t0, tau](real time) { time -= t0; return amp * std::exp(-time * time / (2 * tau * tau)); };
	}
	if (kind == "until") {  // value while t < t1, then 0: the loading pulse of the reference's cube tasks (src/launcher/main.cpp:604-618)
		const real t1 = t.num(), value = t.num();
		return [t1, value](real time) { return (time < t1) ? value : real(0); };
	}
	throw Exception(GCMB_E_INVALID_ARG, "task text: unknown time dependency " + kind);
}

}  // namespace

Task parseTaskText(const std::string& text) {
	Task task;
	task.globalSettings.gridId = Grids::T::CUBIC;
	task.globalSettings.verboseTimeSteps = false;
	task.globalSettings.stepsPerSnap = 1;
	std::istringstream in(text);
	std::string line;
	while (std::getline(in, line)) {
		const size_t hash = line.find('#');
		if (hash != std::string::npos) { line.resize(hash); }
		std::istringstream ls(line);
		Words t;
		for (std::string w; ls >> w;) { t.w.push_back(w); }
		if (t.done()) { continue; }
		const std::string key = t.next();
		const int D = task.globalSettings.dimensionality;
		if (key == "dimensionality") { task.globalSettings.dimensionality = t.inum(); }
		else if (key == "courant") { task.globalSettings.CourantNumber = t.num(); }
		else if (key == "border_size") { task.cubicGrid.borderSize = t.inum(); }
		else if (key == "h") { task.cubicGrid.h.clear(); while (!t.done()) { task.cubicGrid.h.push_back(t.num()); } }
		else if (key == "steps") { task.globalSettings.numberOfSnaps = t.inum(); }
		else if (key == "required_time") { task.globalSettings.numberOfSnaps = 0; task.globalSettings.requiredTime = t.num(); }
		else if (key == "body") {
			const size_t id = (size_t) t.inum();
			const std::string model = t.next(), mat = t.next();
			Task::Body body;
			body.modelId = model == "acoustic" ? Models::T::ACOUSTIC : Models::T::ELASTIC;
			body.materialId = mat == "orthotropic" ? Materials::T::ORTHOTROPIC : Materials::T::ISOTROPIC;
			Task::CubicGrid::Cube cube;
			while (!t.done()) {
				const std::string sub = t.next();
				if (sub == "sizes") { for (int i = 0; i < D; i++) { cube.sizes.push_back(t.inum()); } }
				else if (sub == "start") { for (int i = 0; i < D; i++) { cube.start.push_back(t.inum()); } }
				else if (sub == "ode") { t.next(); body.odes.push_back(Odes::T::MAXWELL_VISCOSITY); }
				else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown body option " + sub); }
			}
			task.bodies[id] = body;
			task.cubicGrid.cubics[id] = cube;
		} else if (key == "material") {
			const std::string how = t.next();
			if (how == "default") { task.materialConditions.byAreas.defaultMaterial = material(t); }
			else if (how == "area") {
				Task::MaterialCondition::ByAreas::Inhomogenity inh;
				inh.area = area(t);
				inh.material = material(t);
				task.materialConditions.byAreas.materials.push_back(inh);
			} else if (how == "body") {
				task.materialConditions.type = Task::MaterialCondition::Type::BY_BODIES;
				const size_t id = (size_t) t.inum();
				task.materialConditions.byBodies.bodyMaterialMap[id] = material(t);
			} else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown material clause " + how); }
		} else if (key == "initial") {
			const std::string what = t.next();
			if (what == "quantity") {
				Task::InitialCondition::Quantity q;
				q.physicalQuantity = quantity(t.next());
				q.value = t.num();
				q.area = area(t);
				task.initialCondition.quantities.push_back(q);
			} else if (what == "wave") {
				Task::InitialCondition::Wave w;
				w.waveType = wave(t.next());
				w.direction = t.inum();
				w.quantity = quantity(t.next());
				w.quantityValue = t.num();
				w.area = area(t);
				task.initialCondition.waves.push_back(w);
			} else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown initial clause " + what); }
		} else if (key == "border") {
			const size_t id = (size_t) t.inum();
			Task::CubicBorderCondition bc;
			bc.direction = t.inum();
			bc.area = area(t);
			while (!t.done()) {
				const PhysicalQuantities::T q = quantity(t.next());
				bc.values[q] = timeDependency(t);
			}
			task.cubicBorderConditions[id].push_back(bc);
		} else if (key == "grid") {
			task.globalSettings.gridId = t.next() == "simplex" ? Grids::T::SIMPLEX : Grids::T::CUBIC;
		} else if (key == "gcm_type") {
			const std::string type = t.next();
			if (type == "riemann_invariants") { task.globalSettings.gcmType = GcmType::ADVECT_RIEMANN_INVARIANTS; }
			else if (type == "pde_vectors") { task.globalSettings.gcmType = GcmType::ADVECT_PDE_VECTORS; }
			else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown gcm_type " + type); }
		} else if (key == "border_calc_mode") {
			const std::string mode = t.next();
			if (mode == "global") { task.simplexGrid.borderCalcMode = BorderCalcMode::GLOBAL_BASIS; }
			else if (mode == "local") { task.simplexGrid.borderCalcMode = BorderCalcMode::LOCAL_BASIS; }
			else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown border_calc_mode " + mode); }
		} else if (key == "splitting") {
			const std::string type = t.next();
			if (type == "product") { task.globalSettings.splittingType = SplittingType::PRODUCT; }
			else if (type == "summ") { task.globalSettings.splittingType = SplittingType::SUMM; }
			else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown splitting " + type); }
		} else if (key == "simplex_box") {
			Task::SimplexGrid& g = task.simplexGrid;
			for (size_t i = 0; i < 3; i++) { g.boxCubes[i] = t.inum(); }
			g.boxOrigin = t.vec();
			g.spatialStep = t.num();
			while (!t.done()) {
				const std::string sub = t.next();
				if (sub == "jitter") { g.jitter = t.num(); }
				else if (sub == "seed") { g.seed = (unsigned) t.inum(); }
				else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown simplex_box option " + sub); }
			}
		} else if (key == "simplex_mesh") {
			// simplex_mesh FILE [scale S]: an INM mesh file (Task::SimplexGrid::Mesher::INM_MESHER)
			Task::SimplexGrid& g = task.simplexGrid;
			g.mesher = Task::SimplexGrid::Mesher::INM_MESHER;
			g.fileName = t.next();
			if (t.peek() == "scale") { t.next(); g.scale = t.num(); }
		} else if (key == "region") {
			Task::SimplexGrid::BodyRegion r;
			r.id = (size_t) t.inum();
			r.area = area(t);
			task.simplexGrid.bodies.push_back(r);
		} else if (key == "cavity") {
			task.simplexGrid.cavities.push_back(area(t));
		} else if (key == "basis") {
			task.calculationBasis.clear();
			if (t.peek() == "random") { t.next(); if (!t.done()) { task.randomBasisSeed = (unsigned) t.inum(); } }
			else { for (int i = 0; i < 9; i++) { task.calculationBasis.push_back(t.num()); } }
		} else if (key == "border_condition") {
			Task::BorderCondition bc;
			bc.area = area(t);
			const std::string type = t.next();
			if (type == "fixed_force") { bc.type = BorderConditions::T::FIXED_FORCE; }
			else if (type == "fixed_velocity") { bc.type = BorderConditions::T::FIXED_VELOCITY; }
			else { throw Exception(GCMB_E_INVALID_ARG, "task text: unknown border condition type " + type); }
			if (t.peek() == "no_multicontact") { t.next(); bc.useForMulticontactNodes = false; }
			while (!t.done()) { bc.values.push_back(timeDependency(t)); }
			task.borderConditions.push_back(bc);
		} else if (key == "contact") {
			const std::string type = t.next();
			const ContactConditions::T c = type == "slide" ? ContactConditions::T::SLIDE : ContactConditions::T::ADHESION;
			if (t.done()) { task.contactCondition.defaultCondition = c; }
			else { const size_t a = (size_t) t.inum(), b = (size_t) t.inum(); task.contactCondition.gridToGridConditions[{a, b}] = c; }
		} else if (key == "vtk") {
			// vtk [every N] Q...: Snapshotters::T::VTK with vtkSnapshotter.quantitiesToSnap
			task.globalSettings.snapshottersId.push_back(Snapshotters::T::VTK);
			if (t.peek() == "every") { t.next(); task.globalSettings.stepsPerSnap = t.inum(); }
			while (!t.done()) { task.vtkSnapshotter.quantitiesToSnap.push_back(quantity(t.next())); }
		} else if (key == "output") {
			task.globalSettings.outputDirectory = t.next();
		} else if (key == "detector") {
			task.detector.gridId = (size_t) t.inum();
			task.detector.quantities = {quantity(t.next())};
			task.detector.area = area(t);
			task.globalSettings.snapshottersId.push_back(Snapshotters::T::SLICESNAP);
			task.globalSettings.outputDirectory = t.done() ? "" : t.next();
		} else {
			throw Exception(GCMB_E_INVALID_ARG, "task text: unknown key " + key);
		}
	}
	return task;
}

}  // namespace gcmb
