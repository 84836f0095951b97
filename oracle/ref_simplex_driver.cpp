// TEST INFRASTRUCTURE ONLY — never linked into or called from the product (gcm_b200/).
//
// Driver around the UNMODIFIED reference simplex engine: simplex::Engine<3, CgalTriangulation>
// (src/libgcm/engine/simplex/Engine.cpp), SimplexGrid.cpp, LineWalker.hpp, the GCM in Riemann invariants / PDE
// vectors, the border and contact correctors, the interpolators and linal -- all compiled where they lie, with CGAL
// replaced by the flat container of shim/CGAL/flat_triangulation_3.h (CGAL is only a container + point location
// service for this code; the cell location itself is the reference's own line walk).  The mesh is a "flat" dump
// written by the tests from the product's box mesher, so both sides work on the very same triangulation.
//
// usage: gcm_ref_simplex <task-file> <dump-prefix> [--locate <queries-file>]
//   task file: the simplex subset of the plain-text task format (gcm_b200/host/task_file.cpp) with
//              "simplex_flat FILE" naming the mesh dump (simplex_box/region/cavity lines are ignored)
//   <dump-prefix>.body<ID>.f64  per local vertex: x y z, then the M PDE values (raw doubles)
//   <dump-prefix>.meta          steps, time, tau, per body: vertices, M, average/minimal height
//   <dump-prefix>.cells         grid id of every finite cell after the reference's clean-up of the triangulation
#include <chrono>
#include <cstdio>

#include <libgcm/engine/simplex/DefaultMesh.hpp>
#include <libgcm/rheology/models/models.hpp>

using namespace gcm;

namespace {

struct Tokens {
	std::vector<std::string> t;
	size_t pos = 0;
	bool done() const { return pos >= t.size(); }
	std::string next() {
		if (done()) { THROW_INVALID_ARG("task file: unexpected end of line"); }
		return t[pos++];
	}
	std::string peek() const { return done() ? std::string() : t[pos]; }
	real num() { return std::stod(next()); }
	int inum() { return std::stoi(next()); }
	Real3 vec() { Real3 v; for (int i = 0; i < 3; i++) { v(i) = num(); } return v; }
};

std::shared_ptr<Area> parseArea(Tokens& tk) {
	const std::string kind = tk.next();
	if (kind == "infinite") { return std::make_shared<InfiniteArea>(); }
	if (kind == "box") { const Real3 a = tk.vec(), b = tk.vec(); return std::make_shared<AxisAlignedBoxArea>(a, b); }
	if (kind == "sphere") { const real r = tk.num(); const Real3 c = tk.vec(); return std::make_shared<SphereArea>(r, c); }
	THROW_INVALID_ARG("task file: unknown area " + kind);
}

PhysicalQuantities::T parseQuantity(const std::string& s) {
	typedef PhysicalQuantities::T Q;
	static const std::map<std::string, Q> m = {
		{"Vx", Q::Vx}, {"Vy", Q::Vy}, {"Vz", Q::Vz}, {"Sxx", Q::Sxx}, {"Sxy", Q::Sxy}, {"Sxz", Q::Sxz},
		{"Syy", Q::Syy}, {"Syz", Q::Syz}, {"Szz", Q::Szz}, {"PRESSURE", Q::PRESSURE}};
	return m.at(s);
}

Task::TimeDependency parseTimeDependency(Tokens& tk) {
	const std::string kind = tk.next();
	if (kind == "const") { const real c = tk.num(); return [c](real) { return c; }; }
	if (kind == "sin") { const real amp = tk.num(), omega = tk.num(); return [amp, omega](real t) { return amp * sin(omega * t); }; }
	if (kind == "until") { const real t1 = tk.num(), value = tk.num(); return [t1, value](real t) { return (t < t1) ? value : real(0); }; }
	THROW_INVALID_ARG("task file: unknown time dependency " + kind);
}

void parseTaskFile(const std::string& fileName, Task& task, std::map<size_t, bool>& acousticBodies) {
	task.globalSettings.gridId = Grids::T::SIMPLEX;
	task.globalSettings.verboseTimeSteps = false;
	task.globalSettings.stepsPerSnap = 1;
	task.materialConditions.type = Task::MaterialCondition::Type::BY_BODIES;
	task.simplexGrid.mesher = Task::SimplexGrid::Mesher::CGAL_MESHER;   // -> the stand-in loader
	task.contactCondition.defaultCondition = ContactConditions::T::ADHESION;
	std::ifstream in(fileName);
	if (!in.is_open()) { THROW_INVALID_ARG("cannot open task file " + fileName); }
	std::string line;
	while (std::getline(in, line)) {
		const size_t hash = line.find('#');
		if (hash != std::string::npos) { line = line.substr(0, hash); }
		std::istringstream ss(line);
		Tokens tk;
		std::string w;
		while (ss >> w) { tk.t.push_back(w); }
		if (tk.done()) { continue; }
		const std::string key = tk.next();
		if (key == "grid" || key == "simplex_box" || key == "region" || key == "cavity") { continue; }
		else if (key == "dimensionality") { task.globalSettings.dimensionality = tk.inum(); }
		else if (key == "courant") { task.globalSettings.CourantNumber = tk.num(); }
		else if (key == "steps") { task.globalSettings.numberOfSnaps = tk.inum(); }
		else if (key == "simplex_flat") { task.simplexGrid.fileName = tk.next(); }
		else if (key == "body") {
			const size_t id = (size_t) tk.inum();
			const bool acoustic = tk.next() == "acoustic";
			acousticBodies[id] = acoustic;
			task.bodies[id] = {Materials::T::ISOTROPIC, acoustic ? Models::T::ACOUSTIC : Models::T::ELASTIC, {}};
			tk.next();  // isotropic
			if (tk.peek() == "ode") { tk.next(); tk.next(); task.bodies[id].odes.push_back(Odes::T::MAXWELL_VISCOSITY); }
		} else if (key == "material") {
			if (tk.next() != "body") { THROW_INVALID_ARG("simplex tasks take materials by bodies"); }
			const size_t id = (size_t) tk.inum();
			tk.next();  // isotropic
			const real rho = tk.num(), la = tk.num(), mu = tk.num();
			real tau0 = 0;
			if (tk.peek() == "tau0") { tk.next(); tau0 = tk.num(); }
			task.materialConditions.byBodies.bodyMaterialMap[id] = std::make_shared<IsotropicMaterial>(rho, la, mu, 0, 0, 0, tau0);
		} else if (key == "basis") {
			task.calculationBasis.clear();
			for (int i = 0; i < 9; i++) { task.calculationBasis.push_back(tk.num()); }
		} else if (key == "border_condition") {
			Task::BorderCondition bc;
			bc.area = parseArea(tk);
			bc.type = tk.next() == "fixed_velocity" ? BorderConditions::T::FIXED_VELOCITY : BorderConditions::T::FIXED_FORCE;
			if (tk.peek() == "no_multicontact") { tk.next(); bc.useForMulticontactNodes = false; }
			while (!tk.done()) { bc.values.push_back(parseTimeDependency(tk)); }
			task.borderConditions.push_back(bc);
		} else if (key == "contact") {
			task.contactCondition.defaultCondition = tk.next() == "slide" ? ContactConditions::T::SLIDE : ContactConditions::T::ADHESION;
		} else if (key == "initial") {
			if (tk.next() != "quantity") { THROW_INVALID_ARG("only 'initial quantity' is supported here"); }
			Task::InitialCondition::Quantity q;
			q.physicalQuantity = parseQuantity(tk.next());
			q.value = tk.num();
			q.area = parseArea(tk);
			task.initialCondition.quantities.push_back(q);
		} else if (key == "border_calc_mode") {
			task.simplexGrid.borderCalcMode = tk.next() == "local" ? BorderCalcMode::LOCAL_BASIS : BorderCalcMode::GLOBAL_BASIS;
		} else if (key == "splitting") {
			task.globalSettings.splittingType = tk.next() == "summ" ? SplittingType::SUMM : SplittingType::PRODUCT;
		} else if (key == "gcm_type") {
			task.globalSettings.gcmType = tk.next() == "pde_vectors" ? GcmType::ADVECT_PDE_VECTORS : GcmType::ADVECT_RIEMANN_INVARIANTS;
		} else {
			THROW_INVALID_ARG("task file: unknown key " + key);
		}
	}
}

typedef simplex::Engine<3, CgalTriangulation> Engine3;
typedef SimplexGrid<3, CgalTriangulation> Grid3;

template<typename Model>
void dumpBody(const Engine3& engine, const size_t id, const std::string& prefix, std::ofstream& meta) {
	typedef simplex::DefaultMesh<Model, Grid3, IsotropicMaterial> Mesh;
	auto mesh = std::dynamic_pointer_cast<const Mesh>(engine.getMesh(id));
	assert_true(mesh);
	const int M = Mesh::PdeVector::M;
	std::vector<double> data;
	size_t n = 0;
	for (auto it : *mesh) {
		const Real3 x = mesh->coordsD(it);
		for (int i = 0; i < 3; i++) { data.push_back(x(i)); }
		for (int i = 0; i < M; i++) { data.push_back(mesh->pde(it)(i)); }
		n++;
	}
	const std::string name = prefix + ".body" + std::to_string(id) + ".f64";
	FILE* f = fopen(name.c_str(), "wb");
	assert_true(f);
	fwrite(data.data(), sizeof(double), data.size(), f);
	fclose(f);
	meta << "body " << id << " M " << M << " vertices " << n << " average_height " << mesh->getAverageHeight()
	     << " minimal_height " << mesh->getMinimalHeight() << "\n";
}

}  // namespace


/// --locate: SimplexGrid::findCellCrossedByTheRay (grid/simplex/SimplexGrid.cpp:61-112) of the reference for a list of
/// queries "body local_vertex sx sy sz"; one answer "n p0 p1 p2 p3" per line (local vertex indices, -1 padding), or
/// "throw" where the reference throws.  The grids are built on the reference-cleaned triangulation, like the engine's.
int locateMode(const Task& task, const std::string& queriesFile, const std::string& outFile) {
	CgalTriangulation<3, VertexInfo, CellInfoT<4>> triangulation(task);
	std::map<size_t, std::shared_ptr<Grid3>> grids;
	for (const auto& b : task.bodies) { grids[b.first] = std::make_shared<Grid3>(b.first, Grid3::ConstructionPack({&triangulation})); }
	std::ifstream in(queriesFile);
	std::ofstream out(outFile);
	size_t body, vertex;
	Real3 shift;
	while (in >> body >> vertex >> shift(0) >> shift(1) >> shift(2)) {
		try {
			const Grid3::Cell c = grids.at(body)->findCellCrossedByTheRay(Grid3::Iterator(vertex), shift);
			out << c.n;
			for (int i = 0; i < 4; i++) { out << " " << (i < c.n ? (long) c(i).iter : -1L); }
			out << "\n";
		} catch (Exception&) {
			out << "throw\n";
		}
	}
	return 0;
}

int main(int argc, char** argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: %s <task-file> <dump-prefix> [--locate <queries-file>]\n", argv[0]);
		return 2;
	}
	MPI_Init(&argc, &argv);
	try {
		Task task;
		std::map<size_t, bool> acousticBodies;
		parseTaskFile(argv[1], task, acousticBodies);
		if (argc > 4 && std::string(argv[3]) == "--locate") { return locateMode(task, argv[4], std::string(argv[2]) + ".located"); }
		{
			// grid ids of the finite cells after the reference's own clean-up of the triangulation
			// (CgalTriangulation.cpp:8-112: hanged cells, disconnected cell sets), in file order
			CgalTriangulation<3, VertexInfo, CellInfoT<4>> cleaned(task);
			std::ofstream cells(std::string(argv[2]) + ".cells");
			for (auto c = cleaned.allCellsBegin(); c != cleaned.allCellsEnd(); ++c) {
				if (cleaned.isInfinite(c)) { continue; }
				const GridId id = c->info().getGridId();
				cells << (id == (GridId) (-1) ? -1L : (long) id) << "\n";
			}
		}
		Engine3 engine(task);
		const auto t0 = std::chrono::high_resolution_clock::now();
		engine.run();
		const auto t1 = std::chrono::high_resolution_clock::now();
		std::ofstream meta(std::string(argv[2]) + ".meta");
		meta.precision(17);
		meta << "time " << Clock::Time() << "\n";
		meta << "tau " << Clock::TimeStep() << "\n";
		meta << "steps " << (long) llround(Clock::Time() / Clock::TimeStep()) << "\n";
		meta << "run_seconds " << std::chrono::duration<double>(t1 - t0).count() << "\n";
		for (const auto& b : acousticBodies) {
			if (b.second) { dumpBody<AcousticModel<3>>(engine, b.first, argv[2], meta); }
			else { dumpBody<ElasticModel<3>>(engine, b.first, argv[2], meta); }
		}
	} catch (Exception& e) {
		fprintf(stderr, "gcm::Exception: %s\n", e.what().c_str());
		return 1;
	}
	MPI_Finalize();
	return 0;
}
