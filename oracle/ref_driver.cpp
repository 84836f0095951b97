// TEST INFRASTRUCTURE ONLY — never linked into or called from the product (gcm_b200/).
//
// Driver around the UNMODIFIED reference sources (/root/reference/src/libgcm), compiled by
// oracle/Makefile into oracle/_ref/gcm_ref.  It parses a plain-text task file (the same format
// gcm_b200's host layer and the C restatement read: see DESIGN.md "task files"), fills the
// reference's own gcm::Task (src/libgcm/util/task/Task.hpp:24-234), runs the reference's own
// cubic::Engine<D> (src/libgcm/engine/cubic/Engine.cpp:13-35, AbstractEngine.cpp:30-46), and dumps
// raw fp64 PDE values of the real nodes of every body (x slowest, M values per node).
//
// usage: gcm_ref <task-file> <dump-prefix> [--matrices]
//   <dump-prefix>.body<ID>.f64   raw doubles, real nodes only, reference iteration order
//   <dump-prefix>.meta           text: steps taken, tau, run() wall seconds, per body sizes/M
//   <dump-prefix>.mat<ID>.f64    (--matrices) for each material condition of body ID, per
//                                direction: U (M*M), U1 (M*M), L (M)
//
#include <chrono>
#include <cstdio>

#include <libgcm/engine/cubic/Engine.hpp>
#include <libgcm/engine/cubic/DefaultMesh.hpp>
#include <libgcm/rheology/models/models.hpp>
#include <libgcm/util/task/MaterialsCondition.hpp>

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
};

std::shared_ptr<Area> parseArea(Tokens& tk) {
	const std::string kind = tk.next();
	if (kind == "infinite") { return std::make_shared<InfiniteArea>(); }
	if (kind == "box") {
		Real3 a, b;
		for (int i = 0; i < 3; i++) { a(i) = tk.num(); }
		for (int i = 0; i < 3; i++) { b(i) = tk.num(); }
		return std::make_shared<AxisAlignedBoxArea>(a, b);
	}
	if (kind == "sphere") {
		real r = tk.num();
		Real3 c;
		for (int i = 0; i < 3; i++) { c(i) = tk.num(); }
		return std::make_shared<SphereArea>(r, c);
	}
	if (kind == "cylinder") {
		real r = tk.num();
		Real3 a, b;
		for (int i = 0; i < 3; i++) { a(i) = tk.num(); }
		for (int i = 0; i < 3; i++) { b(i) = tk.num(); }
		return std::make_shared<StraightBoundedCylinderArea>(r, a, b);
	}
	THROW_INVALID_ARG("task file: unknown area " + kind);
}

Task::MaterialCondition::Material parseMaterial(Tokens& tk) {
	const std::string kind = tk.next();
	if (kind == "isotropic") {
		real rho = tk.num(), la = tk.num(), mu = tk.num();
		real tau0 = 0;
		if (tk.peek() == "tau0") { tk.next(); tau0 = tk.num(); }
		return std::make_shared<IsotropicMaterial>(rho, la, mu, 0, 0, 0, tau0);
	}
	if (kind == "orthotropic") {
		real rho = tk.num();
		real c[9];
		for (int i = 0; i < 9; i++) { c[i] = tk.num(); }
		real tau0 = 0;
		Real3 angles = Real3::Zeros();
		if (tk.peek() == "angles") { tk.next(); for (int i = 0; i < 3; i++) { angles(i) = tk.num(); } }
		if (tk.peek() == "tau0") { tk.next(); tau0 = tk.num(); }
		return std::make_shared<OrthotropicMaterial>(rho,
				std::initializer_list<real>({c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]}),
				0, 0, angles, tau0);
	}
	THROW_INVALID_ARG("task file: unknown material " + kind);
}

PhysicalQuantities::T parseQuantity(const std::string& s) {
	typedef PhysicalQuantities::T Q;
	static const std::map<std::string, Q> m = {
		{"Vx", Q::Vx}, {"Vy", Q::Vy}, {"Vz", Q::Vz},
		{"Sxx", Q::Sxx}, {"Sxy", Q::Sxy}, {"Sxz", Q::Sxz},
		{"Syy", Q::Syy}, {"Syz", Q::Syz}, {"Szz", Q::Szz},
		{"PRESSURE", Q::PRESSURE}};
	return m.at(s);
}

Waves::T parseWave(const std::string& s) {
	typedef Waves::T W;
	static const std::map<std::string, W> m = {
		{"P_FORWARD", W::P_FORWARD}, {"P_BACKWARD", W::P_BACKWARD},
		{"S1_FORWARD", W::S1_FORWARD}, {"S1_BACKWARD", W::S1_BACKWARD},
		{"S2_FORWARD", W::S2_FORWARD}, {"S2_BACKWARD", W::S2_BACKWARD}};
	return m.at(s);
}

Task::TimeDependency parseTimeDependency(Tokens& tk) {
	const std::string kind = tk.next();
	if (kind == "const") {
		real c = tk.num();
		return [c](real) { return c; };
	}
	if (kind == "sin") {
		real amp = tk.num(), omega = tk.num();
		return [amp, omega](real t) { return amp * sin(omega * t); };
	}
	if (kind == "gauss") {
		real amp = tk.num(), t0 = tk.num(), tau = tk.num();
		return [amp, t0, tau](real t) { t -= t0; return amp * exp(-t * t / (2 * tau * tau)); };
	}
	if (kind == "until") {
		real t1 = tk.num(), value = tk.num();
		return [t1, value](real t) { return (t < t1) ? value : real(0); };
	}
	THROW_INVALID_ARG("task file: unknown time dependency " + kind);
}

struct Parsed {
	Task task;
	std::map<size_t, std::pair<std::string, std::string>> bodyKinds; // id -> (model, material)
	std::vector<std::vector<real>> initialVectors; // keep initializer data alive
};

void parseTaskFile(const std::string& fileName, Parsed& out) {
	Task& task = out.task;
	task.globalSettings.gridId = Grids::T::CUBIC;
	task.globalSettings.verboseTimeSteps = false;
	task.globalSettings.stepsPerSnap = 1;
	task.materialConditions.type = Task::MaterialCondition::Type::BY_AREAS;
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
		if (key == "dimensionality") {
			task.globalSettings.dimensionality = tk.inum();
		} else if (key == "courant") {
			task.globalSettings.CourantNumber = tk.num();
		} else if (key == "border_size") {
			task.cubicGrid.borderSize = tk.inum();
		} else if (key == "h") {
			task.cubicGrid.h.clear();
			while (!tk.done()) { task.cubicGrid.h.push_back(tk.num()); }
		} else if (key == "steps") {
			task.globalSettings.numberOfSnaps = tk.inum();
		} else if (key == "required_time") {
			task.globalSettings.numberOfSnaps = 0;
			task.globalSettings.requiredTime = tk.num();
		} else if (key == "body") {
			const size_t id = (size_t) tk.inum();
			const std::string model = tk.next(), material = tk.next();
			Task::Body body;
			body.modelId = (model == "acoustic") ? Models::T::ACOUSTIC : Models::T::ELASTIC;
			body.materialId = (material == "orthotropic") ?
					Materials::T::ORTHOTROPIC : Materials::T::ISOTROPIC;
			Task::CubicGrid::Cube cube;
			while (!tk.done()) {
				const std::string sub = tk.next();
				const int D = task.globalSettings.dimensionality;
				if (sub == "sizes") { for (int i = 0; i < D; i++) { cube.sizes.push_back(tk.inum()); } }
				else if (sub == "start") { for (int i = 0; i < D; i++) { cube.start.push_back(tk.inum()); } }
				else if (sub == "ode") { tk.next(); body.odes.push_back(Odes::T::MAXWELL_VISCOSITY); }
				else { THROW_INVALID_ARG("task file: unknown body option " + sub); }
			}
			task.bodies[id] = body;
			task.cubicGrid.cubics[id] = cube;
			out.bodyKinds[id] = {model, material};
		} else if (key == "material") {
			const std::string how = tk.next();
			if (how == "default") {
				task.materialConditions.byAreas.defaultMaterial = parseMaterial(tk);
			} else if (how == "area") {
				Task::MaterialCondition::ByAreas::Inhomogenity inh;
				inh.area = parseArea(tk);
				inh.material = parseMaterial(tk);
				task.materialConditions.byAreas.materials.push_back(inh);
			} else if (how == "body") {
				task.materialConditions.type = Task::MaterialCondition::Type::BY_BODIES;
				const size_t id = (size_t) tk.inum();
				task.materialConditions.byBodies.bodyMaterialMap[id] = parseMaterial(tk);
			} else { THROW_INVALID_ARG("task file: unknown material clause " + how); }
		} else if (key == "initial") {
			const std::string what = tk.next();
			if (what == "quantity") {
				Task::InitialCondition::Quantity q;
				q.physicalQuantity = parseQuantity(tk.next());
				q.value = tk.num();
				q.area = parseArea(tk);
				task.initialCondition.quantities.push_back(q);
			} else if (what == "wave") {
				Task::InitialCondition::Wave wv;
				wv.waveType = parseWave(tk.next());
				wv.direction = tk.inum();
				wv.quantity = parseQuantity(tk.next());
				wv.quantityValue = tk.num();
				wv.area = parseArea(tk);
				task.initialCondition.waves.push_back(wv);
			} else { THROW_INVALID_ARG("task file: unsupported initial clause " + what); }
		} else if (key == "border") {
			const size_t id = (size_t) tk.inum();
			Task::CubicBorderCondition bc;
			bc.direction = tk.inum();
			bc.area = parseArea(tk);
			while (!tk.done()) {
				PhysicalQuantities::T q = parseQuantity(tk.next());
				bc.values[q] = parseTimeDependency(tk);
			}
			task.cubicBorderConditions[id].push_back(bc);
		} else if (key == "detector") {
			task.detector.gridId = (size_t) tk.inum();
			task.detector.quantities = {parseQuantity(tk.next())};
			task.detector.area = parseArea(tk);
			task.globalSettings.snapshottersId.push_back(Snapshotters::T::SLICESNAP);
			task.globalSettings.outputDirectory = tk.done() ? "" : tk.next();
		} else {
			THROW_INVALID_ARG("task file: unknown key " + key);
		}
	}
}


template<int D, typename Model, typename Material>
void dumpBody(const cubic::Engine<D>& engine, const size_t id,
		const std::string& prefix, std::ofstream& meta) {
	typedef cubic::DefaultMesh<Model, CubicGrid<D>, Material> Mesh;
	auto mesh = std::dynamic_pointer_cast<const Mesh>(engine.getMesh(id));
	assert_true(mesh);
#ifdef GCMB_GPU_BACKEND
	// gcm_ref_gpu: the state lives on the device; DefaultMesh is its host mirror
	dynamic_cast<const cubic::gpu::GpuMeshBase&>(*mesh).downloadToHost();
#endif
	const int M = Mesh::PdeVector::M;
	std::vector<double> data;
	data.reserve(mesh->sizeOfRealNodes() * (size_t) M);
	for (auto it : *mesh) {
		for (int i = 0; i < M; i++) { data.push_back(mesh->pde(it)(i)); }
	}
	const std::string name = prefix + ".body" + std::to_string(id) + ".f64";
	FILE* f = fopen(name.c_str(), "wb");
	assert_true(f);
	fwrite(data.data(), sizeof(double), data.size(), f);
	fclose(f);
	meta << "body " << id << " M " << M << " sizes";
	for (int i = 0; i < D; i++) { meta << " " << mesh->sizes(i); }
	meta << "\n";
}

template<int D, typename Model, typename Material>
void dumpMatrices(const Task& task, const size_t id, const std::string& prefix) {
	typedef MaterialsCondition<Model, CubicGrid<D>, Material, cubic::DefaultMesh> MC;
	auto conditions = MC::convertToLocalFormat(task, id);
	const int M = Model::PDE_SIZE;
	std::vector<double> data;
	for (const auto& c : conditions) {
		for (int s = 0; s < D; s++) {
			const auto& g = (*c.matrices)(s);
			for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) { data.push_back(g.U(i, j)); }
			for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) { data.push_back(g.U1(i, j)); }
			for (int i = 0; i < M; i++) { data.push_back(g.L(i)); }
		}
	}
	const std::string name = prefix + ".mat" + std::to_string(id) + ".f64";
	FILE* f = fopen(name.c_str(), "wb");
	assert_true(f);
	fwrite(data.data(), sizeof(double), data.size(), f);
	fclose(f);
}

template<int D, typename Model, typename Material>
void dumpAll(const cubic::Engine<D>& engine, const Task& task, const size_t id,
		const std::string& prefix, std::ofstream& meta, const bool matrices) {
	dumpBody<D, Model, Material>(engine, id, prefix, meta);
	if (matrices) { dumpMatrices<D, Model, Material>(task, id, prefix); }
}

template<int D>
int runD(Parsed& parsed, const std::string& prefix, const bool matrices) {
	const Task& task = parsed.task;
	cubic::Engine<D> engine(task);
	const auto t0 = std::chrono::high_resolution_clock::now();
	engine.run();
	const auto t1 = std::chrono::high_resolution_clock::now();
	const double seconds = std::chrono::duration<double>(t1 - t0).count();

	std::ofstream meta(prefix + ".meta");
	meta.precision(17);
	meta << "time " << Clock::Time() << "\n";
	meta << "tau " << Clock::TimeStep() << "\n";
	meta << "steps " << (long) llround(Clock::Time() / Clock::TimeStep()) << "\n";
	meta << "run_seconds " << seconds << "\n";
	for (const auto& b : parsed.bodyKinds) {
		const bool acoustic = b.second.first == "acoustic";
		const bool ortho = b.second.second == "orthotropic";
		if (acoustic) {
			dumpAll<D, AcousticModel<D>, IsotropicMaterial>(engine, task, b.first, prefix, meta, matrices);
		} else if (ortho) {
			dumpAll<D, ElasticModel<D>, OrthotropicMaterial>(engine, task, b.first, prefix, meta, matrices);
		} else {
			dumpAll<D, ElasticModel<D>, IsotropicMaterial>(engine, task, b.first, prefix, meta, matrices);
		}
	}
	return 0;
}

}  // namespace


int main(int argc, char** argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: %s <task-file> <dump-prefix> [--matrices]\n", argv[0]);
		return 2;
	}
	const bool matrices = (argc > 3 && std::string(argv[3]) == "--matrices");
	MPI_Init(&argc, &argv);
	try {
		Parsed parsed;
		parseTaskFile(argv[1], parsed);
		switch (parsed.task.globalSettings.dimensionality) {
			case 1: return runD<1>(parsed, argv[2], matrices);
			case 2: return runD<2>(parsed, argv[2], matrices);
			case 3: return runD<3>(parsed, argv[2], matrices);
			default: THROW_INVALID_ARG("bad dimensionality");
		}
	} catch (Exception& e) {
		fprintf(stderr, "gcm::Exception: %s\n", e.what().c_str());
		return 1;
	}
	MPI_Finalize();
	return 0;
}
