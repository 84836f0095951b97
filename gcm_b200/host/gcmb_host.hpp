// gcm_b200 host layer: the reference's Task / Engine entry points on top of the C ABI.
//
// Mirrors the public interface of libgcm for the cubic hot path so that a user of the reference can
// switch by changing the namespace:  gcm::Task -> gcmb::Task  (reference util/task/Task.hpp:24-234),
// gcm::createEngine -> gcmb::createEngine (engine/EngineFactory.hpp:11-36), AbstractEngine::run()
// (engine/AbstractEngine.cpp:30-46), cubic::Engine<D>::getMesh(id)->pde(it) (engine/cubic/Engine.hpp:34-36,
// engine/cubic/DefaultMesh.hpp:70-81).  All grid data lives on the GPU; the host keeps only the task
// description, the per-material eigen-systems and lazily downloaded copies for accessors/snapshots.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gcm_b200.h"

namespace gcmb {

typedef double real;
typedef std::array<real, 3> Real3;
typedef size_t GridId;

/// thrown for every failure; code() follows gcm::Exception codes (Exception.hpp:94-126)
class Exception : public std::runtime_error {
public:
	Exception(int code, const std::string& what) : std::runtime_error(what), code_(code) { }
	int code() const { return code_; }
private:
	int code_;
};

// ---- enumerations: same names and order as reference util/Enum.hpp ---------------------------
struct PhysicalQuantities {
	enum class T { VELOCITY, FORCE, Vx, Vy, Vz, Sxx, Sxy, Sxz, Syy, Syz, Szz, RHO, PRESSURE, DAMAGE_MEASURE };
};
struct Waves { enum class T { P_FORWARD, P_BACKWARD, S1_FORWARD, S1_BACKWARD, S2_FORWARD, S2_BACKWARD }; };
struct ContactConditions { enum class T { ADHESION, SLIDE }; };
struct Materials { enum class T { ISOTROPIC, ORTHOTROPIC }; };
struct Grids { enum class T { CUBIC, SIMPLEX }; };
struct Models { enum class T { ELASTIC, ACOUSTIC, MAXWELL_ACOUSTIC }; };
struct Snapshotters { enum class T { VTK, DETECTOR, SLICESNAP }; };
struct Odes { enum class T { MAXWELL_VISCOSITY, CONTINUAL_DAMAGE, IDEAL_PLASTIC_FLOW }; };
struct BorderConditions { enum class T { FIXED_FORCE, FIXED_VELOCITY }; };
/// reference util/task/Task.hpp:12-22
enum class BorderCalcMode { GLOBAL_BASIS, LOCAL_BASIS };
enum class GcmType { ADVECT_RIEMANN_INVARIANTS, ADVECT_PDE_VECTORS };
enum class SplittingType { PRODUCT, SUMM };

// ---- areas (reference util/math/Area.hpp) ------------------------------------------------------
struct Area {
	virtual ~Area() { }
	virtual bool contains(const Real3& coords) const = 0;
	virtual void move(const Real3& shift) = 0;
	/// description for the device-side evaluation: kind + parameters (see gcm_b200.h)
	virtual int deviceKind() const = 0;
	virtual std::vector<double> deviceParams() const = 0;
};
struct InfiniteArea : Area {
	bool contains(const Real3&) const override { return true; }
	void move(const Real3&) override { }
	int deviceKind() const override { return 0; }
	std::vector<double> deviceParams() const override { return {}; }
};
struct AxisAlignedBoxArea : Area {
	AxisAlignedBoxArea(const Real3& min_, const Real3& max_);
	bool contains(const Real3& c) const override;
	void move(const Real3& shift) override;
	int deviceKind() const override { return 1; }
	std::vector<double> deviceParams() const override;
	Real3 getMin() const { return min; }
	Real3 getMax() const { return max; }
private:
	Real3 min, max;
};
struct SphereArea : Area {
	SphereArea(const real& radius_, const Real3& center_);
	bool contains(const Real3& c) const override;
	void move(const Real3& shift) override;
	int deviceKind() const override { return 2; }
	std::vector<double> deviceParams() const override;
private:
	real radius;
	Real3 center;
};
struct StraightBoundedCylinderArea : Area {
	StraightBoundedCylinderArea(const real& radius_, const Real3& begin_, const Real3& end_);
	bool contains(const Real3& c) const override;
	void move(const Real3& shift) override;
	int deviceKind() const override { return 3; }
	std::vector<double> deviceParams() const override;
private:
	real radius;
	Real3 begin, end, axis;
};

// ---- materials (reference rheology/materials/*.hpp) --------------------------------------------
struct AbstractMaterial {
	virtual ~AbstractMaterial() { }
	int materialNumber = 0;
};
struct IsotropicMaterial : AbstractMaterial {
	static const Materials::T Type = Materials::T::ISOTROPIC;
	real rho = 0, lambda = 0, mu = 0, yieldStrength = 0, continualDamageParameter = 0, tau0 = 0;
	IsotropicMaterial(real rho_ = 0, real lambda_ = 0, real mu_ = 0, real yieldStrength_ = 0,
	                  real continualDamageParameter_ = 0, int materialNumber_ = 0, real tau0_ = 0);
};
struct OrthotropicMaterial : AbstractMaterial {
	static const Materials::T Type = Materials::T::ORTHOTROPIC;
	real rho = 0;
	/// c11, c12, c13, c22, c23, c33, c44, c55, c66
	real c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	real yieldStrength = 0, continualDamageParameter = 0, tau0 = 0;
	Real3 anglesOfRotation = {{0, 0, 0}};
	OrthotropicMaterial(real rho_ = 0, std::initializer_list<real> c_ = {0, 0, 0, 0, 0, 0, 0, 0, 0},
	                    real yieldStrength_ = 0, real continualDamageParameter_ = 0,
	                    Real3 phi = {{0, 0, 0}}, real tau0_ = 0);
	explicit OrthotropicMaterial(const IsotropicMaterial& isotropic);
};

// ---- eigen-systems (reference util/math/GridCharacteristicMethod.hpp:33-111) -------------------
/// U, U1 row-major [D][M*M]; L [D][M] -- directly the layout gcmb_cubic_set_materials takes
struct GcmMatrices {
	int D = 0, M = 0;
	std::vector<real> U, U1, L;
	real getMaximalEigenvalue() const;
	const real* u(int s) const { return U.data() + (size_t) s * M * M; }
	const real* u1(int s) const { return U1.data() + (size_t) s * M * M; }
	const real* l(int s) const { return L.data() + (size_t) s * M; }
	/// U*U1 == I etc. within eps; throws Exception like checkDecomposition
	void checkDecomposition(real eps) const;
};
int pdeSize(Models::T model, int D);
/// position of sigma_ij in the elastic PDE vector (velocities first, then the packed upper triangle)
int sigmaComponent(int D, int i, int j);
/// Model::constructGcmMatrices in the global basis (reference ElasticModel.hpp:56-65,
/// ElasticModel3D.cpp:432-448, ElasticModel2D.cpp:79-91, AcousticModel.hpp:57-65)
GcmMatrices constructGcmMatrices(Models::T model, int D, const AbstractMaterial& material);
/// the same for an isotropic material in any orthonormal calculation basis (row-major DxD; stage i runs along
/// column i): ElasticModel.hpp:56-65, AcousticModel.hpp:57-65
GcmMatrices constructGcmMatrices(Models::T model, int D, const IsotropicMaterial& material, const real* basis);
/// column of U1 holding the given wave (reference rheology/models/Model.cpp:6-63)
int waveColumn(Models::T model, Materials::T material, int D, Waves::T wave);
/// component index, or GCMB_Q_PRESSURE_TRACE; throws if the model has no such quantity
int quantityCode(Models::T model, int D, PhysicalQuantities::T q);

// ---- Task: field-for-field the reference's (util/task/Task.hpp:24-234) --------------------------
struct Task {
	typedef std::function<real(real)> TimeDependency;

	struct Body {
		Materials::T materialId;
		Models::T modelId;
		std::vector<Odes::T> odes;
	};
	std::map<size_t, Body> bodies;

	struct GlobalSettings {
		int dimensionality = 0;
		Grids::T gridId = Grids::T::CUBIC;
		bool forceSequence = false;
		std::vector<Snapshotters::T> snapshottersId;
		std::string outputDirectory = "";
		real CourantNumber = 0;
		int numberOfSnaps = 0;
		int stepsPerSnap = 1;
		real requiredTime = 0;
		bool verboseTimeSteps = true;
		GcmType gcmType = GcmType::ADVECT_RIEMANN_INVARIANTS;
		SplittingType splittingType = SplittingType::PRODUCT;
	} globalSettings;

	struct CubicGrid {
		struct Cube {
			std::vector<int> sizes;
			std::vector<int> start;
		};
		std::vector<real> h;
		int borderSize = 0;
		std::map<size_t, Cube> cubics;
	} cubicGrid;

	/// reference util/task/Task.hpp:84-122.  The reference meshes with CGAL (not available here); the mesher
	/// this build provides is a structured box cut into tetrahedra (host/simplex_mesh.cpp).
	struct SimplexGrid {
		enum class Mesher { CGAL_MESHER, INM_MESHER, BOX_MESHER } mesher = Mesher::BOX_MESHER;
		real spatialStep = 0;            ///< edge of the box mesher's cubes
		std::string fileName;            ///< INM_MESHER: the mesh file
		real scale = 1;                  ///< denominator to scale the points after loading
		bool movable = false;
		BorderCalcMode borderCalcMode = BorderCalcMode::GLOBAL_BASIS;
		/// BOX_MESHER @{
		std::array<int, 3> boxCubes = {{0, 0, 0}};
		Real3 boxOrigin = {{0, 0, 0}};
		real jitter = 0;                 ///< displacement of interior points, fraction of spatialStep
		unsigned seed = 1;
		struct BodyRegion {
			size_t id;                   ///< cells whose centroid lies in the area belong to this body
			std::shared_ptr<Area> area;  ///< later regions override earlier ones
		};
		std::vector<BodyRegion> bodies;  ///< cells in no region are empty space
		std::vector<std::shared_ptr<Area>> cavities;  ///< cells whose centroid lies here are empty space
		/// @}
	} simplexGrid;

	/// calculation basis of the simplex engine, row-major DxD; stage i runs along column i.  Empty = a new
	/// random basis at every time step (reference util/task/Task.hpp:125-129)
	std::vector<real> calculationBasis;
	unsigned randomBasisSeed = 1;        ///< gcm_b200: the random bases are reproducible

	struct MaterialCondition {
		typedef std::shared_ptr<AbstractMaterial> Material;
		enum class Type { BY_AREAS, BY_BODIES } type = Type::BY_AREAS;
		struct ByAreas {
			Material defaultMaterial;
			struct Inhomogenity {
				std::shared_ptr<Area> area;
				Material material;
			};
			std::vector<Inhomogenity> materials;
		} byAreas;
		struct ByBodies {
			typedef size_t BodyId;
			std::map<BodyId, Material> bodyMaterialMap;
		} byBodies;
	} materialConditions;

	struct InitialCondition {
		struct Vector {
			std::shared_ptr<Area> area;
			std::vector<real> list;
		};
		std::vector<Vector> vectors;
		struct Wave {
			std::shared_ptr<Area> area;
			Waves::T waveType;
			int direction;
			PhysicalQuantities::T quantity;
			real quantityValue;
		};
		std::vector<Wave> waves;
		struct Quantity {
			std::shared_ptr<Area> area;
			PhysicalQuantities::T physicalQuantity;
			real value;
		};
		std::vector<Quantity> quantities;
	} initialCondition;

	struct CubicBorderCondition {
		int direction;
		std::shared_ptr<Area> area;
		typedef std::map<PhysicalQuantities::T, TimeDependency> Values;
		Values values;
	};
	std::map<size_t, std::vector<CubicBorderCondition>> cubicBorderConditions;

	/// simplex grids (reference util/task/Task.hpp:204-213)
	struct BorderCondition {
		std::shared_ptr<Area> area;
		bool useForMulticontactNodes = true;
		BorderConditions::T type = BorderConditions::T::FIXED_FORCE;
		std::vector<TimeDependency> values;  ///< as many as the model has outer characteristics
	};
	std::vector<BorderCondition> borderConditions;

	struct ContactCondition {
		typedef std::pair<size_t, size_t> GridsContact;
		ContactConditions::T defaultCondition = ContactConditions::T::ADHESION;
		std::map<GridsContact, ContactConditions::T> gridToGridConditions;
	} contactCondition;

	struct VtkSnapshotter {
		std::vector<PhysicalQuantities::T> quantitiesToSnap;
	} vtkSnapshotter;

	struct Detector {
		std::vector<PhysicalQuantities::T> quantities;
		std::shared_ptr<Area> area;
		size_t gridId = 0;
	} detector;

	/// gcm_b200 extension: CUDA device to run on (default 0) and slab decomposition along x
	struct Device {
		int device = 0;
		int slabRank = 0, slabCount = 1;   ///< this process owns slab slabRank of slabCount
		const void* ncclUniqueId = nullptr; ///< 128 bytes shared by all slabs when slabCount > 1
		int realBytes = 8;                  ///< 8: double (reference default), 4: float (LIBGCM_DOUBLE_PRECISION off)
		bool fma = false;                   ///< fp64 stage kernels with FMA contraction (<= 1e-12 of the reference, not bitwise)
	} device;
};

// ---- engine ---------------------------------------------------------------------------------
/// time and step of the RUNNING engine (reference engine/GlobalVariables.hpp:16-40).  The reference keeps them in
/// process-wide statics; here every engine owns its clock (AbstractEngine::time / timeStep) and publishes it through this
/// interface while one of its methods runs, so that several engines can live in one process.
struct Clock {
	static real Time() { return time; }
	static real TimeStep() { return timeStep; }
private:
	static real time, timeStep;
	static void setZero() { time = timeStep = 0; }
	static void tickTack() { time += timeStep; }
	friend class AbstractEngine;
};

class AbstractEngine {
public:
	AbstractEngine(const Task& task);
	virtual ~AbstractEngine() { }
	AbstractEngine(const AbstractEngine&) = delete;
	AbstractEngine& operator=(const AbstractEngine&) = delete;
	/// perform all calculations (reference engine/AbstractEngine.cpp:30-46)
	void run();
	/// gcm_b200 extension: perform exactly n more time steps (the body of run()'s loop), for benchmarks
	void advance(int n);
	/// steps performed by run() so far
	int stepsDone() const { return step; }
	/// this engine's own clock (what Clock::Time() / Clock::TimeStep() show while it runs)
	real currentTime() const { return time; }
	real currentTimeStep() const { return timeStep; }
protected:
	real time = 0, timeStep = 0;
	/// publish / take back this engine's clock around everything that reads the global Clock
	void activateClock() const;
	void storeClock();
	const real CourantNumber = 0;
	real requiredTime = 0;
	bool verboseTimeSteps = true;
	int step = 0;
	void afterConstruction(const Task& task);
	virtual void nextTimeStep() = 0;
	virtual real estimateTimeStep() = 0;
	virtual void writeSnapshots(const int step) = 0;
	/// end of run() / advance(): snapshots still in flight are written, deferred error checks are made
	virtual void finishRun() { }
};

namespace cubic {

/// host view of one body (reference cubic/AbstractMesh.hpp + DefaultMesh.hpp accessors)
class Mesh {
public:
	typedef std::array<int, 3> Iterator;  ///< only the first D entries are used
	GridId id = 0;
	int D = 0, M = 0, borderSize = 0;
	std::array<int, 3> sizes = {{1, 1, 1}}, start = {{0, 0, 0}};
	/// the whole body when this process holds a slab of it (== sizes / start otherwise)
	std::array<int, 3> globalSizes = {{1, 1, 1}}, globalStart = {{0, 0, 0}};
	int realBytes = 8;
	std::array<real, 3> h = {{1, 1, 1}};
	Models::T modelType = Models::T::ELASTIC;
	Materials::T materialType = Materials::T::ISOTROPIC;

	~Mesh();
	/// PDE vector (M values) of a real node; downloads the current layer on first use after a step
	const real* pde(const Iterator& it) const;
	/// all real nodes, x slowest, M per node
	const std::vector<real>& pdeRealNodes() const;
	Real3 coords(const Iterator& it) const;
	size_t sizeOfRealNodes() const;
	real getMaximalEigenvalue() const { return maximalEigenvalue; }
	real getMinimalSpatialStep() const;
	gcmb_body* handle() const { return body; }
	const std::vector<std::shared_ptr<AbstractMaterial>>& tableMaterials() const { return materials; }
	const std::vector<GcmMatrices>& tableMatrices() const { return matrices; }
	void invalidateHostCopy() const { hostValid = false; }

	// filled by the engine
	gcmb_body* body = nullptr;
	real maximalEigenvalue = 0;
	std::vector<std::shared_ptr<AbstractMaterial>> materials;  ///< one per table id
	std::vector<GcmMatrices> matrices;
private:
	mutable std::vector<real> host;
	mutable bool hostValid = false;
};

class EngineBase : public AbstractEngine {
public:
	EngineBase(const Task& task, int dimensionality);
	virtual ~EngineBase();
	std::shared_ptr<const Mesh> getMesh(const GridId gridId) const;
	/// detector history written so far: (time, value) as the SliceSnapshotter records them
	const std::vector<std::pair<real, float>>& seismogram() const { return seismo; }
	gcmb_ctx* context() const { return ctx; }
protected:
	void nextTimeStep() override;
	real estimateTimeStep() override;
	void writeSnapshots(const int step) override;
private:
	struct Contact {
		GridId neighborId;
		int direction;
		int boxA[3], boxB[3], extent[3];
	};
	struct Border {
		int direction;
		std::vector<Task::TimeDependency> values;  ///< std::map order
	};
	struct Body {
		std::shared_ptr<Mesh> mesh;
		std::vector<Contact> contacts;
		std::vector<Border> borders;
		std::vector<Odes::T> odes;
		bool borderFilledByStage = false;  ///< the last direction's ghost nodes were written by the previous stage kernel
	};
	int D;
	gcmb_ctx* ctx = nullptr;
	std::vector<Body> bodies;
	Task taskCopy;
	std::vector<std::pair<real, float>> seismo;
	int slabRank = 0, slabCount = 1;

	Body& getBody(const GridId id);
	const Body& getBody(const GridId id) const;
	void createGridsAndContacts(const Task& task);
	void setUpPde(const Task& task, Body& body);
	void setUpBorders(const Task& task, Body& body);
	void sliceSnapshot(const int step);
	void vtkSnapshot(const int step);
	/// VTK snapshots whose read-back is still in flight (gcmb_cubic_download_box_begin) or waiting to be written
	struct PendingSnapshot {
		int step = 0;
		size_t body = 0;
		bool async = false;
		std::vector<real> values;
	};
	std::vector<PendingSnapshot> pending;
	std::vector<void*> pinned;   ///< page-locked read-back buffer per body
	void finishPendingSnapshots();
	/// seismogram taps of a snapshot step whose read-back is in flight (gcmb_cubic_seismo_begin)
	bool seismoPending = false;
	bool seismoDirsMade = false;
	int pendingSeismoStep = 0;
	real pendingSeismoTime = 0;
	void finishPendingSeismo();
	void finishRun() override;
};

template<int Dimensionality>
class Engine : public EngineBase {
public:
	static const int DIMENSIONALITY = Dimensionality;
	Engine(const Task& task) : EngineBase(task, Dimensionality) { }
};

}  // namespace cubic

namespace simplex {

/// CGAL-free triangulation consumed by the simplex path (see host/simplex_mesh.cpp)
static const int EmptySpaceFlag = -1;
struct FlatTriangulation {
	int nV = 0, nC = 0;
	std::vector<double> xyz;     ///< [nV][3]
	std::vector<int> cellV;      ///< [nC][4]
	std::vector<int> cellN;      ///< [nC][4] neighbour opposite vertex i, -1 outside the hull
	std::vector<int> cellGrid;   ///< [nC] body id, -1 = empty space
	std::vector<int> incOff;     ///< [nV+1]
	std::vector<int> incCell;    ///< incident cells, ascending id
	void buildTopology();        ///< orientation, neighbours, incidence from xyz + cellV (+ cellGrid)
	int cleanBodyIds();          ///< the reference's clean-up of hanged cells and disconnected cell sets
};
FlatTriangulation makeBoxMesh(int nx, int ny, int nz, const Real3& origin, real h, real jitter, unsigned seed,
		const Real3* voidMin, const Real3* voidMax, int gridId);

/// INM mesh files (reference grid/simplex/mesh_loaders/InmMeshLoader.hpp): the file's cells are the triangulation
FlatTriangulation loadInmMesh(const std::string& fileName, real scale = 1);
void saveInmMesh(const FlatTriangulation& t, const std::string& fileName);

/// host view of one body of the simplex engine (reference engine/simplex/AbstractMesh.hpp + DefaultMesh.hpp)
class Mesh {
public:
	typedef int Iterator;   ///< local vertex index
	GridId id = 0;
	int M = 0;
	Models::T modelType = Models::T::ELASTIC;
	Materials::T materialType = Materials::T::ISOTROPIC;
	~Mesh();
	size_t sizeOfRealNodes() const { return globalOf.size(); }
	Real3 coords(const Iterator it) const;
	/// PDE vector of a vertex; the current layer is downloaded on first use after a step
	const real* pde(const Iterator it) const;
	const std::vector<real>& pdeAll() const;
	real getMaximalEigenvalue() const { return maximalEigenvalue; }
	real getAverageHeight() const { return averageHeight; }
	real getMinimalHeight() const { return minimalHeight; }
	gcmb_sbody* handle() const { return body; }
	void invalidateHostCopy() const { hostValid = false; }

	// filled by the engine
	const FlatTriangulation* triangulation = nullptr;
	gcmb_sbody* body = nullptr;
	std::vector<int> globalOf, localOf;
	std::vector<uint8_t> state;                 ///< 0 inner, 1 border, 2 contact, 3 multicontact
	std::vector<real> borderNormals, commonNormals;
	std::shared_ptr<AbstractMaterial> material;
	GcmMatrices matrices;
	real maximalEigenvalue = 0, averageHeight = 0, minimalHeight = 0;
private:
	mutable std::vector<real> host;
	mutable bool hostValid = false;
};

/// reference engine/simplex/Engine.hpp:19-258 for Dimensionality = 3, on the flat triangulation
class Engine : public AbstractEngine {
public:
	typedef std::pair<GridId, GridId> GridsPair;
	Engine(const Task& task);
	virtual ~Engine();
	std::shared_ptr<const Mesh> getMesh(const GridId gridId) const { return getBody(gridId).mesh; }
	const FlatTriangulation& getTriangulation() const { return triangulation; }
	gcmb_ctx* context() const { return ctx; }
	/// node computations the reference would have thrown on, over all bodies so far (sync)
	int errorCount() const;
	size_t numberOfContactNodes(const GridsPair& pair) const;
	size_t numberOfBorderNodes(const GridId id, const size_t condition) const;
	/// what Engine::addBorderNode / addContactNode collected (for tests and logs): local ids and normals
	void borderNodes(const GridId id, const size_t condition, std::vector<int>& nodes, std::vector<real>& normals) const;
	void contactNodes(const GridsPair& pair, std::vector<int>& first, std::vector<int>& second, std::vector<real>& normals) const;
	size_t numberOfBorderConditions(const GridId id) const { return getBody(id).borders.size(); }
	const real* calculationBasis() const { return basis; }
	/// Where the reference throws in the middle of a step (a characteristic foot on a mesh edge, interpolation weights
	/// outside the cell, a failed least-squares fit, ...), the kernels count the event and go on; run() / advance() and
	/// every snapshot then throw BAD_MESH here, so that a run the reference would have aborted never ends with rc 0.
	/// false: keep going and leave the count to errorCount() (benchmarks on deliberately rough meshes).
	bool throwOnNodeErrors = true;
protected:
	void nextTimeStep() override;
	real estimateTimeStep() override;
	void writeSnapshots(const int step) override;
	void finishRun() override;
private:
	void checkNodeErrors() const;
	Task::VtkSnapshotter vtkSettings;
	std::vector<Snapshotters::T> snapshotters;
	std::string outputDirectory;
	int stepsPerSnap = 1;
	struct Border {
		std::shared_ptr<Area> correctionArea;
		bool useForMulticontactNodes = true;
		BorderConditions::T type = BorderConditions::T::FIXED_FORCE;
		std::vector<Task::TimeDependency> values;
		std::vector<int> nodes;
		std::vector<real> normals;
	};
	struct Body {
		std::shared_ptr<Mesh> mesh;
		std::vector<Border> borders;
		std::vector<Odes::T> odes;
	};
	struct Contact {
		ContactConditions::T condition = ContactConditions::T::ADHESION;
		std::vector<int> first, second;
		std::vector<real> normals;
		gcmb_scontact* handle = nullptr;
	};
	gcmb_ctx* ctx = nullptr;
	FlatTriangulation triangulation;
	std::vector<Body> bodies;
	std::map<GridsPair, Contact> contacts;
	std::map<GridsPair, std::vector<real>> contactNormalsCache;  ///< set-up only
	real basis[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	bool createNewRandomAtEachTimeStep = false;
	bool summSplitting = false;
	bool localBasis = false;
	unsigned long long randomState = 0;

	Body& getBody(const GridId id);
	const Body& getBody(const GridId id) const;
	void createTriangulation(const Task& task);
	void createMeshes(const Task& task);
	void createContacts(const Task& task);
	void addBorderOrContact(const int vertex);
	void setMaterial(Body& body);
	void applyInitialConditions(const Task& task, Body& body);
	void changeCalculationBasis();
	void gcmStage(const int stage, const real currentTime, const real timeStep);
	void applyPlainBorderContactCorrection(const real timeForBorderCondition);
	std::vector<real> borderValues(const Body& body, const real time) const;
	real randomReal(real lo, real hi);
};

}  // namespace simplex

/// VTK XML snapshots (host/vtk_writer.cpp; reference util/snapshot/VtkSnapshotter.hpp, VtkUtils.hpp)
namespace vtk {
struct Field {
	std::string name;
	int components = 1;
	std::vector<float> data;
};
std::string snapshotFileName(const std::string& outputDirectory, const std::string& folder, size_t meshId, int rank, int step,
		const std::string& extension);
void writeStructuredGrid(const std::string& path, const int (&n)[3], const std::vector<float>& points, const std::vector<Field>& fields);
void writeUnstructuredGrid(const std::string& path, const std::vector<float>& points, const std::vector<int32_t>& tetrahedra,
		const std::vector<Field>& fields);
std::vector<Field> snapshotFields(Models::T model, int D, int M, const std::vector<real>& pde, const std::vector<size_t>& order,
		const std::vector<PhysicalQuantities::T>& quantities, const std::vector<float>& materialIndex);
}  // namespace vtk

/// reference engine/EngineFactory.hpp:11-36
std::shared_ptr<AbstractEngine> createEngine(const Task& task);

/// plain-text task files (DESIGN.md "task files")
Task parseTaskText(const std::string& text);

}  // namespace gcmb
