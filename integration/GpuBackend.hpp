// gcm_b200 as a backend of the UNMODIFIED libgcm: the classes cubic::Engine<D> obtains from its factory
// (reference engine/cubic/AbstractFactory.hpp:36-53), implemented on top of the C ABI (include/gcm_b200.h).
//
//   reference interface (file:line)                                             this file
//   cubic::AbstractMesh<TGrid>        engine/cubic/AbstractMesh.hpp:17-51       GpuMesh      (DefaultMesh + device body)
//   cubic::GridCharacteristicMethodBase  .../GridCharacteristicMethod.hpp:13-17 GpuGcm       -> gcmb_cubic_stage
//   cubic::AbstractBorderConditions   engine/cubic/BorderConditions.hpp:17-20   GpuBorder    -> gcmb_cubic_border_set / _apply
//   cubic::AbstractContactCopier      engine/cubic/ContactConditions.hpp:20-45  GpuContact   -> gcmb_cubic_contact_apply
//   AbstractOde                       rheology/ode/Ode.hpp:17-20                GpuMaxwellOde -> gcmb_cubic_ode_maxwell
//   Snapshotter                       util/snapshot/Snapshotter.hpp:19-74       GpuSnapshotter (reads the state back, then the
//                                                                               reference's own snapshotter writes the files)
//
// The mesh keeps the reference's host storage (DefaultMesh) as a mirror: it is filled by the reference's own
// MaterialsCondition / InitialCondition, uploaded once, and refreshed from the device only when somebody reads it
// (snapshots, Engine::getMesh()->pde(it) via downloadToHost()).  The time loop itself never crosses PCIe except for the
// border values (a few doubles per stage).  The engine is compiled against this header through
// integration/shadow/libgcm/engine/cubic/AbstractFactory.hpp; no reference file is edited.
#ifndef GCM_B200_INTEGRATION_GPUBACKEND_HPP
#define GCM_B200_INTEGRATION_GPUBACKEND_HPP

#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include <gcm_b200.h>

#include <libgcm/engine/GlobalVariables.hpp>
#include <libgcm/engine/cubic/BorderConditions.hpp>
#include <libgcm/engine/cubic/ContactConditions.hpp>
#include <libgcm/engine/cubic/DefaultMesh.hpp>
#include <libgcm/engine/cubic/GridCharacteristicMethod.hpp>
#include <libgcm/rheology/ode/Ode.hpp>
#include <libgcm/util/snapshot/snapshotters.hpp>

namespace gcm {
namespace cubic {
namespace gpu {

inline void check(const int rc) {
	if (rc != GCMB_OK) { THROW_INVALID_OP(std::string("gcm_b200: ") + gcmb_last_error()); }
}

/// one device context per process (the reference engine is one process = one MPI rank); GCM_B200_DEVICE picks the GPU
inline gcmb_ctx* context() {
	static gcmb_ctx* ctx = nullptr;
	if (!ctx) {
		const char* dev = std::getenv("GCM_B200_DEVICE");
		check(gcmb_create(dev ? std::atoi(dev) : 0, (int) sizeof(real), &ctx));
	}
	return ctx;
}

/// what the other backend classes need from a mesh, whatever its model and material
struct GpuMeshBase {
	virtual ~GpuMeshBase() { }
	virtual gcmb_body* handle() const = 0;
	/// refresh the host mirror (DefaultMesh::pdeVariables) from the device
	virtual void downloadToHost() const = 0;
	/// exp(-tau / tau0) per material table (MaxwellViscosityOde, rheology/ode/Ode.hpp:28-38)
	virtual std::vector<double> maxwellDecay(const real timeStep) const = 0;
};

template<typename TModel, typename TGrid, typename TMaterial>
class GpuMesh : public DefaultMesh<TModel, TGrid, TMaterial>, public GpuMeshBase {
public:
	typedef DefaultMesh<TModel, TGrid, TMaterial>   Base;
	typedef typename Base::ConstructionPack         ConstructionPack;
	typedef typename Base::PdeVariables             PdeVariables;
	typedef typename Base::GcmMatricesPtr           GcmMatricesPtr;
	typedef typename Base::ConstMaterialPtr         ConstMaterialPtr;
	static const int D = Base::DIMENSIONALITY;
	static const int M = TModel::PDE_SIZE;

	GpuMesh(const Task& task, const GridId gridId_, const ConstructionPack& constructionPack,
			const size_t numberOfNextPdeTimeLayers_) :
			Base(task, gridId_, constructionPack, numberOfNextPdeTimeLayers_) { }
	virtual ~GpuMesh() { if (body) { gcmb_cubic_body_destroy(body); } }

	/// the reference fills its host storage (materials by areas, initial conditions); the result goes to the device
	virtual void setUpPde(const Task& task) override {
		Base::setUpPde(task);
		static_assert(sizeof(PdeVariables) == M * sizeof(real), "PDE vectors are plain arrays of M reals");
		int sizes[3], start[3];
		double h[3];
		for (int i = 0; i < D; i++) { sizes[i] = this->sizes(i); start[i] = this->start(i); h[i] = this->h(i); }
		check(gcmb_cubic_body_create(context(), D, M, sizes, start, h, this->borderSize, &body));
		// one eigen-system table per distinct GcmMatrices object, in order of first appearance
		std::map<const void*, int> tableOf;
		std::vector<double> U, U1, L;
		std::vector<uint8_t> ids;
		for (auto it : *this) {
			const GcmMatricesPtr& m = this->gcmMatrices[this->getIndex(it)];
			auto found = tableOf.find(m.get());
			if (found == tableOf.end()) {
				found = tableOf.insert({m.get(), (int) tableOf.size()}).first;
				tableMaterials.push_back(this->material(it));
				for (int s = 0; s < D; s++) {
					const auto& g = (*m)(s);
					for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) { U.push_back(g.U(i, j)); }
					for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) { U1.push_back(g.U1(i, j)); }
					for (int i = 0; i < M; i++) { L.push_back(g.L(i)); }
				}
			}
			ids.push_back((uint8_t) found->second);
		}
		if (tableOf.size() > GCMB_MAX_TABLES) { THROW_UNSUPPORTED("too many materials in one body"); }
		check(gcmb_cubic_set_materials(body, (int) tableOf.size(), U.data(), U1.data(), L.data(), ids.data()));
		check(gcmb_cubic_upload_state(body, this->pdeVariables.data(), 1));
	}

	/// gcmb_cubic_stage swaps the device layers itself (cubic/Engine.cpp:110-111 calls stage, then this)
	virtual void swapCurrAndNextPdeTimeLayer(const int) override { }

	virtual gcmb_body* handle() const override { return body; }
	virtual void downloadToHost() const override {
		check(gcmb_cubic_download_state(body, const_cast<PdeVariables*>(this->pdeVariables.data()), 1));
	}
	virtual std::vector<double> maxwellDecay(const real timeStep) const override {
		std::vector<double> decay;
		for (const ConstMaterialPtr& m : tableMaterials) { decay.push_back(exp(-timeStep / m->tau0)); }
		return decay;
	}

private:
	gcmb_body* body = nullptr;
	std::vector<ConstMaterialPtr> tableMaterials;
};

inline const GpuMeshBase& device(const AbstractGrid& mesh) { return dynamic_cast<const GpuMeshBase&>(mesh); }

/// cubic::GridCharacteristicMethod<Mesh>::stage (engine/cubic/GridCharacteristicMethod.hpp:42-52)
class GpuGcm : public GridCharacteristicMethodBase {
public:
	virtual void stage(const int s, const real& timeStep, AbstractGrid& mesh_) const override {
		check(gcmb_cubic_stage(device(mesh_).handle(), s, timeStep));
	}
};

/// cubic::BorderConditions<Mesh> (engine/cubic/BorderConditions.hpp:46-114)
template<typename Mesh>
class GpuBorder : public AbstractBorderConditions {
public:
	typedef typename Mesh::PdeVariables          PdeVariables;
	typedef typename Mesh::PdeVector             PdeVector;
	typedef typename Mesh::Grid::PartIterator    PartIterator;
	typedef std::function<real(real)>            TimeDependency;

	GpuBorder(const Task& task, const AbstractGrid& mesh_) {
		const Mesh& mesh = dynamic_cast<const Mesh&>(mesh_);
		const auto meshConditions = task.cubicBorderConditions.find(mesh.id);
		if (meshConditions == task.cubicBorderConditions.end()) { return; }
		int number = 0;
		for (const Task::CubicBorderCondition& bc : meshConditions->second) {
			Condition c;
			c.direction = bc.direction;
			std::vector<int> codes;
			for (const auto& q : bc.values) {
				assert_false(PdeVariables::QUANTITIES.find(q.first) == PdeVariables::QUANTITIES.end());
				codes.push_back(quantityCode(q.first));
				c.values.push_back(q.second);
			}
			// face nodes inside the condition's area, in the order of CubicGrid::slice (x slowest)
			std::vector<uint8_t> left, right;
			for (PartIterator n = mesh.leftBorder(c.direction); n != n.end(); ++n) { left.push_back(bc.area->contains(mesh.coords(n)) ? 1 : 0); }
			for (PartIterator n = mesh.rightBorder(c.direction); n != n.end(); ++n) { right.push_back(bc.area->contains(mesh.coords(n)) ? 1 : 0); }
			check(gcmb_cubic_border_set(device(mesh_).handle(), number++, c.direction, left.data(), right.data(),
					(int) codes.size(), codes.data()));
			conditions.push_back(c);
		}
	}

	virtual void apply(AbstractGrid& mesh_, const int direction) const override {
		if (conditions.empty()) { return; }
		std::vector<double> values;
		for (const Condition& c : conditions) if (c.direction == direction) {
			for (const TimeDependency& f : c.values) { values.push_back(f(Clock::Time())); }
		}
		check(gcmb_cubic_border_apply(device(mesh_).handle(), direction, (int) values.size(), values.data()));
	}

private:
	struct Condition {
		int direction;
		std::vector<TimeDependency> values;  ///< std::map order of the task's quantities
	};
	std::vector<Condition> conditions;

	/// component index of a quantity, or GCMB_Q_PRESSURE_TRACE: found by letting the reference's own setter write 1
	static int quantityCode(const PhysicalQuantities::T quantity) {
		PdeVector v = PdeVector::Zeros();
		PdeVariables::QUANTITIES.at(quantity).Set(1, v);
		int code = -2, touched = 0;
		for (int i = 0; i < PdeVector::M; i++) { if (v(i) != 0) { touched++; code = i; } }
		if (touched == 1 && v(code) == 1) { return code; }
		if (quantity == PhysicalQuantities::T::PRESSURE) { return GCMB_Q_PRESSURE_TRACE; }
		THROW_UNSUPPORTED("border quantity is neither a PDE component nor the pressure");
	}
};

/// cubic::ContactCopier (engine/cubic/ContactConditions.hpp:56-68): ghost nodes of a <- real nodes of b
template<typename TGrid>
class GpuContact : public AbstractContactCopier<TGrid> {
public:
	typedef AbstractContactCopier<TGrid>         Base;
	typedef typename Base::PartIterator          PartIterator;

	GpuContact(const PartIterator& boxA_, const PartIterator& boxB_) : Base(boxA_, boxB_) {
		corners(this->boxA, minA, extent);
		int extentB[3];
		corners(this->boxB, minB, extentB);
		for (int i = 0; i < TGrid::DIMENSIONALITY; i++) { assert_eq(extent[i], extentB[i]); }
	}

	virtual void apply(AbstractMesh<TGrid>& a, const AbstractMesh<TGrid>& b) override {
		check(gcmb_cubic_contact_apply(device(a).handle(), device(b).handle(), minA, minB, extent));
	}

private:
	int minA[3], minB[3], extent[3];
	static void corners(const PartIterator& box, int (&lo)[3], int (&ext)[3]) {
		int hi[3];
		bool first = true;
		for (PartIterator it = box.begin(); it != it.end(); ++it) {
			for (int i = 0; i < TGrid::DIMENSIONALITY; i++) {
				if (first || it(i) < lo[i]) { lo[i] = it(i); }
				if (first || it(i) > hi[i]) { hi[i] = it(i); }
			}
			first = false;
		}
		for (int i = 0; i < TGrid::DIMENSIONALITY; i++) { ext[i] = hi[i] - lo[i] + 1; }
	}
};

/// MaxwellViscosityOde (rheology/ode/Ode.hpp:28-38)
class GpuMaxwellOde : public AbstractOde {
public:
	virtual void apply(AbstractGrid& mesh_, const real timeStep) override {
		const std::vector<double> decay = device(mesh_).maxwellDecay(timeStep);
		check(gcmb_cubic_ode_maxwell(device(mesh_).handle(), decay.data()));
	}
};

/// the reference's own snapshotter, fed from the device: the host mirror is refreshed on the steps it writes
class GpuSnapshotter : public Snapshotter {
public:
	GpuSnapshotter(const Task& task, const std::shared_ptr<Snapshotter> inner_) : Snapshotter(task), inner(inner_) { }
protected:
	virtual void snapshotImpl(const AbstractGrid* grid, const int step) override {
		device(*grid).downloadToHost();
		inner->snapshot(grid, step);
	}
private:
	std::shared_ptr<Snapshotter> inner;
};

} // namespace gpu
} // namespace cubic
} // namespace gcm

#endif // GCM_B200_INTEGRATION_GPUBACKEND_HPP
