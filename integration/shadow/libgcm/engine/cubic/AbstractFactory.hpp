// Stands in front of the reference's engine/cubic/AbstractFactory.hpp on the include path (-I integration/shadow before
// -I <reference>/src), so that the UNMODIFIED cubic::Engine<D> (engine/cubic/Engine.cpp:157-188 names
// AbstractFactory<Model, Grid, Material, DefaultMesh>) receives GPU-backed objects.  The reference's own header is
// included underneath with its factory renamed; nothing in the reference tree is edited.
#ifndef GCM_B200_SHADOW_CUBIC_ABSTRACTFACTORY_HPP
#define GCM_B200_SHADOW_CUBIC_ABSTRACTFACTORY_HPP

#define AbstractFactory ReferenceHostFactory
#include_next <libgcm/engine/cubic/AbstractFactory.hpp>
#undef AbstractFactory

#include <GpuBackend.hpp>

namespace gcm {
namespace cubic {

template<typename TModel, typename TGrid, typename TMaterial,
         template<typename, typename, typename> class TMesh>
class AbstractFactory : public ReferenceHostFactory<TModel, TGrid, TMaterial, TMesh> {
public:
	typedef ReferenceHostFactory<TModel, TGrid, TMaterial, TMesh>   Host;
	typedef gpu::GpuMesh<TModel, TGrid, TMaterial>                  Mesh;   ///< is-a TMesh<TModel, TGrid, TMaterial>
	typedef typename Host::MeshPtr                 MeshPtr;
	typedef typename Host::GcmPtr                  GcmPtr;
	typedef typename Host::OdePtr                  OdePtr;
	typedef typename Host::SnapPtr                 SnapPtr;
	typedef typename Host::GridConstructionPack    GridConstructionPack;
	typedef typename Host::ContactPtr              ContactPtr;
	typedef typename Host::BorderPtr               BorderPtr;
	typedef typename Host::PartIterator            PartIterator;

	virtual MeshPtr createMesh(const Task& task, const GridId gridId, const GridConstructionPack& constructionPack,
			const size_t numberOfNextPdeTimeLayers) override {
		return std::make_shared<Mesh>(task, gridId, constructionPack, numberOfNextPdeTimeLayers);
	}
	virtual GcmPtr createGcm(const Task&) override { return std::make_shared<gpu::GpuGcm>(); }
	virtual BorderPtr createBorder(const Task& task, const MeshPtr mesh) override {
		return std::make_shared<gpu::GpuBorder<TMesh<TModel, TGrid, TMaterial>>>(task, *mesh);
	}
	virtual OdePtr createOde(const Odes::T type) override {
		assert_true(Odes::T::MAXWELL_VISCOSITY == type);
		return std::make_shared<gpu::GpuMaxwellOde>();
	}
	/// the reference's snapshotters, reading a host mirror that is refreshed from the device when they write
	virtual SnapPtr createSnapshotter(const Task& task, const Snapshotters::T type) override {
		return std::make_shared<gpu::GpuSnapshotter>(task, Host::createSnapshotter(task, type));
	}
	virtual ContactPtr createContact(const PartIterator& iterA, const PartIterator& iterB,
			const ContactConditions::T condition, const Models::T neighborModel, const Materials::T) override {
		assert_true(condition == ContactConditions::T::ADHESION);
		assert_true(TModel::Type == neighborModel);
		return std::make_shared<gpu::GpuContact<TGrid>>(iterA, iterB);
	}
};

} // namespace cubic
} // namespace gcm

#endif
