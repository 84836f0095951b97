// Shadows <libgcm/util/snapshot/snapshotters.hpp> (same guard): VTK is absent, so the VTK
// snapshotter becomes a no-op; the text SliceSnapshotter is the reference's own.
#ifndef LIBGCM_SNAPSHOTTERS_HPP
#define LIBGCM_SNAPSHOTTERS_HPP
#include <libgcm/util/snapshot/Snapshotter.hpp>
#include <libgcm/util/snapshot/SliceSnapshotter.hpp>
namespace gcm {
template<typename TMesh>
class VtkSnapshotter : public Snapshotter {
public:
	VtkSnapshotter(const Task& t) : Snapshotter(t) { }
	virtual void snapshotImpl(const AbstractGrid*, const int) override { }
};
}
#endif
