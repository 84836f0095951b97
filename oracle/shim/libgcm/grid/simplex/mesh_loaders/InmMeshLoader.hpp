// TEST INFRASTRUCTURE ONLY.  Shadows the reference's InmMeshLoader.hpp, which inserts points into a CGAL Delaunay
// triangulation (not available); the oracle build feeds meshes through the Cgal3DMesher stand-in instead.
#ifndef LIBGCM_INMMESHLOADER_HPP
#define LIBGCM_INMMESHLOADER_HPP
#include <string>
#include <libgcm/util/infrastructure/infrastructure.hpp>
namespace gcm {
struct InmMeshLoader {
	template<typename Triangulation>
	static void load(const std::string, Triangulation&) { THROW_UNSUPPORTED("INM loading needs CGAL"); }
};
}
#endif
