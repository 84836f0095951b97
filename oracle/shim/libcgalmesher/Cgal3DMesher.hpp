// TEST INFRASTRUCTURE ONLY.  Stands in for src/libcgalmesher/Cgal3DMesher.hpp (CGAL::make_mesh_3 is not available):
// "meshing" loads the flat triangulation dump named by task.simplexGrid.fileName (see shim/CGAL/flat_triangulation_3.h).
#ifndef LIBCGALMESHER_CGAL3DMESHER_HPP
#define LIBCGALMESHER_CGAL3DMESHER_HPP
#include <string>
namespace cgalmesher {
struct Cgal3DMesher {
	template<typename Triangulation>
	static void triangulate(const double /*spatialStep*/, const bool /*detectSharpEdges*/, const std::string fileName,
			Triangulation& triangulation) {
		triangulation.loadFlat(fileName, -1);
	}
};
}
#endif
