// TEST INFRASTRUCTURE ONLY.  The 2-D CGAL wrapper is not part of the oracle build (3-D bodies only).
#ifndef LIBGCM_CGAL2DTRIANGULATION_HPP
#define LIBGCM_CGAL2DTRIANGULATION_HPP
namespace gcm {
template<typename VertexInfo, typename CellInfo> class Cgal2DTriangulation;
}
#endif
