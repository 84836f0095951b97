// TEST INFRASTRUCTURE ONLY.  Force-included in front of the reference's simplex translation units
// (grid/simplex/SimplexGrid.cpp, grid/simplex/cgal/CgalTriangulation.cpp, engine/simplex/Engine.cpp).  Those files end
// with explicit instantiations for 2 AND 3 dimensions; the 2-D ones would drag in the CGAL 2-D wrapper.  Declaring
// explicit specialisations of the 2-D classes first turns those instantiation statements into no-ops
// ([temp.explicit]: an explicit instantiation of an explicitly specialised template has no effect), so the files
// themselves stay untouched and their 3-D halves are compiled as they are.
#pragma once
#include "compat_prelude.hpp"
#include <libgcm/grid/simplex/VertexInfoAndCellInfo.hpp>
#include <libgcm/grid/simplex/cgal/CgalTriangulation.hpp>
#include <libgcm/grid/simplex/SimplexGrid.hpp>
#include <libgcm/engine/simplex/Engine.hpp>
namespace gcm {
template<> class CgalTriangulation<2, VertexInfo, CellInfoT<3>> { };
template<> class SimplexGrid<2, CgalTriangulation> { };
namespace simplex {
template<> class Engine<2, CgalTriangulation> { };
}
}
