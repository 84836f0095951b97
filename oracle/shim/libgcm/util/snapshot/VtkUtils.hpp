// TEST INFRASTRUCTURE ONLY.  Shadows the reference's VtkUtils.hpp (needs the VTK library): debugging dumps do nothing.
#ifndef LIBGCM_VTKUTILS_HPP
#define LIBGCM_VTKUTILS_HPP
#include <string>
#include <vector>
#include <libgcm/util/Elements.hpp>
namespace gcm {
namespace vtk_utils {
template<typename TElement>
void drawCellsToVtk(const std::vector<TElement>&, const std::string& = "cells") { }
}
}
#endif
