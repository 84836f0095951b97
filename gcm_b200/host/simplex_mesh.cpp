// Flat (CGAL-free) tetrahedral triangulation for the simplex path and a structured box mesher.
//
// The reference builds its triangulation with CGAL (src/libcgalmesher, grid/simplex/cgal/*), which is outside
// the hot path and unavailable here; what the hot path consumes is only: points, the four vertices and four
// neighbours of every cell (neighbour i opposite vertex i), the grid id of every cell
// (src/libgcm/grid/simplex/VertexInfoAndCellInfo.hpp) and the cells incident to every vertex.  This file
// provides exactly that as plain arrays.  Cells outside any body carry EmptySpace (-1) like the reference's
// CellInfo::EmptySpaceFlag; the exterior of the convex hull is neighbour -1 (CGAL's infinite cells).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>

#include "gcmb_host.hpp"

namespace gcmb {
namespace simplex {

static double orientedVolume6(const double* a, const double* b, const double* c, const double* d) {
	const double ba[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
	const double ca[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
	const double da[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
	return ba[0] * (ca[1] * da[2] - ca[2] * da[1]) - ba[1] * (ca[0] * da[2] - ca[2] * da[0]) +
	       ba[2] * (ca[0] * da[1] - ca[1] * da[0]);
}

void FlatTriangulation::buildTopology() {
	nV = (int) (xyz.size() / 3);
	nC = (int) (cellV.size() / 4);
	// positive orientation of every cell (CGAL's convention)
	for (int c = 0; c < nC; c++) {
		int* v = &cellV[(size_t) 4 * c];
		if (orientedVolume6(&xyz[(size_t) 3 * v[0]], &xyz[(size_t) 3 * v[1]], &xyz[(size_t) 3 * v[2]], &xyz[(size_t) 3 * v[3]]) < 0) {
			std::swap(v[2], v[3]);
		}
	}
	// neighbour i is the cell sharing the facet opposite vertex i
	cellN.assign((size_t) 4 * nC, -1);
	std::map<std::array<int, 3>, std::pair<int, int>> open;
	for (int c = 0; c < nC; c++) {
		for (int i = 0; i < 4; i++) {
			std::array<int, 3> f;
			int n = 0;
			for (int k = 0; k < 4; k++) { if (k != i) { f[(size_t) n++] = cellV[(size_t) 4 * c + k]; } }
			std::sort(f.begin(), f.end());
			auto it = open.find(f);
			if (it == open.end()) { open[f] = {c, i}; }
			else {
				cellN[(size_t) 4 * c + i] = it->second.first;
				cellN[(size_t) 4 * it->second.first + it->second.second] = c;
				open.erase(it);
			}
		}
	}
	// incident cells of every vertex, ascending cell id
	incOff.assign((size_t) nV + 1, 0);
	for (int c = 0; c < nC; c++) for (int i = 0; i < 4; i++) { incOff[(size_t) cellV[(size_t) 4 * c + i] + 1]++; }
	for (int v = 0; v < nV; v++) { incOff[(size_t) v + 1] += incOff[(size_t) v]; }
	incCell.assign((size_t) incOff[(size_t) nV], 0);
	std::vector<int> fill(incOff.begin(), incOff.end() - 1);
	for (int c = 0; c < nC; c++) for (int i = 0; i < 4; i++) { incCell[(size_t) fill[(size_t) cellV[(size_t) 4 * c + i]]++] = c; }
}

/// nx x ny x nz cubes of edge h from `origin`, each cut into six tetrahedra around its main diagonal (Kuhn),
/// interior vertices displaced by jitter*h*(u-0.5) per axis (deterministic), cells whose centroid lies
/// strictly inside [voidMin, voidMax] tagged EmptySpace (a cavity with a free border, like the inner
/// tetrahedron of meshes/layers_with_fracture.off ends up after the reference's mesher)
FlatTriangulation makeBoxMesh(int nx, int ny, int nz, const Real3& origin, real h, real jitter, unsigned seed,
		const Real3* voidMin, const Real3* voidMax, int gridId) {
	FlatTriangulation t;
	const int px = nx + 1, py = ny + 1, pz = nz + 1;
	t.xyz.resize((size_t) px * py * pz * 3);
	unsigned long long state = 0x9E3779B97F4A7C15ull ^ seed;
	auto uniform = [&state]() {
		state = state * 6364136223846793005ull + 1442695040888963407ull;
		return (double) ((state >> 11) & ((1ull << 53) - 1)) / (double) (1ull << 53);
	};
	auto vid = [=](int i, int j, int k) { return (i * py + j) * pz + k; };
	for (int i = 0; i < px; i++) for (int j = 0; j < py; j++) for (int k = 0; k < pz; k++) {
		double p[3] = {origin[0] + i * h, origin[1] + j * h, origin[2] + k * h};
		const bool interior = i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz;
		const double d[3] = {uniform(), uniform(), uniform()};
		if (interior) { for (int a = 0; a < 3; a++) { p[a] += jitter * h * (d[a] - 0.5); } }
		std::memcpy(&t.xyz[(size_t) 3 * vid(i, j, k)], p, sizeof p);
	}
	static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
	for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) {
		for (int p = 0; p < 6; p++) {
			int at[3] = {i, j, k};
			int v[4];
			v[0] = vid(at[0], at[1], at[2]);
			for (int s = 0; s < 3; s++) { at[perm[p][s]]++; v[s + 1] = vid(at[0], at[1], at[2]); }
			double c[3] = {0, 0, 0};
			for (int m = 0; m < 4; m++) for (int a = 0; a < 3; a++) { c[a] += t.xyz[(size_t) 3 * v[m] + a] / 4; }
			bool hole = voidMin && voidMax;
			for (int a = 0; a < 3 && hole; a++) { hole = c[a] > (*voidMin)[(size_t) a] && c[a] < (*voidMax)[(size_t) a]; }
			for (int m = 0; m < 4; m++) { t.cellV.push_back(v[m]); }
			t.cellGrid.push_back(hole ? -1 : gridId);
		}
	}
	t.buildTopology();
	return t;
}

/// INM mesh files (the reference's grid/simplex/mesh_loaders/InmMeshLoader.hpp:96-168): number of points, one
/// "x y z" line per point, number of cells, one "v1 v2 v3 v4 material" line per cell (1-based vertex numbers), a
/// closing "0".  The reference feeds the points to CGAL's Delaunay triangulation and keeps the materials of the
/// cells it finds again (:36-94); here the file's own cells ARE the triangulation (nothing is re-meshed, no cell
/// is missed): the material number becomes the body id of the cell, and facets without a neighbour border on
/// empty space.
FlatTriangulation loadInmMesh(const std::string& fileName, real scale) {
	std::ifstream in(fileName);
	if (!in) { throw Exception(GCMB_E_INVALID_ARG, "cannot open mesh file " + fileName); }
	FlatTriangulation t;
	size_t nPoints = 0, nCells = 0;
	if (!(in >> nPoints) || nPoints < 4) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: bad number of points"); }
	t.xyz.resize(nPoints * 3);
	for (size_t i = 0; i < nPoints * 3; i++) {
		if (!(in >> t.xyz[i])) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: bad point"); }
		t.xyz[i] /= scale;   // Task::SimplexGrid::scale: "denominator to scale the points after meshing"
	}
	if (!(in >> nCells) || nCells < 1) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: bad number of cells"); }
	t.cellV.resize(nCells * 4);
	t.cellGrid.resize(nCells);
	for (size_t c = 0; c < nCells; c++) {
		for (int k = 0; k < 4; k++) {
			long long v = 0;
			if (!(in >> v) || v < 1 || (size_t) v > nPoints) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: bad cell vertex"); }
			t.cellV[4 * c + (size_t) k] = (int) (v - 1);
		}
		if (!(in >> t.cellGrid[c])) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: bad cell material"); }
	}
	int zero = -1;
	if (!(in >> zero) || zero != 0) { throw Exception(GCMB_E_BAD_MESH, "INM mesh: the closing 0 is missing"); }
	t.buildTopology();
	return t;
}

void saveInmMesh(const FlatTriangulation& t, const std::string& fileName) {
	FILE* f = std::fopen(fileName.c_str(), "w");
	if (!f) { throw Exception(GCMB_E_INVALID_OP, "cannot write " + fileName); }
	std::fprintf(f, "%d\n", t.nV);
	for (int v = 0; v < t.nV; v++) { std::fprintf(f, "%.17e %.17e %.17e\n", t.xyz[(size_t) 3 * v], t.xyz[(size_t) 3 * v + 1], t.xyz[(size_t) 3 * v + 2]); }
	int cells = 0;
	for (int c = 0; c < t.nC; c++) { if (t.cellGrid[(size_t) c] != EmptySpaceFlag) { cells++; } }
	std::fprintf(f, "%d\n", cells);
	for (int c = 0; c < t.nC; c++) {
		if (t.cellGrid[(size_t) c] == EmptySpaceFlag) { continue; }
		const int* v = &t.cellV[(size_t) 4 * c];
		std::fprintf(f, "%d %d %d %d %d\n", v[0] + 1, v[1] + 1, v[2] + 1, v[3] + 1, t.cellGrid[(size_t) c]);
	}
	std::fprintf(f, "0\n");
	std::fclose(f);
}

}  // namespace simplex
}  // namespace gcmb

extern "C" {

/// box mesher through the C boundary: call with null outputs to get the sizes, then with buffers
/// sizes_out = {nV, nC, nIncident}
int gcmb_host_simplex_box_mesh(int nx, int ny, int nz, const double* origin, double h, double jitter, unsigned seed,
		const double* void_box /* 6 doubles or null */, int grid_id, int* sizes_out,
		double* xyz, int* cell_v, int* cell_n, int* cell_grid, int* inc_off, int* inc_cell) {
	using namespace gcmb;
	const Real3 o = {{origin[0], origin[1], origin[2]}};
	Real3 lo, hi;
	if (void_box) { lo = {{void_box[0], void_box[1], void_box[2]}}; hi = {{void_box[3], void_box[4], void_box[5]}}; }
	simplex::FlatTriangulation t = simplex::makeBoxMesh(nx, ny, nz, o, h, jitter, seed, void_box ? &lo : nullptr,
			void_box ? &hi : nullptr, grid_id);
	sizes_out[0] = t.nV; sizes_out[1] = t.nC; sizes_out[2] = (int) t.incCell.size();
	if (!xyz) { return 0; }
	std::memcpy(xyz, t.xyz.data(), t.xyz.size() * sizeof(double));
	std::memcpy(cell_v, t.cellV.data(), t.cellV.size() * sizeof(int));
	std::memcpy(cell_n, t.cellN.data(), t.cellN.size() * sizeof(int));
	std::memcpy(cell_grid, t.cellGrid.data(), t.cellGrid.size() * sizeof(int));
	std::memcpy(inc_off, t.incOff.data(), t.incOff.size() * sizeof(int));
	std::memcpy(inc_cell, t.incCell.data(), t.incCell.size() * sizeof(int));
	return 0;
}

}  // extern "C"
