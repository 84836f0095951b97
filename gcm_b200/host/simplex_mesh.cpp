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
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <list>
#include <map>
#include <set>

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

/// The clean-up of the body ids that the reference's triangulation constructor runs after meshing
/// (grid/simplex/cgal/CgalTriangulation.cpp:8-112 with CgalTriangulation.hpp:206-290): cells none of whose four
/// neighbours carries their id take the most common neighbouring id ("hanged cells"), and around every vertex
/// where some id forms several disconnected sets of cells, all but the largest set of the id with the most sets are
/// handed to another id; repeated until nothing changes, at most ten times.  The hull's outside counts as empty
/// space (CGAL's infinite cells: one per hull facet, joined across the hull edges).  Ids compare as the reference's
/// GridId (size_t): empty space, (size_t)(-1), is the largest.  Where the reference picks "a random other id" it
/// calls libc's rand(); so does this (the outcome then depends on the process' rand() state, as the reference's).
/// Returns the number of cells whose id changed.
int FlatTriangulation::cleanBodyIds() {
	typedef unsigned long long Id;
	auto idOf = [this](int cell) -> Id { return (Id) (long long) cellGrid[(size_t) cell]; };   // -1 -> max
	// infinite cells: ids nC.., vertices (a, b, c, infinite = -1), neighbours like CGAL's
	std::vector<std::array<int, 4>> infV, infN;
	std::vector<std::array<int, 4>> nb((size_t) nC);
	std::map<std::array<int, 2>, std::vector<std::array<int, 2>>> hullEdges;
	for (int c = 0; c < nC; c++) for (int k = 0; k < 4; k++) { nb[(size_t) c][(size_t) k] = cellN[(size_t) 4 * c + k]; }
	for (int c = 0; c < nC; c++) for (int k = 0; k < 4; k++) {
		if (cellN[(size_t) 4 * c + k] >= 0) { continue; }
		std::array<int, 4> v = {{-1, -1, -1, -1}};
		int m = 0;
		for (int j = 0; j < 4; j++) { if (j != k) { v[(size_t) m++] = cellV[(size_t) 4 * c + j]; } }
		const int me = nC + (int) infV.size();
		infV.push_back(v);
		infN.push_back({{-1, -1, -1, c}});
		nb[(size_t) c][(size_t) k] = me;
		for (int a = 0; a < 3; a++) {
			std::array<int, 2> e = {{v[(size_t) ((a + 1) % 3)], v[(size_t) ((a + 2) % 3)]}};
			if (e[1] < e[0]) { std::swap(e[0], e[1]); }
			hullEdges[e].push_back({{me, a}});
		}
	}
	for (const auto& e : hullEdges) {
		if (e.second.size() != 2) { throw Exception(GCMB_E_BAD_MESH, "the hull of the triangulation is not a closed surface"); }
		infN[(size_t) (e.second[0][0] - nC)][(size_t) e.second[0][1]] = e.second[1][0];
		infN[(size_t) (e.second[1][0] - nC)][(size_t) e.second[1][1]] = e.second[0][0];
	}
	const Id EMPTY = (Id) (long long) EmptySpaceFlag;
	auto id = [&](int cell) -> Id { return cell >= nC ? EMPTY : idOf(cell); };
	auto hasVertex = [&](int cell, int v) {
		if (cell >= nC) { const auto& x = infV[(size_t) (cell - nC)]; return x[0] == v || x[1] == v || x[2] == v; }
		const int* x = &cellV[(size_t) 4 * cell];
		return x[0] == v || x[1] == v || x[2] == v || x[3] == v;
	};
	auto neighbor = [&](int cell, int k) { return cell >= nC ? infN[(size_t) (cell - nC)][(size_t) k] : nb[(size_t) cell][(size_t) k]; };
	// incident cells of every vertex, finite ones first in ascending id, then the infinite ones
	std::vector<std::vector<int>> incident((size_t) nV);
	for (int v = 0; v < nV; v++) { incident[(size_t) v].assign(incCell.begin() + incOff[(size_t) v], incCell.begin() + incOff[(size_t) v + 1]); }
	for (size_t i = 0; i < infV.size(); i++) for (int k = 0; k < 3; k++) { incident[(size_t) infV[i][(size_t) k]].push_back(nC + (int) i); }

	const std::vector<int> before = cellGrid;
	int hanged = 1, disconnected = 1, iterations = 0;
	long long guard = 0;
	while (disconnected > 0 || hanged > 0) {
		if (++iterations > 10) { break; }
		// correctHangedCells (CgalTriangulation.cpp:42-66)
		hanged = 0;
		for (int c = 0; c < nC; c++) {
			std::multiset<Id> around;
			for (int k = 0; k < 4; k++) { around.insert(id(neighbor(c, k))); }
			if (around.find(idOf(c)) != around.end()) { continue; }
			++hanged;
			Id common = *around.begin();
			for (const Id x : around) { if (around.count(x) > around.count(common)) { common = x; } }
			cellGrid[(size_t) c] = (int) (long long) common;
		}
		// clearFromDisconnectedCellSets for every vertex (CgalTriangulation.cpp:69-111)
		disconnected = 0;
		for (int v = 0; v < nV; v++) {
			while (true) {
				std::set<Id> unique;
				for (const int c : incident[(size_t) v]) { unique.insert(id(c)); }
				if (unique.size() == 1) { break; }
				// connected sets of equal-id cells around the vertex, in the order the reference builds them
				struct Set { Id id; std::set<int> cells; };
				std::vector<Set> sets;   // kept sorted by id, equal ids in insertion order (std::multiset)
				std::set<int> all(incident[(size_t) v].begin(), incident[(size_t) v].end());
				while (!all.empty()) {
					Set s;
					s.id = id(*all.begin());
					std::vector<int> stack = {*all.begin()};
					while (!stack.empty()) {
						const int c = stack.back();
						stack.pop_back();
						if (!(hasVertex(c, v) && id(c) == s.id) || !s.cells.insert(c).second) { continue; }
						for (int k = 0; k < 4; k++) { stack.push_back(neighbor(c, k)); }
					}
					size_t at = sets.size();
					while (at > 0 && s.id < sets[at - 1].id) { at--; }
					sets.insert(sets.begin() + (long) at, s);
					for (const int c : s.cells) { all.erase(c); }
				}
				struct Info { Id id; size_t nSets = 0, nCells = 0; };
				std::list<Info> byCells;
				for (const Id x : unique) {
					Info info;
					info.id = x;
					for (const Set& s : sets) { if (s.id == x) { ++info.nSets; info.nCells += s.cells.size(); } }
					byCells.push_back(info);
				}
				byCells.sort([](const Info& a, const Info& b) { return a.nCells > b.nCells; });
				std::list<Info> bySets(byCells);
				bySets.sort([](const Info& a, const Info& b) { return a.nSets < b.nSets; });
				const Info toRemove = bySets.back();
				if (toRemove.nSets == 1) { break; }
				// Utils::chooseRandomElementExceptSpecified (util/Utils.hpp:57-76): libc rand(), which the reference has not
				// seeded yet when its triangulation is constructed (AbstractEngine::run seeds it later)
				std::set<Id> others(unique);
				others.erase(toRemove.id);
				const real upTo = (1 - 1e-9) * (real) others.size();
				const int pick = (int) ((upTo * std::rand()) / RAND_MAX + 0);
				auto chosen = others.begin();
				for (int i = 0; i < pick; i++) { ++chosen; }
				const Id toInsert = *chosen;
				if (++guard > 100000) { throw Exception(GCMB_E_BAD_MESH, "the clean-up of body ids does not terminate"); }
				// removeDCS: every set of that id but the (first) largest changes its id
				const Set* largest = nullptr;
				for (const Set& s : sets) { if (s.id == toRemove.id && (!largest || largest->cells.size() < s.cells.size())) { largest = &s; } }
				for (const Set& s : sets) {
					if (s.id != toRemove.id || &s == largest) { continue; }
					for (const int c : s.cells) { if (c < nC) { cellGrid[(size_t) c] = (int) (long long) toInsert; } }
				}
				++disconnected;
			}
		}
	}
	int changed = 0;
	for (int c = 0; c < nC; c++) { if (cellGrid[(size_t) c] != before[(size_t) c]) { changed++; } }
	return changed;
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
