// TEST INFRASTRUCTURE ONLY.  A stand-in for the subset of CGAL's Delaunay_triangulation_3 interface that the
// reference's simplex engine touches (grid/simplex/cgal/Cgal3DTriangulation.hpp, CgalTriangulation.hpp/.cpp,
// SimplexGrid.cpp, LineWalker.hpp), so that the UNMODIFIED reference simplex sources compile and run without CGAL.
// Nothing is triangulated here: the cells come from a file ("flat" dump written by the tests from the product's box
// mesher): points, 4 vertices + 4 neighbours (neighbour i opposite vertex i, -1 = outside the hull) + grid id per
// cell.  The hull's outside is represented the way CGAL does it: one infinite cell per hull facet, incident to the
// infinite vertex, carrying whatever the reference writes into its info (EmptySpaceFlag).  The incident cells of a
// vertex are reported finite cells first in ascending index, then infinite ones (CGAL's order is unspecified).
#ifndef GCM_B200_ORACLE_FLAT_TRIANGULATION_3_H
#define GCM_B200_ORACLE_FLAT_TRIANGULATION_3_H
#include <algorithm>
#include <array>
#include <cstddef>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace CGAL {

struct Exact_predicates_inexact_constructions_kernel {
	struct Vector_3 {
		double c[3];
		Vector_3() : c{0, 0, 0} { }
		Vector_3(double x_, double y_, double z_) : c{x_, y_, z_} { }
		double x() const { return c[0]; }
		double y() const { return c[1]; }
		double z() const { return c[2]; }
	};
	struct Point_3 {
		double c[3];
		Point_3() : c{0, 0, 0} { }
		Point_3(double x_, double y_, double z_) : c{x_, y_, z_} { }
		double x() const { return c[0]; }
		double y() const { return c[1]; }
		double z() const { return c[2]; }
		Point_3 operator+(const Vector_3& v) const { return Point_3(c[0] + v.c[0], c[1] + v.c[1], c[2] + v.c[2]); }
	};
};

template<typename Info, typename K> struct Triangulation_vertex_base_with_info_3 { typedef Info InfoType; };
template<typename Info, typename K> struct Triangulation_cell_base_with_info_3 { typedef Info InfoType; };
template<typename Vb, typename Cb> struct Triangulation_data_structure_3 {
	typedef typename Vb::InfoType VertexInfo;
	typedef typename Cb::InfoType CellInfo;
};

template<typename K, typename Tds>
class Delaunay_triangulation_3 {
public:
	typedef K Geom_traits;
	typedef typename K::Point_3 Point;
	struct Cell;
	struct Vertex {
		Point p;
		typename Tds::VertexInfo i;
		Cell* c = nullptr;
		Point& point() { return p; }
		const Point& point() const { return p; }
		typename Tds::VertexInfo& info() { return i; }
		Cell* cell() const { return c; }
	};
	struct Cell {
		Vertex* v[4] = {nullptr, nullptr, nullptr, nullptr};
		Cell* n[4] = {nullptr, nullptr, nullptr, nullptr};
		typename Tds::CellInfo i;
		bool infinite = false;
		Vertex* vertex(int k) const { return v[k]; }
		Cell* neighbor(int k) const { return n[k]; }
		int index(const Cell* other) const {
			for (int k = 0; k < 4; k++) { if (n[k] == other) { return k; } }
			throw std::runtime_error("flat triangulation: not a neighbour");
		}
		int index(const Vertex* x) const {
			for (int k = 0; k < 4; k++) { if (v[k] == x) { return k; } }
			throw std::runtime_error("flat triangulation: not a vertex of the cell");
		}
		bool has_vertex(const Vertex* x) const { return v[0] == x || v[1] == x || v[2] == x || v[3] == x; }
		typename Tds::CellInfo& info() { return i; }
	};
	typedef Vertex* Vertex_handle;
	typedef Cell* Cell_handle;
	typedef Cell* All_cells_iterator;
	typedef Vertex* Finite_vertices_iterator;

	All_cells_iterator all_cells_begin() const { return const_cast<Cell*>(cells.data()); }
	All_cells_iterator all_cells_end() const { return const_cast<Cell*>(cells.data()) + cells.size(); }
	Finite_vertices_iterator finite_vertices_begin() const { return const_cast<Vertex*>(vertices.data()); }
	Finite_vertices_iterator finite_vertices_end() const { return const_cast<Vertex*>(vertices.data()) + nFiniteVertices; }
	size_t number_of_vertices() const { return nFiniteVertices; }
	size_t number_of_cells() const { return nFiniteCells; }
	bool is_infinite(const Cell* c) const { return c->infinite; }
	bool is_infinite(const Vertex* x) const { return x == vertices.data() + nFiniteVertices; }
	template<typename Out>
	void incident_cells(Vertex_handle x, Out out) const {
		for (Cell* c : incident[(size_t) (x - vertices.data())]) { *out++ = c; }
	}
	Cell_handle locate(const Point&, Cell_handle) const { throw std::runtime_error("flat triangulation: locate() is not provided"); }

	/// file: "nV nC", nV lines "x y z", nC lines "v0 v1 v2 v3 n0 n1 n2 n3 grid" (grid -1 = empty space)
	void loadFlat(const std::string& fileName, long long emptyFlag) {
		std::ifstream in(fileName);
		if (!in) { throw std::runtime_error("flat triangulation: cannot open " + fileName); }
		size_t nV = 0, nC = 0;
		in >> nV >> nC;
		nFiniteVertices = nV;
		nFiniteCells = nC;
		std::vector<std::array<long long, 9>> raw(nC);
		std::vector<double> xyz(3 * nV);
		for (double& x : xyz) { in >> x; }
		size_t hullFacets = 0;
		for (auto& r : raw) {
			for (auto& x : r) { in >> x; }
			for (int k = 0; k < 4; k++) { if (r[4 + k] < 0) { hullFacets++; } }
		}
		if (!in) { throw std::runtime_error("flat triangulation: bad file " + fileName); }
		vertices.assign(nV + 1, Vertex());             // the last one is the infinite vertex
		cells.assign(nC + hullFacets, Cell());
		incident.assign(nV + 1, std::vector<Cell*>());
		for (size_t v = 0; v < nV; v++) { vertices[v].p = Point(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]); }
		Vertex* inf = &vertices[nV];
		size_t nextInfinite = nC;
		std::map<std::pair<Vertex*, Vertex*>, std::vector<std::pair<Cell*, int>>> hullEdges;
		for (size_t c = 0; c < nC; c++) {
			Cell& cell = cells[c];
			for (int k = 0; k < 4; k++) { cell.v[k] = &vertices[(size_t) raw[c][(size_t) k]]; }
			cell.i.setGridId(raw[c][8] < 0 ? (decltype(cell.i.getGridId())) emptyFlag : (decltype(cell.i.getGridId())) raw[c][8]);
		}
		for (size_t c = 0; c < nC; c++) {
			Cell& cell = cells[c];
			for (int k = 0; k < 4; k++) {
				if (raw[c][(size_t) (4 + k)] >= 0) { cell.n[k] = &cells[(size_t) raw[c][(size_t) (4 + k)]]; continue; }
				Cell& out = cells[nextInfinite++];
				out.infinite = true;
				out.i.setGridId((decltype(out.i.getGridId())) emptyFlag);
				int m = 0;
				for (int j = 0; j < 4; j++) { if (j != k) { out.v[m++] = cell.v[j]; } }
				out.v[3] = inf;
				out.n[3] = &cell;
				cell.n[k] = &out;
				for (int a = 0; a < 3; a++) {
					Vertex* p = out.v[(a + 1) % 3];
					Vertex* q = out.v[(a + 2) % 3];
					if (q < p) { std::swap(p, q); }
					hullEdges[{p, q}].push_back({&out, a});   // the neighbour opposite out.v[a] shares edge (p, q)
				}
			}
		}
		for (auto& e : hullEdges) {
			if (e.second.size() != 2) { throw std::runtime_error("flat triangulation: the hull is not a closed surface"); }
			e.second[0].first->n[e.second[0].second] = e.second[1].first;
			e.second[1].first->n[e.second[1].second] = e.second[0].first;
		}
		for (size_t c = 0; c < cells.size(); c++) {
			for (int k = 0; k < 4; k++) {
				Vertex* x = cells[c].v[k];
				incident[(size_t) (x - vertices.data())].push_back(&cells[c]);
				if (!x->c) { x->c = &cells[c]; }
			}
		}
	}

private:
	std::vector<Vertex> vertices;
	std::vector<Cell> cells;
	std::vector<std::vector<Cell*>> incident;
	size_t nFiniteVertices = 0, nFiniteCells = 0;
};

}  // namespace CGAL
#endif
