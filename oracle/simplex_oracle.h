/*
 * TEST INFRASTRUCTURE ONLY — plain-C CPU restatement of the reference's simplex (tetrahedral) GCM path.
 * Nothing under gcm_b200/ may include, link or call this.
 *
 * Parity status: **pinned**.  The reference's simplex engine needs CGAL (absent from this image and not vendored:
 * `find_package(CGAL 4.8)`, /root/reference/CMakeLists.txt:41-43), but CGAL is only its container: oracle/Makefile
 * builds oracle/_ref/gcm_ref_simplex from the UNMODIFIED reference sources (engine/simplex/Engine.cpp,
 * grid/simplex/SimplexGrid.cpp, grid/simplex/cgal/CgalTriangulation.cpp + .hpp, Cgal3DTriangulation.hpp, LineWalker.hpp,
 * both GCMs, the correctors, interpolators, linal) against a stand-in for the handful of CGAL calls they make
 * (shim/CGAL/flat_triangulation_3.h: cells from a file instead of a Delaunay triangulation) and a restatement of
 * GSL's LU (shim/libgcm/util/math/GslUtils.hpp).  tests/golden/simplex_*.npz hold that engine's results on five
 * scenarios (cavity, two bodies in contact, rotated basis, both GCM types, elastic and acoustic); this restatement,
 * driven in the reference's order, reproduces every value bit for bit (tests/test_host_logic_emul.py
 * ::test_simplex_oracle_matches_reference_bitwise), and so does the product (…::test_simplex_engine_matches_
 * reference_bitwise on the stepping harness, tests/test_gpu_parity.py::test_simplex_cuda_engine_matches_reference_
 * bitwise on the GPU).  What stays outside the pin: CGAL's own meshing/Delaunay code and its (unspecified) order of
 * incident cells, which both sides replace by ascending cell ids.  The reference tests' PROPERTIES are asserted as
 * well: containment of the query in the returned cell for every vertex x 16x16 directions x 9 lengths
 * (src/test/sequence/TestLineWalkSearch3D.cpp:120-154), gradient exactness on linear fields, zero stays zero
 * (TestSimplexGcm.cpp:29-67).
 *
 * Topology is a flat triangulation (our own, CGAL-free): points, 4 vertex ids and 4 neighbour ids per cell
 * (neighbour i is opposite vertex i, -1 outside the hull), a grid id per cell (EMPTY = no body), and the
 * incident cells of every vertex in ascending cell id (the reference's order is unspecified,
 * grid/simplex/cgal/Cgal3DTriangulation.hpp:88-101; both sides freeze this one).
 */
#ifndef GCM_SIMPLEX_ORACLE_H
#define GCM_SIMPLEX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCMO_EMPTY_SPACE (-1)
#define GCMO_MAX_NEIGHBORS 20 /* Cgal3DTriangulation.hpp:53 */

typedef struct {
	int nV, nC;
	const double* xyz;       /* [nV][3] */
	const int* cell_v;       /* [nC][4] global vertex ids */
	const int* cell_n;       /* [nC][4] neighbour cell ids, -1 = outside the convex hull */
	const int* cell_grid;    /* [nC] */
	const int* inc_off;      /* [nV+1] */
	const int* inc_cell;     /* incident cells of every vertex, ascending cell id */
	int grid_id;             /* the body this view works on */
	const int* local_of;     /* [nV] local index of a global vertex in this body, -1 if not in the body */
	const int* global_of;    /* [nLocal] */
	int n_local;
} gcmo_tri;

/* border state of a local vertex: 0 inner, 1 border, 2 contact, 3 multicontact (SimplexGrid.hpp:385-393) */
int gcmo_simplex_border_state(const gcmo_tri* t, int local_vertex);

/* border (which=0) or common (which=1) unit normal of a local vertex; returns 0 when there is none */
int gcmo_simplex_normal(const gcmo_tri* t, int local_vertex, int which, double out[3]);

/* SimplexGrid::findCellCrossedByTheRay (grid/simplex/SimplexGrid.cpp:61-164).  out[0] = n (4 cell, 3/2/1
 * border facet/edge/vertex, 0 none), out[1..n] = local vertex ids, rest -1.  Returns 0, or 1 when the reference
 * would have thrown (degenerate linear system / failed assertion). */
int gcmo_simplex_locate(const gcmo_tri* t, int local_vertex, const double shift[3], int out[5]);

/* neighbours of a local vertex: ascending local id, itself excluded (SimplexGrid.hpp:226-236); returns count */
int gcmo_simplex_neighbors(const gcmo_tri* t, int local_vertex, int* out, int capacity);

/* Differentiation::estimateGradient (util/math/Differentiation.hpp:33-63): values [nLocal][M] ->
 * grad [nLocal][3][M] */
int gcmo_simplex_gradient(const gcmo_tri* t, int M, const double* values, double* grad);

/* TetrahedronInterpolator::hybridInterpolate of component k in a located cell (…Interpolator.hpp:93-104) */
double gcmo_simplex_hybrid_interpolate(const gcmo_tri* t, int M, const double* values, const double* grad,
                                       const int cell[4], int k, const double q[3], int* err);

/* geometry helpers exported for the known-answer tests (linal/geometry.hpp) */
double gcmo_oriented_volume(const double a[3], const double b[3], const double c[3], const double d[3]);
int gcmo_barycentric4(const double a[3], const double b[3], const double c[3], const double d[3],
                      const double q[3], double lambda[4]);

/* One stage of the simplex GCM in Riemann invariants with border correction, single body, GLOBAL_BASIS,
 * PRODUCT splitting (engine/simplex/Engine.cpp:121-141 and everything it calls).
 *   U, U1 [3][M][M], L [3][M] in the calculation basis `basis` (row-major 3x3, stage s runs along column s)
 *   border nodes: n_border, node ids (local), normals [n][3], condition id per node;
 *   conditions: type (0 FIXED_FORCE, 1 FIXED_VELOCITY), b values at t+tau [n_cond][outer_number]
 *   model: 0 elastic (M=9, outer 3), 1 acoustic (M=4, outer 1)
 * cur [nLocal][M] -> next [nLocal][M].  Returns 0 or an error count. */
int gcmo_simplex_stage(const gcmo_tri* t, int model, int M, int s, double tau,
                       const double* U, const double* U1, const double* L, const double* basis,
                       int n_border, const int* border_node, const double* border_normal, const int* border_cond,
                       int n_cond, const int* cond_type, const double* cond_b,
                       const double* cur, double* next);

/* the same stage in the four phases simplex::Engine::gcmStage runs over ALL bodies (Engine.cpp:118-141), so that
 * several bodies with contacts can be driven in the reference's order:
 *   begin (all bodies) -> nodes(pass 0) (all) -> contact_correct (every contact) -> border_correct (all)
 *   -> nodes(pass 1) (all) -> end (all) */
typedef struct gcmo_sstage gcmo_sstage;
/* gcm_type: 0 GcmType::ADVECT_RIEMANN_INVARIANTS (…InRiemannInvariants.hpp), 1 ADVECT_PDE_VECTORS (…InPdeVectors.hpp) */
gcmo_sstage* gcmo_sx_begin(const gcmo_tri* t, int model, int M, int s, double tau, const double* U, const double* U1,
                           const double* L, const double* basis, const double* cur, double* next, int gcm_type);
void gcmo_sx_nodes(gcmo_sstage* h, int pass);
void gcmo_sx_border_correct(gcmo_sstage* h, int n_border, const int* border_node, const double* border_normal,
                            const int* border_cond, int n_cond, const int* cond_type, const double* cond_b);
/* contact of body a with body b (ContactCorrector.hpp:133-253,334-410): node pairs (local ids), normals a -> b.
 * Elastic bodies: ADHESION; acoustic bodies: SLIDE (the only combinations the reference's factory offers). */
void gcmo_sx_contact_correct(gcmo_sstage* a, gcmo_sstage* b, int n, const int* node_a, const int* node_b, const double* normals);
int gcmo_sx_end(gcmo_sstage* h);

/* normal of a local vertex towards the body `neighbor` only (SimplexGrid.hpp:141-144) */
int gcmo_simplex_contact_normal(const gcmo_tri* t, int local_vertex, int neighbor, double out[3]);

/* AbstractContactCorrector::applyPlainCorrection: averages of the two nodes' values (ContactCorrector.hpp:256-269) */
void gcmo_simplex_plain_contact(int model, int M, int n, const int* node_a, const int* node_b, const double* normal,
                                double* pde_a, double* pde_b);

/* out[0] = SimplexGrid::getAverageHeight (mean of a 100-bin histogram of minimal cell heights), out[1] = minimal */
void gcmo_simplex_heights(const gcmo_tri* t, double out[2]);

/* Engine::applyPlainBorderContactCorrection for border nodes (Engine.cpp:197-214) */
void gcmo_simplex_plain_border(int model, int M, int n_border, const int* border_node, const double* border_normal,
                               const int* border_cond, const int* cond_type, const double* cond_b, double* pde);

#ifdef __cplusplus
}
#endif
/* TriangleInterpolator<real> (util/math/interpolation/TriangleInterpolator.hpp:8-130) for n independent queries.
 * mode 0 interpolate (linear), 1 interpolate (quadratic), 2 minMaxInterpolate, 3 hybridInterpolate: points [n][3][2],
 * values [n][3], grads [n][3][2] (modes 1-3); mode 4 interpolateInOwner: points [n][4][2], values [n][4].
 * queries [n][2] -> out [n], status [n] (1 where the reference throws). */
void gcmo_triangle_interpolate(int mode, int n, const double* points, const double* values, const double* grads,
                               const double* queries, double* out, int* status);

#endif
