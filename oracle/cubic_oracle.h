/*
 * TEST INFRASTRUCTURE ONLY — a plain-C CPU restatement of the reference's cubic GCM hot path.
 * Nothing under gcm_b200/ may include, link or call this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit
 * against outputs of the unmodified reference (oracle/_ref/gcm_ref, built from /root/reference by
 * oracle/Makefile) committed as fixtures under tests/golden/, and against the known-answer vectors
 * of the reference's own tests (TestInterpolator.cpp:68-77, TestGridCharacteristicMethod.cpp:15-70,
 * TestCubicGrid.cpp:27).
 *
 * Memory layout everywhere: the reference's (src/libgcm/grid/cubic/CubicGrid.hpp:141-147,204-226):
 * array of nodes, x slowest / last axis fastest, `bs` ghost nodes on both sides of every axis,
 * M doubles per node (AoS).
 */
#ifndef GCM_CUBIC_ORACLE_H
#define GCM_CUBIC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCMO_MAX_M 9

/* quantity codes for border conditions / detectors: 0..M-1 = PDE component, or: */
#define GCMO_Q_PRESSURE_TRACE (-1) /* VelocitySigmaVariables::Get/SetPressure (…Variables.hpp:100-111) */

/* ---- eigen-systems (rheology/models): out arrays are [D][M*M] (row-major), [D][M*M], [D][M] ---- */
int gcmo_pde_size(int model /*0 elastic,1 acoustic*/, int D);
void gcmo_elastic_isotropic(int D, double rho, double lambda, double mu,
                            double* U, double* U1, double* L);
void gcmo_elastic_orthotropic(int D, double rho, const double c[9], double* U, double* U1, double* L);
/* rotated material axes, 3-D only (ElasticModel3D.cpp:151-283); 0 = ok */
int gcmo_elastic_orthotropic_rotated(double rho, const double c[9], const double angles[3], double* U, double* U1, double* L);
void gcmo_acoustic(int D, double rho, double lambda, double* U, double* U1, double* L);

/* ---- interpolation (util/math/interpolation/EqualDistanceLineInterpolator.hpp:18-71) ---- */
/* src: (n) vectors of m doubles, overwritten; returns into out[m]; rc!=0 when the reference would throw */
int gcmo_minmax_interpolate(int m, int n, double* src, double q, double* out);
int gcmo_interpolate(int m, int n, double* src, double q, double* out);

/* ---- grid ---- */
size_t gcmo_all_nodes(int D, const int* sizes, int bs);
size_t gcmo_index(int D, const int* sizes, int bs, const int* it);

/* ---- one splitting stage over all real nodes (engine/cubic/GridCharacteristicMethod.hpp:42-87) ----
 * tables: n_tables x [D][..] as produced above; node_table: one byte per node INCLUDING ghosts
 * (ghost entries unused).  cur is read, next gets the real nodes written (ghosts untouched). */
int gcmo_stage(int D, int M, const int* sizes, int bs, const double* h, int s, double tau,
               int n_tables, const double* U, const double* U1, const double* L,
               const uint8_t* node_table, const double* cur, double* next);

/* ---- ghost fill by mirroring (engine/cubic/BorderConditions.hpp:81-114) ----
 * left_mask/right_mask: one byte per node of the face slice(dir,0)/slice(dir,size-1) in SlowXFastZ
 * order (1 = node listed in the condition); q[]: quantity codes in std::map order; val[]: b_q(t_n). */
void gcmo_border_apply(int D, int M, const int* sizes, int bs, int dir,
                       const uint8_t* left_mask, const uint8_t* right_mask,
                       int nq, const int* q, const double* val, double* pde);

/* ---- ghost copy from a neighbour body (engine/cubic/ContactConditions.hpp:56-68) ----
 * boxes in LOCAL indices (may be negative = ghosts), extent nodes per axis */
void gcmo_contact_copy(int D, int M, const int* sizesA, const int* sizesB, int bs,
                       const int* boxA_min, const int* boxB_min, const int* extent,
                       double* pdeA, const double* pdeB);

/* ---- Maxwell viscosity (rheology/ode/Ode.hpp:28-38): sigma *= decay[table] on real nodes ---- */
void gcmo_ode_maxwell(int D, int M, int model, const int* sizes, int bs,
                      const double* decay_per_table, const uint8_t* node_table, double* pde);

/* get a quantity from one node */
double gcmo_get_quantity(int D, int M, int code, const double* node);
void gcmo_set_quantity(int D, int M, int code, double value, double* node);

#ifdef __cplusplus
}
#endif
#endif
