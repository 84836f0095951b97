/*
 * gcm_b200 — C ABI of the B200-native grid-characteristic time-stepping engine.
 *
 * This is the drop-in boundary for the hot path of AlexanderKazakov/gcm (libgcm): everything
 * cubic::Engine<D>::nextTimeStep (reference src/libgcm/engine/cubic/Engine.cpp:92-121) calls through its
 * per-body virtual interfaces is available here as a plain-C entry point working on device-resident
 * state.  No torch / C++ types cross this boundary: opaque handles, plain pointers and sizes.
 *
 * Conventions
 *   - every call returns 0 on success or a GCMB_E_* code; gcmb_last_error() gives the message
 *     (codes follow gcm::Exception, reference util/infrastructure/Exception.hpp:94-126);
 *   - one host thread per context; all device work is enqueued on the context's CUDA stream and is
 *     asynchronous unless the comment says "sync";
 *   - host-side array layout is the reference's: nodes x-slowest / last axis fastest, `border_size`
 *     ghost nodes on both sides of every axis, M reals per node (array of structures)
 *     (reference grid/cubic/CubicGrid.hpp:141-147,204-226);
 *   - there is NO CPU fallback: without a CUDA device gcmb_create fails with GCMB_E_NO_DEVICE.
 */
#ifndef GCM_B200_H
#define GCM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCMB_MAX_M 9       /* largest PDE vector: 3-D velocity + symmetric stress */
#define GCMB_MAX_BORDER 8  /* largest border size (= interpolation order) */
#define GCMB_MAX_TABLES 255

enum {
	GCMB_OK = 0,
	GCMB_E_INVALID_ARG = 1,  /* gcm::Exception::INVALID_ARG */
	GCMB_E_INVALID_OP = 3,   /* gcm::Exception::INVALID_OP */
	GCMB_E_BAD_MESH = 4,     /* gcm::Exception::BAD_MESH */
	GCMB_E_UNSUPPORTED = -1, /* gcm::Exception::UNSUPPORTED */
	GCMB_E_CUDA = 100,
	GCMB_E_NO_DEVICE = 101,
	GCMB_E_NCCL = 102,
};

/* quantity codes (border conditions, detectors): 0..M-1 selects a PDE component;
 * GCMB_Q_PRESSURE_TRACE is VelocitySigmaVariables::Get/SetPressure
 * (reference rheology/variables/VelocitySigmaVariables.hpp:100-111: set clears the vector first). */
#define GCMB_Q_PRESSURE_TRACE (-1)

typedef struct gcmb_ctx gcmb_ctx;
typedef struct gcmb_body gcmb_body;

const char* gcmb_last_error(void);
const char* gcmb_version(void);

/* ---- context ------------------------------------------------------------------------------- */
/* real_bytes: 8 = fp64 or 4 = fp32 -- the reference's compile-time `real` (util/infrastructure/Types.hpp:8-14,
 * LIBGCM_DOUBLE_PRECISION) chosen per context.  With 4 the grid state, the eigen-system tables and all stage
 * arithmetic are float; `void*` state arrays of gcmb_cubic_upload/download_state, download_box and halo_get/put hold
 * floats; everything typed `double` in this header (tables, tau, border values, sums) stays double and is rounded
 * on the way in.  The simplex entry points need real_bytes == 8. */
int gcmb_create(int device, int real_bytes, gcmb_ctx** out);
int gcmb_real_bytes(gcmb_ctx* ctx);
/* fp64 contexts: on != 0 selects the stage kernels compiled WITH floating-point contraction (FMA).  Their results
 * agree with the reference to <= 1e-12 relative (north_star's fp64 tolerance) but not bit for bit; the default
 * (off, or environment GCMB_FMA=1 to switch it on at gcmb_create) is the bit-exact set, built like the reference
 * without contraction (CMakeLists.txt:6-7).  fp32 contexts always contract. */
int gcmb_set_fma(gcmb_ctx* ctx, int on);
void gcmb_destroy(gcmb_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the own stream */
int gcmb_set_stream(gcmb_ctx* ctx, void* cuda_stream);
int gcmb_sync(gcmb_ctx* ctx); /* sync */
/* device-side timing of everything enqueued between start and stop (CUDA events on the context's
 * stream); stop is sync and returns milliseconds */
int gcmb_timer_start(gcmb_ctx* ctx);
int gcmb_timer_stop(gcmb_ctx* ctx, float* ms);
/* per-kernel-class accumulated device time (CUDA events around every launch when enabled):
 * classes: 0..2 stage along internal axis 0,1,2 (2 = contiguous axis); 3 border; 4 contact; 5 ode;
 * 6 halo pack/unpack; 7 seismo.  get is sync. */
int gcmb_profile_enable(gcmb_ctx* ctx, int on);
int gcmb_profile_get(gcmb_ctx* ctx, int n_classes, double* ms, long long* launches);
/* number of kernels of this library launched on the context since creation */
long long gcmb_launch_count(gcmb_ctx* ctx);
/* bytes of device memory currently held by the context's bodies */
size_t gcmb_device_bytes(gcmb_ctx* ctx);

/* ---- cubic body = one CubicGrid + DefaultMesh storage (engine/cubic/DefaultMesh.hpp:142-167) -- */
/* D in 1..3, M <= GCMB_MAX_M, sizes/start/h have D entries, 1 <= border_size <= GCMB_MAX_BORDER,
 * sizes[i] >= border_size (CubicGrid.hpp:186-199).  Both time layers start zeroed. */
int gcmb_cubic_body_create(gcmb_ctx* ctx, int D, int M, const int* sizes, const int* start,
                           const double* h, int border_size, gcmb_body** out);
void gcmb_cubic_body_destroy(gcmb_body* body);

/* current time layer <-> host array in the reference layout.  with_ghosts=1: all
 * prod(sizes+2*border) nodes; 0: real nodes only (prod(sizes)).  Both sync. */
int gcmb_cubic_upload_state(gcmb_body* body, const void* aos_pde, int with_ghosts);
int gcmb_cubic_download_state(gcmb_body* body, void* aos_pde, int with_ghosts);
/* Asynchronous read-back of a box of nodes of the current layer (snapshots that do not stall the time loop,
 * util/snapshot/VtkSnapshotter.hpp:20-61): begin enqueues a gather + device-to-host copy on a side stream behind the
 * work enqueued so far and returns at once; kernels enqueued later that overwrite the layer wait for the gather only,
 * the PCIe transfer overlaps them.  host: pinned memory of prod(extent)*M reals, x slowest, M per node (the layout of
 * download_state over the box); valid after end (sync).  box_min/extent: D entries, local node indices (ghosts allowed).
 * One read-back in flight per body. */
int gcmb_cubic_download_box_begin(gcmb_body* body, const int* box_min, const int* extent, void* host);
int gcmb_cubic_download_box_end(gcmb_body* body);
/* page-locked host memory for the asynchronous read-backs (cudaHostAlloc / cudaFreeHost) */
int gcmb_host_alloc_pinned(size_t bytes, void** out);
void gcmb_host_free_pinned(void* p);
/* material table id of every REAL node, x slowest (DefaultMesh::material(it), engine/cubic/DefaultMesh.hpp:96-110;
 * read by VtkSnapshotter for its material_index field, util/snapshot/VtkSnapshotter.hpp:49-56). sync */
int gcmb_cubic_download_tables(gcmb_body* body, uint8_t* node_table_id);

/* Eigen-system tables of the materials present in the body (GcmMatrices, reference
 * util/math/GridCharacteristicMethod.hpp:33-111): U, U1 are [n_tables][D][M][M] row-major,
 * L is [n_tables][D][M].  node_table_id: one byte per REAL node (x slowest), NULL = all nodes use
 * table 0 (MaterialsCondition.hpp:23-36).  The library analyses the sparsity of the tables and picks
 * the specialised stage kernels (or the dense one).  sync. */
int gcmb_cubic_set_materials(gcmb_body* body, int n_tables, const double* U, const double* U1,
                             const double* L, const uint8_t* node_table_id);

/* Device-side evaluation of Area-based setup (util/math/Area.hpp), so that large grids never exist
 * on the host.  area_kind: 0 infinite, 1 axis-aligned box {min[3],max[3]}, 2 sphere {r, c[3]},
 * 3 straight bounded cylinder {r, begin[3], end[3]}; params as listed. */
int gcmb_cubic_assign_table_in_area(gcmb_body* body, int table_id, int area_kind, const double* params);
/* pde(node) += vector for real nodes inside the area (InitialCondition.hpp:28-35) */
int gcmb_cubic_add_vector_in_area(gcmb_body* body, const double* vector_M, int area_kind,
                                  const double* params);

/* ---- border conditions = ghost mirroring (engine/cubic/BorderConditions.hpp:46-114) ----------- */
/* Registers condition number `cond` (conditions are applied in increasing `cond`, later ones
 * overwrite): direction, masks over the left/right face nodes (slice(dir,0) / slice(dir,size-1) in
 * x-slowest order, 1 = node inside the condition's area; NULL = no node), the quantity codes in the
 * reference's std::map order.  sync. */
int gcmb_cubic_border_set(gcmb_body* body, int cond, int dir, const uint8_t* left_mask,
                          const uint8_t* right_mask, int n_q, const int* q_codes);
/* same, the masks evaluated on the device from an Area.  sides: bit 0 = left face, bit 1 = right face
 * (3 = both, the reference's behaviour; slabs of a decomposed grid switch off their inner x faces) */
int gcmb_cubic_border_set_area(gcmb_body* body, int cond, int dir, int sides, int area_kind,
                               const double* params, int n_q, const int* q_codes);
/* apply all registered conditions of direction `dir`; values: for every registered condition of that
 * direction in order, its n_q border values b_q(t_n) evaluated by the caller (time dependencies are
 * host functors in the reference); n_values = total count, checked.
 * For the faces across the LAST direction (the contiguous axis), when the surviving condition of each face covers
 * the whole face with plain components, no kernel runs here: the fill is deferred to gcmb_cubic_stage of that
 * direction, whose tile kernel mirrors the ghost nodes inside the shared-memory copy of every row (no pass over the
 * faces in HBM).  Every other call that reads or writes ghost nodes of the body (contact copy, halo exchange, transfers
 * with ghosts, another border/stage call ...) first runs the deferred fill as a kernel, so the reference's order
 * border conditions -> contact copies -> stage (cubic/Engine.cpp:94-111) holds in every case.  What differs from the
 * reference: when the stage consumed the condition, these two faces' ghost nodes in HBM keep their old values (the
 * other directions' stages read real nodes only; nothing observes them). */
int gcmb_cubic_border_apply(gcmb_body* body, int dir, int n_values, const double* values);

/* ---- contact = ghost copy from the neighbour body (engine/cubic/ContactConditions.hpp:56-68) --- */
/* boxes in LOCAL node indices of each body (ghosts are negative / >= size), `extent` nodes per axis */
int gcmb_cubic_contact_apply(gcmb_body* a, const gcmb_body* b, const int* boxA_min,
                             const int* boxB_min, const int* extent);

/* ---- one stage of the dimensional splitting (engine/cubic/GridCharacteristicMethod.hpp:42-87) -- */
/* computes the next time layer of all real nodes along reference direction `dir` and swaps the two
 * layers, like cubic/Engine.cpp:110-111 */
int gcmb_cubic_stage(gcmb_body* body, int dir, double tau);

/* gcmb_cubic_stage(body, dir, tau) followed by gcmb_cubic_border_apply(body, next_dir, n_values, values) on the new
 * layer, as ONE kernel when it can be (*fused = 1): next_dir is the contiguous axis, on each of its faces the last
 * registered condition covers the whole face with plain components, and dir's stage is a marching kernel -- the rows it
 * writes then leave with their ghost nodes, and the caller must NOT call border_apply for next_dir on this layer.
 * With *fused = 0 only the stage was done and the caller applies the border as usual (cubic/Engine.cpp:94-98). */
int gcmb_cubic_stage_fill_next_border(gcmb_body* body, int dir, double tau, int next_dir, int n_values,
                                      const double* values, int* fused);

/* gcmb_cubic_border_apply(body, dir, n_values, values) followed by gcmb_cubic_stage(body, dir, tau) -- the reference's
 * order for one direction of a body without contacts across it (cubic/Engine.cpp:94-111) -- in one call;
 * *fused = 1 when the stage kernel produced the ghost nodes itself (see gcmb_cubic_border_apply), 0 when the fill
 * kernel ran (masked faces, pressure-trace quantities, dense kernels, rows not longer than the border size). */
int gcmb_cubic_stage_with_border(gcmb_body* body, int dir, double tau, int n_values, const double* values, int* fused);

/* ---- Maxwell viscosity (rheology/ode/Ode.hpp:28-38): sigma *= decay[table of the node];
 * decay = exp(-tau/tau0) is evaluated by the caller with the host libm, like the reference ------- */
int gcmb_cubic_ode_maxwell(gcmb_body* body, const double* decay_per_table);

/* ---- seismogram taps (util/snapshot/SliceSnapshotter.hpp:44-82) ------------------------------- */
/* register the detector: nodes of the right border of the last axis inside the area */
int gcmb_cubic_detector_set_area(gcmb_body* body, int q_code, int area_kind, const double* params);
int gcmb_cubic_detector_set_mask(gcmb_body* body, int q_code, const uint8_t* face_mask);
/* sum of the quantity over the detector nodes and their number (the reference writes sum/count);
 * n_line values of component `line_comp` along the last axis through the centre node (sizes/2)
 * -> line (may be NULL).  sync. */
int gcmb_cubic_seismo(gcmb_body* body, double* sum, long long* count, int line_comp, double* line,
                      int n_line);
/* the same with the line through the given LOCAL node (D-1 entries: all axes but the last) -- slabs of a decomposed
 * grid pass the global centre translated into their own indices */
int gcmb_cubic_seismo_at(gcmb_body* body, double* sum, long long* count, int line_comp, double* line,
                         int n_line, const int* line_node);

/* The same taps without stalling the time loop: begin enqueues the detector reduction (over all slabs of a decomposed body:
 * an all-reduce on the stream), the line gather and ONE asynchronous copy into page-locked memory, and returns at once;
 * end (sync on that copy only) hands out the values.  line_node: D-1 local indices or NULL for no line; one in flight per body. */
int gcmb_cubic_seismo_begin(gcmb_body* body, int with_detector, int line_comp, const int* line_node);
int gcmb_cubic_seismo_end(gcmb_body* body, double* sum, long long* count, double* line, int n_line);

/* ---- multi-GPU: slab decomposition along x (the slowest axis), one process per GPU ------------- */
/* NCCL communicator shared by all slabs: rank 0 makes the id (128 bytes), the launcher distributes it */
int gcmb_comm_unique_id(void* id128);
int gcmb_comm_init(gcmb_ctx* ctx, int n_ranks, int rank, const void* id128);
/* exchange `border_size` x-planes with the slab neighbours (rank-1 on the left, rank+1 on the right,
 * none at the ends): my ghost planes <- neighbour's outermost real planes.  Equivalent to two
 * ContactCopiers between neighbouring slabs; call before the direction-0 stage. */
int gcmb_cubic_halo_exchange(gcmb_body* body);
/* the same for several bodies of one context (a decomposed multi-body task) in ONE NCCL group: one exchange in flight
 * for all of them, overlapped with the interior of every body's direction-0 stage */
int gcmb_halo_exchange_bodies(gcmb_body* const* bodies, int n);
/* Host-staged variant for transports that are not GPU-aware (the reference's own MPI_Sendrecv of raw
 * PdeVector bytes, src/test/TestMPI.cpp:31-51): get copies the outermost `border_size` REAL x-planes of
 * side (0 left, 1 right) into a host buffer of gcmb_cubic_halo_bytes(body) bytes, put writes such a buffer
 * into the GHOST planes of that side.  The buffer layout is opaque but identical for slabs of equal y/z
 * extent.  Both sync. */
size_t gcmb_cubic_halo_bytes(gcmb_body* body);
int gcmb_cubic_halo_get(gcmb_body* body, int side, void* host_buffer);
int gcmb_cubic_halo_put(gcmb_body* body, int side, const void* host_buffer);
/* in-place sum over all ranks of n doubles held on the host (detector sums of decomposed grids). sync */
int gcmb_comm_allreduce_sum(gcmb_ctx* ctx, double* host_values, int n);

/* ---- simplex (tetrahedral) path: engine/simplex/*, grid/simplex/* of the reference ------------------------
 * The triangulation is passed as flat arrays (what the reference's CGAL structure holds for the hot path):
 * xyz[nV][3]; cell_v[nC][4]; cell_n[nC][4] (neighbour i opposite vertex i, -1 outside the hull);
 * cell_grid[nC] (body id, -1 = empty space); the incident cells of every vertex (inc_off[nV+1], inc_cell,
 * ascending cell id).  A body is the set of cells with grid_id; its vertices get local indices in ascending
 * global id.  model: 0 elastic (M = 9), 1 acoustic (M = 4); isotropic materials (BorderCorrector.hpp:299-302). */
typedef struct gcmb_sbody gcmb_sbody;
int gcmb_simplex_body_create(gcmb_ctx* ctx, int model, int nV, int nC, const double* xyz, const int* cell_v,
                             const int* cell_n, const int* cell_grid, const int* inc_off, const int* inc_cell,
                             int grid_id, gcmb_sbody** out);
void gcmb_simplex_body_destroy(gcmb_sbody* body);
int gcmb_simplex_info(gcmb_sbody* body, int* n_local, int* M, int* n_border_vertices);
/* per local vertex (any output may be NULL): global id, border state (0 inner, 1 border, 2 contact,
 * 3 multicontact: SimplexGrid.hpp:385-393), border and common unit normals (SimplexGrid.hpp:141-160,427-444). sync */
int gcmb_simplex_vertices(gcmb_sbody* body, int* global_of, uint8_t* state, double* border_normal, double* common_normal);
/* SimplexGrid::findCellCrossedByTheRay (SimplexGrid.cpp:61-164) for nq (vertex, shift) pairs on the GPU:
 * out5[q] = {n, p0..p3}: n = 4 cell, 3/2/1 border facet/edge/vertex, 0 none; local vertex ids, -1 padded. sync */
int gcmb_simplex_locate(gcmb_sbody* body, int nq, const int* vertex, const double* shift, int* out5);
/* number of node computations in which the reference would have thrown (degenerate systems, failed asserts). sync */
int gcmb_simplex_errors(gcmb_sbody* body, int* count);
/* eigen-system in the calculation basis: U, U1 [3][M][M], L [3][M], basis row-major 3x3 (stage s runs along
 * column s; Task::calculationBasis) */
int gcmb_simplex_set_material(gcmb_sbody* body, const double* U, const double* U1, const double* L, const double* basis);
int gcmb_simplex_upload_state(gcmb_sbody* body, const double* pde /* [n_local][M] */);  /* sync */
int gcmb_simplex_download_state(gcmb_sbody* body, double* pde);                          /* sync */
/* border nodes as Engine::addBorderNode collects them (engine/simplex/Engine.cpp:288-309): condition types
 * (0 FIXED_FORCE, 1 FIXED_VELOCITY), and for every border node its local id, normal and condition */
int gcmb_simplex_border_set(gcmb_sbody* body, int n_cond, const int* types, int n_border, const int* node,
                            const double* normal, const int* cond_of_node);
/* Engine::applyPlainBorderContactCorrection (Engine.cpp:197-214); values [n_cond][outer] at the given time */
int gcmb_simplex_plain_border(gcmb_sbody* body, const double* values);
/* Engine::gcmStage (Engine.cpp:121-141), Riemann-invariant GCM, GLOBAL_BASIS, PRODUCT splitting: invariants,
 * gradients, border vertices, border correction with values [n_cond][outer] at t+tau, inner vertices, back to
 * PDE variables, swap */
int gcmb_simplex_stage(gcmb_sbody* body, int s, double tau, const double* values);
/* The same stage as the four calls simplex::Engine::gcmStage makes over ALL bodies
 * (engine/simplex/Engine.cpp:118-141; simplex::GridCharacteristicMethodBase, engine/simplex/common.hpp:16-41), so
 * that contacts between bodies are corrected in the reference's order:
 *   before_stage (every body) -> border_contact_stage (every body) -> contact_correct (every contact)
 *   -> border_correct (every body) -> inner_stage (every body) -> after_stage (every body; swaps the layers) */
/* GcmType of the task (util/task/Task.hpp:15-18,62): 0 ADVECT_RIEMANN_INVARIANTS (default,
 * engine/simplex/GridCharacteristicMethodInRiemannInvariants.hpp), 1 ADVECT_PDE_VECTORS (…InPdeVectors.hpp with
 * BorderCorrectorInPdeVectors / ContactCorrectorInPdeVectors) */
int gcmb_simplex_set_gcm_type(gcmb_sbody* body, int gcm_type);
/* SplittingType of the task (util/task/Task.hpp:19-22,64; engine/simplex/Engine.hpp:135-160): 0 PRODUCT -- every stage
 * starts from the result of the previous one (default); 1 SUMM -- every stage starts from the same layer and fills a
 * layer of its own, gcmb_simplex_average_layers then averages them into the current layer
 * (DefaultMesh::averageNewPdeLayersToCurrent, engine/simplex/DefaultMesh.hpp:160-169; Engine.cpp:104-108) */
/* MaxwellViscosityOde::apply on a simplex mesh (rheology/ode/Ode.hpp:28-38, engine/simplex/Engine.cpp:110-114):
 * stress components *= decay, decay = exp(-tau / tau0) evaluated by the caller's libm like the reference's */
int gcmb_simplex_ode_maxwell(gcmb_sbody* body, double decay);
int gcmb_simplex_set_splitting(gcmb_sbody* body, int splitting);
/* BorderCalcMode::LOCAL_BASIS (util/task/Task.hpp:12-14; engine/simplex/DefaultMesh.hpp:245-266): the listed border and
 * contact vertices get eigen-systems of their own, U/U1 [n][3][M][M] and L [n][3][M], written in the bases [n][9]
 * (row-major; stage s runs along column s, the first column is the vertex' normal).  From then on
 * gcmb_simplex_border_correct / contact_correct act as applyInLocalBasis and only at stage 0
 * (Engine::correctContactsAndBorders, engine/simplex/Engine.cpp:170-187).  n = 0 returns to GLOBAL_BASIS. */
int gcmb_simplex_set_local_bases(gcmb_sbody* body, int n, const int* nodes, const double* U, const double* U1,
                                 const double* L, const double* bases);
int gcmb_simplex_average_layers(gcmb_sbody* body);
int gcmb_simplex_before_stage(gcmb_sbody* body, int s, double tau);
int gcmb_simplex_border_contact_stage(gcmb_sbody* body);
int gcmb_simplex_border_correct(gcmb_sbody* body, const double* values /* [n_cond][outer] at t+tau */);
int gcmb_simplex_inner_stage(gcmb_sbody* body);
int gcmb_simplex_after_stage(gcmb_sbody* body);
/* SimplexGrid::contactNormal(it, neighborId) of every local vertex (grid/simplex/SimplexGrid.hpp:141-144):
 * unit normal over the faces shared with body `neighbor_grid_id`, zero where there is none. sync */
int gcmb_simplex_contact_normals(gcmb_sbody* body, int neighbor_grid_id, double* normals /* [n_local][3] */);
/* A contact of two bodies as Engine::addContactNode collects it (Engine.cpp:273-285): pairs of local vertex ids
 * and the normal from a to b.  Elastic bodies are glued (ADHESION), acoustic bodies slide (SLIDE) — the two
 * combinations the reference's ContactCorrectorFactory offers (ContactCorrector.hpp:484-560). */
typedef struct gcmb_scontact gcmb_scontact;
int gcmb_simplex_contact_create(gcmb_sbody* a, gcmb_sbody* b, int n, const int* node_a, const int* node_b,
                                const double* normals, gcmb_scontact** out);
void gcmb_simplex_contact_destroy(gcmb_scontact* contact);
/* AbstractContactCorrector::applyPlainCorrection on the current layers (ContactCorrector.hpp:256-269) */
int gcmb_simplex_contact_plain(gcmb_scontact* contact);
/* ContactCorrectorInRiemannInvariants::applyInGlobalBasis on the stage in flight (ContactCorrector.hpp:334-356) */
int gcmb_simplex_contact_correct(gcmb_scontact* contact);
/* Differentiation::estimateGradient of a host field [n_local][M] -> [n_local][3][M] (test hook). sync */
int gcmb_simplex_gradient(gcmb_sbody* body, const double* values, double* grad);

/* ---- TriangleInterpolator (util/math/interpolation/TriangleInterpolator.hpp:8-130), the 2-D member of the simplex
 * interpolators, for n independent scalar queries on the GPU (test hook; the 2-D simplex engine itself is not built).
 * mode 0: interpolate, linear (points [n][3][2], values [n][3]); 1: interpolate, quadratic (+ gradients [n][3][2]);
 * 2: minMaxInterpolate; 3: hybridInterpolate; 4: interpolateInOwner (points [n][4][2], values [n][4], no gradients).
 * queries [n][2] -> out [n]; status[n]: 0, or 1 where the reference throws (degenerate triangle, query outside). sync */
int gcmb_triangle_interpolate(gcmb_ctx* ctx, int mode, int n, const double* points, const double* values,
                              const double* gradients, const double* queries, double* out, int* status);

/* ---- checksum of the current layer over real nodes: sum_nodes sum_i (i+1)*u_i (sync) ---------- */
int gcmb_cubic_checksum(gcmb_body* body, double* out);

/* description of the stage kernel LAST LAUNCHED for a direction (for logs/tests), e.g. "sparse:elastic3d_iso_x/bs2",
 * "sparse:acoustic2d_x/bs2+k0" (foot cells read from the table: Courant number >= 1), "dense_k0_one:M9/bs2";
 * "unset" before the first stage of that direction */
const char* gcmb_cubic_stage_kernel_name(gcmb_body* body, int dir);

#ifdef __cplusplus
}
#endif
#endif /* GCM_B200_H */
