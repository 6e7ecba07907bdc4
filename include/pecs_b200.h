/* pecs_b200.h -- C ABI of the B200-native PECS per-step IMEX path.
 *
 * The reference (mdh266/PECS) has no plugin / FFI layer: its seam is the five argument-less private methods the
 * time loop calls (reference source/SolarCell.cpp:2057-2075), which talk through public members of the
 * Carrier / CarrierPair / PoissonData structs.  This header is the C boundary a host that owns
 * those structs (deal.II in the reference; pecs_b200/csrc/host in this repository) binds instead:
 *
 *   one-time  : pecs_ctx_create()  <- everything the reference has after setup_dofs / setup_mappings /
 *               assemble_Poisson_matrix / assemble_LDG_system (reference source/SolarCell.cpp:1933-1959).
 *               The LU factorisation the reference does in set_solvers (source/SolarCell.cpp:1964,
 *               source/Carrier.cpp:26-32, source/Poisson.cpp:92-96) happens inside.
 *   per step  : the five calls below, 1:1 with the reference methods, or pecs_step() for all five, N times.
 *
 * Conventions: plain pointers and sizes, caller-owned host buffers (copied during the call), context-owned
 * device memory, fp64 everywhere, int32 indices.  Vectors use the reference's component-wise block layout:
 * carriers [Jx | Jy | rho] with 4 nodal values per cell per block (source/CarrierPair.cpp:29-33), Poisson
 * [RT0 fluxes | potentials] (source/Poisson.cpp:25-28).  Every function returns a pecs_status; no exception
 * crosses the boundary; pecs_last_error() gives the message of the calling thread's last failure.
 * A context is not re-entrant; different contexts (e.g. one per applied bias) are independent.
 * There is NO CPU fallback: without a CUDA device pecs_ctx_create fails with PECS_ERR_NO_DEVICE.
 */
#ifndef PECS_B200_H
#define PECS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PECS_OK = 0,
  PECS_ERR_INVALID = 1,   /* bad argument / inconsistent tables */
  PECS_ERR_NO_DEVICE = 2, /* no CUDA device: there is no CPU fallback */
  PECS_ERR_CUDA = 3,      /* a CUDA runtime call or kernel failed */
  PECS_ERR_SINGULAR = 4,  /* a pivot block of a fixed matrix could not be inverted */
  PECS_ERR_INTERNAL = 5
} pecs_status;

/* boundary ids, reference include/Grid.hpp:141-147 */
enum { PECS_INTERFACE = 0, PECS_DIRICHLET = 1, PECS_NEUMANN = 2, PECS_SCHOTTKY = 3 };

/* species / vector selectors */
enum { PECS_ELECTRONS = 0, PECS_HOLES = 1, PECS_REDUCTANTS = 2, PECS_OXIDANTS = 3, PECS_POISSON = 4 };

/* which right-hand-side functors the kernels evaluate: the production ones (reference
 * source/SolarCell.cpp:1073-1726, 487-815) or the manufactured-solution ones of the reference's convergence
 * tests (source/LDG.cpp:681-982, source/MixedFEM.cpp:166-254, source/SolarCell.cpp:2105-2330). */
enum {
  PECS_KIND_PRODUCTION = 0,
  PECS_KIND_TEST_STEADY = 1,    /* tests/Poisson_test.cpp   -> test_steady_state  */
  PECS_KIND_TEST_TRANSIENT = 2, /* tests/IMEX_LDG_test.cpp  -> test_transient     */
  PECS_KIND_TEST_DD_POISSON = 3 /* tests/DD_Poisson_test.cpp-> test_DD_Poisson    */
};

/* scaled scalar parameters, reference include/Parameters.hpp:181-242 and the constants of
 * source/InitialConditions.cpp:20,43,61,79; indices into pecs_problem_desc.params */
enum {
  PECS_P_DELTA_T = 0,
  PECS_P_PENALTY,      /* tau, reference source/SolarCell.cpp:1944-1945 */
  PECS_P_MU_N, PECS_P_MU_P, PECS_P_MU_R, PECS_P_MU_O,
  PECS_P_EPS_S, PECS_P_EPS_E,
  PECS_P_LAMBDA2,      /* scaled_debeye_length */
  PECS_P_K_ET, PECS_P_K_HT,
  PECS_P_V_N, PECS_P_V_P, /* scaled recombination velocities (Schottky) */
  PECS_P_GEN_FLUX, PECS_P_GEN_ALPHA, PECS_P_GEN_LOCATION, /* Generation, reference source/Generation.cpp:14-44 */
  PECS_P_RHO_N_E, PECS_P_RHO_P_E, PECS_P_RHO_R_E, PECS_P_RHO_O_E,
  PECS_P_PHI_BI, PECS_P_PHI_APP, PECS_P_PHI_SCH, PECS_P_SCH_LOCATION,
  PECS_P_TRANSIENT,    /* transient_or_steady of assemble_LDG_system */
  /* Shockley-Read-Hall recombination R(rho_n, rho_p) (reference include/SolarCell.hpp:86-98; the reference's function
     returns 0.0 with the formula commented out, which is PECS_P_SRH == 0, the default):
     R = (n_i^2 - rho_n rho_p) / (tau_n (rho_n - n_i) + tau_p (rho_p - n_i)), added to both semiconductor carriers */
  PECS_P_SRH, PECS_P_N_INTRINSIC, PECS_P_TAU_N, PECS_P_TAU_P,
  PECS_P_COUNT
};

typedef struct {
  int32_t n;              /* rows = cols */
  const int32_t* row_ptr; /* n+1 */
  const int32_t* col;     /* nnz, sorted within a row */
  const double* val;      /* nnz */
} pecs_csr;

/* one carrier subdomain = one reference triangulation + DoFHandler + CarrierPair
 * (reference include/SolarCell.hpp:351-367, include/CarrierPair.hpp:72-108) */
typedef struct {
  int32_t n_cells;
  const double* vertices;        /* [n_cells][4][2], deal.II lexicographic vertex order */
  const int32_t* poisson_cell;   /* [n_cells] matched Poisson cell: s_2_p_map / e_2_p_map, SolarCell.cpp:308-369 */
  int32_t n_boundary_faces;      /* faces with at_boundary() */
  const int32_t* bface_cell;     /* [n_boundary_faces] */
  const int32_t* bface_face;     /* [n_boundary_faces] local face number 0..3 */
  const int32_t* bface_id;       /* [n_boundary_faces] boundary id */
  pecs_csr system_matrix[2];     /* Carrier::system_matrix of carrier_1, carrier_2 (12*n_cells rows) */
} pecs_domain_desc;

/* PoissonData (reference include/Poisson.hpp:44-83) on the Poisson triangulation */
typedef struct {
  int32_t n_cells;
  const double* vertices;        /* [n_cells][4][2] */
  int32_t n_rt;                  /* number of RT0 flux dofs; potential of cell c is dof n_rt + c */
  const int32_t* face_dof;       /* [n_cells][4] global flux dof of each face */
  int32_t n_boundary_faces;
  const int32_t* bface_cell;
  const int32_t* bface_face;
  const int32_t* bface_id;
  pecs_csr system_matrix;        /* constraint-condensed, as assembled by SolarCell.cpp:377-417 */
  int32_t n_constraints;         /* ConstraintMatrix lines, reference source/Poisson.cpp:33-49 */
  const int32_t* constraint_dof;
  const int32_t* constraint_master; /* -1: dof = 0 */
  const double* constraint_weight;
} pecs_poisson_desc;

/* matched interface faces, reference source/SolarCell.cpp:165-262 */
typedef struct {
  int32_t n_pairs;
  const int32_t* semi_cell;
  const int32_t* semi_face;
  const int32_t* elec_cell;
  const int32_t* elec_face;
} pecs_interface_desc;

typedef struct {
  int32_t kind;         /* PECS_KIND_* */
  int32_t full_system;  /* 0: semiconductor + Poisson only (the manufactured tests), 1: both subdomains */
  int32_t device;       /* CUDA device ordinal */
  int32_t owned_species; /* bit k set: carrier k (PECS_ELECTRONS..PECS_OXIDANTS) is factorised and solved by this
                          * context; 0 = all four.  A context that owns a subset is one shard of a step spread over
                          * several GPUs (pecs_step_local / pecs_step_finish below); the others' densities arrive from
                          * their owners through pecs_density_block */
  double params[32];    /* PECS_P_* */
  pecs_domain_desc semiconductor;
  pecs_domain_desc electrolyte; /* ignored unless full_system */
  pecs_poisson_desc poisson;
  pecs_interface_desc interface_pairs;
} pecs_problem_desc;

typedef struct pecs_ctx pecs_ctx;

const char* pecs_last_error(void);
/* number of CUDA devices visible to the library (0 when there is none or the driver is missing) */
int32_t pecs_device_count(void);
/* Optional: pays the one-time costs of a device's first pecs_ctx_create ahead of it -- CUDA context, the library's kernel
 * image, the cuSOLVER / cuBLAS handles of the setup factorisation -- so that a caller can overlap them with its own host
 * work (SolarCellProblem::setup_full_system calls it on a second thread while it builds grids and matrices).  Has no
 * counterpart in the reference and changes no result; pecs_ctx_create does the same work itself when it was not called. */
pecs_status pecs_device_warmup(int32_t device);

/* replaces: end of setup + SolarCellProblem::set_solvers, reference source/SolarCell.cpp:1733-1747 */
pecs_status pecs_ctx_create(const pecs_problem_desc* desc, pecs_ctx** out);
void pecs_ctx_destroy(pecs_ctx* ctx);

/* host <-> device copies of Carrier::solution / PoissonData::solution / ::system_rhs; which = PECS_ELECTRONS..PECS_POISSON */
pecs_status pecs_set_state(pecs_ctx* ctx, int32_t which, const double* solution);
pecs_status pecs_get_state(pecs_ctx* ctx, int32_t which, double* solution);
pecs_status pecs_get_rhs(pecs_ctx* ctx, int32_t which, double* system_rhs);
pecs_status pecs_set_rhs(pecs_ctx* ctx, int32_t which, const double* system_rhs);
int32_t pecs_n_dofs(const pecs_ctx* ctx, int32_t which);

/* time of the manufactured right-hand sides (Function::set_time in the reference's tests) */
pecs_status pecs_set_time(pecs_ctx* ctx, double time);

/* replaces SolarCellProblem::assemble_semiconductor_rhs, reference source/SolarCell.cpp:1037-1414 */
pecs_status pecs_assemble_semiconductor_rhs(pecs_ctx* ctx);
/* replaces SolarCellProblem::assemble_electrolyte_rhs, reference source/SolarCell.cpp:1417-1726 */
pecs_status pecs_assemble_electrolyte_rhs(pecs_ctx* ctx);
/* replaces SolarCellProblem::solve_full_system -> Carrier::solve x4, source/SolarCell.cpp:1758-1782, source/Carrier.cpp:34-40 */
pecs_status pecs_solve_full_system(pecs_ctx* ctx);
/* one species only (the manufactured tests call carrier_1.solve() directly, source/SolarCell.cpp:2934, 3077) */
pecs_status pecs_solve_species(pecs_ctx* ctx, int32_t which);
/* replaces SolarCellProblem::assemble_Poisson_rhs, reference source/SolarCell.cpp:430-815 */
pecs_status pecs_assemble_poisson_rhs(pecs_ctx* ctx);
/* replaces SolarCellProblem::solve_Poisson -> PoissonData::solve, source/SolarCell.cpp:1750-1756, source/Poisson.cpp:98-105 */
pecs_status pecs_solve_poisson(pecs_ctx* ctx);
/* n_steps iterations of the body of the time loop (reference source/SolarCell.cpp:2055-2080), captured in a CUDA graph */
pecs_status pecs_step(pecs_ctx* ctx, int32_t n_steps);
/* One step cut at its only exchange point, for contexts that own a subset of the carriers (owned_species):
 *   pecs_step_local : both RHS assemblies of the subdomains with owned carriers + the owned solves   (graph replay)
 *   -- the caller exchanges the density blocks (pecs_density_block) between the owners, on pecs_stream --
 *   pecs_step_finish: Poisson RHS + Poisson solve (every shard keeps its own copy of the potential)   (graph replay)
 * The reference's decomposition is the same: separate triangulations and CarrierPairs per subdomain
 * (include/SolarCell.hpp:351-367), four independent solve tasks (source/SolarCell.cpp:1763-1781). */
pecs_status pecs_step_local(pecs_ctx* ctx);
pecs_status pecs_step_finish(pecs_ctx* ctx);
/* The same exchange WITHOUT a collective: fused into the solves over peer memory (NVLink / NVSwitch).
 *   pecs_p2p_export : fills `blob` with CUDA IPC handles of this context's state vectors and flag block; returns the
 *                     number of bytes written (-1 on error).  Every rank passes its blob to every other rank
 *                     (any transport; pecs_b200/shard.py uses torch.distributed.all_gather_object, once, at setup).
 *   pecs_p2p_connect: blobs = the `world` blobs in rank order.  From then on the backward sweep of every owned
 *                     carrier stores each finished density straight into the other ranks' vectors while it runs,
 *                     single-thread flag kernels order it against the peers' assembly ("rank r no longer reads the old
 *                     densities" before the first remote store, "carrier s is complete" before the Poisson assembly),
 *                     and pecs_step_local / pecs_step_finish need nothing in between.  Ranks must step in lockstep. */
int64_t pecs_p2p_export(pecs_ctx* ctx, void* blob, int64_t capacity);
pecs_status pecs_p2p_connect(pecs_ctx* ctx, int32_t rank, int32_t world, const void* blobs);
/* device pointer to the density block of carrier `which` (4 * n_cells doubles, *n_doubles receives the count) */
double* pecs_density_block(pecs_ctx* ctx, int32_t which, int64_t* n_doubles);
/* the context's main CUDA stream (cudaStream_t): everything above is enqueued on it */
void* pecs_stream(pecs_ctx* ctx);
/* all calls above return after enqueueing; this waits for the context's streams */
pecs_status pecs_synchronize(pecs_ctx* ctx);

/* The same n_steps with HOST-resident state, as a host that keeps Carrier::solution / PoissonData::solution in its
 * own memory would call it (states[PECS_ELECTRONS..PECS_POISSON], NULL entries are skipped).  Uploads what a step
 * reads of the caller's state -- the density block of every carrier and the Poisson vector; the LDG currents are
 * outputs only (the assembly reads densities, reference source/SolarCell.cpp:1146-1193; Carrier::solve overwrites the
 * whole solution, source/Carrier.cpp:34-40) -- runs the steps, downloads all five vectors and waits.  With pinned
 * buffers (pecs_host_alloc) every species is downloaded on its own stream as soon as its solve has finished, while the
 * other solves and the Poisson part still run; pageable buffers are copied after the last step. */
pecs_status pecs_step_host(pecs_ctx* ctx, int32_t n_steps, double* const states[5]);
/* page-locked host memory for the above */
void* pecs_host_alloc(uint64_t bytes);
void pecs_host_free(void* p);

/* ---- output path: what the reference's print_results does per time stamp (source/SolarCell.cpp:1826-1858):
 * LDG::output_rescaled_results (source/LDG.cpp:1195-1232) and MixedFEM::output_rescaled_results
 * (source/MixedFEM.cpp:297-320) run DataOut::build_patches over every cell and PostProcessor
 * (source/PostProcessor.cpp:80-123) rescales the values to physical units.  Here the rescaled patch values (one patch of
 * 4 vertices per cell, deal.II vertex order) are produced on the device in the layout of the VTU data arrays and copied
 * to the caller's buffers on a separate stream, so the time loop does not stall on output.
 *   host[0] semiconductor pair, host[1] electrolyte pair: n = cells, doubles
 *       current_1 [4n][3] | density_1 [4n] | current_2 [4n][3] | density_2 [4n]        (pecs_output_doubles = 32 n)
 *   host[2] Poisson: field [4n][3] | potential [4n]                                     (16 n)
 * scales = {potential, field, density (unused, as in the reference), current}.  NULL entries of host are skipped.
 * pecs_output_snapshot returns after enqueueing: the values are those of the state after the steps enqueued so far;
 * the buffers (page-locked: pecs_host_alloc, for the copy to be asynchronous) are complete after pecs_output_wait. */
int64_t pecs_output_doubles(const pecs_ctx* ctx, int32_t which);
pecs_status pecs_output_snapshot(pecs_ctx* ctx, const double scales[4], double* const host[3]);
pecs_status pecs_output_wait(pecs_ctx* ctx);

/* I-V post-processing on the device: the two charge-transfer currents through the
 * semiconductor-electrolyte interface in scaled units, integrals of the step's interface terms (reference
 * source/SolarCell.cpp:1265-1347) over the current device state:
 *   out = { int k_et (rho_n - rho_n^e) rho_o ds,  int k_ht (rho_p - rho_p^e) rho_r ds }.
 * One small kernel over the interface cells + an ordered sum of the per-cell values; synchronises the context. */
pecs_status pecs_interface_currents(pecs_ctx* ctx, double out[2]);

/* measurement support for bench.py: run n_steps and report device times measured with CUDA events on the
 * context's own streams.  ms[0] = whole region.  sectioned == 0: n_steps replays of the step graph; sectioned == 1:
 * ms[1..5] = the reference's five TimerOutput sections summed over the steps, launched one by one
 * (serialised, includes launch gaps); sectioned == 2: n_steps replays of a graph that holds only the five solves
 * (four concurrent carrier solves, then Poisson); sectioned == 3: n_steps replays of a graph that holds only the
 * assembly passes (fused carrier RHS, Poisson RHS).  In the step graph the assembly passes run strictly before / between
 * the solves, so (mode 0) - (mode 3) is the device time of the solves inside a real step: the denominator of the solve
 * roofline. */
pecs_status pecs_step_timed(pecs_ctx* ctx, int32_t n_steps, int32_t sectioned, double ms[6]);
/* repeat one kernel class in isolation: which = 0 carrier RHS (both subdomains), 1 Poisson RHS, 2 carrier solves,
 * 3 Poisson solve; returns average ms per launch group and the number of kernel launches in one group */
pecs_status pecs_time_kernel(pecs_ctx* ctx, int32_t which, int32_t repeats, double* avg_ms, int32_t* launches);

/* integer facts about the context: see PECS_INFO_* */
enum {
  PECS_INFO_LAUNCHES_PER_STEP = 0,
  PECS_INFO_FACTOR_BYTES = 1,      /* bytes of all factor tables resident in HBM */
  PECS_INFO_SOLVE_BYTES_PER_STEP = 2, /* factor bytes streamed by the five solves of one step */
  PECS_INFO_TREE_LEVELS_MAX = 3,
  PECS_INFO_RHS_BYTES_PER_STEP = 4, /* algorithmic bytes of the carrier + Poisson RHS kernels of one step */
  PECS_INFO_HOST_STEP_H2D_BYTES = 5, /* bytes pecs_step_host uploads per call */
  PECS_INFO_HOST_STEP_D2H_BYTES = 6, /* bytes pecs_step_host downloads per call */
  PECS_INFO_SOLVE_WAIT_ERRORS = 7,   /* systems whose level kernels ever gave up waiting for a front (must be 0; synchronises) */
  PECS_INFO_SHARED_FACTOR_PAIRS = 8  /* carrier pairs whose two identical matrices share one factorisation */
};
int64_t pecs_get_info(const pecs_ctx* ctx, int32_t what);

#ifdef __cplusplus
}
#endif
#endif /* PECS_B200_H */
