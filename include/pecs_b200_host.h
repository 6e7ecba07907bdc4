/* pecs_b200_host.h -- C bindings of the host-side mirror of the reference's class API.
 *
 * pecs_b200.h is the device boundary (what a deal.II host would bind).  This header exposes the C++ host classes
 * of pecs_b200/csrc/host (SOLARCELL::SolarCellProblem and the tables it owns) to non-C++ callers -- the Python
 * package, the parity tests and bench.py -- with the reference's method names.  Everything here is one-time
 * setup, table export or post-processing; the per-step work goes through pecs_b200.h on the pecs_ctx that
 * pecs_solarcell_ctx() returns.
 */
#ifndef PECS_B200_HOST_H
#define PECS_B200_HOST_H

#include "pecs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pecs_solarcell pecs_solarcell;

/* main(): ParameterReader::read_parameters + SolarCellProblem<2>(1, prm)  (reference source/main/main.cpp:5-24).
 * prm_text: contents of input_file.prm (NULL or "" -> declared defaults); test_defaults != 0 declares the
 * test_file.prm defaults instead (reference source/ParameterReader.cpp:209-392). */
pecs_status pecs_solarcell_create(const char* prm_text, int32_t test_defaults, int32_t device, pecs_solarcell** out);
void pecs_solarcell_destroy(pecs_solarcell* p);
/* before setup_*: which carriers this process' context factorises and solves (pecs_problem_desc::owned_species) */
pecs_status pecs_solarcell_set_owned_species(pecs_solarcell* p, int32_t mask);

/* staged setup: *_host builds grids, dofs, mappings and the constant matrices on the host (no device needed);
 * the variants without _host additionally create the device context (set_solvers), upload the initial state and,
 * for the production problem, do the initial Poisson solve (reference source/SolarCell.cpp:1898-2034). */
pecs_status pecs_solarcell_setup_full_system_host(pecs_solarcell* p);
pecs_status pecs_solarcell_setup_full_system(pecs_solarcell* p);
pecs_status pecs_solarcell_setup_test_host(pecs_solarcell* p, int32_t kind, int32_t n_refine);
pecs_status pecs_solarcell_setup_test(pecs_solarcell* p, int32_t kind, int32_t n_refine);
/* the reference's public entry points */
pecs_status pecs_solarcell_run_full_system(pecs_solarcell* p);
/* output path: where print_results / run_full_system put the .vtu files (default "."), whether run_full_system writes
 * them at every time stamp (default yes, as the reference does: source/SolarCell.cpp:2037-2087) */
pecs_status pecs_solarcell_set_output(pecs_solarcell* p, const char* directory, int32_t write_output);
/* print_results(time_step_number) of the reference (source/SolarCell.cpp:1826-1858): Poisson-NNN.vtu,
 * Semiconductor-NNN.vtu, Electrolyte-NNN.vtu, rescaled to physical units.  Returns once the snapshot is enqueued; the
 * files are complete after pecs_solarcell_finish_output. */
pecs_status pecs_solarcell_print_results(pecs_solarcell* p, int32_t time_step_number);
pecs_status pecs_solarcell_finish_output(pecs_solarcell* p);
/* host half of the output path alone (no device): write the file of mesh `which` (0 Semiconductor-, 1 Electrolyte-,
 * 2 Poisson-) for caller-provided patch values in the pecs_output_snapshot layout */
pecs_status pecs_solarcell_write_patches(pecs_solarcell* p, int32_t which, const double* patches, int32_t time_step_number,
                                         const char* directory);
/* I-V post-processing: the two charge-transfer currents through the interface in scaled units,
 * out = { int k_et (rho_n - rho_n^e) rho_o ds, int k_ht (rho_p - rho_p^e) rho_r ds }, from host state vectors
 * states[PECS_ELECTRONS..PECS_OXIDANTS] or, with states == NULL, from the current device state */
pecs_status pecs_solarcell_interface_currents(pecs_solarcell* p, const double* const states[4], double out[2]);
/* the four PostProcessor scales {potential, field, density, current} (reference source/PostProcessor.cpp:14-18) */
pecs_status pecs_solarcell_output_scales(const pecs_solarcell* p, double scales[4]);
/* test_steady_state / test_transient / test_DD_Poisson at one refinement level; errors[4] = {u, J, Phi, D} */
pecs_status pecs_solarcell_run_test(pecs_solarcell* p, int32_t kind, int32_t n_refine, double errors[4]);

pecs_ctx* pecs_solarcell_ctx(pecs_solarcell* p);
pecs_status pecs_solarcell_get_params(const pecs_solarcell* p, double params[32]);
double pecs_solarcell_delta_t(const pecs_solarcell* p);

/* mesh tables; which: 0 semiconductor, 1 electrolyte, 2 Poisson */
int32_t pecs_solarcell_n_cells(const pecs_solarcell* p, int32_t which);
pecs_status pecs_solarcell_get_mesh(const pecs_solarcell* p, int32_t which, double* vertices, int32_t* material_id,
                                    int32_t* level, int32_t* face_kind, int32_t* neighbor, int32_t* neighbor2,
                                    int32_t* boundary_id, double* nb_parent_diameter);
/* Poisson dofs and constraints */
int32_t pecs_solarcell_n_rt(const pecs_solarcell* p);
pecs_status pecs_solarcell_get_poisson_face_dofs(const pecs_solarcell* p, int32_t* face_dof);
int32_t pecs_solarcell_n_constraints(const pecs_solarcell* p);
pecs_status pecs_solarcell_get_constraints(const pecs_solarcell* p, int32_t* dof, int32_t* master, double* weight);
/* cell maps (which: 0 s_2_p, 1 e_2_p) and interface pairs */
pecs_status pecs_solarcell_get_cell_map(const pecs_solarcell* p, int32_t which, int32_t* map);
int32_t pecs_solarcell_n_interface_pairs(const pecs_solarcell* p);
pecs_status pecs_solarcell_get_interface_pairs(const pecs_solarcell* p, int32_t* semi_cell, int32_t* semi_face,
                                               int32_t* elec_cell, int32_t* elec_face);
/* constant matrices; which: 0..3 species system matrix, 4 Poisson, 5 semiconductor mass, 6 electrolyte mass */
int64_t pecs_solarcell_matrix_nnz(const pecs_solarcell* p, int32_t which);
pecs_status pecs_solarcell_get_matrix(const pecs_solarcell* p, int32_t which, int32_t* row_ptr, int32_t* col, double* val);

/* initial conditions computed on the host into the mirrors (and pushed to the device if a context exists) */
pecs_status pecs_solarcell_project_initial_conditions(pecs_solarcell* p);
pecs_status pecs_solarcell_project_test_initial_condition(pecs_solarcell* p);
/* host mirror access: which 0..3 */
pecs_status pecs_solarcell_get_host_solution(const pecs_solarcell* p, int32_t which, double* out);

/* post-processing of the manufactured tests */
pecs_status pecs_solarcell_ldg_errors(pecs_solarcell* p, int32_t which, double time, double out[2]);
pecs_status pecs_solarcell_mixed_errors(pecs_solarcell* p, double out[2]);

/* ---- verification of the setup tables (CPU tests only; not reachable from any per-step entry point) ----
 * Builds the nested-dissection plan of one constant matrix exactly as pecs_ctx_create does and reports its size:
 * stats[0] fronts, [1] levels, [2] max np, [3] max nb, [4] forward-table entries, [5] backward-table entries,
 * [6] update-buffer entries.  which: 0..3 species, 4 Poisson. */
pecs_status pecs_solarcell_plan_stats(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int64_t stats[8]);
/* per-level breakdown of the same plan: for level d (0 = root) out[6*d..6*d+5] = fronts, forward entries, backward
 * entries, max np, max nb, sum of np; returns the number of levels (<= max_levels written) */
int32_t pecs_solarcell_plan_levels(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int64_t* out, int32_t max_levels);
/* per-front listing of the same plan: out[8*f..8*f+7] = depth, np, nb, log2 of the forward / backward panel height,
 * forward / backward "one warp per front" flags, parent; returns the number of fronts (<= max_fronts written) */
int64_t pecs_solarcell_plan_fronts(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int32_t* out, int64_t max_fronts);
#ifdef __cplusplus
}
#endif
#endif /* PECS_B200_HOST_H */
