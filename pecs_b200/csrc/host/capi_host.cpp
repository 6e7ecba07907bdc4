// capi_host.cpp -- extern "C" bindings of the host classes (include/pecs_b200_host.h).
#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../../include/pecs_b200_host.h"
#include "../error.hpp"
#include "capi_internal.hpp"
#include "SolarCell.hpp"
#include "SolverSetup.hpp"

namespace {
using SOLARCELL::SolarCellProblem;
using pecs::capi::guarded;
using pecs::capi::tria;
ChargeCarrierSpace::Carrier& carrier(const pecs_solarcell* p, int which) {
  SolarCellProblem& s = *p->problem;
  switch (which) {
    case PECS_ELECTRONS: return s.electron_hole_pair.carrier_1;
    case PECS_HOLES: return s.electron_hole_pair.carrier_2;
    case PECS_REDUCTANTS: return s.redox_pair.carrier_1;
    case PECS_OXIDANTS: return s.redox_pair.carrier_2;
  }
  throw pecs::StatusError(PECS_ERR_INVALID, "species selector must be 0..3");
}
const pecs::CsrMatrix& matrix(const pecs_solarcell* p, int which) {
  if (which >= 0 && which <= 3) return carrier(p, which).system_matrix;
  if (which == 4) return p->problem->Poisson_object.system_matrix;
  if (which == 5) return p->problem->electron_hole_pair.mass_matrix;
  if (which == 6) return p->problem->redox_pair.mass_matrix;
  throw pecs::StatusError(PECS_ERR_INVALID, "matrix selector must be 0..6");
}
template <class T>
void copy_out(const std::vector<T>& v, T* out) {
  if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
}
} // namespace

extern "C" {

pecs_status pecs_solarcell_create(const char* prm_text, int32_t test_defaults, int32_t device, pecs_solarcell** out) {
  return guarded([&] {
    if (!out) throw pecs::StatusError(PECS_ERR_INVALID, "out is NULL");
    std::unique_ptr<pecs_solarcell> p(new pecs_solarcell());
    ParameterSpace::ParameterReader reader(p->prm);
    if (test_defaults)
      reader.declare_test_parameters();
    else
      reader.declare_parameters();
    if (prm_text && *prm_text) p->prm.read_input_from_string(prm_text);
    p->problem.reset(new SolarCellProblem(1, p->prm));
    p->problem->device = device;
    *out = p.release();
  });
}
void pecs_solarcell_destroy(pecs_solarcell* p) { delete p; }
pecs_status pecs_solarcell_set_owned_species(pecs_solarcell* p, int32_t mask) {
  return guarded([&] {
    if (mask < 0 || mask > 0xF) throw pecs::StatusError(PECS_ERR_INVALID, "owned species mask must be 0..15");
    p->problem->owned_species = mask;
  });
}

#define PECS_FORWARD(name, expr) \
  pecs_status name(pecs_solarcell* p) { return guarded([&] { expr; }); }
PECS_FORWARD(pecs_solarcell_setup_full_system_host, p->problem->setup_full_system_host())
PECS_FORWARD(pecs_solarcell_setup_full_system, p->problem->setup_full_system())
PECS_FORWARD(pecs_solarcell_run_full_system, p->problem->run_full_system())

PECS_FORWARD(pecs_solarcell_finish_output, p->problem->finish_output())
pecs_status pecs_solarcell_set_output(pecs_solarcell* p, const char* directory, int32_t write_output) {
  return guarded([&] {
    if (directory && *directory) p->problem->output_directory = directory;
    p->problem->write_output = write_output != 0;
  });
}
pecs_status pecs_solarcell_print_results(pecs_solarcell* p, int32_t time_step_number) {
  return guarded([&] {
    if (time_step_number < 0) throw pecs::StatusError(PECS_ERR_INVALID, "time_step_number must be >= 0");
    p->problem->print_results((unsigned)time_step_number);
  });
}
pecs_status pecs_solarcell_write_patches(pecs_solarcell* p, int32_t which, const double* patches, int32_t time_step_number,
                                         const char* directory) {
  return guarded([&] {
    if (!patches || time_step_number < 0) throw pecs::StatusError(PECS_ERR_INVALID, "write_patches: bad argument");
    SolarCellProblem& s = *p->problem;
    if (which < 0 || which > 2) throw pecs::StatusError(PECS_ERR_INVALID, "write_patches: which must be 0..2");
    if (!p->patch_mesh[which] || p->patch_mesh[which]->n_cells() != tria(p, which).tables().n_cells)
      p->patch_mesh[which].reset(new pecs::VtuMesh(tria(p, which).tables()));
    const pecs::VtuMesh& mesh = *p->patch_mesh[which];
    const std::string dir = directory && *directory ? directory : ".";
    if (which == 2)
      s.Mixed_Assembler.output_rescaled_results(mesh, patches, s.sim_params, (unsigned)time_step_number, dir);
    else
      s.LDG_Assembler.output_rescaled_results(mesh, which == 0 ? s.electron_hole_pair : s.redox_pair, s.sim_params, patches,
                                              (unsigned)time_step_number, dir);
  });
}
pecs_status pecs_solarcell_interface_currents(pecs_solarcell* p, const double* const states[4], double out[2]) {
  return guarded([&] {
    if (!out) throw pecs::StatusError(PECS_ERR_INVALID, "interface_currents: out is NULL");
    p->problem->interface_currents(states, out);
  });
}
pecs_status pecs_solarcell_output_scales(const pecs_solarcell* p, double scales[4]) {
  return guarded([&] { PostProcessor(p->problem->sim_params, true, "").get_scales(scales); });
}

pecs_status pecs_solarcell_setup_test_host(pecs_solarcell* p, int32_t kind, int32_t n_refine) {
  return guarded([&] { p->problem->setup_test_host(kind, (unsigned)n_refine); });
}
pecs_status pecs_solarcell_setup_test(pecs_solarcell* p, int32_t kind, int32_t n_refine) {
  return guarded([&] { p->problem->setup_test(kind, (unsigned)n_refine); });
}
pecs_status pecs_solarcell_run_test(pecs_solarcell* p, int32_t kind, int32_t n_refine, double errors[4]) {
  return guarded([&] {
    SOLARCELL::ConvergenceTable mixed, ldg;
    errors[0] = errors[1] = errors[2] = errors[3] = 0.0;
    if (kind == PECS_KIND_TEST_STEADY)
      p->problem->test_steady_state((unsigned)n_refine, mixed, ldg);
    else if (kind == PECS_KIND_TEST_TRANSIENT)
      p->problem->test_transient((unsigned)n_refine, ldg);
    else if (kind == PECS_KIND_TEST_DD_POISSON)
      p->problem->test_DD_Poisson((unsigned)n_refine, mixed, ldg);
    else
      throw pecs::StatusError(PECS_ERR_INVALID, "unknown test kind");
    errors[0] = ldg.column("u").back();
    errors[1] = ldg.column("J").back();
    if (kind != PECS_KIND_TEST_TRANSIENT) {
      errors[2] = mixed.column("Phi").back();
      errors[3] = mixed.column("D").back();
    }
  });
}

pecs_ctx* pecs_solarcell_ctx(pecs_solarcell* p) { return p->problem->ctx; }
pecs_status pecs_solarcell_get_params(const pecs_solarcell* p, double params[32]) {
  return guarded([&] { p->problem->fill_params(params); });
}
double pecs_solarcell_delta_t(const pecs_solarcell* p) { return p->problem->delta_t; }

int32_t pecs_solarcell_n_cells(const pecs_solarcell* p, int32_t which) {
  try {
    return tria(p, which).n_active_cells();
  } catch (...) {
    return -1;
  }
}
pecs_status pecs_solarcell_get_mesh(const pecs_solarcell* p, int32_t which, double* vertices, int32_t* material_id,
                                    int32_t* level, int32_t* face_kind, int32_t* neighbor, int32_t* neighbor2,
                                    int32_t* boundary_id, double* nb_parent_diameter) {
  return guarded([&] {
    const pecs::MeshTables& t = tria(p, which).tables();
    copy_out(t.vertices, vertices);
    copy_out(t.material_id, material_id);
    copy_out(t.level, level);
    copy_out(t.face_kind, face_kind);
    copy_out(t.neighbor, neighbor);
    copy_out(t.neighbor2, neighbor2);
    copy_out(t.boundary_id, boundary_id);
    copy_out(t.nb_parent_diameter, nb_parent_diameter);
  });
}
int32_t pecs_solarcell_n_rt(const pecs_solarcell* p) { return p->problem->Poisson_object.dofs.n_rt; }
pecs_status pecs_solarcell_get_poisson_face_dofs(const pecs_solarcell* p, int32_t* face_dof) {
  return guarded([&] { copy_out(p->problem->Poisson_object.dofs.face_dof, face_dof); });
}
int32_t pecs_solarcell_n_constraints(const pecs_solarcell* p) {
  return (int32_t)p->problem->Poisson_object.dofs.constraints.size();
}
pecs_status pecs_solarcell_get_constraints(const pecs_solarcell* p, int32_t* dof, int32_t* master, double* weight) {
  return guarded([&] {
    size_t k = 0;
    for (const pecs::ConstraintLine& l : p->problem->Poisson_object.dofs.constraints) {
      dof[k] = l.dof;
      master[k] = l.master;
      weight[k] = l.weight;
      ++k;
    }
  });
}
pecs_status pecs_solarcell_get_cell_map(const pecs_solarcell* p, int32_t which, int32_t* map) {
  return guarded([&] { copy_out(which == 0 ? p->problem->s_2_p_map : p->problem->e_2_p_map, map); });
}
int32_t pecs_solarcell_n_interface_pairs(const pecs_solarcell* p) {
  return (int32_t)p->problem->semi_interface_cells.size();
}
pecs_status pecs_solarcell_get_interface_pairs(const pecs_solarcell* p, int32_t* semi_cell, int32_t* semi_face,
                                               int32_t* elec_cell, int32_t* elec_face) {
  return guarded([&] {
    copy_out(p->problem->semi_interface_cells, semi_cell);
    copy_out(p->problem->semi_interface_faces, semi_face);
    copy_out(p->problem->elec_interface_cells, elec_cell);
    copy_out(p->problem->elec_interface_faces, elec_face);
  });
}
int64_t pecs_solarcell_matrix_nnz(const pecs_solarcell* p, int32_t which) {
  try {
    return (int64_t)matrix(p, which).nnz();
  } catch (...) {
    return -1;
  }
}
pecs_status pecs_solarcell_get_matrix(const pecs_solarcell* p, int32_t which, int32_t* row_ptr, int32_t* col, double* val) {
  return guarded([&] {
    const pecs::CsrMatrix& A = matrix(p, which);
    copy_out(A.row_ptr, row_ptr);
    copy_out(A.col, col);
    copy_out(A.val, val);
  });
}

pecs_status pecs_solarcell_project_initial_conditions(pecs_solarcell* p) {
  return guarded([&] {
    SolarCellProblem& s = *p->problem;
    s.project_initial_conditions();
    if (s.ctx) {
      s.electron_hole_pair.carrier_1.push_solution();
      s.electron_hole_pair.carrier_2.push_solution();
      if (s.full_system) {
        s.redox_pair.carrier_1.push_solution();
        s.redox_pair.carrier_2.push_solution();
      }
    }
  });
}
pecs_status pecs_solarcell_project_test_initial_condition(pecs_solarcell* p) {
  return guarded([&] {
    p->problem->project_test_initial_condition();
    if (p->problem->ctx) p->problem->electron_hole_pair.carrier_1.push_solution();
  });
}
pecs_status pecs_solarcell_get_host_solution(const pecs_solarcell* p, int32_t which, double* out) {
  return guarded([&] { copy_out(carrier(p, which).solution, out); });
}
pecs_status pecs_solarcell_ldg_errors(pecs_solarcell* p, int32_t which, double time, double out[2]) {
  return guarded([&] { p->problem->ldg_errors(which, time, out[0], out[1]); });
}
pecs_status pecs_solarcell_mixed_errors(pecs_solarcell* p, double out[2]) {
  return guarded([&] { p->problem->mixed_errors(out[0], out[1]); });
}

pecs_status pecs_solarcell_plan_stats(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int64_t stats[8]) {
  return guarded([&] {
    const pecs::SolvePlan plan = pecs::plan_for_system(*p->problem, which, leaf_nodes);
    stats[0] = (int64_t)plan.fronts.size();
    stats[1] = (int64_t)plan.levels.size();
    stats[2] = plan.max_np;
    stats[3] = plan.max_nb;
    stats[4] = plan.fwd_entries;
    stats[5] = plan.bwd_entries;
    stats[6] = plan.upd_entries;
    stats[7] = 0;
  });
}
int32_t pecs_solarcell_plan_levels(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int64_t* out, int32_t max_levels) {
  try {
    const pecs::SolvePlan plan = pecs::plan_for_system(*p->problem, which, leaf_nodes);
    for (int d = 0; d < (int)plan.levels.size() && d < max_levels; ++d) {
      int64_t* o = out + 6 * d;
      o[0] = (int64_t)plan.levels[d].size();
      o[1] = o[2] = o[3] = o[4] = o[5] = 0;
      for (int f : plan.levels[d]) {
        const pecs::Front& F = plan.fronts[f];
        o[1] += F.fwd.size();
        o[2] += F.bwd.size();
        o[3] = std::max<int64_t>(o[3], F.np);
        o[4] = std::max<int64_t>(o[4], F.nb);
        o[5] += F.np;
      }
    }
    return (int32_t)plan.levels.size();
  } catch (const std::exception& e) {
    pecs::set_last_error(e.what());
    return -1;
  }
}
int64_t pecs_solarcell_plan_fronts(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, int32_t* out, int64_t max_fronts) {
  try {
    const pecs::SolvePlan plan = pecs::plan_for_system(*p->problem, which, leaf_nodes);
    const int64_t nf = (int64_t)plan.fronts.size();
    for (int64_t f = 0; f < nf && f < max_fronts; ++f) {
      const pecs::Front& F = plan.fronts[f];
      int32_t* o = out + 8 * f;
      o[0] = F.depth;
      o[1] = F.np;
      o[2] = F.nb;
      o[3] = F.fwd.log2P;
      o[4] = F.bwd.log2P;
      o[5] = F.fwd.small;
      o[6] = F.bwd.small;
      o[7] = F.parent < 0 ? -1 : (int32_t)F.parent;
    }
    return nf;
  } catch (const std::exception& e) {
    pecs::set_last_error(e.what());
    return -1;
  }
}
} // extern "C"
