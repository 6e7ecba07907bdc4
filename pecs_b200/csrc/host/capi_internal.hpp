// capi_internal.hpp -- what the C-ABI translation units of the host classes share (the product's capi_host.cpp and the
// test library's selftest/selftest.cpp): the handle type and the exception -> status wrapper.
#pragma once
#include <memory>
#include <stdexcept>

#include "../../../include/pecs_b200_host.h"
#include "../error.hpp"
#include "Parameters.hpp"
#include "SolarCell.hpp"

struct pecs_solarcell {
  ParameterSpace::ParameterHandler prm;
  std::unique_ptr<SOLARCELL::SolarCellProblem> problem;
  // geometry of the three meshes for pecs_solarcell_write_patches, encoded at the first call (the meshes do not change
  // after setup)
  std::unique_ptr<pecs::VtuMesh> patch_mesh[3];
};

namespace pecs {
namespace capi {
template <class F>
pecs_status guarded(F&& f) {
  try {
    f();
    return PECS_OK;
  } catch (const StatusError& e) {
    set_last_error(e.what());
    return e.status;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return PECS_ERR_INTERNAL;
  }
}
inline const Triangulation& tria(const pecs_solarcell* p, int which) {
  if (which == 0) return p->problem->semiconductor_triangulation;
  if (which == 1) return p->problem->electrolyte_triangulation;
  if (which == 2) return p->problem->Poisson_triangulation;
  throw StatusError(PECS_ERR_INVALID, "mesh selector must be 0, 1 or 2");
}
} // namespace capi
} // namespace pecs
