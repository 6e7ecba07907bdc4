// test_functions.hpp -- manufactured solutions of the reference's convergence tests, host/device inline.
//
// Closed forms taken from reference source/test_functions.cpp: test_Poisson (:13-93), test_LDG_IMEX (:99-220),
// test_DD_Poisson (:227-361) and the initial condition every transient test uses,
// test_interface_problem::InitialConditions (:429-448).  The reference evaluates them through virtual
// dealii::Function calls per quadrature point; here they are inlined into the kernels that need them.
#pragma once
#include <math.h>

#include "fe.hpp"

namespace pecs {
namespace testfn {

#define PECS_TWO_PI 6.283185307179586476925286766559

// Phi = cos 2pi y - sin 2pi x - x ; D = -grad Phi ; f = -div grad Phi
PECS_HD double poisson_rhs(double x, double y) { return PECS_TWO_PI * PECS_TWO_PI * (cos(PECS_TWO_PI * y) - sin(PECS_TWO_PI * x)); }
PECS_HD double poisson_bc(double x, double y) { return cos(PECS_TWO_PI * y) - sin(PECS_TWO_PI * x) - x; }
PECS_HD void poisson_solution(double x, double y, double v[3]) {
  v[0] = 1.0 + PECS_TWO_PI * cos(PECS_TWO_PI * x);
  v[1] = PECS_TWO_PI * sin(PECS_TWO_PI * y);
  v[2] = poisson_bc(x, y);
}

// u = e^-t + cos 2pi x + cos 2pi y
PECS_HD double density(double x, double y, double t) { return exp(-t) + cos(PECS_TWO_PI * x) + cos(PECS_TWO_PI * y); }

// LDG-IMEX with the fixed field E = (1,0)
PECS_HD double ldg_rhs(double x, double y, double t) {
  return -exp(-t) + PECS_TWO_PI * PECS_TWO_PI * (cos(PECS_TWO_PI * x) + cos(PECS_TWO_PI * y)) + PECS_TWO_PI * sin(PECS_TWO_PI * x);
}
PECS_HD double ldg_interface(double, double y, double t) { return -exp(-t) - cos(PECS_TWO_PI * y) - 1.0; }
PECS_HD void ldg_solution(double x, double y, double t, double v[3]) {
  v[2] = density(x, y, t);
  v[0] = PECS_TWO_PI * sin(PECS_TWO_PI * x) - v[2];
  v[1] = PECS_TWO_PI * sin(PECS_TWO_PI * y);
}

// drift-diffusion coupled to Poisson: E = D of the Poisson problem above
PECS_HD double dd_rhs(double x, double y, double t) {
  const double u = density(x, y, t);
  const double sx = sin(PECS_TWO_PI * x), sy = sin(PECS_TWO_PI * y), cx = cos(PECS_TWO_PI * x), cy = cos(PECS_TWO_PI * y);
  const double div_E_u = PECS_TWO_PI * PECS_TWO_PI * (cy - sx) * u - PECS_TWO_PI * (PECS_TWO_PI * cx + 1.0) * sx -
                         PECS_TWO_PI * PECS_TWO_PI * sy * sy;
  return -exp(-t) + PECS_TWO_PI * PECS_TWO_PI * cx + PECS_TWO_PI * PECS_TWO_PI * cy - div_E_u;
}
PECS_HD double dd_poisson_rhs(double x, double y, double t) { return poisson_rhs(x, y) + density(x, y, t); }
PECS_HD void dd_solution(double x, double y, double t, double v[3]) {
  v[2] = density(x, y, t);
  v[0] = PECS_TWO_PI * sin(PECS_TWO_PI * x) - (PECS_TWO_PI * cos(PECS_TWO_PI * x) + 1.0) * v[2];
  v[1] = PECS_TWO_PI * sin(PECS_TWO_PI * y) - PECS_TWO_PI * sin(PECS_TWO_PI * y) * v[2];
}

PECS_HD double initial_condition(double x, double y) { return 1.0 + cos(PECS_TWO_PI * x) * cos(PECS_TWO_PI * y); }

} // namespace testfn
} // namespace pecs
