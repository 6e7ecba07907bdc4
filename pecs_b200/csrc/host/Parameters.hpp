// Parameters.hpp -- ParameterSpace::{ParameterHandler, ParameterReader, Parameters}
//
// Host-side mirror of the reference's configuration layer:
//   * ParameterReader  : reference source/ParameterReader.cpp:13-392 (declares every entry with its
//                        default, then reads a deal.II-style .prm file: `subsection X` / `set k = v` / `end`)
//   * Parameters       : reference include/Parameters.hpp:27-287 (member names kept, including the
//                        reference's spellings, so code written against the reference reads the same)
// deal.II's ParameterHandler is replaced by a small own class with the subset of the interface the
// reference uses (enter_subsection / declare_entry / get_double / get_integer / get_bool / read_input).
#pragma once
#include <map>
#include <string>

namespace ParameterSpace {

namespace PhysicalConstants {
// reference include/Parameters.hpp:9-11
const double thermal_voltage = 0.02585;      // [V]
const double electron_charge = 1.62e-19;     // [C]
const double vacuum_permittivity = 8.85e-14; // [A s V^-1 cm^-1]
} // namespace PhysicalConstants

class ParameterHandler {
public:
  void enter_subsection(const std::string& name) { current_ = name; }
  void leave_subsection() { current_.clear(); }
  void declare_entry(const std::string& key, const std::string& default_value, const std::string& doc = "");
  void set(const std::string& key, const std::string& value); // in the current subsection
  double get_double(const std::string& key) const;
  long get_integer(const std::string& key) const;
  bool get_bool(const std::string& key) const;
  // parse a .prm file; unknown entries are an error (as in deal.II)
  void read_input(const std::string& file_name);
  void read_input_from_string(const std::string& text);

private:
  const std::string& lookup(const std::string& key) const;
  std::string current_;
  std::map<std::string, std::map<std::string, std::string>> entries_;
};

class ParameterReader {
public:
  explicit ParameterReader(ParameterHandler& param_handler) : prm(param_handler) {}
  // reference ParameterReader.cpp:13-19: declare, then read (a missing file leaves the defaults)
  void read_parameters(const std::string& parameter_file);
  // reference ParameterReader.cpp:21-27
  void read_test_parameters(const std::string& parameter_file);
  void declare_parameters();      // reference ParameterReader.cpp:29-208
  void declare_test_parameters(); // reference ParameterReader.cpp:209-392

private:
  ParameterHandler& prm;
};

struct Parameters {
  // computational
  unsigned int n_global_refine = 0;
  unsigned int n_local_refine = 0;
  unsigned int time_stamps = 0;
  double h_max = 0, h_min = 0;
  double t_end = 0, t_end_2 = 0, delta_t = 0, penalty = 0;
  // electrons / holes
  double scaled_electron_mobility = 0, scaled_electron_recombo_t = 0, scaled_electron_recombo_v = 0, scaled_k_et = 0;
  double scaled_hole_mobility = 0, scaled_hole_recombo_t = 0, scaled_hole_recombo_v = 0, scaled_k_ht = 0;
  double scaled_intrinsic_density = 0, semiconductor_permittivity = 0;
  // redox
  double scaled_reductant_mobility = 0, scaled_oxidant_mobility = 0, electrolyte_permittivity = 0;
  // physical
  double scaled_absorption_coeff = 0, scaled_photon_flux = 0;
  double scaled_debeye_length = 0, scaled_boundary_layer = 0;
  double characteristic_length = 0, characteristic_time = 0, characteristic_denisty = 0;
  double scaled_domain_length = 0, scaled_domain_height = 0, scaled_radius_one = 0, scaled_radius_two = 0;
  bool illum_or_dark = false, insulated = false, restart_status = false, schottky_status = false;
  bool srh_recombination = false; // extension: the reference's SRH_Recombination returns 0.0 (include/SolarCell.hpp:86-98)
  double scaled_applied_bias = 0, scaled_built_in_bias = 0, scaled_schottky_bias = 0;
  double rescale_current = 0, rescaled_k_et = 0, rescaled_k_ht = 0;

  // reference include/Parameters.hpp:88-107
  void set_params_for_testing(const unsigned int& n_refine);
  // reference include/Parameters.hpp:117-287: read every entry and apply the singular-perturbation scaling
  void parse_and_scale_parameters(ParameterHandler& prm);
};

} // namespace ParameterSpace
