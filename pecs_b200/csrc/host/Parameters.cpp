// Parameters.cpp -- see Parameters.hpp.
#include "Parameters.hpp"

#include <fstream>
#include <sstream>
#include <stdexcept>

namespace ParameterSpace {

namespace {
std::string trim(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

struct Entry {
  const char* section;
  const char* key;
  const char* production_default; // reference ParameterReader.cpp:29-208
  const char* test_default;       // reference ParameterReader.cpp:209-392
};
// One table for both declaration sets; the (section, key) universe is identical in the reference.
const Entry kEntries[] = {
    {"computational", "global refinements", "4", "2"},
    {"computational", "local refinements", "0", "0"},
    {"computational", "time step size", "0.05", "0.01"},
    {"computational", "end time", "10", "1"},
    {"computational", "end time 2", "20", "20"},
    {"computational", "time stamps", "50", "100"},
    {"computational", "restart status", "false", "false"},
    {"physical", "applied bias", "0.0", "0.0"},
    {"physical", "built in bias", "0.41", "0.0"},
    {"physical", "schottky bias", "0.0", "0.0"},
    {"physical", "illumination status", "false", "false"},
    {"physical", "insulated", "false", "false"},
    {"physical", "schottky status", "false", "false"},
    // not a key of the reference (its SRH_Recombination returns 0.0, include/SolarCell.hpp:86-98): switches the
    // Shockley-Read-Hall formula that function carries as a comment on
    {"physical", "srh recombination", "false", "false"},
    {"physical", "characteristic length", "1.0e-4", "1.0"},
    {"physical", "characteristic density", "1.0e16", "1.0"},
    {"physical", "characteristic time", "1.0e-12", "1.0"},
    {"physical", "intrinsic density", "2.564e9", "1.0"},
    {"physical", "semiconductor permittivity", "11.9", "1.0"},
    {"physical", "electrolyte permittivity", "100", "1.0"},
    {"physical", "photon flux", "1.2e17", "0.0"},
    {"physical", "absorption coefficient", "1.74974e5", "0.0"},
    {"mesh", "mesh length", "1.0", "2"},
    {"mesh", "mesh height", "1", "1"},
    {"mesh", "radius one", "0.5", "1.0"},
    {"mesh", "radius two", "0.5", "1.0"},
    {"mesh", "boundary layer", "0.1", "0.1"},
    {"electrons", "mobility", "1350.0", "1.0"},
    {"electrons", "transfer rate", "1e-19", "0"},
    {"electrons", "recombination time", "5e-5", "0"},
    {"electrons", "recombination velocity", "3e5", "3e5"},
    {"holes", "mobility", "480.0", "1.0"},
    {"holes", "transfer rate", "1e-16", "0"},
    {"holes", "recombination time", "5e-5", "0"},
    {"holes", "recombination velocity", "2.9e-2", "2.9e-2"},
    {"reductants", "mobility", "1.0", "1.0"},
    {"oxidants", "mobility", "1.0", "1.0"},
};
} // namespace

void ParameterHandler::declare_entry(const std::string& key, const std::string& default_value, const std::string&) {
  entries_[current_][key] = default_value;
}

void ParameterHandler::set(const std::string& key, const std::string& value) {
  auto sec = entries_.find(current_);
  if (sec == entries_.end() || sec->second.find(key) == sec->second.end())
    throw std::runtime_error("ParameterHandler: undeclared entry '" + key + "' in subsection '" + current_ + "'");
  sec->second[key] = value;
}

const std::string& ParameterHandler::lookup(const std::string& key) const {
  auto sec = entries_.find(current_);
  if (sec == entries_.end()) throw std::runtime_error("ParameterHandler: no subsection '" + current_ + "'");
  auto it = sec->second.find(key);
  if (it == sec->second.end())
    throw std::runtime_error("ParameterHandler: no entry '" + key + "' in subsection '" + current_ + "'");
  return it->second;
}

double ParameterHandler::get_double(const std::string& key) const { return std::stod(lookup(key)); }
long ParameterHandler::get_integer(const std::string& key) const { return std::stol(lookup(key)); }
bool ParameterHandler::get_bool(const std::string& key) const {
  const std::string& v = lookup(key);
  if (v == "true" || v == "yes" || v == "on") return true;
  if (v == "false" || v == "no" || v == "off") return false;
  throw std::runtime_error("ParameterHandler: '" + key + "' is not a bool: " + v);
}

void ParameterHandler::read_input_from_string(const std::string& text) {
  std::istringstream in(text);
  std::string line;
  const std::string saved = current_;
  current_.clear();
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    line = trim(line);
    if (line.empty()) continue;
    if (line.rfind("subsection", 0) == 0) {
      current_ = trim(line.substr(10));
    } else if (line == "end") {
      current_.clear();
    } else if (line.rfind("set", 0) == 0) {
      const size_t eq = line.find('=');
      if (eq == std::string::npos) throw std::runtime_error("ParameterHandler: malformed line: " + line);
      set(trim(line.substr(3, eq - 3)), trim(line.substr(eq + 1)));
    } else {
      throw std::runtime_error("ParameterHandler: cannot parse line: " + line);
    }
  }
  current_ = saved;
}

void ParameterHandler::read_input(const std::string& file_name) {
  std::ifstream f(file_name);
  if (!f) return; // the reference creates a default file when it is missing: the defaults stay in effect
  std::stringstream ss;
  ss << f.rdbuf();
  read_input_from_string(ss.str());
}

void ParameterReader::declare_parameters() {
  for (const Entry& e : kEntries) {
    prm.enter_subsection(e.section);
    prm.declare_entry(e.key, e.production_default);
    prm.leave_subsection();
  }
}

void ParameterReader::declare_test_parameters() {
  for (const Entry& e : kEntries) {
    prm.enter_subsection(e.section);
    prm.declare_entry(e.key, e.test_default);
    prm.leave_subsection();
  }
}

void ParameterReader::read_parameters(const std::string& parameter_file) {
  declare_parameters();
  prm.read_input(parameter_file);
}

void ParameterReader::read_test_parameters(const std::string& parameter_file) {
  declare_test_parameters();
  prm.read_input(parameter_file);
}

void Parameters::set_params_for_testing(const unsigned int& n_refine) {
  n_global_refine = n_refine;
  t_end = 1.0;
  scaled_electron_mobility = scaled_hole_mobility = 1.0;
  scaled_oxidant_mobility = scaled_reductant_mobility = 1.0;
  scaled_absorption_coeff = 0.0;
  scaled_domain_height = scaled_domain_length = 1.0;
  scaled_radius_one = scaled_radius_two = 0.5;
  scaled_debeye_length = 1.0;
  characteristic_length = characteristic_denisty = characteristic_time = 1.0;
}

void Parameters::parse_and_scale_parameters(ParameterHandler& prm) {
  using namespace PhysicalConstants;
  prm.enter_subsection("computational");
  n_global_refine = (unsigned)prm.get_integer("global refinements");
  n_local_refine = (unsigned)prm.get_integer("local refinements");
  delta_t = prm.get_double("time step size");
  t_end = prm.get_double("end time");
  t_end_2 = prm.get_double("end time 2");
  time_stamps = (unsigned)prm.get_integer("time stamps");
  restart_status = prm.get_bool("restart status");
  prm.leave_subsection();

  prm.enter_subsection("mesh");
  scaled_domain_height = prm.get_double("mesh height");
  scaled_domain_length = prm.get_double("mesh length");
  scaled_radius_one = prm.get_double("radius one");
  scaled_radius_two = prm.get_double("radius two");
  scaled_boundary_layer = prm.get_double("boundary layer");
  prm.leave_subsection();

  prm.enter_subsection("physical");
  illum_or_dark = prm.get_bool("illumination status");
  insulated = prm.get_bool("insulated");
  schottky_status = prm.get_bool("schottky status");
  srh_recombination = prm.get_bool("srh recombination");
  scaled_applied_bias = prm.get_double("applied bias");
  scaled_built_in_bias = prm.get_double("built in bias");
  scaled_schottky_bias = prm.get_double("schottky bias");
  characteristic_length = prm.get_double("characteristic length");
  characteristic_denisty = prm.get_double("characteristic density");
  characteristic_time = prm.get_double("characteristic time");
  scaled_intrinsic_density = prm.get_double("intrinsic density");
  scaled_photon_flux = prm.get_double("photon flux");
  scaled_absorption_coeff = prm.get_double("absorption coefficient");
  semiconductor_permittivity = prm.get_double("semiconductor permittivity");
  electrolyte_permittivity = prm.get_double("electrolyte permittivity");
  prm.leave_subsection();

  prm.enter_subsection("electrons");
  scaled_electron_mobility = prm.get_double("mobility");
  scaled_k_et = prm.get_double("transfer rate");
  scaled_electron_recombo_t = prm.get_double("recombination time");
  scaled_electron_recombo_v = prm.get_double("recombination velocity");
  prm.leave_subsection();

  prm.enter_subsection("holes");
  scaled_hole_mobility = prm.get_double("mobility");
  scaled_k_ht = prm.get_double("transfer rate");
  scaled_hole_recombo_t = prm.get_double("recombination time");
  scaled_hole_recombo_v = prm.get_double("recombination velocity");
  prm.leave_subsection();

  prm.enter_subsection("reductants");
  scaled_reductant_mobility = prm.get_double("mobility");
  prm.leave_subsection();
  prm.enter_subsection("oxidants");
  scaled_oxidant_mobility = prm.get_double("mobility");
  prm.leave_subsection();

  // singular-perturbation scaling, reference Parameters.hpp:181-242 (same operation order)
  const double L = characteristic_length, T = characteristic_time, C = characteristic_denisty;
  scaled_intrinsic_density /= C;
  scaled_electron_recombo_t /= T;
  scaled_hole_recombo_t /= T;
  scaled_electron_recombo_v *= (T / L);
  scaled_hole_recombo_v *= (T / L);
  scaled_photon_flux *= (T / C);
  scaled_absorption_coeff *= L;
  // the material permittivity is deliberately NOT part of the Debye length (Parameters.hpp:197-205)
  scaled_debeye_length = (thermal_voltage * vacuum_permittivity) / (electron_charge * C * L * L);
  const double mobility_scale = (T * thermal_voltage) / (L * L);
  scaled_electron_mobility *= mobility_scale;
  scaled_hole_mobility *= mobility_scale;
  scaled_reductant_mobility *= mobility_scale;
  scaled_oxidant_mobility *= mobility_scale;
  rescaled_k_et = electron_charge * scaled_k_et * C * C;
  rescaled_k_ht = electron_charge * scaled_k_ht * C * C;
  scaled_k_et *= (T * C / L);
  scaled_k_ht *= (T * C / L);
  scaled_applied_bias /= thermal_voltage;
  scaled_built_in_bias /= thermal_voltage;
  scaled_schottky_bias /= thermal_voltage;
  rescale_current = (electron_charge * C * L) / T;
}

} // namespace ParameterSpace
