// PostProcessor.hpp -- rescaling to physical units and the asynchronous VTU writer of the output path.
//
// Mirror of reference include/PostProcessor.hpp / source/PostProcessor.cpp:9-123 (the scale factors and the names of
// the output fields) and of what dealii::DataOut::build_patches + write_vtu do for the reference in
// LDG::output_rescaled_results (source/LDG.cpp:1195-1232) and MixedFEM::output_rescaled_results
// (source/MixedFEM.cpp:297-320).  The arithmetic -- evaluating the solutions at the patch vertices and multiplying by
// the scales -- runs on the device (cuda/output_kernels.cu, include/pecs_b200.h pecs_output_snapshot); this file only
// names things and puts finished arrays into files, on a thread of its own so that the time loop never waits for a disk.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "Parameters.hpp"
#include "Triangulation.hpp"

// reference PostProcessor<dim> (include/PostProcessor.hpp): scales and names of one DataOut vector
class PostProcessor {
public:
  PostProcessor(const ParameterSpace::Parameters& sim_params, const bool& print_carrier, const std::string& name);
  // reference PostProcessor.cpp:34-55: dim x vector-part name, then the scalar name
  std::vector<std::string> get_names() const;
  // {potential, field, density, current} in the order pecs_output_snapshot takes them
  void get_scales(double scales[4]) const;

  double scale_potential, scale_elec_field, scale_density, scale_current;
  bool printing_carrier;
  std::string density_name, current_name;
};

namespace pecs {

// one point-data array of a VTU piece: 3-vectors or scalars over the 4 n patch vertices
struct VtuField {
  std::string name;
  int components; // 3 or 1
  const double* data;
};

// Geometry of the patches of one mesh (points, connectivity, offsets, types), encoded once: DataOut writes every patch
// with its own four vertices (no sharing), as VTK_QUAD with deal.II's vertex order 0,1,3,2.
class VtuMesh {
public:
  explicit VtuMesh(const MeshTables& mesh);
  int n_cells() const { return n_cells_; }
  // complete .vtu file (UnstructuredGrid, inline base64 binary, uncompressed, little endian, UInt32 headers)
  void write(const std::string& file, const std::vector<VtuField>& fields) const;

private:
  int n_cells_;
  // The file image.  Everything up to <PointData> (header, points, cells) depends on the mesh only: it is encoded once
  // and stays at the head of the buffer; every write() encodes the fields behind it and hands the whole image to the
  // operating system in one call.  One writer per mesh at a time (the output queue runs one stamp after the other).
  mutable std::vector<char> image_;
  size_t prefix_bytes_ = 0;
};

std::string base64_with_header(const void* data, size_t bytes);

// single background thread running file-writing jobs in submission order
class OutputQueue {
public:
  OutputQueue();
  ~OutputQueue(); // drains
  // returns the job's ticket (1, 2, ...); jobs finish in ticket order
  unsigned long submit(std::function<void()> job);
  void wait_for(unsigned long ticket); // returns at once for ticket 0
  void wait_idle();
  // first exception text a job raised (empty: none)
  std::string error() const;

private:
  void run();
  mutable std::mutex m_;
  std::condition_variable cv_, idle_;
  std::deque<std::function<void()>> jobs_;
  bool stop_ = false, busy_ = false;
  unsigned long submitted_ = 0, completed_ = 0;
  std::string error_;
  std::thread worker_;
};

} // namespace pecs
