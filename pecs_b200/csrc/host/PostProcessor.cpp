// PostProcessor.cpp -- see PostProcessor.hpp.
#include "PostProcessor.hpp"

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>

PostProcessor::PostProcessor(const ParameterSpace::Parameters& sim_params, const bool& print_carrier,
                             const std::string& name) {
  // reference PostProcessor.cpp:14-18
  scale_potential = 0.02585;
  scale_elec_field = 0.2585 / sim_params.characteristic_length;
  scale_density = sim_params.characteristic_denisty;
  scale_current = 1.6e-19 * scale_density * sim_params.characteristic_length / sim_params.characteristic_time;
  printing_carrier = print_carrier;
  if (print_carrier) {
    density_name = name + " Density";
    current_name = name + " Current";
  }
}

std::vector<std::string> PostProcessor::get_names() const {
  std::vector<std::string> names;
  if (printing_carrier) {
    names.push_back(current_name);
    names.push_back(current_name);
    names.push_back(density_name);
  } else {
    names.push_back("Field");
    names.push_back("Field");
    names.push_back("Potential");
  }
  return names;
}

void PostProcessor::get_scales(double scales[4]) const {
  scales[0] = scale_potential;
  scales[1] = scale_elec_field;
  scales[2] = scale_density;
  scales[3] = scale_current;
}

namespace pecs {

namespace {
const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

// base64 of two byte ranges taken as one stream
std::string base64_stream(const unsigned char* a, size_t na, const unsigned char* b, size_t nb) {
  const size_t n = na + nb;
  std::string out;
  out.resize((n + 2) / 3 * 4);
  auto at = [&](size_t i) -> unsigned { return i < na ? a[i] : b[i - na]; };
  size_t o = 0, i = 0;
  // the header is 4 bytes: run the generic accessor over the first 6 bytes, then the fast path on b alone
  const size_t slow_end = std::min(n / 3 * 3, (size_t)6);
  for (; i < slow_end; i += 3) {
    const unsigned v = (at(i) << 16) | (at(i + 1) << 8) | at(i + 2);
    out[o++] = kB64[v >> 18];
    out[o++] = kB64[(v >> 12) & 63];
    out[o++] = kB64[(v >> 6) & 63];
    out[o++] = kB64[v & 63];
  }
  if (i >= na) {
    const unsigned char* p = b + (i - na);
    const size_t full = (n - i) / 3 * 3;
    for (size_t k = 0; k < full; k += 3) {
      const unsigned v = ((unsigned)p[k] << 16) | ((unsigned)p[k + 1] << 8) | p[k + 2];
      out[o++] = kB64[v >> 18];
      out[o++] = kB64[(v >> 12) & 63];
      out[o++] = kB64[(v >> 6) & 63];
      out[o++] = kB64[v & 63];
    }
    i += full;
  }
  for (; i + 2 < n; i += 3) {
    const unsigned v = (at(i) << 16) | (at(i + 1) << 8) | at(i + 2);
    out[o++] = kB64[v >> 18];
    out[o++] = kB64[(v >> 12) & 63];
    out[o++] = kB64[(v >> 6) & 63];
    out[o++] = kB64[v & 63];
  }
  if (i < n) {
    const unsigned b0 = at(i), b1 = i + 1 < n ? at(i + 1) : 0;
    const unsigned v = (b0 << 16) | (b1 << 8);
    out[o++] = kB64[v >> 18];
    out[o++] = kB64[(v >> 12) & 63];
    out[o++] = i + 1 < n ? kB64[(v >> 6) & 63] : '=';
    out[o++] = '=';
  }
  out.resize(o);
  return out;
}
} // namespace

std::string base64_with_header(const void* data, size_t bytes) {
  if (bytes > 0xffffffffull) throw std::runtime_error("VTU array larger than the UInt32 header allows");
  const uint32_t header = (uint32_t)bytes;
  return base64_stream(reinterpret_cast<const unsigned char*>(&header), sizeof(header),
                       reinterpret_cast<const unsigned char*>(data), bytes);
}

VtuMesh::VtuMesh(const MeshTables& mesh) : n_cells_(mesh.n_cells) {
  const size_t n = (size_t)mesh.n_cells;
  std::vector<double> points(12 * n);
  std::vector<int32_t> conn(4 * n), offs(n);
  std::vector<uint8_t> types(n, 9); // VTK_QUAD
  for (size_t c = 0; c < n; ++c) {
    const double* v = mesh.vtx((int)c);
    for (int a = 0; a < 4; ++a) {
      points[3 * (4 * c + a) + 0] = v[2 * a];
      points[3 * (4 * c + a) + 1] = v[2 * a + 1];
      points[3 * (4 * c + a) + 2] = 0.0;
    }
    conn[4 * c + 0] = (int32_t)(4 * c + 0); // deal.II lexicographic -> VTK counter-clockwise
    conn[4 * c + 1] = (int32_t)(4 * c + 1);
    conn[4 * c + 2] = (int32_t)(4 * c + 3);
    conn[4 * c + 3] = (int32_t)(4 * c + 2);
    offs[c] = (int32_t)(4 * (c + 1));
  }
  points_ = base64_with_header(points.data(), points.size() * sizeof(double));
  connectivity_ = base64_with_header(conn.data(), conn.size() * sizeof(int32_t));
  offsets_ = base64_with_header(offs.data(), offs.size() * sizeof(int32_t));
  types_ = base64_with_header(types.data(), types.size());
}

void VtuMesh::write(const std::string& file, const std::vector<VtuField>& fields) const {
  std::ofstream out(file.c_str(), std::ios::binary);
  if (!out) throw std::runtime_error("cannot open output file " + file);
  const size_t np = 4 * (size_t)n_cells_;
  out << "<?xml version=\"1.0\" ?>\n"
      << "<!-- pecs_b200 output path; patches and field names as the reference writes them through deal.II DataOut -->\n"
      << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n"
      << "<UnstructuredGrid>\n<Piece NumberOfPoints=\"" << np << "\" NumberOfCells=\"" << n_cells_ << "\">\n"
      << "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"binary\">\n"
      << points_ << "\n</DataArray>\n</Points>\n<Cells>\n"
      << "<DataArray type=\"Int32\" Name=\"connectivity\" format=\"binary\">\n" << connectivity_ << "\n</DataArray>\n"
      << "<DataArray type=\"Int32\" Name=\"offsets\" format=\"binary\">\n" << offsets_ << "\n</DataArray>\n"
      << "<DataArray type=\"UInt8\" Name=\"types\" format=\"binary\">\n" << types_ << "\n</DataArray>\n</Cells>\n"
      << "<PointData Scalars=\"scalars\">\n";
  for (const VtuField& f : fields) {
    out << "<DataArray type=\"Float64\" Name=\"" << f.name << "\"";
    if (f.components > 1) out << " NumberOfComponents=\"" << f.components << "\"";
    out << " format=\"binary\">\n" << base64_with_header(f.data, np * (size_t)f.components * sizeof(double))
        << "\n</DataArray>\n";
  }
  out << "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
  if (!out) throw std::runtime_error("write failed: " + file);
}

OutputQueue::OutputQueue() : worker_([this] { run(); }) {}
OutputQueue::~OutputQueue() {
  {
    std::lock_guard<std::mutex> l(m_);
    stop_ = true;
  }
  cv_.notify_all();
  worker_.join();
}
unsigned long OutputQueue::submit(std::function<void()> job) {
  unsigned long ticket;
  {
    std::lock_guard<std::mutex> l(m_);
    jobs_.push_back(std::move(job));
    ticket = ++submitted_;
  }
  cv_.notify_one();
  return ticket;
}
void OutputQueue::wait_for(unsigned long ticket) {
  std::unique_lock<std::mutex> l(m_);
  idle_.wait(l, [&] { return completed_ >= ticket; });
}
void OutputQueue::wait_idle() {
  std::unique_lock<std::mutex> l(m_);
  idle_.wait(l, [this] { return jobs_.empty() && !busy_; });
}
std::string OutputQueue::error() const {
  std::lock_guard<std::mutex> l(m_);
  return error_;
}
void OutputQueue::run() {
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> l(m_);
      cv_.wait(l, [this] { return stop_ || !jobs_.empty(); });
      if (jobs_.empty()) return; // stop requested and drained
      job = std::move(jobs_.front());
      jobs_.pop_front();
      busy_ = true;
    }
    try {
      job();
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> l(m_);
      if (error_.empty()) error_ = e.what();
    }
    {
      std::lock_guard<std::mutex> l(m_);
      busy_ = false;
      ++completed_;
    }
    idle_.notify_all();
  }
}

} // namespace pecs
