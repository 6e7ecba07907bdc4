// PostProcessor.cpp -- see PostProcessor.hpp.
#include "PostProcessor.hpp"

#include <cstdint>
#include <cstdio>
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

// two output characters per 12 input bits (little-endian pair)
const uint16_t* base64_pairs() {
  static const std::vector<uint16_t> table = [] {
    std::vector<uint16_t> t(4096);
    for (unsigned v = 0; v < 4096; ++v) t[v] = (uint16_t)((unsigned char)kB64[v >> 6] | ((unsigned char)kB64[v & 63] << 8));
    return t;
  }();
  return table.data();
}

// base64 of two byte ranges taken as one stream.  The bulk (everything of b after the first two groups, in whole
// groups of three bytes) is encoded by several threads, each on its own range of the output: a cfg3 snapshot is 47 MB
// per time stamp and the writer must keep up with a time loop that needs 20 ms for the steps between two stamps.
// writes (na + nb + 2) / 3 * 4 characters to out and returns that number
size_t base64_stream(const unsigned char* a, size_t na, const unsigned char* b, size_t nb, char* out) {
  const size_t n = na + nb;
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
    const size_t groups = (n - i) / 3;
    const uint16_t* pair = base64_pairs();
    char* dst = out + o;
    const long long n_blocks = (long long)((groups + 65535) / 65536);
#pragma omp parallel for schedule(static) num_threads(4) if (n_blocks > 16)
    for (long long blk = 0; blk < n_blocks; ++blk) {
      const size_t g0 = (size_t)blk * 65536, g1 = std::min(groups, g0 + 65536);
      const unsigned char* q = p + 3 * g0;
      char* w = dst + 4 * g0;
      for (size_t g = g0; g < g1; ++g, q += 3, w += 4) {
        const unsigned v = ((unsigned)q[0] << 16) | ((unsigned)q[1] << 8) | q[2];
        const uint16_t hi = pair[v >> 12], lo = pair[v & 4095];
        std::memcpy(w, &hi, 2);
        std::memcpy(w + 2, &lo, 2);
      }
    }
    o += 4 * groups;
    i += 3 * groups;
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
  return o;
}
} // namespace

size_t base64_size_with_header(size_t bytes) { return (sizeof(uint32_t) + bytes + 2) / 3 * 4; }

// VTK inline binary: a UInt32 byte count, then the data, as ONE base64 stream; returns the characters written
size_t base64_with_header(const void* data, size_t bytes, char* out) {
  if (bytes > 0xffffffffull) throw std::runtime_error("VTU array larger than the UInt32 header allows");
  const uint32_t header = (uint32_t)bytes;
  return base64_stream(reinterpret_cast<const unsigned char*>(&header), sizeof(header),
                       reinterpret_cast<const unsigned char*>(data), bytes, out);
}

std::string base64_with_header(const void* data, size_t bytes) {
  std::string out(base64_size_with_header(bytes), '\0');
  out.resize(base64_with_header(data, bytes, &out[0]));
  return out;
}

namespace {
void append(std::vector<char>& image, const std::string& text) { image.insert(image.end(), text.begin(), text.end()); }
void append_array(std::vector<char>& image, const void* data, size_t bytes) {
  const size_t at = image.size();
  image.resize(at + base64_size_with_header(bytes));
  image.resize(at + base64_with_header(data, bytes, image.data() + at));
}
} // namespace

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
  const size_t np = 4 * n;
  append(image_, "<?xml version=\"1.0\" ?>\n"
                 "<!-- pecs_b200 output path; patches and field names as the reference writes them through deal.II DataOut -->\n"
                 "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n"
                 "<UnstructuredGrid>\n<Piece NumberOfPoints=\"" + std::to_string(np) + "\" NumberOfCells=\"" +
                     std::to_string(n_cells_) + "\">\n"
                 "<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"binary\">\n");
  append_array(image_, points.data(), points.size() * sizeof(double));
  append(image_, "\n</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"binary\">\n");
  append_array(image_, conn.data(), conn.size() * sizeof(int32_t));
  append(image_, "\n</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"binary\">\n");
  append_array(image_, offs.data(), offs.size() * sizeof(int32_t));
  append(image_, "\n</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"binary\">\n");
  append_array(image_, types.data(), types.size());
  append(image_, "\n</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n");
  prefix_bytes_ = image_.size();
}

void VtuMesh::write(const std::string& file, const std::vector<VtuField>& fields) const {
  const size_t np = 4 * (size_t)n_cells_;
  image_.resize(prefix_bytes_); // keeps the capacity of the previous stamp: no allocation, no page faults
  for (const VtuField& f : fields) {
    std::string open = "<DataArray type=\"Float64\" Name=\"" + f.name + "\"";
    if (f.components > 1) open += " NumberOfComponents=\"" + std::to_string(f.components) + "\"";
    append(image_, open + " format=\"binary\">\n");
    append_array(image_, f.data, np * (size_t)f.components * sizeof(double));
    append(image_, "\n</DataArray>\n");
  }
  append(image_, "</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
  std::FILE* out = std::fopen(file.c_str(), "wb");
  if (!out) throw std::runtime_error("cannot open output file " + file);
  const size_t written = std::fwrite(image_.data(), 1, image_.size(), out);
  const bool closed = std::fclose(out) == 0;
  if (written != image_.size() || !closed) throw std::runtime_error("write failed: " + file);
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
