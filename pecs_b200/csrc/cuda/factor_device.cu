// factor_device.cu -- numeric multifrontal factorisation on the device (one-time setup, see host/SparseDirect.hpp).
//
// Replaces SparseDirectUMFPACK::initialize (reference source/Carrier.cpp:26-32, source/Poisson.cpp:92-96).
// Level by level from the leaves: assemble every frontal matrix F = [F_PP F_PB; F_BP F_BB] (original entries +
// the children's Schur complements), then per front
//     Inv = F_PP^-1,   -H = -Inv F_PB,   G = F_BP Inv,   F_BB <- F_BB - G F_PB  (the Schur complement for the parent)
// Small fronts (the vast majority) are handled by one thread block each with Gauss-Jordan elimination; large fronts
// use cuSOLVER getrf/getrs for the inverse and cuBLAS DGEMM for the three products (plain library GEMMs, setup only).
// All matrices are row-major; a row-major matrix handed to a column-major library is its transpose, and
// (F_PP^T)^-1 = Inv^T, so the library results land directly in row-major scratch operators (derivation in the
// comments); a last kernel per level packs them into the panel layout the solve kernels stream.
#include "factor_device.cuh"

#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>

#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "../error.hpp"
#include "device_util.cuh"

namespace pecs {

namespace {

struct FactorFront {
  int np, nb, p0;
  int child[2];            // indices into the PREVIOUS level's FactorFront array, -1 if none
  long long bd_off;        // into bd_index
  long long F_off;         // into this level's frontal buffer (m x m, row-major)
  long long Gs_off, Bs_off; // into this level's scratch operators: G (nb x np) and [Inv | -H] (np x m), row-major
  long long cl_off[2];     // into the child-local index scratch
  // destination panels (host/SparseDirect.hpp PanelTable)
  long long fwd_off, bwd_off;
  int fwd_log2P, fwd_cols_pad, bwd_log2P, bwd_cols_pad;
};

__device__ __forceinline__ int local_index(int pos, int p0, int np, const int* bd, int nb) {
  if (pos < p0 + np) return pos - p0;
  int lo = 0, hi = nb;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bd[mid] < pos)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < nb && bd[lo] == pos) ? np + lo : -1;
}

// original matrix entries: one block per front
__global__ void assemble_kernel(const FactorFront* __restrict__ fronts, const int* __restrict__ bd_index,
                                const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
                                const int* __restrict__ trp, const int* __restrict__ tcol, const double* __restrict__ tval,
                                double* Fbuf, int* error) {
  const FactorFront F = fronts[blockIdx.x];
  const int m = F.np + F.nb;
  const int* bd = bd_index + F.bd_off;
  double* M = Fbuf + F.F_off;
  for (int i = threadIdx.x; i < F.np; i += blockDim.x) {
    const int r = F.p0 + i;
    for (int k = rp[r]; k < rp[r + 1]; ++k) {
      const int q = col[k];
      if (q < F.p0) continue;
      const int l = local_index(q, F.p0, F.np, bd, F.nb);
      if (l < 0) {
        atomicExch(error, 1);
        continue;
      }
      M[(size_t)i * m + l] = val[k];
    }
    for (int k = trp[r]; k < trp[r + 1]; ++k) {
      const int q = tcol[k];
      if (q < F.p0 + F.np) continue;
      const int l = local_index(q, F.p0, F.np, bd, F.nb);
      if (l < 0) {
        atomicExch(error, 1);
        continue;
      }
      M[(size_t)l * m + i] = tval[k];
    }
  }
}

// cl[s] = local index in the parent front of the child's boundary slot s
__global__ void child_local_kernel(const FactorFront* __restrict__ fronts, const FactorFront* __restrict__ child_fronts,
                                   const int* __restrict__ bd_index, int which, int* cl, int* error) {
  const FactorFront F = fronts[blockIdx.x];
  if (F.child[which] < 0) return;
  const FactorFront C = child_fronts[F.child[which]];
  const int* cbd = bd_index + C.bd_off;
  const int* bd = bd_index + F.bd_off;
  int* out = cl + F.cl_off[which];
  for (int s = threadIdx.x; s < C.nb; s += blockDim.x) {
    const int l = local_index(cbd[s], F.p0, F.np, bd, F.nb);
    if (l < 0) atomicExch(error, 2);
    out[s] = l;
  }
}

// F[cl[s]][cl[t]] += U_child[s][t]; grid.x = front, grid.y = row chunk of the child's Schur complement
__global__ void extend_add_kernel(const FactorFront* __restrict__ fronts, const FactorFront* __restrict__ child_fronts,
                                  int which, const int* __restrict__ cl, const double* __restrict__ child_Fbuf, double* Fbuf) {
  const FactorFront F = fronts[blockIdx.x];
  if (F.child[which] < 0) return;
  const FactorFront C = child_fronts[F.child[which]];
  const int m = F.np + F.nb, mc = C.np + C.nb;
  const int* map = cl + F.cl_off[which];
  const double* U = child_Fbuf + C.F_off + (size_t)C.np * mc + C.np;
  double* M = Fbuf + F.F_off;
  for (int s = blockIdx.y; s < C.nb; s += gridDim.y) {
    double* row = M + (size_t)map[s] * m;
    const double* u = U + (size_t)s * mc;
    for (int t = threadIdx.x; t < C.nb; t += blockDim.x) row[map[t]] += u[t];
  }
}

// One block per small front: Gauss-Jordan inverse of F_PP (partial pivoting) straight into the backward table,
// then -H, G and the Schur complement.  Everything stays in L1/L2 for these sizes.
__global__ void __launch_bounds__(256) small_front_kernel(const FactorFront* __restrict__ fronts,
                                                          const int* __restrict__ small_list, double* Fbuf, double* Gs,
                                                          double* Bs, int* error) {
  const FactorFront F = fronts[small_list[blockIdx.x]];
  const int np = F.np, nb = F.nb, m = np + nb;
  double* M = Fbuf + F.F_off;   // rows 0..np-1 hold [F_PP | F_PB]
  double* B = Bs + F.Bs_off;    // [Inv | -H], row stride m
  double* G = Gs + F.Gs_off;    // nb x np, row-major
  const int ldb = m;
  const size_t g_rs = (size_t)np, g_cs = 1;
  __shared__ int s_piv;
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < np * np; e += nt) B[(size_t)(e / np) * ldb + (e % np)] = (e / np == e % np) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 0; k < np; ++k) {
    // pivot search in column k, rows k..np-1
    double best = -1.0;
    int bi = k;
    for (int r = k + tid; r < np; r += nt) {
      const double a = fabs(M[(size_t)r * m + k]);
      if (a > best) {
        best = a;
        bi = r;
      }
    }
    s_val[tid] = best;
    s_idx[tid] = bi;
    __syncthreads();
    for (int o = nt >> 1; o > 0; o >>= 1) {
      if (tid < o && (s_val[tid + o] > s_val[tid] || (s_val[tid + o] == s_val[tid] && s_idx[tid + o] < s_idx[tid]))) {
        s_val[tid] = s_val[tid + o];
        s_idx[tid] = s_idx[tid + o];
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_piv = s_idx[0];
      if (!(s_val[0] > 0.0) || !isfinite(s_val[0])) atomicExch(error, 3);
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k) {
      for (int j = tid; j < np; j += nt) {
        const double a = M[(size_t)k * m + j];
        M[(size_t)k * m + j] = M[(size_t)p * m + j];
        M[(size_t)p * m + j] = a;
        const double b = B[(size_t)k * ldb + j];
        B[(size_t)k * ldb + j] = B[(size_t)p * ldb + j];
        B[(size_t)p * ldb + j] = b;
      }
      __syncthreads();
    }
    const double inv = 1.0 / M[(size_t)k * m + k];
    __syncthreads();
    for (int j = tid; j < np; j += nt) {
      M[(size_t)k * m + j] *= inv;
      B[(size_t)k * ldb + j] *= inv;
    }
    __syncthreads();
    // eliminate column k from every other row; each thread owns (row, column) pairs, column k is read first
    for (int e = tid; e < np * np; e += nt) {
      const int r = e / np, j = e % np;
      if (r == k) continue;
      const double f = M[(size_t)r * m + k];
      if (j != k) M[(size_t)r * m + j] -= f * M[(size_t)k * m + j];
      B[(size_t)r * ldb + j] -= f * B[(size_t)k * ldb + j];
    }
    __syncthreads();
    for (int r = tid; r < np; r += nt)
      if (r != k) M[(size_t)r * m + k] = 0.0;
    __syncthreads();
  }
  // NOTE: row swaps were applied to F_PP only; F_PB rows were not swapped, which is right: P F_PP = LU-like
  // elimination acts on [F_PP | I], and B now holds F_PP^-1 exactly (no permutation left over).
  // -H = -Inv F_PB, G = F_BP Inv and the Schur complement F_BB -= G F_PB follow in small_front_products_kernel.
}

// The three products of the small fronts of a level, tiled over the whole device.  (They used to be the tail of
// small_front_kernel, one thread block per front, one dot product from global memory per thread: fine for thousands of
// leaf fronts, but a front with 100 pivots and 1 000-1 400 boundary unknowns then kept ONE block busy for 60-90 ms:
// levels 6-8 of a carrier tree were 220 of the 390 ms of a factorisation.)  A tile is 64 x 64 entries of the result,
// 256 threads x (4 x 4), the operands staged through shared memory 16 columns of the inner dimension at a time.  Every
// entry is still the fused-multiply-add chain over k = 0, 1, 2, ... of the old loops: the tables are bit-identical.
//   stage 0: -H = -Inv F_PB (tiles first)  and  G = F_BP Inv;    stage 1: F_BB -= G F_PB
__global__ void __launch_bounds__(256) small_front_products_kernel(const FactorFront* __restrict__ fronts,
                                                                   const int* __restrict__ small_list, int stage, double* Fbuf,
                                                                   double* Gs, double* Bs) {
  const FactorFront F = fronts[small_list[blockIdx.x]];
  const int np = F.np, nb = F.nb, m = np + nb;
  if (nb == 0) return;
  double* Mf = Fbuf + F.F_off; // rows 0..np-1: [. | F_PB], rows np..: [F_BP | F_BB]
  double* Bt = Bs + F.Bs_off;  // [Inv | -H], row stride m
  double* Gt = Gs + F.Gs_off;  // nb x np
  const int tp = (np + 63) / 64, tb = (nb + 63) / 64;
  // C (rows x cols) = sign * A (rows x inner, lda) * Bm (inner x cols, ldbm)  [+ C when accumulate]
  const double *A, *Bm;
  double* C;
  int rows, cols, lda, ldbm, ldc, tile = (int)blockIdx.y;
  bool negate, accumulate;
  const int inner = np;
  if (stage == 0) {
    if (tile < tp * tb) { // -H
      A = Bt, lda = m, Bm = Mf + np, ldbm = m, C = Bt + np, ldc = m;
      rows = np, cols = nb, negate = true, accumulate = false;
    } else if ((tile -= tp * tb) < tb * tp) { // G
      A = Mf + (size_t)np * m, lda = m, Bm = Bt, ldbm = m, C = Gt, ldc = np;
      rows = nb, cols = np, negate = false, accumulate = false;
    } else {
      return;
    }
  } else {
    if (tile >= tb * tb) return;
    A = Gt, lda = np, Bm = Mf + np, ldbm = m, C = Mf + (size_t)np * m + np, ldc = m;
    rows = nb, cols = nb, negate = true, accumulate = true;
  }
  const int tiles_across = (cols + 63) / 64;
  const int i0 = (tile / tiles_across) * 64, j0 = (tile % tiles_across) * 64;
  __shared__ double As[16][65], Bsm[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < inner; k0 += 16) {
    const int kn = min(16, inner - k0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q;
      {
        const int r = e >> 4, c = e & 15; // A: 64 rows x 16 inner, the inner index fastest
        As[c][r] = (i0 + r < rows && c < kn) ? A[(size_t)(i0 + r) * lda + k0 + c] : 0.0;
      }
      {
        const int r = e >> 6, c = e & 63; // Bm: 16 inner x 64 columns
        Bsm[r][c] = (r < kn && j0 + c < cols) ? Bm[(size_t)(k0 + r) * ldbm + j0 + c] : 0.0;
      }
    }
    __syncthreads();
    for (int k = 0; k < kn; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][4 * ty + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bsm[k][4 * tx + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = i0 + 4 * ty + i, c = j0 + 4 * tx + j;
      if (r < rows && c < cols) {
        double* out = C + (size_t)r * ldc + c;
        if (accumulate)
          *out -= acc[i][j];
        else
          *out = negate ? -acc[i][j] : acc[i][j];
      }
    }
}

// row-major scratch operators -> panels; grid.x = front, grid.y strides over the entries
__global__ void pack_kernel(const FactorFront* __restrict__ fronts, const double* __restrict__ Gs, const double* __restrict__ Bs,
                            double* __restrict__ fwd, double* __restrict__ bwd) {
  const FactorFront F = fronts[blockIdx.x];
  const int np = F.np, nb = F.nb, m = np + nb;
  const long long stride = (long long)gridDim.y * blockDim.x;
  const long long first = (long long)blockIdx.y * blockDim.x + threadIdx.x;
  {
    const double* G = Gs + F.Gs_off;
    const int P = 1 << F.fwd_log2P;
    const long long ps = (long long)F.fwd_cols_pad << F.fwd_log2P;
    for (long long e = first; e < (long long)nb * np; e += stride) {
      const int i = (int)(e / np), j = (int)(e % np);
      fwd[F.fwd_off + (long long)(i >> F.fwd_log2P) * ps + ((long long)j << F.fwd_log2P) + (i & (P - 1))] = G[e];
    }
  }
  {
    const double* B = Bs + F.Bs_off;
    const int P = 1 << F.bwd_log2P;
    const long long ps = (long long)F.bwd_cols_pad << F.bwd_log2P;
    for (long long e = first; e < (long long)np * m; e += stride) {
      const int i = (int)(e / m), j = (int)(e % m);
      bwd[F.bwd_off + (long long)(i >> F.bwd_log2P) * ps + ((long long)j << F.bwd_log2P) + (i & (P - 1))] = B[e];
    }
  }
}

__global__ void check_info_kernel(const int* info, int* error) {
  if (*info != 0) atomicExch(error, 3);
}

__global__ void set_identity_kernel(double* B, int np, int ld) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < np * np) B[(size_t)(e / np) * ld + (e % np)] = (e / np == e % np) ? 1.0 : 0.0;
}

void cublas_check(cublasStatus_t s, const char* what) {
  if (s != CUBLAS_STATUS_SUCCESS) throw StatusError(PECS_ERR_CUDA, std::string(what) + ": cuBLAS status " + std::to_string((int)s));
}
void cusolver_check(cusolverStatus_t s, const char* what) {
  if (s != CUSOLVER_STATUS_SUCCESS)
    throw StatusError(PECS_ERR_CUDA, std::string(what) + ": cuSOLVER status " + std::to_string((int)s));
}

// The large fronts of a level are independent: they are factorised on kLanes streams side by side, each with its own
// cuSOLVER / cuBLAS handle, work space, pivot vector and info word.  (Round 1 ran them one after the other on the
// default stream: getrf of a 200..1300-row block fills a fraction of the device, and the numeric factorisation was 60 %
// of pecs_ctx_create at cfg3: 1.5-2.3 s per carrier system, profiles/r02_setup_timing_q.log.)  The lanes are BLOCKING
// streams: the legacy default stream, on which the assembly / extend-add / small-front / pack kernels of the level run,
// orders itself with them in both directions, so no events are needed.
constexpr int kLanes = 8;
struct Lane {
  cudaStream_t stream = nullptr;
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
  DeviceBuffer<double> work;
  DeviceBuffer<int> ipiv, info;
};
struct Handles {
  Lane lane[kLanes];
  Handles() {
    // every lane's handles on a thread of its own: cusolverDnCreate + cublasCreate take 0.1-0.2 s each
    int device = 0;
    PECS_CUDA(cudaGetDevice(&device));
    std::string failure[kLanes];
    std::thread maker[kLanes];
    for (int k = 0; k < kLanes; ++k)
      maker[k] = std::thread([this, k, device, &failure] {
        try {
          Lane& l = lane[k];
          PECS_CUDA(cudaSetDevice(device));
          PECS_CUDA(cudaStreamCreate(&l.stream)); // blocking with respect to the legacy default stream, on purpose
          cublas_check(cublasCreate(&l.blas), "cublasCreate");
          cusolver_check(cusolverDnCreate(&l.solver), "cusolverDnCreate");
          cublas_check(cublasSetStream(l.blas, l.stream), "cublasSetStream");
          cusolver_check(cusolverDnSetStream(l.solver, l.stream), "cusolverDnSetStream");
          l.info.resize(1);
        } catch (const std::exception& e) {
          failure[k] = e.what();
        }
      });
    for (std::thread& t : maker) t.join();
    for (const std::string& f : failure)
      if (!f.empty()) throw StatusError(PECS_ERR_CUDA, "factorize_device: " + f);
  }
};

// cuSOLVER / cuBLAS initialisation costs about a second: one set of handles per process and device
Handles& handles_of_current_device() {
  static std::mutex guard;
  static Handles* handles_of[64] = {};
  int device = 0;
  PECS_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(guard);
  if (!handles_of[device & 63]) handles_of[device & 63] = new Handles();
  return *handles_of[device & 63];
}

} // namespace

void warm_factor_handles() { handles_of_current_device(); }

bool device_factorization_enabled() {
  const char* e = std::getenv("PECS_B200_HOST_FACTOR");
  return !(e && e[0] == '1');
}

void factorize_device(const SolvePlan& plan, const CsrMatrix& A, double* d_fwd, double* d_bwd) {
  const CsrMatrix Ap = permute_csr(A, plan.perm, false), Apt = permute_csr(A, plan.perm, true);
  factorize_device(plan, Ap, Apt, d_fwd, d_bwd);
}

void factorize_device(const SolvePlan& plan, const CsrMatrix& Ap, const CsrMatrix& Apt, double* d_fwd, double* d_bwd) {
  DeviceBuffer<int> bd_index_buf;
  bd_index_buf.upload(plan.bd_index.data(), std::max<size_t>(plan.bd_index.size(), 1));
  const int* d_bd_index = bd_index_buf.get();
  DeviceBuffer<int> rp, col, trp, tcol, d_error(1);
  DeviceBuffer<double> val, tval;
  rp.upload(Ap.row_ptr);
  col.upload(Ap.col);
  val.upload(Ap.val);
  trp.upload(Apt.row_ptr);
  tcol.upload(Apt.col);
  tval.upload(Apt.val);
  d_error.zero();
  PECS_CUDA(cudaMemset(d_fwd, 0, (size_t)std::max<int64_t>(plan.fwd_entries, 2) * sizeof(double)));
  PECS_CUDA(cudaMemset(d_bwd, 0, (size_t)std::max<int64_t>(plan.bwd_entries, 2) * sizeof(double)));

  Handles& h = handles_of_current_device();
  for (Lane& l : h.lane)
    if (l.ipiv.size() < (size_t)std::max(plan.max_np, 1)) l.ipiv.resize((size_t)std::max(plan.max_np, 1));
  // work buffers sized ONCE for the largest level: an allocation or a release per level is a device-wide
  // synchronisation and the most expensive host call of the whole loop
  DeviceBuffer<double> Fcur, Fchild, Gs, Bs;
  DeviceBuffer<FactorFront> d_cur, d_child;
  DeviceBuffer<int> cl, small_list;
  {
    size_t F_max = 1, Gs_max = 1, Bs_max = 1, cl_max = 1, fronts_max = 1;
    for (const std::vector<int>& lvl : plan.levels) {
      size_t F_total = 0, Gs_total = 0, Bs_total = 0, cl_total = 0;
      for (int f : lvl) {
        const Front& F = plan.fronts[f];
        F_total += (size_t)(F.np + F.nb) * (F.np + F.nb);
        Gs_total += (size_t)F.nb * F.np;
        Bs_total += (size_t)F.np * (F.np + F.nb);
        for (int c = 0; c < 2; ++c)
          if (F.child[c] >= 0) cl_total += (size_t)plan.fronts[F.child[c]].nb;
      }
      F_max = std::max(F_max, F_total);
      Gs_max = std::max(Gs_max, Gs_total);
      Bs_max = std::max(Bs_max, Bs_total);
      cl_max = std::max(cl_max, cl_total);
      fronts_max = std::max(fronts_max, lvl.size());
    }
    Fcur.resize(F_max);
    Fchild.resize(F_max);
    Gs.resize(Gs_max);
    Bs.resize(Bs_max);
    cl.resize(cl_max);
    d_cur.resize(fronts_max);
    d_child.resize(fronts_max);
    small_list.resize(fronts_max);
  }
  std::vector<int> index_in_level(plan.fronts.size(), -1);

  // PECS_B200_SETUP_TIMING=2: wall-clock per level on stderr (each level ends with a device synchronisation anyway)
  const char* timing_env = std::getenv("PECS_B200_SETUP_TIMING");
  const bool level_timing = timing_env && std::atoi(timing_env) >= 2;
  for (int d = (int)plan.levels.size() - 1; d >= 0; --d) {
    const auto level_begin = std::chrono::steady_clock::now();
    const std::vector<int>& lvl = plan.levels[d];
    std::vector<FactorFront> ff(lvl.size());
    std::vector<int> small, large;
    long long F_total = 0, cl_total = 0, Gs_total = 0, Bs_total = 0;
    int max_child_nb = 0;
    for (size_t k = 0; k < lvl.size(); ++k) {
      const Front& F = plan.fronts[lvl[k]];
      FactorFront& x = ff[k];
      x.np = F.np;
      x.nb = F.nb;
      x.p0 = F.p0;
      x.bd_off = F.bd_off;
      x.fwd_off = F.fwd.off;
      x.bwd_off = F.bwd.off;
      x.fwd_log2P = F.fwd.log2P;
      x.fwd_cols_pad = F.fwd.cols_pad;
      x.bwd_log2P = F.bwd.log2P;
      x.bwd_cols_pad = F.bwd.cols_pad;
      x.Gs_off = Gs_total;
      Gs_total += (long long)F.nb * F.np;
      x.Bs_off = Bs_total;
      Bs_total += (long long)F.np * (F.np + F.nb);
      x.F_off = F_total;
      F_total += (long long)(F.np + F.nb) * (F.np + F.nb);
      for (int c = 0; c < 2; ++c) {
        x.child[c] = F.child[c] >= 0 ? index_in_level[F.child[c]] : -1;
        x.cl_off[c] = cl_total;
        if (F.child[c] >= 0) {
          cl_total += plan.fronts[F.child[c]].nb;
          max_child_nb = std::max(max_child_nb, plan.fronts[F.child[c]].nb);
        }
      }
      (F.np <= kSmallFrontMaxNp ? small : large).push_back((int)k);
    }
    for (size_t k = 0; k < lvl.size(); ++k) index_in_level[lvl[k]] = (int)k;
    if (F_total > 0) PECS_CUDA(cudaMemsetAsync(Fcur.get(), 0, (size_t)F_total * sizeof(double)));
    PECS_CUDA(cudaMemcpy(d_cur.get(), ff.data(), ff.size() * sizeof(FactorFront), cudaMemcpyHostToDevice));
    const int nfl = (int)lvl.size();
    assemble_kernel<<<nfl, 128>>>(d_cur.get(), d_bd_index, rp.get(), col.get(), val.get(), trp.get(), tcol.get(),
                                  tval.get(), Fcur.get(), d_error.get());
    if (cl_total > 0) {
      for (int c = 0; c < 2; ++c) {
        child_local_kernel<<<nfl, 128>>>(d_cur.get(), d_child.get(), d_bd_index, c, cl.get(), d_error.get());
        const dim3 grid(nfl, std::max(1, std::min(max_child_nb, 1024)));
        extend_add_kernel<<<grid, 256>>>(d_cur.get(), d_child.get(), c, cl.get(), Fchild.get(), Fcur.get());
      }
    }
    PECS_CUDA(cudaGetLastError());
    if (!small.empty()) {
      PECS_CUDA(cudaMemcpy(small_list.get(), small.data(), small.size() * sizeof(int), cudaMemcpyHostToDevice));
      small_front_kernel<<<(int)small.size(), 256>>>(d_cur.get(), small_list.get(), Fcur.get(), Gs.get(), Bs.get(),
                                                     d_error.get());
      int tiles0 = 0, tiles1 = 0;
      for (int k : small) {
        const int tp = (ff[k].np + 63) / 64, tb = (ff[k].nb + 63) / 64;
        tiles0 = std::max(tiles0, 2 * tp * tb);
        tiles1 = std::max(tiles1, tb * tb);
      }
      if (tiles0 > 0) {
        small_front_products_kernel<<<dim3((unsigned)small.size(), (unsigned)tiles0), 256>>>(d_cur.get(), small_list.get(), 0,
                                                                                              Fcur.get(), Gs.get(), Bs.get());
        small_front_products_kernel<<<dim3((unsigned)small.size(), (unsigned)tiles1), 256>>>(d_cur.get(), small_list.get(), 1,
                                                                                              Fcur.get(), Gs.get(), Bs.get());
      }
      PECS_CUDA(cudaGetLastError());
    }
    if (!large.empty()) {
      // work space of the lanes: sized once per level (a resize is a device-wide synchronisation)
      int lwork_max = 0;
      for (int k : large) {
        int lwork = 0;
        cusolver_check(cusolverDnDgetrf_bufferSize(h.lane[0].solver, ff[k].np, ff[k].np, Fcur.get() + ff[k].F_off,
                                                   ff[k].np + ff[k].nb, &lwork),
                       "getrf_bufferSize");
        lwork_max = std::max(lwork_max, lwork);
      }
      for (Lane& l : h.lane)
        if (l.work.size() < (size_t)lwork_max) l.work.resize((size_t)lwork_max);
    }
    // Every front costs about a millisecond of HOST time (getrf alone is dozens of launches), far more than the device
    // needs for it: the lanes are fed by one host thread each.  Ordering: the level's assembly kernels were issued to
    // the legacy default stream before the threads start and the pack kernel is issued after they have joined, so the
    // blocking lane streams order themselves with both.
    if (!large.empty()) {
      int device = 0;
      PECS_CUDA(cudaGetDevice(&device));
      const int n_feeders = (int)std::min<size_t>(kLanes, large.size());
      std::exception_ptr failure[kLanes];
      auto feed = [&](int lane_index) {
        try {
          PECS_CUDA(cudaSetDevice(device));
          Lane& l = h.lane[lane_index];
          for (size_t q = (size_t)lane_index; q < large.size(); q += (size_t)n_feeders) {
            const int k = large[q];
            const FactorFront& x = ff[k];
            const int np = x.np, nb = x.nb, m = np + nb;
            double* M = Fcur.get() + x.F_off;
            double* B = Bs.get() + x.Bs_off;
            double* G = Gs.get() + x.Gs_off;
            // col-major view of the row-major F_PP (ld m) is F_PP^T; getrf/getrs on it give (F_PP^T)^-1 = Inv^T, whose
            // col-major storage with leading dimension ldb IS the row-major Inv with row stride ldb: it lands in the table.
            const int ldb = m, ldf = np;
            cusolver_check(cusolverDnDgetrf(l.solver, np, np, M, m, l.work.get(), l.ipiv.get(), l.info.get()), "getrf");
            check_info_kernel<<<1, 1, 0, l.stream>>>(l.info.get(), d_error.get());
            set_identity_kernel<<<(np * np + 255) / 256, 256, 0, l.stream>>>(B, np, ldb);
            cusolver_check(cusolverDnDgetrs(l.solver, CUBLAS_OP_N, np, np, M, m, l.ipiv.get(), B, ldb, l.info.get()), "getrs");
            if (nb > 0) {
              const double one = 1.0, zero = 0.0, minus = -1.0;
              // (-H)^T = -F_PB^T Inv^T : C(nb x np, ld ldb) = -A(nb x np: F_PB memory, ld m) * B(np x np: Inv memory, ld ldb)
              cublas_check(cublasDgemm(l.blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, np, np, &minus, M + np, m, B, ldb, &zero, B + np, ldb),
                           "dgemm H");
              // G^T = Inv^T F_BP^T : C(np x nb, ld ldf) = A(np x np: Inv memory, ld ldb) * B(np x nb: F_BP memory, ld m)
              cublas_check(cublasDgemm(l.blas, CUBLAS_OP_N, CUBLAS_OP_N, np, nb, np, &one, B, ldb, M + (size_t)np * m, m, &zero,
                                       G, ldf),
                           "dgemm G");
              // U^T = F_BB^T - F_PB^T G^T : C(nb x nb, ld m) -= A(nb x np: F_PB memory, ld m) * B(np x nb: G memory, ld ldf)
              cublas_check(cublasDgemm(l.blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, nb, np, &minus, M + np, m, G, ldf, &one,
                                       M + (size_t)np * m + np, m),
                           "dgemm U");
            }
          }
        } catch (...) {
          failure[lane_index] = std::current_exception();
        }
      };
      std::thread feeder[kLanes];
      for (int t = 1; t < n_feeders; ++t) feeder[t] = std::thread(feed, t);
      feed(0);
      for (int t = 1; t < n_feeders; ++t) feeder[t].join();
      for (int t = 0; t < n_feeders; ++t)
        if (failure[t]) std::rethrow_exception(failure[t]);
    }
    {
      long long max_entries = 1;
      for (const FactorFront& x : ff) max_entries = std::max(max_entries, (long long)x.np * (x.np + x.nb));
      const dim3 grid(nfl, (unsigned)std::max<long long>(1, std::min<long long>(512, max_entries / 1024)));
      pack_kernel<<<grid, 256>>>(d_cur.get(), Gs.get(), Bs.get(), d_fwd, d_bwd);
    }
    PECS_CUDA(cudaDeviceSynchronize());
    int herr = 0;
    PECS_CUDA(cudaMemcpy(&herr, d_error.get(), sizeof(int), cudaMemcpyDeviceToHost));
    if (herr == 3) throw StatusError(PECS_ERR_SINGULAR, "factorize_device: singular pivot block in a small front");
    if (herr != 0) throw StatusError(PECS_ERR_INTERNAL, "factorize_device: matrix entry outside the symbolic front structure");
    if (level_timing) {
      int max_np = 0, max_nb = 0;
      for (const FactorFront& x : ff) {
        max_np = std::max(max_np, x.np);
        max_nb = std::max(max_nb, x.nb);
      }
      std::fprintf(stderr, "factorize_device: level %2d: %6zu fronts (%zu by cuSOLVER), np <= %4d, nb <= %4d, %7.2f ms\n", d,
                   ff.size(), large.size(), max_np, max_nb,
                   1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - level_begin).count());
    }
    // this level becomes the child level of the next one
    std::swap(Fcur, Fchild);
    std::swap(d_cur, d_child);
  }
}

} // namespace pecs
