// extern "C" entry points of libqgrad_b200.so: argument validation and dispatch only.
// Declarations and the reference function each entry replaces: include/qgrad_b200.h.
#include <map>
#include <mutex>
#include <utility>
#include "common.cuh"

namespace qg {
std::atomic<int64_t> g_launch_count{0};

// expm_small.cu
int expm_small(int dtype, const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n,
               int adjoint, bool vjp, cudaStream_t st);
// expm_large.cu
size_t expm_large_workspace_bytes(int dtype, int n, int64_t batch, int mode, int max_sq);
int expm_large_fwd(int dtype, const void* A, void* Y, int64_t batch, int n, void* ws, size_t ws_bytes,
                   int max_sq, bool keep, cudaStream_t st);
int expm_set_plan_rigorous(int on);
int gemm_tc_set_enabled(int on);   // gemm_tc.cu
int expm_large_bwd(int dtype, const void* Ybar, void* Abar, int64_t batch, int n, int adjoint, void* ws,
                   size_t ws_bytes, int max_sq, cudaStream_t st);

int gemm_plain(int dtype, const void* A, const void* B, void* C, int64_t batch, int m, int n, int k, int64_t lda,
               int64_t ldb, int64_t ldc, int64_t sA, int64_t sB, int64_t sC, cudaStream_t st);
// state_ops.cu
int state_op(int dtype, int op, const void* p0, const void* p1, const void* p2, void* o0, void* o1, int64_t batch,
             int n, int ob, cudaStream_t st);
// builders.cu
int builder_displace(int dtype, bool bwd, const void* V, const void* lam, const void* alpha, void* D, const void* Dbar,
                     void* abar, int64_t batch, int n, cudaStream_t st);
int builder_displace_apply(int dtype, bool bwd, const void* V, const void* lam, const void* alpha, const void* x, void* y,
                           const void* ybar, void* abar, void* xbar, int64_t batch, int n, cudaStream_t st);
int builder_givens(int dtype, bool bwd, const void* th, const void* ph, const void* om, void* U, const void* Ubar, void* tb,
                   void* pb, void* ob, int64_t batch, int N, cudaStream_t st);
int builder_unitary_cost(int dtype, const void* th, const void* ph, const void* om, const void* kin, const void* kout, int M,
                         void* loss, void* tb, void* pb, void* ob, int64_t batch, int N, cudaStream_t st);
int builder_misc(int dtype, int op, const void* p0, const void* p1, void* o0, void* o1, int64_t batch, int N, int i, int j,
                 int flag, cudaStream_t st);
// learn.cu
int learn_hamiltonian(int dtype, const void* theta, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set,
                      int nq, double t, void* A, cudaStream_t st);
int learn_project(int dtype, const void* Abar, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set, int nq,
                  double t, void* tb, int accumulate, cudaStream_t st);
int learn_overlap(int dtype, const void* phi, const void* out, void* G, void* loss, int64_t batch, int n, int m_local,
                  double inv_m, cudaStream_t st);
int learn_conj_transpose(int dtype, const void* src, void* dst, int64_t batch, int rows, int cols, cudaStream_t st);
int learn_adam(int dtype, void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2,
               double eps, int64_t i, cudaStream_t st);
int learn_adam_device(int dtype, void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2,
                      double eps, int64_t* step, const void* loss_partial, double inv_m_total, void* loss_out, cudaStream_t st);
// rand.cu
int rand_op(int dtype, int what, void* out, int64_t batch, int n, uint64_t seed, uint64_t first, int kind, double a, double b,
            cudaStream_t st);
// snap.cu
int snap_cost(int dtype, const void* V, const void* lam, const void* alphas, const void* thetas, const void* init, const void* target,
              void* cost, void* alphabar, void* thetabar, int64_t batch, int n, int nblocks, cudaStream_t st);
// spectral.cu
int spectral_op(int dtype, int mode, const void* in0, const void* in1, void* out0, void* out1, int64_t batch, int n, cudaStream_t st);

// stream-K scratch of the GEMM kernel (gemm.cuh): accumulator images + ready flags + the CTA ticket, one per
// (device, stream), allocated on first use (one synchronous cudaMalloc pair -- the only non-stream-ordered thing this
// library ever does, see qgrad_b200.h; the flags are cleared ON THE STREAM) and kept until qg_release_scratch()
// (the library may cache device buffers behind a mutex).
namespace gemm {
struct SplitScratch { void* partial; unsigned* counters; };
constexpr int kSplitSlotsApi = 148 * 8;
static std::mutex g_scratch_mu;
static std::map<std::pair<int, cudaStream_t>, SplitScratch> g_scratch_pool;

int device_sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = cached[dev & 63].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  cached[dev & 63].store(v, std::memory_order_relaxed);
  return v;
}

int split_scratch(cudaStream_t st, size_t image_bytes, SplitScratch* out) {
  int dev = 0;
  QG_RETURN_IF_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  auto key = std::make_pair(dev, st);
  auto it = g_scratch_pool.find(key);
  if (it == g_scratch_pool.end()) {
    SplitScratch sc{nullptr, nullptr};
    const size_t cbytes = (size_t)(kSplitSlotsApi + 2) * sizeof(unsigned);
    cudaError_t e = cudaMalloc(&sc.partial, (size_t)kSplitSlotsApi * image_bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMalloc((void**)&sc.counters, cbytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(sc.counters, 0, cbytes, st);   // ordered before the first kernel on `st`
    if (e != cudaSuccess) { cudaFree(sc.partial); if (sc.counters) cudaFree(sc.counters); return (int)e; }
    it = g_scratch_pool.emplace(key, sc).first;
  }
  *out = it->second;
  return QG_OK;
}
static int release_scratch() {
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  int dev0 = 0;
  cudaGetDevice(&dev0);
  cudaError_t first = cudaSuccess;
  for (auto& kv : g_scratch_pool) {
    cudaSetDevice(kv.first.first);
    cudaError_t e = cudaFree(kv.second.partial); if (e != cudaSuccess && first == cudaSuccess) first = e;
    e = cudaFree(kv.second.counters); if (e != cudaSuccess && first == cudaSuccess) first = e;
  }
  g_scratch_pool.clear();
  cudaSetDevice(dev0);
  return (int)first;
}
}  // namespace gemm

static inline bool bad_dtype(int d) { return d != QG_C64 && d != QG_C128; }
static inline int norm_max_sq(int m) { return m <= 0 ? QG_DEFAULT_MAX_SQUARINGS : (m > 60 ? 60 : m); }
constexpr int kSmallMax = 32;
constexpr int64_t kLargeChunkMax = 16384;   // grid.z budget of the GEMM launches
// how many matrices of the batch the given workspace covers at once (0: not even one)
static inline int64_t chunk_size(int dtype, int n, int64_t batch, int mode, int ms, const void* ws, size_t ws_bytes) {
  const size_t per = expm_large_workspace_bytes(dtype, n, 1, mode, ms);
  if (!ws || ws_bytes < per) return 0;
  const size_t fixed = expm_large_workspace_bytes(dtype, n, 0, mode, ms);
  int64_t chunk = (int64_t)((ws_bytes - fixed) / (per - fixed));
  if (chunk > batch) chunk = batch;
  if (chunk > kLargeChunkMax) chunk = kLargeChunkMax;
  return chunk;
}
}  // namespace qg

using namespace qg;

extern "C" {

int qg_version(void) { return QG_VERSION; }

const char* qg_status_string(int status) {
  switch (status) {
    case QG_OK: return "ok";
    case QG_ERR_ARG: return "invalid argument";
    case QG_ERR_DIM: return "dimension not supported by this kernel family";
    case QG_ERR_WORKSPACE: return "workspace too small";
    case QG_ERR_NO_DEVICE: return "no CUDA device (qgrad_b200 has no CPU path)";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown status";
  }
}

int64_t qg_launch_count(void) { return g_launch_count.load(); }

int qg_release_scratch(void) { return gemm::release_scratch(); }

int qg_set_plan_rigorous(int on) { return expm_set_plan_rigorous(on); }

int qg_set_c64_tensor(int on) { return gemm_tc_set_enabled(on); }

size_t qg_expm_workspace_bytes(int dtype, int n, int64_t batch, int mode, int max_squarings) {
  if (bad_dtype(dtype) || n <= 0 || batch <= 0) return 0;
  if (n <= kSmallMax) return 0;
  return expm_large_workspace_bytes(dtype, n, batch, mode, norm_max_sq(max_squarings));
}

int qg_expm(int dtype, const void* A, void* Y, int64_t batch, int n, void* ws, size_t ws_bytes,
            int max_squarings, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!A || !Y) return QG_ERR_ARG;
  if (n <= kSmallMax) return expm_small(dtype, A, nullptr, Y, nullptr, batch, n, 0, false, as_stream(stream));
  const int ms = norm_max_sq(max_squarings);
  int64_t chunk = chunk_size(dtype, n, batch, QG_EXPM_FWD, ms, ws, ws_bytes);
  if (chunk < 1) return QG_ERR_WORKSPACE;
  const size_t esz = (dtype == QG_C128 ? 16 : 8) * (size_t)n * n;
  for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
    const int64_t nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    int rc = expm_large_fwd(dtype, (const char*)A + b0 * esz, (char*)Y + b0 * esz, nb, n, ws, ws_bytes, ms, false, as_stream(stream));
    if (rc) return rc;
  }
  return QG_OK;
}

int qg_expm_fwd_vjp(int dtype, const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n,
                    int adjoint, void* ws, size_t ws_bytes, int max_squarings, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG;
  if (adjoint != QG_ADJ_CONJ && adjoint != QG_ADJ_TRANS) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!A || !Ybar || !Abar) return QG_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (n <= kSmallMax) return expm_small(dtype, A, Ybar, Y, Abar, batch, n, adjoint, true, st);
  const int ms = norm_max_sq(max_squarings);
  // chunk over the batch if the workspace only covers part of it
  int64_t chunk = chunk_size(dtype, n, batch, QG_EXPM_FWD_VJP, ms, ws, ws_bytes);
  if (chunk < 1) return QG_ERR_WORKSPACE;
  const size_t esz = (dtype == QG_C128 ? 16 : 8) * (size_t)n * n;
  for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
    const int64_t nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    const char* a = (const char*)A + b0 * esz;
    char* y = Y ? (char*)Y + b0 * esz : nullptr;
    int rc = expm_large_fwd(dtype, a, y, nb, n, ws, ws_bytes, ms, true, st);
    if (rc) return rc;
    rc = expm_large_bwd(dtype, (const char*)Ybar + b0 * esz, (char*)Abar + b0 * esz, nb, n, adjoint, ws, ws_bytes, ms, st);
    if (rc) return rc;
  }
  return QG_OK;
}

int qg_expm_fwd_save(int dtype, const void* A, void* Y, int64_t batch, int n, void* ws, size_t ws_bytes,
                     int max_squarings, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!A || !Y) return QG_ERR_ARG;
  if (n <= kSmallMax) return expm_small(dtype, A, nullptr, Y, nullptr, batch, n, 0, false, as_stream(stream));
  return expm_large_fwd(dtype, A, Y, batch, n, ws, ws_bytes, norm_max_sq(max_squarings), true, as_stream(stream));
}

int qg_expm_bwd_saved(int dtype, const void* A, const void* Ybar, void* Abar, int64_t batch, int n, int adjoint,
                      void* ws, size_t ws_bytes, int max_squarings, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG;
  if (adjoint != QG_ADJ_CONJ && adjoint != QG_ADJ_TRANS && adjoint != QG_ADJ_SKEW_BODY) return QG_ERR_ARG;
  if (adjoint == QG_ADJ_SKEW_BODY && n <= kSmallMax) return QG_ERR_ARG;       // GEMM-chain path only
  if (batch == 0) return QG_OK;
  if (!A || !Ybar || !Abar) return QG_ERR_ARG;
  // small dims keep nothing: the fused kernel recomputes the value recurrences from A
  if (n <= kSmallMax) return expm_small(dtype, A, Ybar, nullptr, Abar, batch, n, adjoint, true, as_stream(stream));
  return expm_large_bwd(dtype, Ybar, Abar, batch, n, adjoint, ws, ws_bytes, norm_max_sq(max_squarings), as_stream(stream));
}

// ---------------------------------------------------------------- expect / fidelity
// (an empty batch is a no-op whatever the pointers are: empty device tensors have null data pointers)
#define QG_STATE_GUARD(cond) if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG; if (batch == 0) return QG_OK; if (cond) return QG_ERR_ARG

int qg_expect_ket_fwd(int dtype, const void* oper, const void* ket, void* out, int64_t batch, int n, int ob, void* stream) {
  QG_STATE_GUARD(!oper || !ket || !out);
  return state_op(dtype, 0, oper, ket, nullptr, out, nullptr, batch, n, ob, as_stream(stream));
}
int qg_expect_ket_bwd(int dtype, const void* oper, const void* ket, const void* outbar, void* ketbar, void* operbar,
                      int64_t batch, int n, int ob, void* stream) {
  QG_STATE_GUARD(!oper || !ket || !outbar);
  return state_op(dtype, 1, oper, ket, outbar, ketbar, operbar, batch, n, ob, as_stream(stream));
}
int qg_expect_dm_fwd(int dtype, const void* oper, const void* rho, void* out, int64_t batch, int n, int ob, void* stream) {
  QG_STATE_GUARD(!oper || !rho || !out);
  return state_op(dtype, 2, oper, rho, nullptr, out, nullptr, batch, n, ob, as_stream(stream));
}
int qg_expect_dm_bwd(int dtype, const void* oper, const void* rho, const void* outbar, void* rhobar, void* operbar,
                     int64_t batch, int n, int ob, void* stream) {
  QG_STATE_GUARD(!oper || !rho || !outbar);
  return state_op(dtype, 3, oper, rho, outbar, rhobar, operbar, batch, n, ob, as_stream(stream));
}
int qg_fidelity_ket_fwd(int dtype, const void* a, const void* b, void* out, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!a || !b || !out);
  return state_op(dtype, 4, a, b, nullptr, out, nullptr, batch, n, 1, as_stream(stream));
}
int qg_fidelity_ket_bwd(int dtype, const void* a, const void* b, const void* outbar, void* abar, void* bbar, int64_t batch,
                        int n, void* stream) {
  QG_STATE_GUARD(!a || !b || !outbar);
  return state_op(dtype, 5, a, b, outbar, abar, bbar, batch, n, 1, as_stream(stream));
}
int qg_fidelity_dm_fwd(int dtype, const void* r1, const void* r2, void* out, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!r1 || !r2 || !out);
  return state_op(dtype, 6, r1, r2, nullptr, out, nullptr, batch, n, 1, as_stream(stream));
}
int qg_fidelity_dm_bwd(int dtype, const void* r1, const void* r2, const void* outbar, void* r1bar, void* r2bar,
                       int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!r1 || !r2 || !outbar);
  return state_op(dtype, 7, r1, r2, outbar, r1bar, r2bar, batch, n, 1, as_stream(stream));
}

// ---------------------------------------------------------------- builders
int qg_displace_fwd(int dtype, const void* evecs, const void* evals, const void* alpha, void* D, int64_t batch, int n,
                    void* stream) {
  QG_STATE_GUARD(!evecs || !evals || !alpha || !D);
  return builder_displace(dtype, false, evecs, evals, alpha, D, nullptr, nullptr, batch, n, as_stream(stream));
}
int qg_displace_bwd(int dtype, const void* evecs, const void* evals, const void* alpha, const void* Dbar, void* alphabar,
                    int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!evecs || !evals || !alpha || !Dbar || !alphabar);
  return builder_displace(dtype, true, evecs, evals, alpha, nullptr, Dbar, alphabar, batch, n, as_stream(stream));
}
int qg_displace_apply_fwd(int dtype, const void* evecs, const void* evals, const void* alpha, const void* x, void* y,
                          int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!evecs || !evals || !alpha || !x || !y);
  return builder_displace_apply(dtype, false, evecs, evals, alpha, x, y, nullptr, nullptr, nullptr, batch, n, as_stream(stream));
}
int qg_displace_apply_bwd(int dtype, const void* evecs, const void* evals, const void* alpha, const void* x,
                          const void* ybar, void* alphabar, void* xbar, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!evecs || !evals || !alpha || !x || !ybar);
  return builder_displace_apply(dtype, true, evecs, evals, alpha, x, nullptr, ybar, alphabar, xbar, batch, n, as_stream(stream));
}
int qg_givens_unitary_fwd(int dtype, const void* thetas, const void* phis, const void* omegas, void* U, int64_t batch, int N,
                          void* stream) {
  const int n = N;
  QG_STATE_GUARD(!thetas || !phis || !omegas || !U);
  return builder_givens(dtype, false, thetas, phis, omegas, U, nullptr, nullptr, nullptr, nullptr, batch, N, as_stream(stream));
}
int qg_givens_unitary_bwd(int dtype, const void* thetas, const void* phis, const void* omegas, const void* Ubar,
                          void* thetasbar, void* phisbar, void* omegasbar, int64_t batch, int N, void* stream) {
  const int n = N;
  QG_STATE_GUARD(!thetas || !phis || !omegas || !Ubar || !thetasbar || !phisbar || !omegasbar);
  return builder_givens(dtype, true, thetas, phis, omegas, nullptr, Ubar, thetasbar, phisbar, omegasbar, batch, N, as_stream(stream));
}
int qg_make_rot_fwd(int dtype, const void* params, void* R, int64_t batch, int N, int i, int j, void* stream) {
  const int n = N;
  QG_STATE_GUARD(!params || !R || i < 0 || j < 0 || i >= N || j >= N);
  return builder_misc(dtype, 0, params, nullptr, R, nullptr, batch, N, i, j, 0, as_stream(stream));
}
int qg_make_rot_bwd(int dtype, const void* params, const void* Rbar, void* paramsbar, int64_t batch, int N, int i, int j,
                    void* stream) {
  const int n = N;
  QG_STATE_GUARD(!params || !Rbar || !paramsbar || i < 0 || j < 0 || i >= N || j >= N);
  return builder_misc(dtype, 1, params, Rbar, paramsbar, nullptr, batch, N, i, j, 0, as_stream(stream));
}
int qg_rot_zyz_fwd(int dtype, const void* angles, void* R, int64_t batch, void* stream) {
  const int n = 2;
  QG_STATE_GUARD(!angles || !R);
  return builder_misc(dtype, 2, angles, nullptr, R, nullptr, batch, 2, 0, 0, 0, as_stream(stream));
}
int qg_rot_zyz_bwd(int dtype, const void* angles, const void* Rbar, void* anglesbar, int64_t batch, void* stream) {
  const int n = 2;
  QG_STATE_GUARD(!angles || !Rbar || !anglesbar);
  return builder_misc(dtype, 3, angles, Rbar, anglesbar, nullptr, batch, 2, 0, 0, 0, as_stream(stream));
}
int qg_rot_fidelity_fwd_grad(int dtype, const void* angles, const void* ket, int ket_batched, void* fid, void* grad,
                             int64_t batch, void* stream) {
  const int n = 2;
  QG_STATE_GUARD(!angles || !ket || !fid);
  return builder_misc(dtype, 4, angles, ket, fid, grad, batch, 2, 0, 0, ket_batched, as_stream(stream));
}

int qg_unitary_cost_fwd_grad(int dtype, const void* thetas, const void* phis, const void* omegas, const void* kets_in,
                             const void* kets_out, int n_kets, void* loss, void* thetasbar, void* phisbar, void* omegasbar,
                             int64_t batch, int N, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || N <= 0 || n_kets <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!thetas || !phis || !omegas || !kets_in || !kets_out || !loss) return QG_ERR_ARG;
  if ((thetasbar || phisbar || omegasbar) && !(thetasbar && phisbar && omegasbar)) return QG_ERR_ARG;
  return builder_unitary_cost(dtype, thetas, phis, omegas, kets_in, kets_out, n_kets, loss, thetasbar, phisbar, omegasbar, batch, N,
                              as_stream(stream));
}

int qg_snap_cost_fwd_grad(int dtype, const void* evecs, const void* evals, const void* alphas, const void* thetas,
                          const void* init, const void* target, int n_blocks, void* cost, void* alphabar, void* thetabar,
                          int64_t batch, int n, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0 || n_blocks <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!evecs || !evals || !alphas || !thetas || !init || !target || !cost) return QG_ERR_ARG;
  if ((alphabar != nullptr) != (thetabar != nullptr)) return QG_ERR_ARG;
  return snap_cost(dtype, evecs, evals, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, n_blocks, as_stream(stream));
}

// ---------------------------------------------------------------- learning step pieces
int qg_pauli_hamiltonian(int dtype, const void* theta, const void* ctrl, const int32_t* xmask, const int32_t* zmask,
                         int n_terms, int n_set, int n_qubits, double t, void* A, void* stream) {
  if (bad_dtype(dtype) || !theta || !ctrl || !xmask || !zmask || !A || n_terms <= 0 || n_set <= 0 || n_qubits < 1 || n_qubits > 14)
    return QG_ERR_ARG;
  return learn_hamiltonian(dtype, theta, ctrl, xmask, zmask, n_terms, n_set, n_qubits, t, A, as_stream(stream));
}
int qg_pauli_project(int dtype, const void* Abar, const void* ctrl, const int32_t* xmask, const int32_t* zmask, int n_terms,
                     int n_set, int n_qubits, double t, void* thetabar, int accumulate, void* stream) {
  if (bad_dtype(dtype) || !Abar || !ctrl || !xmask || !zmask || !thetabar || n_terms <= 0 || n_set <= 0 || n_qubits < 1 || n_qubits > 14)
    return QG_ERR_ARG;
  return learn_project(dtype, Abar, ctrl, xmask, zmask, n_terms, n_set, n_qubits, t, thetabar, accumulate, as_stream(stream));
}
int qg_gemm(int dtype, const void* A, const void* B, void* C, int64_t batch, int m, int n, int k, int64_t lda, int64_t ldb,
            int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || m <= 0 || n <= 0 || k <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!A || !B || !C) return QG_ERR_ARG;
  return gemm_plain(dtype, A, B, C, batch, m, n, k, lda, ldb, ldc, strideA, strideB, strideC, as_stream(stream));
}
int qg_conj_transpose(int dtype, const void* src, void* dst, int64_t batch, int rows, int cols, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || rows <= 0 || cols <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!src || !dst || src == dst) return QG_ERR_ARG;
  return learn_conj_transpose(dtype, src, dst, batch, rows, cols, as_stream(stream));
}

int qg_overlap_loss(int dtype, const void* phi, const void* out, void* G, void* loss_partial, int64_t batch, int n,
                    int m_local, double inv_m_total, void* stream) {
  if (bad_dtype(dtype) || !phi || !out || !G || !loss_partial || batch <= 0 || n <= 0 || m_local <= 0) return QG_ERR_ARG;
  return learn_overlap(dtype, phi, out, G, loss_partial, batch, n, m_local, inv_m_total, as_stream(stream));
}
int qg_adam_step(int dtype, void* x, void* m, void* v, const void* grad, int64_t count, double lr, double b1, double b2,
                 double eps, int64_t i, void* stream) {
  if (bad_dtype(dtype) || !x || !m || !v || !grad || count < 0) return QG_ERR_ARG;
  if (count == 0) return QG_OK;
  return learn_adam(dtype, x, m, v, grad, count, lr, b1, b2, eps, i, as_stream(stream));
}
int qg_adam_step_device(int dtype, void* x, void* m, void* v, const void* grad, int64_t count, double lr, double b1,
                        double b2, double eps, int64_t* step, const void* loss_partial, double inv_m_total, void* loss_out,
                        void* stream) {
  if (bad_dtype(dtype) || !x || !m || !v || !grad || !step || count < 0) return QG_ERR_ARG;
  if ((loss_out != nullptr) != (loss_partial != nullptr)) return QG_ERR_ARG;
  return learn_adam_device(dtype, x, m, v, grad, count, lr, b1, b2, eps, step, loss_partial, inv_m_total, loss_out, as_stream(stream));
}

// ---------------------------------------------------------------- Hermitian eigen-problems / density-matrix predicates
int qg_eigh(int dtype, const void* A, void* evals, void* evecs, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!A || !evals);
  return spectral_op(dtype, 0, A, nullptr, evals, evecs, batch, n, as_stream(stream));
}
int qg_sqrtm_psd(int dtype, const void* A, void* out, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!A || !out);
  return spectral_op(dtype, 1, A, nullptr, out, nullptr, batch, n, as_stream(stream));
}
int qg_fidelity_uhlmann(int dtype, const void* rho, const void* sigma, void* out, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!rho || !sigma || !out);
  return spectral_op(dtype, 2, rho, sigma, out, nullptr, batch, n, as_stream(stream));
}
int qg_herm_stats(int dtype, const void* A, void* stats, int64_t batch, int n, void* stream) {
  QG_STATE_GUARD(!A || !stats);
  return spectral_op(dtype, 3, A, nullptr, stats, nullptr, batch, n, as_stream(stream));
}

// ---------------------------------------------------------------- device-side generators
int qg_rand_ket(int dtype, void* out, int64_t batch, int n, uint64_t seed, uint64_t first_ket, int kind, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0 || (kind != QG_RAND_HAAR && kind != QG_RAND_REFERENCE)) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!out) return QG_ERR_ARG;
  return rand_op(dtype, 0, out, batch, n, seed, first_ket, kind, 0.0, 0.0, as_stream(stream));
}
int qg_rand_hermitian(int dtype, void* H, int64_t batch, int n, uint64_t seed, uint64_t first_matrix, double scale, void* stream) {
  if (bad_dtype(dtype) || batch < 0 || n <= 0) return QG_ERR_ARG;
  if (batch == 0) return QG_OK;
  if (!H) return QG_ERR_ARG;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {                      // grid.y budget
    const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    const size_t off = (size_t)b0 * n * n * (dtype == QG_C128 ? 16 : 8);
    int rc = rand_op(dtype, 1, (char*)H + off, nb, n, seed, first_matrix + (uint64_t)b0, 0, scale, 0.0, as_stream(stream));
    if (rc) return rc;
  }
  return QG_OK;
}
int qg_rand_uniform(int dtype, void* out, int64_t count, uint64_t seed, double lo, double hi, void* stream) {
  if (bad_dtype(dtype) || count < 0) return QG_ERR_ARG;
  if (count == 0) return QG_OK;
  if (!out) return QG_ERR_ARG;
  return rand_op(dtype, 2, out, count, 1, seed, 0, 0, lo, hi, as_stream(stream));
}

}  // extern "C"
