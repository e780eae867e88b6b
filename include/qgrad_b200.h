/* qgrad_b200 -- C ABI of the B200-native qgrad hot path (libqgrad_b200.so).
 *
 * The reference (qgrad/qgrad) is pure Python on JAX: it has no FFI of its own.  The interface
 * this path sits behind is the Python module API of qgrad/qgrad_qutip.py
 * (docs/source/api.rst:11-40); each entry point below names the reference function whose
 * arithmetic it replaces (file:line relative to the reference checkout).  A JAX host binds these
 * through XLA-FFI (qgrad_b200/csrc/xla_ffi_shim.cc); the in-tree host
 * (qgrad_b200/qgrad_qutip.py) binds them with ctypes on raw device pointers.
 *
 * Conventions (all entry points)
 *  - plain pointers and sizes only; the library never owns caller memory.  Inputs, outputs and
 *    workspace are DEVICE pointers allocated by the caller, 16-byte aligned.
 *  - complex data is interleaved (re, im), row-major, batch-major: (batch, n, n) or (batch, n).
 *  - `dtype` selects complex64 (float pairs) or complex128 (double pairs); real side arrays
 *    (angles, losses) use the matching real type.
 *  - stream-ordered and re-entrant: every call takes a cudaStream_t (as void*), enqueues
 *    kernels and returns without synchronising.  Global mutable state is limited to two caches behind
 *    mutexes / atomics: per-device kernel attributes, the planner switch (qg_set_plan_rigorous), and the stream-K scratch of the large GEMMs (~77 MB of
 *    accumulator images per (device, stream) that runs a large-dimension product).  The scratch is
 *    allocated on the FIRST such call on a stream -- one cudaMalloc pair, which synchronises the device
 *    once and is not allowed during CUDA-graph capture: run one call on a stream before capturing on it
 *    (the in-tree learner warms up on its capture stream) -- and lives until qg_release_scratch().
 *  - return value: 0 ok; <0 invalid argument (QG_ERR_*); >0 a cudaError_t from a launch.
 *    Nothing throws.  There is no CPU fallback: without a CUDA device every compute entry
 *    point fails.
 *  - gradients use the Wirtinger-conjugate ("torch") convention: for a real loss L and complex
 *    x, xbar = dL/dRe(x) + i dL/dIm(x).  jax.grad's convention is the complex conjugate of
 *    that; `adjoint = QG_ADJ_TRANS` on the expm pull-backs yields JAX's VJP directly.
 */
#ifndef QGRAD_B200_H
#define QGRAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_VERSION 100 /* 0.1.0 */

typedef enum { QG_C64 = 0, QG_C128 = 1 } qg_dtype;
/* pull-back flavour of Y = exp(A): CONJ -> Abar = L(A^H, Ybar) (torch); TRANS -> L(A^T, Ybar) (JAX) */
typedef enum { QG_ADJ_CONJ = 0, QG_ADJ_TRANS = 1, QG_ADJ_SKEW_BODY = 2 } qg_adjoint;

enum {
  QG_OK = 0,
  QG_ERR_ARG = -1,       /* null pointer / negative size / unknown dtype */
  QG_ERR_DIM = -2,       /* dimension outside what the kernel family supports */
  QG_ERR_WORKSPACE = -3, /* workspace too small (see *_workspace_bytes) */
  QG_ERR_NO_DEVICE = -4  /* no CUDA device: there is no CPU path */
};

int qg_version(void);
const char* qg_status_string(int status);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t qg_launch_count(void);
/* Planner switch (process-wide; also QG_PLAN_RIGOROUS=1 in the environment).  For (skew-)Hermitian inputs with n > 32 the
 * number of squarings normally uses a power-iteration ESTIMATE of rho(A) (two fixed start vectors, 3 % margin, never
 * below the rigorous floor d_k n^(-1/2k)).  on = 1 drops the estimate: s comes from ||A||_1 and d_k = ||A^k||_1^(1/k)
 * alone, both rigorous upper bounds on rho -- at most one squaring level (3 of ~31 products) more on the learning step's
 * Hamiltonians.  Returns the previous setting. */
int qg_set_plan_rigorous(int on);
/* complex64 products of the n > 32 expm chain and of qg_gemm (process-wide; also QG_C64_TENSOR=0 in the environment).
 * on = 1 (default): launches whose shapes qualify (M, N >= 128, K a multiple of 16, no skipped tile rows) run on the
 * tcgen05 tensor cores as split-TF32 products with float32 round-to-nearest accumulation (csrc/gemm_tc.cu); on = 0: every
 * complex64 product runs on the FP32 FMA kernel.  Returns the previous setting. */
int qg_set_c64_tensor(int on);
/* frees the cached stream-K scratch of every (device, stream); no library call may be in flight.  0 or a cudaError_t. */
int qg_release_scratch(void);

/* ---------------------------------------------------------------- batched matrix exponential
 * Y[b] = exp(A[b]),  b < batch,  A (batch,n,n).  Pade scaling-and-squaring, order and number of
 * squarings s chosen per matrix on the device: from ||A||_1 in general; for matrices found to be
 * Hermitian or skew-Hermitian (normal: ||A||_2 = rho(A) <= ||A^6||_1^(1/6)) from the norm of the
 * top even power the Pade evaluation forms anyway and, for n > 32, a power-iteration estimate of
 * rho(A) on that power (inflated by 3 %), which is what expm(-iHt) inputs hit.
 * After a forward call with n > 32 the first `batch` int32 of the workspace (rounded up to a
 * 256-byte address) hold s[b], the next `batch` int32 a flag (1: the matrix needed more than
 * max_squarings squarings and its results are NaN).
 * Replaces: scipy.linalg.expm as called by squeeze (qgrad/qgrad_qutip.py:266-280) and
 * make_unitary (examples/Unitary-Learning-no-qgrad.py:112-141); north star Unitary(H)(t).
 *   n <= 32 : one fused kernel, matrix resident in shared memory, no workspace needed.
 *   n  > 32 : tiled complex GEMM chain over `workspace` (>= qg_expm_workspace_bytes).
 * max_squarings: upper bound on s the caller guarantees (squaring launches are enqueued up to
 *   this bound and predicated per matrix on the device, so no host sync is needed).  Matrices
 *   that would need more are flagged NaN.  Pass <= 0 for the default (QG_DEFAULT_MAX_SQUARINGS).
 */
#define QG_DEFAULT_MAX_SQUARINGS 20
#define QG_EXPM_FWD 0
#define QG_EXPM_FWD_VJP 1
size_t qg_expm_workspace_bytes(int dtype, int n, int64_t batch, int mode, int max_squarings);

int qg_expm(int dtype, const void* A, void* Y, int64_t batch, int n,
            void* workspace, size_t workspace_bytes, int max_squarings, void* stream);

/* Fused value + reverse-mode pull-back: Y = exp(A) and Abar = L(op(A), Ybar) via the
 * block-triangular recurrences (each product costs three n x n products, one inverse shared).
 * Y may be NULL (pull-back only).  Replaces the finite-difference der_cost of
 * examples/Unitary-Learning-no-qgrad.py:213-241 and supplies the squeeze gradient the
 * reference lists as TODO (qgrad/qgrad_qutip.py:265). */
int qg_expm_fwd_vjp(int dtype, const void* A, const void* Ybar, void* Y, void* Abar,
                    int64_t batch, int n, int adjoint,
                    void* workspace, size_t workspace_bytes, int max_squarings, void* stream);

/* Two-pass form for callers whose Ybar depends on Y (a training step): the forward pass keeps
 * its powers, inverse and squaring chain in `workspace`; the backward pass consumes them and
 * runs only the dual recurrences.  The workspace must be untouched in between. */
int qg_expm_fwd_save(int dtype, const void* A, void* Y, int64_t batch, int n,
                     void* workspace, size_t workspace_bytes, int max_squarings, void* stream);
int qg_expm_bwd_saved(int dtype, const void* A, const void* Ybar, void* Abar,
                      int64_t batch, int n, int adjoint,
                      void* workspace, size_t workspace_bytes, int max_squarings, void* stream);
/* adjoint = QG_ADJ_SKEW_BODY (qg_expm_bwd_saved only, n > 32, every A[b] skew-Hermitian, i.e. A = -i t H with H
 * Hermitian -- the north star's Unitary(H)(t)): `Ybar` holds Z[b] = Ybar[b] Y[b]^H, the cotangent in the body
 * frame, and Abar[b] receives the SKEW-HERMITIAN PART of the pull-back, int_0^1 e^{-xA} (Z - Z^H)/2 e^{xA} dx.
 * That part is all a real parameter of a Hermitian generator sees (<Abar, dA/dtheta> with dA/dtheta skew-Hermitian),
 * and every product of the dual recurrences then has a Hermitian or skew-Hermitian result, computed on the upper
 * triangle of tiles only: 12.4 instead of 18 n^3 products at s = 4.  The learning step forms Z as G Phi^H
 * (Phi = Y Psi) without ever building Ybar.  A matrix that is not skew-Hermitian comes back NaN. */

/* dst[b] (cols, rows) = src[b]^H for src (batch, rows, cols): the Phi^H operand of the body-frame cotangent */
int qg_conj_transpose(int dtype, const void* src, void* dst, int64_t batch, int rows, int cols, void* stream);

/* ---------------------------------------------------------------- expectation values
 * expect (qgrad/qgrad_qutip.py:166-203).  out (batch) complex.
 * ket:  out[b] = <psi_b| O |psi_b> = vdot(psi, O psi)      psi (batch,n), O (batch|1, n, n)
 * dm :  out[b] = tr(rho_b O) = sum_ij rho_ij O_ji           rho (batch,n,n)
 * oper_batched = 0 shares one operator across the batch. */
int qg_expect_ket_fwd(int dtype, const void* oper, const void* ket, void* out,
                      int64_t batch, int n, int oper_batched, void* stream);
/* ketbar (batch,n) and/or operbar (batch|1,n,n; accumulated when shared) may be NULL */
int qg_expect_ket_bwd(int dtype, const void* oper, const void* ket, const void* outbar,
                      void* ketbar, void* operbar, int64_t batch, int n, int oper_batched,
                      void* stream);
int qg_expect_dm_fwd(int dtype, const void* oper, const void* rho, void* out,
                     int64_t batch, int n, int oper_batched, void* stream);
int qg_expect_dm_bwd(int dtype, const void* oper, const void* rho, const void* outbar,
                     void* rhobar, void* operbar, int64_t batch, int n, int oper_batched,
                     void* stream);

/* ---------------------------------------------------------------- fidelities
 * fidelity (qgrad/qgrad_qutip.py:11-64).  out (batch) real.
 * ket: out[b] = |a_b^H b_b|^2.     dm: sqrt(1 - (0.5 sum_i |rho1_ii - rho2_ii|)^2), the
 * reference's elementwise-abs surrogate, reproduced as written (only the diagonals enter). */
int qg_fidelity_ket_fwd(int dtype, const void* a, const void* b, void* out,
                        int64_t batch, int n, void* stream);
int qg_fidelity_ket_bwd(int dtype, const void* a, const void* b, const void* outbar,
                        void* abar, void* bbar, int64_t batch, int n, void* stream);
int qg_fidelity_dm_fwd(int dtype, const void* rho1, const void* rho2, void* out,
                       int64_t batch, int n, void* stream);
int qg_fidelity_dm_bwd(int dtype, const void* rho1, const void* rho2, const void* outbar,
                       void* rho1bar, void* rho2bar, int64_t batch, int n, void* stream);

/* ---------------------------------------------------------------- displacement operator
 * Displace.__call__ (qgrad/qgrad_qutip.py:241-261).  evecs (n,n) REAL and evals (n) REAL are the
 * eigensystem of the tridiagonal T built in Displace.__init__ (:225-239), computed once per n
 * by the host.  alpha (batch) complex.  D (batch,n,n):
 *   D_jk = conj(s_j) s_k e^{i phi (j-k)} sum_m V_jm e^{i r lambda_m} V_km,  alpha = r e^{i phi},
 *   s_j = i^(j%2);  alpha == 0 takes the phase factor as 1 (:253-257). */
int qg_displace_fwd(int dtype, const void* evecs, const void* evals, const void* alpha, void* D,
                    int64_t batch, int n, void* stream);
/* alphabar (batch) complex from Dbar (batch,n,n); zero at alpha == 0 */
int qg_displace_bwd(int dtype, const void* evecs, const void* evals, const void* alpha,
                    const void* Dbar, void* alphabar, int64_t batch, int n, void* stream);
/* y_b = D(alpha_b) x_b without forming D (three O(n^2) products): coherent (:302-316) and the
 * D.x steps of apply_blocks (examples/SNAP_gates.py:150-182).  x, y (batch,n). */
int qg_displace_apply_fwd(int dtype, const void* evecs, const void* evals, const void* alpha,
                          const void* x, void* y, int64_t batch, int n, void* stream);
int qg_displace_apply_bwd(int dtype, const void* evecs, const void* evals, const void* alpha,
                          const void* x, const void* ybar, void* alphabar, void* xbar,
                          int64_t batch, int n, void* stream);

/* ---------------------------------------------------------------- Givens-product unitary
 * Unitary.__call__ (qgrad/qgrad_qutip.py:500-555): U = diag(e^{i omega}) prod_{i=2..N}
 * prod_{j=1..i-1} R(-theta_k, -phi_k; (i-1, j-1)), applied as O(N) two-column updates.
 * thetas, phis (batch, N(N-1)/2) real; omegas (batch, N) real; U (batch,N,N). */
int qg_givens_unitary_fwd(int dtype, const void* thetas, const void* phis, const void* omegas,
                          void* U, int64_t batch, int N, void* stream);
int qg_givens_unitary_bwd(int dtype, const void* thetas, const void* phis, const void* omegas,
                          const void* Ubar, void* thetasbar, void* phisbar, void* omegasbar,
                          int64_t batch, int N, void* stream);
/* _make_rot (qgrad/qgrad_qutip.py:424-459): identity with 4 entries replaced.
 * params (batch,2) = (theta, phi) real; R (batch,N,N). */
int qg_make_rot_fwd(int dtype, const void* params, void* R, int64_t batch, int N, int i, int j,
                    void* stream);
int qg_make_rot_bwd(int dtype, const void* params, const void* Rbar, void* paramsbar,
                    int64_t batch, int N, int i, int j, void* stream);
/* rot (examples/Qubit_Rotation.py:40-69): RZ(omega) RY(theta) RZ(phi), 2x2.
 * angles (batch,3) = (phi, theta, omega) real; R (batch,2,2). */
int qg_rot_zyz_fwd(int dtype, const void* angles, void* R, int64_t batch, void* stream);
int qg_rot_zyz_bwd(int dtype, const void* angles, const void* Rbar, void* anglesbar,
                   int64_t batch, void* stream);
/* Fused Qubit_Rotation cost (examples/Qubit_Rotation.py:72-92): F_b = |<0| rot(angles_b) ket_b>|^2
 * = (1 + <sigma_z>)/2 and its gradient w.r.t. the three angles in one pass.  ket (batch|1, 2).
 * grad (batch,3) may be NULL. */
int qg_rot_fidelity_fwd_grad(int dtype, const void* angles, const void* ket, int ket_batched,
                             void* fid, void* grad, int64_t batch, void* stream);

/* Fused Unitary-learning cost (examples/Unitary-Learning-qgrad.py:123-151) for batches of parameter sets that share
 * the n_kets training pairs kets_in / kets_out (n_kets, N):
 *   loss[b] = 1 - (1/n_kets) sum_k | Re < out_k | Unitary(N)(thetas_b, phis_b, omegas_b) | in_k > |
 * and, when the three bar pointers are given, its gradient w.r.t. the parameters in the same pass (U, U psi and the
 * cotangent of U stay in registers).  N <= 16. */
int qg_unitary_cost_fwd_grad(int dtype, const void* thetas, const void* phis, const void* omegas,
                             const void* kets_in, const void* kets_out, int n_kets, void* loss,
                             void* thetasbar, void* phisbar, void* omegasbar, int64_t batch, int N,
                             void* stream);

/* Fused SNAP-gate cost (examples/SNAP_gates.py:150-182 apply_blocks, :231-248 cost) for batches of parameter sets that
 * share the initial and the target state (n):
 *   x = prod_{t < n_blocks} [ D(-alphas[b,t]) diag(e^{i thetas[b,t,:]}) D(alphas[b,t]) ] init,   cost[b] = 1 - |<target|x>|^2
 * and, when alphabar (batch, n_blocks) complex and thetabar (batch, n_blocks, n) real are given, its gradient in the same
 * kernel (state and cotangent stay in shared memory; the reverse sweep un-applies the unitary factors, no tape).
 * thetas rows are full length n (the example pads them with pad_thetas, :84-100).  evecs / evals as qg_displace_fwd. */
int qg_snap_cost_fwd_grad(int dtype, const void* evecs, const void* evals, const void* alphas, const void* thetas,
                          const void* init, const void* target, int n_blocks, void* cost, void* alphabar, void* thetabar,
                          int64_t batch, int n, void* stream);

/* ---------------------------------------------------------------- Hamiltonian-learning step
 * The cost of examples/Unitary-Learning-qgrad.py:123-151 / Unitary-Learning-no-qgrad.py:165-191
 * with U = exp(-i t H(theta)), H_s(theta) = sum_p ctrl[s,p] theta_p P_p, P_p Pauli strings
 * (xmask, zmask): <r|P|c> = i^popcount(x&z) (-1)^popcount(c&z) [r == c^x].
 * These are the pieces around qg_expm_fwd_save / qg_expm_bwd_saved; the host sequences them and
 * all-reduces thetabar across ranks (one NCCL all-reduce of n_terms reals). */
/* A (n_set,n,n) = -i t sum_p ctrl[s,p] theta_p P_p;  n = 2^n_qubits */
int qg_pauli_hamiltonian(int dtype, const void* theta, const void* ctrl, const int32_t* xmask,
                         const int32_t* zmask, int n_terms, int n_set, int n_qubits, double t,
                         void* A, void* stream);
/* thetabar[p] (+)= sum_s ctrl[s,p] Re< dA_s/dtheta_p , Abar_s > */
int qg_pauli_project(int dtype, const void* Abar, const void* ctrl, const int32_t* xmask,
                     const int32_t* zmask, int n_terms, int n_set, int n_qubits, double t,
                     void* thetabar, int accumulate, void* stream);
/* C (batch, m, n) = A (batch, m, k) . B (batch, k, n); ld* are row pitches in elements; used for
 * Phi = U Psi_in and Ubar = G Psi_in^H.  k a multiple of 16; m, n multiples of 64 -- except complex64 with m >= 128,
 * n >= 96 even and even pitches, which the tcgen05 kernel takes at any size (ragged tiles are zero-filled / predicated). */
int qg_gemm(int dtype, const void* A, const void* B, void* C, int64_t batch, int m, int n, int k,
            int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
            int64_t strideC, void* stream);
/* overlaps o_k = <out_k|phi_k> (columns), loss_partial += sum_k |Re o_k|, and the cotangent
 * G[:,k] = -sign(Re o_k)/m_total * out_k.  phi, out, G (batch, n, m_local) column = ket. */
int qg_overlap_loss(int dtype, const void* phi, const void* out, void* G, void* loss_partial,
                    int64_t batch, int n, int m_local, double inv_m_total, void* stream);
/* Adam (jax.experimental.optimizers.adam defaults), in place on device, step index i */
int qg_adam_step(int dtype, void* x, void* m, void* v, const void* grad, int64_t count,
                 double lr, double b1, double b2, double eps, int64_t i, void* stream);
/* The same update with the step index i read from -- and advanced in -- DEVICE memory (*step), so that one captured
 * launch serves every replay of a CUDA graph of the whole training step.  With loss_out != NULL it also writes the
 * step's loss, *loss_out = 1 - *loss_partial * inv_m_total (loss_partial: the all-reduced accumulator of
 * qg_overlap_loss).  Both loss pointers NULL: Adam only. */
int qg_adam_step_device(int dtype, void* x, void* m, void* v, const void* grad, int64_t count,
                        double lr, double b1, double b2, double eps, int64_t* step,
                        const void* loss_partial, double inv_m_total, void* loss_out, void* stream);

/* ---------------------------------------------------------------- Hermitian eigen-problems, true mixed-state fidelity
 * Density-matrix completeness next to the reference's surrogates.  One CTA per matrix, cyclic
 * two-sided Jacobi with complex rotations in shared memory; n <= 64 (QG_ERR_DIM above), any batch.
 * qg_eigh: A (batch,n,n) Hermitian -> evals (batch,n) real ascending, evecs (batch,n,n) columns (NULL: values only).
 *   Replaces jnp.linalg.eigh in Displace.__init__ (qgrad/qgrad_qutip.py:237) and the eigenvalue test of isdm (:392-395). */
int qg_eigh(int dtype, const void* A, void* evals, void* evecs, int64_t batch, int n, void* stream);
/* out[b] = principal square root of the Hermitian positive semi-definite A[b] (eigenvalues below 0 by round-off clamp
 * to 0): what scipy.linalg.sqrtm, imported and unused at qgrad_qutip.py:7, is for. */
int qg_sqrtm_psd(int dtype, const void* A, void* out, int64_t batch, int n, void* stream);
/* out[b] (real) = (tr sqrt( sqrt(rho_b) sigma_b sqrt(rho_b) ))^2, the Uhlmann-Jozsa fidelity of two density matrices
 * -- |<a|b>|^2 for pure states, i.e. the mixed-state continuation of qg_fidelity_ket_fwd; the reference's
 * _fidelity_dm (:50-64) is a diagonal-only surrogate of it (kept, as written, in qg_fidelity_dm_fwd).  Value only. */
int qg_fidelity_uhlmann(int dtype, const void* rho, const void* sigma, void* out, int64_t batch, int n, void* stream);
/* stats (batch,4) real = { #entries with A_rc != conj(A_cr) (exact, as isherm :367), Re tr A, Im tr A,
 * max |A_rc - conj(A_cr)| }: the device half of isherm / isdm (:357-396).  Any n. */
int qg_herm_stats(int dtype, const void* A, void* stats, int64_t batch, int n, void* stream);

/* ---------------------------------------------------------------- device-side generators
 * Synthetic states / Hamiltonians / angles generated where they are used, replacing the host-side draws of
 * rand_ket / rand_dm / rand_unitary (qgrad/qgrad_qutip.py:558-620) and of the examples' QuTiP rand_ket and tenpy GUE
 * (examples/Unitary-Learning-qgrad.py:63,88; examples/Unitary-Learning-no-qgrad.py:106-107).  Counter-based
 * Philox4x32-10: element i of unit u is a pure function of (seed, u, i), so `first_*` lets every rank generate ITS
 * slice of one global sequence (the same data at every world size) and results do not depend on the launch geometry.
 * JAX's threefry stream is not reproduced (the reference's tests pin properties of these draws, not values). */
#define QG_RAND_HAAR 0      /* normalised complex Gaussian: Haar measure on pure states */
#define QG_RAND_REFERENCE 1 /* rand_ket as written (:570-573): uniform(key) + 1j uniform(key) with ONE key, re == im, normalised */
/* out (batch, n): ket first_ket + b of the sequence `seed` */
int qg_rand_ket(int dtype, void* out, int64_t batch, int n, uint64_t seed, uint64_t first_ket, int kind, void* stream);
/* H (batch, n, n) = scale (X + X^H) / 2, X iid complex standard normal (GUE; scale = 1/sqrt(n): the named H/sqrt(n)) */
int qg_rand_hermitian(int dtype, void* H, int64_t batch, int n, uint64_t seed, uint64_t first_matrix, double scale,
                      void* stream);
/* out (count) REAL of the dtype's real type, uniform in [lo, hi): rand_unitary's N^2 angles in [0, 2 pi) (:612-620) */
int qg_rand_uniform(int dtype, void* out, int64_t count, uint64_t seed, double lo, double hi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QGRAD_B200_H */
