/*
 * pcq.h — C-ABI of libpcq: the B200 (sm_100a) polynomial-chaos hot path.
 *
 * This is the drop-in boundary for PyTUQ's PC path.  PyTUQ has no FFI of its
 * own (it is pure Python); every entry point below cites the reference Python
 * method whose arithmetic it replaces (paths relative to the PyTUQ tree,
 * src/pytuq/...).  The Python shim in pytuq_b200/ binds these with ctypes and
 * keeps the reference's PCRV / lreg signatures; INTEGRATION.md shows the stub a
 * PyTUQ maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - every function returns 0 on success, a negative pcq_status otherwise, and
 *     never throws; pcq_last_error() returns a thread-local description.
 *   - "_dev" entry points take DEVICE pointers and a CUDA stream handle
 *     (cudaStream_t passed as void*; NULL = legacy default stream) and are
 *     asynchronous with respect to the host.
 *   - the other entry points take HOST pointers, move data themselves
 *     (chunked H2D -> kernel -> D2H) and return when the result is in host
 *     memory.
 *   - all matrices are row-major (C order) float64, as numpy gives them.
 */
#ifndef PCQ_H_
#define PCQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCQ_VERSION 100

typedef enum pcq_status {
    PCQ_OK = 0,
    PCQ_ERR_ARG = -1,      /* bad argument (shape, null pointer, order out of range) */
    PCQ_ERR_CUDA = -2,     /* a CUDA runtime call failed (see pcq_last_error) */
    PCQ_ERR_NODEV = -3,    /* no CUDA device / device index out of range */
    PCQ_ERR_ALLOC = -4,    /* host or device allocation failed */
    PCQ_ERR_UNSUPPORTED = -5
} pcq_status;

typedef struct pcq_plan pcq_plan;

/* ---- library / device ---------------------------------------------------- */

int pcq_version(void);
const char* pcq_last_error(void);
/* number of visible CUDA devices, or a negative pcq_status */
int pcq_device_count(void);

/* ---- plan ------------------------------------------------------------------
 * A plan is the device-resident form of one (multi-index, 1-D families) pair,
 * i.e. of PCRV.mindices[jdim] + PCRV.PC1ds (rv/pcrv.py:84-115, 745-782).
 *
 *   sdim      s, germ dimension
 *   nterms    K, number of multi-index rows
 *   mi        K x s int32, row-major (host)   -- PCRV.mindices[jdim]
 *   pmax      max over dims of PCRV.maxOrd (rv/pcrv.py:176-184); every mi entry
 *             must be <= pmax
 *   rec_a     (pmax+1) x s, rec_a[n*s+i] = PC1d_i.a(n)  (rv/pcrv.py:755,762,769,776)
 *   rec_b     (pmax+1) x s, rec_b[n*s+i] = PC1d_i.b(n)  (rv/pcrv.py:756,763,770,777)
 *             both evaluated on the host by the caller with the reference's own
 *             Python expressions, so the constants are bit-identical
 *   p1_scale  s, phi_1(x) = x * p1_scale[i]  (1.0 for LU/HG/nHG, sqrt(3) for nLU;
 *             rv/pcrv.py:754,761,768,775)
 *   device    CUDA device ordinal the plan lives on
 */
int pcq_plan_create(pcq_plan** out, int device, int sdim, int nterms,
                    const int32_t* mi, int pmax,
                    const double* rec_a, const double* rec_b,
                    const double* p1_scale);
int pcq_plan_destroy(pcq_plan* plan);
/* queries: 0 sdim, 1 nterms, 2 pmax, 3 width (max non-zero orders per term),
 *          4 device, 5 downward-closed flag, 6 number of SMs, 7 specialised modes (bit 0 exact, bit 1 fast),
 *          8 modes being built in the background */
int64_t pcq_plan_info(const pcq_plan* plan, int what);

/* Opt-in: compiles (NVRTC, at run time) evalPC kernels specialised for this plan's multi-index, the 1-D tables in
 * registers, no shared-memory gathers.  modes: bit 0 exact -- the term list unrolled in k order with the reference's
 * rounding sequence (rv/pcrv.py:486-497): bit-identical to the generic kernels, 4.7x faster at K = 1771; bit 1 fast
 * -- Horner form over the multi-index trie, one fused multiply-add per term, terms re-ordered: rounding-level
 * differences only, 11.5x the generic kernel.  ~0.2 ms of compile time per term on 16 host threads, cubins cached on
 * disk ($PCQ_SPEC_CACHE_DIR), so it pays for propagation jobs and repeated evaluations (MCMC).  Once built,
 * pcq_eval_pc(_dev) (up to 8 outputs, no quadratic-form mode) and pcq_propagate(_dev) use them.
 * Returns PCQ_ERR_UNSUPPORTED when libnvrtc is not available; the generic kernels stay in use. */
int pcq_plan_specialize(pcq_plan* plan, int modes);
/* Same build on a background host thread: returns at once, evaluations keep using the generic kernels until the
 * specialised ones are installed (pcq_plan_info(plan, 7) reports the installed modes, 8 the modes under way).
 * A failed build (no NVRTC) leaves the generic kernels in use.  pcq_plan_destroy waits for the thread. */
int pcq_plan_specialize_async(pcq_plan* plan, int modes);
/* Debugging aid, host only (no device needed): writes the CUDA source that pcq_plan_specialize compiles for the fast
 * mode (Horner form over the multi-index trie, csrc/pcq_spec.cu) and trie positions [j0, j1) into buf; returns the
 * length of the source (call with buf = NULL to size the buffer), < 0 on bad arguments. */
int64_t pcq_debug_horner_source(int sdim, int nterms, const int32_t* mi, int pmax,
                                const double* rec_a, const double* rec_b, const double* p1_scale,
                                int j0, int j1, int spt, char* buf, int64_t bufsize);

/* ---- K1: design matrix -----------------------------------------------------
 * A[n,k] = prod_i phi^{(i)}_{mi[k,i]}(x[n,i])          PCRV.evalBases, rv/pcrv.py:413-442
 * with the 1-D values from the three-term recurrence    PC1d.__call__,   rv/pcrv.py:784-804
 * evaluated with the reference's rounding sequence ((a*x)*p_n)+(b*p_{n-1})
 * (no FMA contraction) and the product taken over the non-zero orders in
 * ascending dimension order, so A is bit-identical to the reference.
 *   x  N x s, leading dimension ldx (>= s)   A  N x K, leading dimension lda (>= K)
 */
int pcq_eval_bases_dev(pcq_plan* plan, const double* x_dev, int64_t n, int64_t ldx,
                       double* a_dev, int64_t lda, void* stream);
int pcq_eval_bases(pcq_plan* plan, const double* x_host, int64_t n, int64_t ldx,
                   double* a_host, int64_t lda);

/* ---- K2: surrogate evaluation (A never materialised) -------------------------
 * y[n,m] = sum_k cf[m,k] * prod_i phi(...)              PCRV.evalPC, rv/pcrv.py:469-498
 * mode 0 (exact): coefficient multiplied first, factors in dimension order,
 *   terms summed sequentially in k order — bit-identical to the reference.
 * mode 1 (fast): one basis product per term shared by all outputs and fused
 *   multiply-adds; differs from mode 0 by rounding only.
 * mode | 2 (quadratic form): instead of the M outputs, y[n] = sum_m (output m)^2 is stored
 *   (one column, ldy >= 1).  With cf = F^T, F a factor of the coefficient covariance
 *   (cf_cov = F F^T), this is the pushed-forward variance || Psi(x_n) F ||^2 of
 *   lreg.compute_stdev / predicta(msc=1) (lreg/lreg.py:112-143,164-198) without forming A.
 *   cf  M x K (row m = coefficients of output m), y  N x M with leading dimension ldy
 */
int pcq_eval_pc_dev(pcq_plan* plan, const double* x_dev, int64_t n, int64_t ldx,
                    const double* cf_dev, int m, double* y_dev, int64_t ldy,
                    int mode, void* stream);
int pcq_eval_pc(pcq_plan* plan, const double* x_host, int64_t n, int64_t ldx,
                const double* cf_host, int m, double* y_host, int64_t ldy, int mode);

/* ---- K3: fused sufficient statistics ------------------------------------------
 * With the augmented row  z_n = [ Psi(x_n) (K) | y_n (M) | 1 ]  of length
 * L = K + M + 1, computes the symmetric L x L matrix  S = sum_n z_n z_n^T :
 *     S[:K,:K]      = A^T A      anl.fita  lreg/anl.py:78 ; bcs_fit lreg/bcs.py:84,95,100
 *     S[:K,K:K+M]   = A^T Y      lreg/anl.py:83 ; lreg/bcs.py:83
 *     S[K:K+M,K:K+M]= Y^T Y      lreg/anl.py:57 ; lreg/bcs.py:229
 *     S[:K,L-1]     = A^T 1,  S[K:K+M,L-1] = sum_n y_n,  S[L-1,L-1] = N
 * The contraction runs on the FP64 tensor pipe (DMMA).  The design matrix is never
 * materialised as a whole: sample chunks of the augmented rows are generated on the
 * device into a fragment-ordered scratch buffer (<= 2 GB per device, PCQ_SYRK_CHUNK_MB)
 * and contracted by a TMA-fed SYRK kernel; with PCQ_GRAM_SYRK=0 they are generated tile
 * by tile in shared memory instead and nothing of A touches HBM.
 *   accumulate != 0 adds to the existing contents of s_out (for streaming and
 *   for sample-sharded partial sums), otherwise s_out is overwritten.
 *   y may be NULL when m == 0.
 *   s_out  L x L, row-major, leading dimension L.
 */
/* Moment form of K3 (csrc/pcq_moment.cu).  The entries of Psi^T Psi are linear combinations of the sums
 *     m_gamma = sum_n prod_i phi^{(i)}_{gamma_i}(x[n,i]),   |gamma| <= 2 * (highest total order of the multi-index),
 * through the linearisation phi_a phi_b = sum_k c_abk phi_k of each 1-D family (d = 20, p = 3: 230 230 moments instead of
 * 1 569 106 Gram entries).  Once a plan has the recurrence tables up to that order, pcq_suffstats(_dev) computes the
 * moments as staircase GEMMs on the FP64 tensor pipe (N |Gamma| multiply-adds instead of N K^2 / 2, operands generated
 * in shared memory) and reconstructs S from them, whenever that is estimated faster than the pack + SYRK form and
 * n >= PCQ_MOMENT_MIN_N (32768).  Same S up to rounding (1e-15 relative to sqrt(S_ii S_jj)); it replaces the same
 * reference lines (lreg/anl.py:78,83, lreg/bcs.py:83-100, lreg/lreg.py:328-339).  PCQ_GRAM_MOMENT=0 disables it, =1
 * uses it for every n.  The rule also keeps the dense form where the reconstruction would amplify rounding errors: Hermite
 * sets of total order >= 7 (moments up to He_14 dwarf the entries they rebuild; PCQ_MOMENT_MAX_AMP, default 500).
 *   nord       >= 2 * max_k sum_i mi[k,i]
 *   rec_a_ext  (nord+1) x s, rec_a_ext[n*s+i] = PC1d_i.a(n);  rec_b_ext likewise   (rv/pcrv.py:755-777)
 */
int pcq_plan_enable_moments(pcq_plan* plan, int nord, const double* rec_a_ext, const double* rec_b_ext);
/* Query: info[0..6] = moments, units, DMMA multiply-adds executed per sample, estimated cycles per sample and SM, usable
 * flag, 1 if pcq_suffstats(_dev) would take the moment form for n samples and m outputs right now, bytes per sample of
 * the 1-D value table the form keeps in scratch.  All zero when the plan has no moment tables. */
int pcq_moment_info(pcq_plan* plan, int m, int64_t n, double* info);   /* info[7] = 1: more than 8 outputs, hybrid route */

/* Host helper of the staging pipelines (the callers of lreg.fit hand over pageable numpy arrays, lreg/lreg.py:62-71):
 * rows x width doubles from src (row stride ld_src) to dst (row stride ld_dst), both HOST memory, split over
 * PCQ_HOST_THREADS threads (default 16) when the block is at least 8 MB.  Pageable -> pinned at ~24 GB/s on the 16-core
 * GPU box where one thread reaches ~10. */
int pcq_host_copy_rows(double* dst, int64_t ld_dst, const double* src, int64_t ld_src, int64_t width, int64_t rows);
/* Debugging aid, host only (no device needed): plans the moment form for this multi-index and m outputs, executes the
 * planned units on the CPU exactly as the kernel does and reconstructs S (L x L, L = nterms + m + 1) into s_out (may be
 * NULL: plan only).  info[0..7] = moments, units, estimated cycles per 16-sample block, multiply-adds per sample,
 * functions / sub-functions of the largest unit, usable flag, tree nodes. */
int pcq_debug_moments_host(int sdim, int nterms, const int32_t* mi, int nord, const double* rec_a_ext,
                           const double* rec_b_ext, const double* p1_scale, const double* x, int64_t n, int64_t ldx,
                           const double* y, int64_t ldy, int m, double* s_out, double* info);

int64_t pcq_suffstats_dim(const pcq_plan* plan, int m);   /* returns L */
int pcq_suffstats_dev(pcq_plan* plan, const double* x_dev, int64_t n, int64_t ldx,
                      const double* y_dev, int64_t ldy, int m,
                      double* s_dev, int accumulate, void* stream);
int pcq_suffstats(pcq_plan* plan, const double* x_host, int64_t n, int64_t ldx,
                  const double* y_host, int64_t ldy, int m, double* s_host);

/* ---- K6: quadratic forms on the tensor pipe ---------------------------------------
 * v[n] = Psi(x_n)^T C Psi(x_n)  for a symmetric K x K matrix C (row-major, leading dimension ldc; only its
 * upper triangle is read).  With C = cf_cov this is the pushed-forward variance of lreg.predict /
 * compute_stdev (lreg/lreg.py:112-143, 164-198) for a full coefficient covariance: N K^2 flops on DMMA
 * (triangular GEMM + row-wise product), the design matrix only ever exists as a packed chunk in scratch.
 * v has stride ldv. */
int pcq_quadform_dev(pcq_plan* plan, const double* x_dev, int64_t n, int64_t ldx,
                     const double* c_dev, int64_t ldc, double* v_dev, int64_t ldv, void* stream);
int pcq_quadform(pcq_plan* plan, const double* x_host, int64_t n, int64_t ldx,
                 const double* c_host, int64_t ldc, double* v_host, int64_t ldv);

/* ---- the same forms on a design matrix handed over by the caller (no plan) ---------------------------------
 * lreg.predicta(Amat, msc) / compute_stdev(Amat) / predicta_samples(Amat) (lreg/lreg.py:112-198; the call pattern of
 * apps/uqpc/uq_pc.py:313-315), where the reference forms Amat @ cf and || Amat chol(cf_cov) ||_row with numpy:
 *   y (n x m, ldy)  = A (n x kcols, lda) CF^T,  CF m x kcols (ldcf)        skipped when m == 0
 *   v[i * ldv]      = a_i^T C a_i,  C kcols x kcols symmetric (ldc; upper triangle read)   skipped when c == NULL
 * Few vectors (m < 48, e.g. the mean) are HBM-bound row dot products; many vectors (predictive samples, the two
 * GEMMs of msc = 2) and the quadratic form run on the FP64 tensor pipe through K6's kernels, the matrix chunk packed
 * by pcq_qf_pack_matrix_kernel. */
int pcq_matrix_forms_dev(int device, const double* a_dev, int64_t n, int64_t lda, int kcols,
                         const double* cf_dev, int64_t ldcf, int m, double* y_dev, int64_t ldy,
                         const double* c_dev, int64_t ldc, double* v_dev, int64_t ldv, void* stream);
int pcq_matrix_forms(int device, const double* a_host, int64_t n, int64_t lda, int kcols,
                     const double* cf_host, int64_t ldcf, int m, double* y_host, int64_t ldy,
                     const double* c_host, int64_t ldc, double* v_host, int64_t ldv);

/* Residual pass of one output: with r_n = y_n - sum_k cf[k] Psi_k(x_n) (surrogate through K2, fast mode),
 *   sums_dev[0] (+)= sum_n r_n^2      bcs_fit's final noise estimate, lreg/bcs.py:229
 *   sums_dev[1] (+)= sum_n y_n r_n    anl._compute_sigmahatsq, lreg/anl.py:61
 * Needed when the fit is so close that  y^T y - 2 cf^T b + cf^T G cf  on the statistics has cancelled
 * (noise-free training data).  cf has K entries (zeros for unused terms); y is one column with stride ldy.
 * Partial sums are combined in a fixed order (deterministic). */
int pcq_residual_ss_dev(pcq_plan* plan, const double* x_dev, int64_t n, int64_t ldx,
                        const double* y_dev, int64_t ldy, const double* cf_dev,
                        double* sums_dev, int accumulate, void* stream);
/* host-pointer form: sums_host[0..1] are overwritten */
int pcq_residual_ss(pcq_plan* plan, const double* x_host, int64_t n, int64_t ldx,
                    const double* y_host, int64_t ldy, const double* cf_host, double* sums_host);

/* ---- K4: sufficient statistics from a materialised design matrix ---------------
 * Same S as above but z_n = [ A[n,:] | y_n | 1 ] with A given (lreg.fita(Amat, y),
 * lreg/lreg.py:328-339, lreg/anl.py:69-105, lreg/bcs.py:37-52).  kcols = A's column count.
 */
int pcq_gram_dev(int device, const double* a_dev, int64_t n, int64_t lda, int kcols,
                 const double* y_dev, int64_t ldy, int m,
                 double* s_dev, int accumulate, void* stream);
int pcq_gram(int device, const double* a_host, int64_t n, int64_t lda, int kcols,
             const double* y_host, int64_t ldy, int m, double* s_host);

/* ---- K5: forward propagation with on-device germ sampling -----------------------
 * Draws n germ samples (U(-1,1) where germ_kind[i]==0, N(0,1) where ==1;
 * PC1d.sample, rv/pcrv.py:757,764,771,778) from a counter-based Philox4x32-10
 * stream keyed by (seed, sample index, dimension), pushes them through evalPC
 * and reduces per-output power sums on the fly (PCRV.sample, rv/pcrv.py:511-525
 * followed by the moment estimates the callers take).  Nothing of size n is
 * written unless y_dev != NULL.
 *   sums_dev: M x 4 doubles  [sum y, sum y^2, min, max] per output (overwritten)
 *   first_index: global index of the first sample (lets shards draw disjoint
 *   sub-streams of one logical stream)
 */
int pcq_propagate_dev(pcq_plan* plan, uint64_t seed, int64_t first_index, int64_t n,
                      const int32_t* germ_kind_host, const double* cf_dev, int m,
                      double* sums_dev, double* y_dev, int64_t ldy, int mode,
                      void* stream);
/* host-pointer form: cf (M x K) and sums (M x 4) in host memory, no per-sample output; returns when the sums
 * are in host memory */
int pcq_propagate(pcq_plan* plan, uint64_t seed, int64_t first_index, int64_t n,
                  const int32_t* germ_kind_host, const double* cf_host, int m,
                  double* sums_host, int mode);
/* the germ alone (same stream as above), x_dev N x s */
int pcq_sample_germ_dev(pcq_plan* plan, uint64_t seed, int64_t first_index, int64_t n,
                        const int32_t* germ_kind_host, double* x_dev, int64_t ldx,
                        void* stream);

/* ---- sample files straight into HBM ---------------------------------------------
 * PyTUQ's applications read their samples with np.loadtxt (apps/uqpc/uq_pc.py:232-265) and hand the
 * matrices to pc_fit.  These three calls put the TEXT of such a file into device memory and convert it
 * there (csrc/pcq_text.cu), so the fit's inputs are born in HBM: no host parse, no upload of the binary
 * matrix.  Values are bit-identical to np.loadtxt's (the conversion is the one of pcq_io.h; the few
 * tokens it cannot decide on the device — ties, subnormals, > 19 digits — are converted by the host's
 * strtod and patched in, and so are inf / nan in numpy's spellings).  Whitespace-separated tables only
 * (what np.savetxt writes: one row per line, blank lines allowed): a file with '#' comments returns
 * PCQ_ERR_UNSUPPORTED and belongs to the host reader pcqio_open (pcq_io.h); ragged rows and tokens
 * numpy would not convert return PCQ_ERR_ARG.
 *   pcq_text_open       maps the file, copies its bytes to `device`, counts *rows x *cols there;
 *   pcq_text_parse_dev  converts into out_dev (rows x cols, leading dimension ld >= cols) on `stream`
 *                       and returns once the result is complete (it synchronises the stream);
 *   pcq_text_close      frees the device text and unmaps the file. */
typedef struct pcq_text pcq_text;
int pcq_text_open(int device, const char* path, pcq_text** out, int64_t* rows, int64_t* cols);
int pcq_text_parse_dev(pcq_text* text, double* out_dev, int64_t ld, void* stream);
int pcq_text_close(pcq_text* text);

/* ---- roofline denominators (micro-benchmarks) ---------------------
 * kind 0: FP64 DFMA register-resident chain   -> TFLOP/s
 * kind 1: FP64 DMMA (mma.sync m8n8k4 f64) chain -> TFLOP/s
 * kind 2: HBM copy (read+write bytes)          -> GB/s
 */
int pcq_measure_peak(int device, int kind, double* result);

/* number of kernel launches issued by this library since load (all threads) */
int64_t pcq_launch_count(void);

/* ---- QR statistic: R factor of a tall matrix by blocked Householder reflections (csrc/pcq_qr.cu) ----------------
 * The sample-axis reduction of the least-squares route that has to follow the reference's SVD solve
 * (scipy.linalg.lstsq(A, y, cond=1e-13), lreg/lreg.py:335) where the normal equations cannot: square,
 * underdetermined, nearly collinear, rank-deficient systems.  z: m x L row-major on the device
 * (leading dimension ldz), overwritten; r: L x L upper triangular (row-major, ldr), R^T R = Z^T Z to rounding,
 * rows >= m are zero.  One call takes at most pcq_qr_max_rows(device) rows (its panel kernel is cooperative);
 * longer matrices are chunked and the triangles combined by factorising [R_a; R_b] with the same call. */
int64_t pcq_qr_max_rows(int device);
int pcq_qr_r_dev(int device, double* z_dev, int64_t m, int64_t ldz, int L, double* r_dev, int64_t ldr, void* stream);
/* One-sided Jacobi SVD of the small (n x n) factor the QR route ends in (csrc/pcq_svd.cu): in  g = A^T (row-major,
 * contiguous); out  g = (U Sigma)^T (row j = sigma_j u_j^T), vt = V^T, sigma (n, unsorted), A = U Sigma V^T.  Sweeps of
 * a round-robin tournament (one kernel launch per round, one CTA per column pair) until no pair is further than tol
 * from orthogonal (or max_sweeps); small singular values come out to high relative accuracy, which the reference's
 * rank cut-off (cond = 1e-13, lreg/lreg.py:335) needs.  Synchronises the stream once per sweep. */
int pcq_svd_jacobi_dev(int device, double* g_dev, int n, double* vt_dev, double* sigma_dev, int max_sweeps, double tol,
                       int* sweeps, void* stream);

/* ---- pcq_group: one process, several GPUs, the same host-pointer calls -----------------------------------------
 * SURVEY.md 8(b) (`ndev, devs` of the plan) / 8(e): the path shards over samples.  A group owns one plan per listed
 * device; every call cuts the caller's rows into contiguous blocks (one per device) and runs them concurrently, one
 * host thread per device.  evalBases / evalPC / quadratic forms / residual sums / propagation need no exchange
 * (propagation blocks are disjoint index ranges of ONE logical Philox stream, so the result does not depend on the
 * number of devices beyond summation order); the statistics pass leaves a partial S on every device and sums them
 * on devices[0] in device order (peer copies over NVLink + an add kernel: deterministic) -- the path's one exchange
 * step.  Same argument meaning and status codes as the single-plan calls above; this is what
 * lreg.fit(x, y) / pc_fit / PCRV.evalPC / PCRV.propagate (lreg/lreg.py:62-71, workflows/fits.py:16-81,
 * rv/pcrv.py:469-525) use in a plain Python process that sees more than one GPU. */
typedef struct pcq_group pcq_group;
int pcq_group_create(pcq_group** out, int ndev, const int* devices, int sdim, int nterms, const int32_t* mi,
                     int pmax, const double* rec_a, const double* rec_b, const double* p1_scale);
int pcq_group_destroy(pcq_group* group);
int pcq_group_size(const pcq_group* group);
int pcq_group_enable_moments(pcq_group* group, int nord, const double* rec_a_ext, const double* rec_b_ext);   /* every member plan */
int pcq_group_set_active(pcq_group* group, int n);               /* spread the next calls over the first n members */
pcq_plan* pcq_group_plan(pcq_group* group, int member);          /* borrowed: the member's plan */
int pcq_group_device(const pcq_group* group, int member);
int pcq_group_specialize(pcq_group* group, int modes);           /* pcq_plan_specialize on every member */
/* S = sum over all rows of [Psi|y|1]^T [Psi|y|1]; written to s_dev0 (device pointer on devices[0]) and / or s_host */
int pcq_group_suffstats(pcq_group* group, const double* x_host, int64_t n, int64_t ldx,
                        const double* y_host, int64_t ldy, int m, double* s_dev0, double* s_host);
int pcq_group_eval_bases(pcq_group* group, const double* x_host, int64_t n, int64_t ldx, double* a_host, int64_t lda);
int pcq_group_eval_pc(pcq_group* group, const double* x_host, int64_t n, int64_t ldx, const double* cf_host, int m,
                      double* y_host, int64_t ldy, int mode);
int pcq_group_quadform(pcq_group* group, const double* x_host, int64_t n, int64_t ldx, const double* c_host,
                       int64_t ldc, double* v_host, int64_t ldv);
int pcq_group_residual_ss(pcq_group* group, const double* x_host, int64_t n, int64_t ldx, const double* y_host,
                          int64_t ldy, const double* cf_host, double* sums_host);
int pcq_group_propagate(pcq_group* group, uint64_t seed, int64_t first_index, int64_t n,
                        const int32_t* germ_kind, const double* cf_host, int m, double* sums_host, int mode);

#ifdef __cplusplus
}
#endif
#endif /* PCQ_H_ */
