/* hvi_b200.h -- C ABI of the B200-native negative-ELBO + gradient path.
 *
 * This is the boundary a Julia host binds with `ccall` (see INTEGRATION.md and
 * julia/HVIB200.jl).  The reference (EarthyScience/HybridVariationalInference.jl, paths
 * below relative to /root/reference) has no FFI of its own; every entry point names the
 * reference interface it stands in for.
 *
 * Conventions: every function returns an `int` status (HVI_OK == 0); no exception crosses
 * the boundary; the caller owns every pointer it passes; a plan owns its device buffers;
 * one in-flight call per plan; calls are synchronous at return unless the name ends in
 * `_async` (or the entry says otherwise).  Arrays are column-major with the reference's shapes
 * (site = column).
 *
 * Stream contract.  A plan works on its own non-blocking stream (or the one given to
 * hvi_plan_set_stream).  Work that an `_async` / `_apply_dev` call enqueues on a CALLER's stream is
 * ordered on the device against the plan's stream in both directions: the caller's stream waits for the
 * plan's earlier work (batch registration, earlier calls), and every later call on the plan waits for what
 * was left on the caller's stream -- no host synchronisation is needed between calls.  What the library
 * cannot see are the kernels that PRODUCE a HVI_MEM_DEVICE argument (phi, zP, zMs, xM, ... written by the
 * caller's own kernels on the caller's own stream): either make the plan work on that stream
 * (hvi_plan_set_stream) or synchronise that stream before the call.
 */
#ifndef HVI_B200_H
#define HVI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVI_ABI_VERSION 1
#define HVI_MAX_PAR 8      /* max n_θP and max n_θM */
#define HVI_MAX_LAYERS 8
#define HVI_MAX_WIDTH 64   /* widest layer of the CUDA-core MLP applicator */

/* status codes (reference: Julia exceptions, src/elbo.jl:187-190,253-258) */
enum {
  HVI_OK = 0,
  HVI_EINVAL = 1,       /* bad argument / unsupported configuration                */
  HVI_ECUDA = 2,        /* CUDA runtime error, see hvi_last_error                   */
  HVI_ENONFINITE = 3,   /* result evaluated but not finite ("inspect non-finite")   */
  HVI_ENODATA = 4,      /* hvi_plan_set_data has not been called                    */
  HVI_ENOMEM = 5,
  HVI_ENCCL = 6         /* NCCL could not be loaded or a collective failed         */
};

enum { HVI_F32 = 0, HVI_F64 = 1 };
enum { HVI_MEM_HOST = 0, HVI_MEM_DEVICE = 1 };
/* bijector per parameter column: identity | Exp | Logistic
 * (src/bijectors_utils.jl:10-23, 38-52; Stacked columns :65-114) */
enum { HVI_TR_IDENTITY = 0, HVI_TR_EXP = 1, HVI_TR_LOGISTIC = 2 };
/* prior family per parameter (Distributions.logpdf call sites src/elbo.jl:217-235) */
enum { HVI_PR_NONE = 0, HVI_PR_NORMAL = 1, HVI_PR_LOGNORMAL = 2 };
/* where the process model takes (r0, r1, K1, K2) from: hcat(θP, θMs, θFix) addressed by
 * name (src/PBMApplicator.jl:234-239, 283-287) */
enum { HVI_SRC_P = 0, HVI_SRC_M = 1, HVI_SRC_FIX = 2 };
enum { HVI_ACT_IDENTITY = 0, HVI_ACT_TANH = 1, HVI_ACT_LOGISTIC = 2 };
/* output scaling of g: NormalScalingModelApplicator (src/ModelApplicator.jl:163-176) or
 * RangeScalingModelApplicator (:191-194) */
enum { HVI_SCALE_NORMAL = 0, HVI_SCALE_RANGE = 1 };
enum { HVI_FLAG_OMIT_PRIORS = 1 };  /* is_omit_priors = Val(true), src/elbo.jl:203-206 */
/* posterior approximation (src/HVIApproximation.jl:7-24): MeanHVIApproximationMat (src/elbo.jl:633-678; the
 * per-block MeanHVIApproximation of src/elbo2.jl is the same mathematics) or MeanVarSepHVIApproximation
 * (src/elbo_sepvec.jl:3-48: one log-variance per site and parameter instead of intercept + slope) */
enum { HVI_APPROX_MEAN = 0, HVI_APPROX_MEANVAR_SEP = 1 };

typedef struct hvi_layer {
  int32_t n_in, n_out, act, has_bias; /* params per layer: vec(W[out x in]) col-major, then b */
} hvi_layer;

typedef struct hvi_prior {
  int32_t kind;
  int32_t _pad;
  double a, b; /* Normal(a=μ, b=σ) | LogNormal(a=μ, b=σ) */
} hvi_prior;

/* Everything the Julia side reads from the problem accessors
 * (get_hybridproblem_* generics, src/AbstractHybridProblem.jl:30-311) and passes as keyword
 * arguments to neg_elbo_gtf (src/elbo.jl:46-66). */
typedef struct hvi_desc {
  int32_t struct_size; /* = sizeof(hvi_desc), checked */
  int32_t dtype;       /* HVI_F32 | HVI_F64 : eltype(ϕ) */
  int32_t device;      /* CUDA device ordinal */
  int32_t flags;
  int32_t n_P, n_M, n_obs, n_covar;
  int32_t n_layers;
  hvi_layer layers[HVI_MAX_LAYERS];
  int32_t scale_kind;
  int32_t _pad0;
  double scale_a[HVI_MAX_PAR]; /* μ_k  (normal) | offset_k (range) */
  double scale_b[HVI_MAX_PAR]; /* σ_k  (normal) | width_k  (range) */
  int32_t trans_P[HVI_MAX_PAR];
  int32_t trans_M[HVI_MAX_PAR];
  hvi_prior prior_P[HVI_MAX_PAR];
  hvi_prior prior_M[HVI_MAX_PAR];
  int32_t n_cor_ends_P; /* cor_ends.P: last column of each block (src/cholesky.jl:262-290) */
  int32_t cor_ends_P[HVI_MAX_PAR];
  int32_t n_cor_ends_M;
  int32_t cor_ends_M[HVI_MAX_PAR];
  int32_t slot_src[4]; /* order: r0, r1, K1, K2 (src/DoubleMM/f_doubleMM.jl:47-65,101-139) */
  int32_t slot_idx[4]; /* 0-based index into θP or θM */
  double slot_fix[4];  /* value if HVI_SRC_FIX */
  int32_t n_pbm_covar; /* pbm_covar_indices (src/elbo.jl:549-557), 0-based into θP */
  int32_t pbm_covar_idx[HVI_MAX_PAR];
  /* site sharding: exactly one plan of a job adds the terms that exist once per
   * evaluation (prior and log-Jacobian of θP, entropy of the P block). */
  int32_t count_globals;
  int32_t approx;        /* HVI_APPROX_* */
  int64_t n_site_total;  /* MeanVarSep only: number of sites of the problem, ϕq holds logσ2_ζMs (n_M x n_site_total)
                          * in place of coef_logσ2_ζMs (src/init_hybrid_params.jl:127-153) */
} hvi_desc;

typedef struct hvi_plan hvi_plan;

/* ---- bookkeeping ------------------------------------------------------------------- */
int hvi_version(void);
/* message of the last failure on this plan (NULL plan: last failure of plan_create) */
const char* hvi_last_error(const hvi_plan* plan);

/* Build a plan for one problem description.  Stands in for the keyword set captured by
 * get_loss_elbo (src/HybridSolver.jl:293-324). */
int hvi_plan_create(const hvi_desc* desc, hvi_plan** out);
int hvi_plan_destroy(hvi_plan* plan);

/* ---- integer index maps (bit-exact contract) ---------------------------------------- */
/* length of ϕ = [ϕg ; ϕq] (src/init_hybrid_params.jl:25-33) */
int64_t hvi_phi_len(const hvi_plan* plan);
/* 0-based offsets/lengths of (ϕg, logσ2_ζP, coef_logσ2_ζMs | logσ2_ζMs, ρsP, ρsM, μP) inside ϕ --
 * what ComponentArrayInterpreter/get_positions yields
 * (src/ComponentArrayInterpreter.jl:281-288; layout src/init_hybrid_params.jl:119-124,
 * src/HybridProblem.jl:101-104) */
int hvi_phi_positions(const hvi_plan* plan, int64_t off[6], int64_t len[6]);
/* src/cholesky.jl:30-37, 42-50, 233-247 (one-based, like the reference) */
int hvi_vec2utri_pos(int64_t s, int64_t* row, int64_t* col);
int64_t hvi_utri2vec_pos(int64_t row, int64_t col);
int64_t hvi_cor_count(const int32_t* cor_ends, int32_t n_cor_ends);

/* ---- data --------------------------------------------------------------------------- */
/* Register one batch (a DataLoader element (xM, xP, y_o, y_unc), src/AbstractHybridProblem.jl:
 * 179-184,206-207): xM (n_covar x n_site), xP (2 n_obs x n_site) rows S1;S2, y_o and y_unc
 * (n_obs x n_site).  The library copies and re-tiles the batch into its own HBM layout.
 * y_o and y_unc may both be NULL (no observations: prediction only). */
int hvi_plan_set_data(hvi_plan* plan, int64_t n_site, const void* xM, const void* xP,
                      const void* y_o, const void* y_unc, int mem_kind);

/* The same registration with arrays that are identical for every site passed as ONE column: xP_shared != 0: xP holds
 * the 2 n_obs (or n_x) driver values shared by all sites (the reference keeps such drivers out of the per-site data with
 * PBMPopulationGlobalApplicator, src/PBMApplicator.jl:305-375; the DoubleMM synthetic data are of this kind,
 * src/DoubleMM/f_doubleMM.jl:12-13,375); y_unc_shared != 0: y_unc holds one column of n_obs log-variances.  The
 * library expands them on the device, so a host batch costs 4·(n_covar + n_obs) instead of 4·(n_covar + 4 n_obs)
 * bytes per site on the PCIe link (52 instead of 148 for DoubleMM). */
int hvi_plan_set_data_bcast(hvi_plan* plan, int64_t n_site, const void* xM, const void* xP, int xP_shared,
                            const void* y_o, const void* y_unc, int y_unc_shared, int mem_kind);

/* Double-buffered registration for a minibatch loop (the DataLoader iteration of src/HybridSolver.jl:240-260):
 * hvi_plan_prefetch_data starts the upload of the NEXT batch from HOST memory on the plan's copy stream and returns
 * at once, so the transfer runs while the current batch is evaluated; hvi_plan_commit_data makes it the current
 * batch (waits for the upload, re-tiles it).  The host arrays must stay valid and unchanged until commit_data
 * returns; page-locked memory is needed for the copy to be asynchronous.  Evaluations between the two calls still
 * use the current batch.  A plain hvi_plan_set_data drops a pending prefetch. */
int hvi_plan_prefetch_data(hvi_plan* plan, int64_t n_site, const void* xM, const void* xP,
                           const void* y_o, const void* y_unc);
int hvi_plan_commit_data(hvi_plan* plan);

/* Device-resident training set with minibatch selection on the device -- gdev_hybridproblem_dataloader
 * (src/AbstractHybridProblem.jl:220-250) moves the whole loader to the GPU once and solve walks its minibatches
 * (src/HybridSolver.jl:259-260; MLUtils.DataLoader(...; batchsize, partial = false), src/AbstractHybridProblem.jl:
 * 206-207).  hvi_plan_set_dataset is the one upload (arrays as in hvi_plan_set_data, n_site_total columns).
 * hvi_plan_select_range makes sites [first, first + n) the registered batch, hvi_plan_select_batch the sites
 * i_sites[0..n) (0-based; host or device array): one gather + re-tile kernel, stream-ordered, no host round trip, no
 * synchronisation.  With MeanVarSep the selected indices also become the plan's `i_sites`. */
int hvi_plan_set_dataset(hvi_plan* plan, int64_t n_site_total, const void* xM, const void* xP,
                         const void* y_o, const void* y_unc, int mem_kind);
int hvi_plan_select_range(hvi_plan* plan, int64_t first, int64_t n);
int hvi_plan_select_batch(hvi_plan* plan, const int64_t* i_sites, int64_t n, int mem_kind);

/* The plan's stream: cuda_stream (a cudaStream_t of the plan's device) replaces the plan's own stream for every
 * later call; use_own != 0 switches back.  The new stream continues behind everything queued on the old one. */
int hvi_plan_set_stream(hvi_plan* plan, void* cuda_stream, int use_own);

/* ---- generic process-based model --------------------------------------------- */
/* The reference's process model is any AbstractPBMApplicator callable f(θP, θMs_tr, xP) (src/PBMApplicator.jl:26-32;
 * per site: PBMSiteApplicator, :122-195), differentiated by Zygote.  hvi_plan_set_pbm replaces the built-in DoubleMM
 * kernel of this plan by a process model given as CUDA C++ source:
 *
 *     struct UserPbm {
 *       template <class S, class T>
 *       __device__ static void f(const S* thP, const S* thM, const T* fix, const T* x, S* y) {
 *         for (int o = 0; o < HVI_NO; o++) y[o] = thP[0] + thM[0] * x[o] / (thM[1] + x[o]);
 *       }
 *     };
 *
 * thP / thM: the constrained parameters of one draw (n_P, n_M values in the problem's order), fix: the n_fix fixed
 * parameters (θFix), x: the site's column of xP (n_x driver rows -- xP is then (n_x x n_site)), y: the n_obs
 * predictions.  S is the working type or a forward-mode dual number over it: the library obtains dy/dθ by
 * instantiating f with duals, so no adjoint is written by hand (arithmetic + - * /, hexp, hlog, hsqrt, htanh; literals
 * as T(2)).  The macros HVI_NP, HVI_NM, HVI_NO, HVI_NX and the type HVI_T are defined.  An optional
 * `static constexpr bool HAS_PENALTY = true;` with `template <class S, class T> __device__ static S penalty(thP, thM,
 * fix, x, y)` adds mean_draws(Σ_sites penalty) as the loss_penalty component (floss_penalty, src/elbo.jl:153,282-286).
 * The source is compiled for this plan's sizes by NVRTC (libnvrtc.so.12 of the CUDA toolkit, loaded on demand); a
 * compiler error is returned as HVI_EINVAL with the log in hvi_last_error.
 *   pbm = "doubleMM"           the hand-written DoubleMM kernel (default)
 *   pbm = "builtin:doubleMM"   f_doubleMM through the generic seam (x = [S1; S2], n_x = 2 n_obs)
 *   pbm = "builtin:mm1"        y_o = θP[0] + θM[0] x_o / (θM[1] + x_o), n_x = n_obs
 *   otherwise                  source text defining `struct UserPbm`
 * An observation is used iff y_o is not NaN (src/logden_normal.jl:34); a non-finite prediction at a used observation
 * makes the result non-finite (HVI_ENONFINITE), as in the reference.  Registers no batch: call hvi_plan_set_data
 * afterwards.  With a generic model the plan evaluates value + gradient (all hvi_neg_elbo_grad* and hvi_adam_* entry
 * points, hvi_generate_zeta, hvi_apply_g); hvi_predict, hvi_loss_gf_grad, prefetch and the device-resident dataset
 * stay DoubleMM-only. */
int hvi_plan_set_pbm(hvi_plan* plan, const char* pbm, int32_t n_x, int32_t n_fix, const double* fix);
/* The same NVRTC compilation without a plan or a device: HVI_OK, or HVI_EINVAL with the compiler's message in log. */
int hvi_pbm_compile_check(const char* pbm, int32_t dtype, int32_t n_P, int32_t n_M, int32_t n_obs, int32_t n_x,
                          char* log, int32_t log_len);

/* MeanVarSep only: the 0-based indices (into 0..n_site_total) of the registered batch's sites -- the
 * `i_sites` argument of loss_elbo (src/HybridSolver.jl:312-322, src/elbo_sepvec.jl:23).  Host array of
 * length n_site; default after set_data is 0..n_site-1. */
int hvi_plan_set_sites(hvi_plan* plan, const int64_t* i_sites, int64_t n);

/* ---- the hot path -------------------------------------------------------------------- */
/* neg_elbo_gtf_components + gradient (src/elbo.jl:30-88 under Zygote, src/HybridSolver.jl:
 * 256-258).  phi: length hvi_phi_len; zP (n_MC x n_P) and zMs (n_MC x n_M*n_site), the
 * standard-normal draws of sample_ζresid_norm (src/elbo.jl:605-606), column-major.
 * comps[6] = (nLjoint, entropy_ζ, loss_penalty, nLy, neg_log_prior, neg_log_jac) and
 * *nL = nLjoint - entropy_ζ + loss_penalty, always double (src/logden_normal.jl:74).
 * grad: same element type and layout as phi; NULL = value only.  All array pointers are
 * host or all device, per mem_kind; comps/nL are host. */
int hvi_neg_elbo_grad(hvi_plan* plan, int32_t n_MC, const void* phi, const void* zP,
                      const void* zMs, int mem_kind, double comps[6], double* nL, void* grad);

/* Same evaluation with the draws produced on the device by a counter-based generator keyed
 * by (seed, global site index, draw, parameter): stands in for CUDA.randn
 * (ext/HybridVariationalInferenceCUDAExt.jl:82-86).  site_offset = first global site of
 * this plan's shard so that results do not depend on the sharding. */
int hvi_neg_elbo_grad_rng(hvi_plan* plan, int32_t n_MC, const void* phi, uint64_t seed,
                          int64_t site_offset, int mem_kind, double comps[6], double* nL,
                          void* grad);

/* Device-resident, stream-ordered variant for multi-GPU jobs: writes
 * out_dev[0 .. n_phi) = gradient (double), out_dev[n_phi .. n_phi+6) = components,
 * out_dev[n_phi+6] = nL, out_dev[n_phi+7] = count of non-finite terms.  The caller sums
 * out_dev across ranks (ncclAllReduce) -- SURVEY 8(e).  zP_dev/zMs_dev NULL = generator. */
int hvi_neg_elbo_grad_async(hvi_plan* plan, int32_t n_MC, const void* phi_dev,
                            const void* zP_dev, const void* zMs_dev, uint64_t seed,
                            int64_t site_offset, double* out_dev, void* cuda_stream);

/* ---- the collective of site-sharded jobs ------------------------------------------------ */
/* Site-sharded jobs: each GPU owns a contiguous site range, ϕ and zP are replicated, and ONE all-reduce(sum)
 * of [∇ϕ; 6 components; nL; non-finite count] (n_phi + 8 doubles) finishes an evaluation -- ncclAllReduce over NVLink, issued by
 * the library on the evaluation's stream (libnccl.so.2 is loaded at run time; HVI_NCCL_LIB overrides the name; HVI_ENCCL when
 * it is missing or a call fails).  The reference is single-device (src/HybridSolver.jl:184-267); a Julia host needs nothing but
 * ccall for the multi-GPU job:
 *   rank 0: hvi_nccl_unique_id(id) and hands the 128 bytes to the other ranks (MPI.jl bcast, a file, a socket);
 *   every rank: hvi_plan_create with desc.count_globals = (rank == 0), hvi_plan_comm_init(plan, id, rank, world),
 *   hvi_plan_set_data(its sites), then hvi_neg_elbo_grad_sharded per evaluation with site_offset = first global site index of
 *   the shard: zMs = this rank's columns of the draws, or NULL for device draws keyed by (seed, global site index, draw) --
 *   identical results for any number of ranks.
 * grad_host, comps, nL are host pointers and receive the job's totals on every rank.  hvi_allreduce_result is the collective
 * alone, for the stream-ordered path (hvi_neg_elbo_grad_async → hvi_allreduce_result → hvi_adam_apply_dev). */
int hvi_nccl_unique_id(void* id128);
int hvi_plan_comm_init(hvi_plan* plan, const void* id128, int32_t rank, int32_t world);
int hvi_plan_comm_destroy(hvi_plan* plan);
int hvi_allreduce_result(hvi_plan* plan, double* out_dev, void* cuda_stream);
int hvi_neg_elbo_grad_sharded(hvi_plan* plan, int32_t n_MC, const void* phi, const void* zP, const void* zMs, uint64_t seed,
                              int64_t site_offset, int mem_kind, double comps[6], double* nL, void* grad_host);

/* ---- point solver ------------------------------------------------------- */
/* loss_gf of get_loss_gf (src/gf.jl:227-295; used by solve(prob, HybridPointSolver), src/HybridSolver.jl:9-140) and
 * its gradient: phi_gf = [ϕg ; ϕP] (ComponentVector(ϕg, ϕP), src/HybridSolver.jl:26), θMs = transM(g([xM; ϕP[covars]], ϕg)'),
 * θP = transP(ϕP), no draws, no clamp.  comps[3] = (nLjoint_pen, nLy, neg_log_prior); grad has the layout of phi_gf
 * (NULL = value only). */
int hvi_loss_gf_grad(hvi_plan* plan, const void* phi_gf, int mem_kind, double comps[3], void* grad);

/* ---- optimiser steps on the device -------------------------------------- */
/* Optimization.solve(optprob, Adam(η); …) of src/HybridSolver.jl:259-260 without the host round trip per
 * iteration: ϕ, the Adam moments and the gradient stay in HBM; the update is the Optimisers.jl Adam rule.
 * hvi_adam_init uploads ϕ0 (host) and resets the moments.  hvi_adam_run enqueues n_iter iterations on the
 * registered batch back to back -- iteration i draws with seed0 + i -- and, if nL_trace (host, n_iter doubles)
 * is given, returns the loss of every iteration (NaN marks an iteration that was skipped because its value or
 * gradient was not finite, or because the batch holds a finite observation under a missing driver; a skipped
 * iteration does not advance the Adam step count); with nL_trace = NULL nothing is synchronised.  hvi_adam_get_phi downloads ϕ.
 * Multi-GPU: evaluate with hvi_neg_elbo_grad_async on phi = *hvi_adam_phi_dev, all-reduce out_dev, then
 * hvi_adam_apply_dev(out_dev) on every rank (identical updates, ϕ stays replicated). */
int hvi_adam_init(hvi_plan* plan, const void* phi0, double eta, double beta1, double beta2, double eps);
int hvi_adam_run(hvi_plan* plan, int32_t n_iter, int32_t n_MC, uint64_t seed0, int64_t site_offset,
                 double* nL_trace);
int hvi_adam_apply_dev(hvi_plan* plan, const double* out_dev, void* cuda_stream);
/* The minibatch loop of solve on the device-resident dataset (hvi_plan_set_dataset): iteration i selects minibatch
 * (first_batch + i) mod n_batches, n_batches = n_site_total / batch_size (consecutive column ranges, last partial
 * batch dropped, no shuffling: src/AbstractHybridProblem.jl:206-207), evaluates value + gradient with draws seeded
 * seed0 + i and applies the Adam update.  Iteration state (step count, minibatch index, seed) lives in device
 * counters, so iterations 2..n replay ONE captured CUDA graph (hvi_plan_use_graphs(plan, 0) turns that off).
 * hvi_adam_run does the same on the registered batch. */
int hvi_adam_run_minibatches(hvi_plan* plan, int32_t n_iter, int32_t n_MC, uint64_t seed0, int64_t batch_size,
                             int64_t first_batch, double* nL_trace);
int hvi_plan_use_graphs(hvi_plan* plan, int on);
/* Small batches (the reference's operating point: 20 sites of 800, n_MC = 3 … 12, src/DoubleMM/f_doubleMM.jl:315-325,
 * src/HybridSolver.jl:147) with device draws run as ONE launch of one block per evaluation / optimiser iteration
 * (minibatch gather → draws → prologue → g → fused streaming phase → g backward → reductions → finalize → Adam)
 * instead of eight launches (same results up to the summation order of the per-block partial sums).
 * hvi_plan_use_fused(plan, on): 0 = always the multi-kernel path; 1 (default) = the single launch where it is the faster one
 * (n_site <= 48 and n_site x n_MC <= 1024: one block against eight latency-bound launches breaks even at about 50 sites);
 * 2 = the single launch wherever it is supported (n_site <= 1024, n_site x n_MC <= 16384).  Supported: the CUDA-core
 * applicator shapes, MeanHVIApproximationMat, no pbm covariates. */
int hvi_plan_use_fused(hvi_plan* plan, int on);
/* Profiling aid of that path: the first call (clocks may be NULL) switches the stamping on; later calls return the
 * SM-clock stamps of the last single-launch evaluation at its phase boundaries: start, gather, global-block draws,
 * prologue, g forward, streaming phase, g backward, reductions, finalize, Adam. */
int hvi_plan_fused_phase_clocks(hvi_plan* plan, int64_t clocks[10]);
int hvi_adam_phi_dev(hvi_plan* plan, void** phi_dev);
int hvi_adam_get_phi(hvi_plan* plan, void* phi_host);

/* generate_ζ (src/elbo.jl:472-521) forward only: ζsP (n_P x n_MC), ζsMs_tr
 * (n_site x n_M x n_MC), σ (n_P + n_M n_site).  Used by sample_posterior (:432-458). */
int hvi_generate_zeta(hvi_plan* plan, int32_t n_MC, const void* phi, const void* zP,
                      const void* zMs, int mem_kind, void* zetaP, void* zetaMs_tr, void* sigma);

/* predict_hvi (src/elbo.jl:320-352): sample_posterior (:381-458: generate_ζ with n_MC = n_sample,
 * transform_ζs :815-835, no clamp) followed by the 3-D PBM applicator (src/PBMApplicator.jl:51-57) on
 * the registered batch (set_data may be called with y_o = y_unc = NULL for prediction-only batches).
 * thetaP (n_P x n_sample), thetaMs_tr (n_site x n_M x n_sample), y (n_obs x n_site x n_sample; NaN where a
 * driver is not finite), *entropy = entropy_ζ.  zP/zMs NULL: draws from (seed, site_offset). */
int hvi_predict(hvi_plan* plan, int32_t n_sample, const void* phi, const void* zP, const void* zMs,
                uint64_t seed, int64_t site_offset, int mem_kind, void* thetaP, void* thetaMs_tr,
                void* y, double* entropy);

/* g(xM, ϕg) alone (src/ModelApplicator.jl:19-23,163-176): mu_zeta (n_M x n_site). */
int hvi_apply_g(hvi_plan* plan, const void* phi, int mem_kind, void* mu_zeta);

/* The library's standard-normal generator on its own: out (n_MC x n_cols, column-major) filled with
 * the draws of columns col_offset .. col_offset+n_cols (column = global site * n_M + k), exactly the
 * values hvi_neg_elbo_grad_rng uses.  Stands in for CUDA.randn (ext/HybridVariationalInferenceCUDAExt.jl:82-86). */
int hvi_randn(hvi_plan* plan, int32_t n_MC, int64_t n_cols, uint64_t seed, int64_t col_offset,
              int mem_kind, void* out);

/* One dense layer of g on the tensor cores (tcgen05.mma kind::tf32 with 3xTF32 error compensation,
 * accumulators in TMEM): C (Mc x N, row-major) = act(A (Mc x K, row-major) * W (N x K, row-major)^T + bias).
 * Host pointers, K % 32 == 0, N % 128 == 0.  Building block of the wide applicator (layers of g wider
 * than HVI_MAX_WIDTH; ext/HybridVariationalInferenceFluxExt.jl:26-31 uses cuBLAS there). */
int hvi_dense_tc(int32_t device, int64_t Mc, int32_t N, int32_t K, const float* A, const float* W,
                 const float* bias, int32_t act, float* C);

/* One GEMM of the TMA-fed chain the wide applicator runs on (cp.async.bulk.tensor → tcgen05.mma kind::f16
 * with bf16 hi/lo operand planes, three MMAs per K step, two TMEM accumulators, persistent CTAs):
 *   mode 0: C = act(A B^T + bias), A (M x K), B (N x K)     (layer forward)
 *   mode 1: C = (A B^T) .* act'(Hd)                          (data gradient)
 *   mode 2: C = A^T B, A (K x M), B (K x N), split over K     (weight gradient; MN-major operands)
 * Hd and C are (M x N); row-major host arrays.  N % 128 == 0; K % 32 == 0 (mode 2: M % 32 == 0).  Modes 1
 * and 2 are the pullbacks of the dense layers (Zygote over ext/HybridVariationalInferenceFluxExt.jl:26-31). */
int hvi_gemm_bf3(int32_t device, int32_t mode, int64_t M, int32_t N, int64_t K, const float* A, const float* B,
                 const float* bias, int32_t act, const float* Hd, float* C);

/* number of kernels launched by this plan since creation (bench.py's gpu_launches) */
int64_t hvi_launch_count(const hvi_plan* plan);

/* Per-kernel device timing of the next evaluations: CUDA events recorded on the launching
 * stream around each kernel.  ms[5] = (prologue, g forward, fused streaming kernel, g backward,
 * finalize) of the last evaluation.  The reference has no tracing. */
int hvi_plan_set_profiling(hvi_plan* plan, int on);
int hvi_plan_kernel_times(hvi_plan* plan, float ms[5]);

#ifdef __cplusplus
}
#endif
#endif /* HVI_B200_H */
