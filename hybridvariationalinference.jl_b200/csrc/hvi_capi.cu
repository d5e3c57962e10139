// hvi_capi.cu -- plan management and the extern "C" boundary (include/hvi_b200.h).
// Reference citations (file:line) are relative to /root/reference.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <utility>
#include <vector>

#include "hvi_internal.cuh"

using namespace hvi;

struct GraphKey { const void* ptrs[32]; int64_t v[8]; };

struct hvi_plan {
  hvi_desc desc;
  int dtype;
  size_t es;  // element size
  Cfg<float> cf;
  Cfg<double> cd;
  cudaStream_t stream;       // the stream every call works on: own_stream, or the caller's (hvi_plan_set_stream)
  cudaStream_t own_stream;
  // ordering against caller streams: ev_user marks the last work an _async / _apply_dev call put on a caller's stream;
  // every later call on `stream` waits for it first.  ev_plan orders a caller's stream behind the plan's own work.
  cudaEvent_t ev_user, ev_plan;
  int user_pending;
  int sm_count;
  int64_t launches;
  std::string err;
  // data
  int64_t nb;
  void *d_xM, *d_rec, *d_mu, *d_mubar;
  void* d_xP;   // raw drivers (2 n_obs x nb), kept for the prediction path
  int64_t* d_isites;   // MeanVarSep: global site index of every site of the batch (NULL: identity)
  int has_obs;
  // next batch being uploaded by hvi_plan_prefetch_data while the current one is evaluated
  cudaStream_t copy_stream;
  // the prologue of an evaluation runs beside the applicator's forward pass (neither needs the other): aux stream + fork/join events
  cudaStream_t aux_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_copied;
  void *d_xM_next, *d_xP_next;
  int64_t cap_cur, cap_next;   // capacity in sites of (d_xM, d_xP) and of their _next twins
  int64_t nb_next;             // > 0: a prefetched batch is pending
  int has_obs_next;
  double* d_bconst;            // device: {½Σy_unc over the finite observations, anomaly count} of the current batch
  double* h_bconst;            // pinned copy (fetched with the result of a synchronous evaluation)
  int64_t anomalies;           // host copy of the anomaly count; -1: not fetched yet for this batch
  // device-resident dataset (hvi_plan_set_dataset): raw arrays, column = site
  int64_t ds_n; int ds_has_obs;
  void *ds_xM, *ds_xP, *ds_yo, *ds_yu;
  int64_t* d_idx; int64_t cap_idx;   // minibatch indices of hvi_plan_select_batch
  int64_t n_isites;                  // length of d_isites
  // generic process-based model (hvi_plan_set_pbm): run-time compiled kernel, driver rows per site, raw observations
  GenericPbm* gpbm; int n_x; void *d_yo, *d_yu; double* d_pen; double pbm_fix[MAXP]; std::string pbm_spec;
  // per-evaluation scratch
  void* d_phi;
  void* d_zP; int64_t cap_zP;
  void* d_zMs; int64_t cap_zMs;
  void* d_ec;
  void* d_pdraw; int64_t cap_pdraw;
  void* d_MU; int64_t cap_MU;        // pbm covariates: per-draw means [M][nM][nb]
  void* d_MUBAR; int64_t cap_MUBAR;  // and their cotangents
  double* d_part_x; int64_t cap_part_x;
  double* d_xred; int64_t cap_xred;
  int wide;                           // wide applicator (forward paths only)
  void* d_wide; int64_t cap_wide;
  void* d_wide_bwd; int64_t cap_wide_bwd;
  void* d_wide_cache; int64_t cap_wide_cache;   // activations kept between g forward and backward (bf3 chain)
  double* d_gacc;                     // wide applicator: ∇ϕg accumulator [nphig] + x̄ accumulators
  void* d_stage; int64_t cap_stage;   // staging of a host batch (xP, y_o, y_unc) for k_preprocess
  double* h_pre;                      // pinned: preprocess partial sums
  double* d_part_stream; int max_rows;
  double* d_part_mlp; int max_blk;
  double* d_out;
  double* d_red;
  void* d_grad;
  double* h_out;  // pinned
  void* nccl_comm = nullptr;   // ncclComm_t of a site-sharded job (hvi_plan_comm_init), NULL otherwise
  int nccl_rank = 0, nccl_world = 1;
  void* h_grad;   // pinned
  double* d_ls; int64_t cap_ls;       // per-warp Σ log σ partials of hvi_predict
  // device-resident Adam state (hvi_adam_*): ϕ in the working type, moments in double
  void* d_phi_opt; double *d_adam_m, *d_adam_v, *d_trace; int64_t cap_trace;
  double adam_eta, adam_b1, adam_b2, adam_eps;
  // device counters of the optimiser loop: [0] applied Adam steps, [1] iteration index inside the running hvi_adam_run,
  // [2] seed of the current iteration, [3] block counter of k_adam
  unsigned long long* d_ctr;
  // one optimiser iteration captured as a CUDA graph (replayed by hvi_adam_run)
  cudaGraphExec_t adam_graph; GraphKey graph_key; int graph_valid, graph_nodes, use_graphs;
  int use_fused;   // small batches: one launch per evaluation / optimiser iteration (k_small_iter)
  long long* d_clocks;   // phase time stamps of the last fused launch (hvi_plan_fused_phase_clocks)
  // optional per-kernel timing (CUDA events on the launching stream)
  int profiling;
  cudaEvent_t ev[8];
  int n_ev;
};

static thread_local std::string g_create_err;

#define PLAN_FAIL(plan, code, ...)                    \
  do {                                                \
    char buf_[2048];                                  \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);         \
    (plan)->err = buf_;                               \
    return (code);                                    \
  } while (0)

#define CU(plan, call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) PLAN_FAIL(plan, HVI_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static int cor_count_n(int n) { return (n - 1) * n / 2; }

// per column: first column of its block, offset of its packed coefficients; returns Σ counts.
// Uses the intended cumulative offsets (src/cholesky.jl:262-290).
static int block_maps(const int32_t* ends, int n_ends, int n, int* blk_start, int* rho_off) {
  int total = 0, start = 0;
  if (n == 0) return 0;
  if (n_ends == 0) {
    for (int k = 0; k < n; k++) { blk_start[k] = k; rho_off[k] = 0; }
    return 0;
  }
  for (int b = 0; b < n_ends; b++) {
    const int end = ends[b];
    for (int k = start; k < end; k++) {
      const int kk = k - start;
      blk_start[k] = start;
      rho_off[k] = total + kk * (kk - 1) / 2;
    }
    total += cor_count_n(end - start);
    start = end;
  }
  return total;
}

template <typename T>
static void fill_cfg(const hvi_desc& d, Cfg<T>& c) {
  memset(&c, 0, sizeof c);
  c.nP = d.n_P; c.nM = d.n_M; c.nO = d.n_obs; c.nC = d.n_covar;
  for (int i = 0; i < MAXP; i++) {
    c.transP[i] = d.trans_P[i]; c.transM[i] = d.trans_M[i];
    c.prP_kind[i] = i < d.n_P ? d.prior_P[i].kind : 0;
    c.prM_kind[i] = i < d.n_M ? d.prior_M[i].kind : 0;
    c.prP_a[i] = (T)d.prior_P[i].a; c.prM_a[i] = (T)d.prior_M[i].a;
    c.prP_ib[i] = c.prP_kind[i] ? (T)(1.0 / d.prior_P[i].b) : T(0);
    c.prM_ib[i] = c.prM_kind[i] ? (T)(1.0 / d.prior_M[i].b) : T(0);
    c.prM_nab[i] = c.prM_kind[i] ? (T)(-d.prior_M[i].a / d.prior_M[i].b) : T(0);
    c.prP_c0[i] = c.prP_kind[i] ? log(d.prior_P[i].b) + 0.9189385332046727418 : 0.0;
    c.prM_c0[i] = c.prM_kind[i] ? log(d.prior_M[i].b) + 0.9189385332046727418 : 0.0;
    c.scale_a[i] = (T)d.scale_a[i]; c.scale_b[i] = (T)d.scale_b[i];
    c.pbm_covar_idx[i] = d.pbm_covar_idx[i];
  }
  for (int s = 0; s < 4; s++) {
    c.slot_code[s] = d.slot_src[s] == HVI_SRC_P ? d.slot_idx[s] : d.slot_src[s] == HVI_SRC_M ? d.n_P + d.slot_idx[s] : -1;
    c.slot_fix[s] = (T)d.slot_fix[s];
  }
  c.omit_priors = (d.flags & HVI_FLAG_OMIT_PRIORS) ? 1 : 0;
  c.count_globals = d.count_globals;
  if (sizeof(T) == 4) { c.clamp_lo = (T)sqrt(1.17549435082228750797e-38); c.clamp_hi = (T)sqrt(3.40282346638528859812e+38); }
  else { c.clamp_lo = (T)sqrt(2.2250738585072014e-308); c.clamp_hi = (T)sqrt(1.7976931348623157e+308); }
  c.ln_clamp_lo = (T)log((double)c.clamp_lo); c.ln_clamp_hi = (T)log((double)c.clamp_hi);
  int p = 0;
  c.nL = d.n_layers; c.sum_width = 0; c.sum_out = 0; c.max_width = 0;
  for (int l = 0; l < d.n_layers; l++) {
    c.l_in[l] = d.layers[l].n_in; c.l_out[l] = d.layers[l].n_out; c.l_act[l] = d.layers[l].act;
    c.l_bias[l] = d.layers[l].has_bias;
    c.woff[l] = p; p += c.l_in[l] * c.l_out[l];
    c.boff[l] = p; if (c.l_bias[l]) p += c.l_out[l];
    c.sum_width += c.l_in[l]; c.sum_out += c.l_out[l];
    if (c.l_in[l] > c.max_width) c.max_width = c.l_in[l];
    if (c.l_out[l] > c.max_width) c.max_width = c.l_out[l];
  }
  c.sum_width += d.layers[d.n_layers - 1].n_out;
  c.nphig = p;
  c.o_ls2P = p; p += c.nP;
  c.sepvar = d.approx == HVI_APPROX_MEANVAR_SEP ? 1 : 0;
  c.n_site_total = (int)d.n_site_total;
  c.o_coef = p; p += c.sepvar ? c.nM * c.n_site_total : 2 * c.nM;
  c.o_rhoP = p; p += block_maps(d.cor_ends_P, d.n_cor_ends_P, c.nP, c.blkP_start, c.rhoP_off);
  c.o_rhoM = p; p += block_maps(d.cor_ends_M, d.n_cor_ends_M, c.nM, c.blkM_start, c.rhoM_off);
  c.o_muP = p; p += c.nP;
  c.nphi = p;
  c.scale_kind = d.scale_kind;
  c.npc = d.n_pbm_covar;
}

// a network with a layer wider than HVI_MAX_WIDTH runs on the wide (tensor-core) applicator: forward only
static bool is_wide(const hvi_desc* d) {
  for (int l = 0; l < d->n_layers; l++)
    if (d->layers[l].n_in > HVI_MAX_WIDTH || d->layers[l].n_out > HVI_MAX_WIDTH) return true;
  return false;
}

static const char* validate(const hvi_desc* d) {
  if (!d) return "desc is NULL";
  if (d->struct_size != (int32_t)sizeof(hvi_desc)) return "hvi_desc.struct_size mismatch (ABI)";
  if (d->dtype != HVI_F32 && d->dtype != HVI_F64) return "dtype must be HVI_F32 or HVI_F64";
  if (d->n_P < 0 || d->n_P > MAXP || d->n_M < 1 || d->n_M > MAXP) return "n_P/n_M out of range";
  if (d->n_layers < 1 || d->n_layers > MAXL) return "n_layers out of range";
  if (d->n_pbm_covar < 0 || d->n_pbm_covar > d->n_P) return "n_pbm_covar out of range";
  for (int c = 0; c < d->n_pbm_covar; c++)
    if (d->pbm_covar_idx[c] < 0 || d->pbm_covar_idx[c] >= d->n_P) return "pbm_covar_idx out of range";
  if (d->layers[0].n_in != d->n_covar + d->n_pbm_covar) return "first layer n_in != n_covar + n_pbm_covar";
  for (int l = 0; l < d->n_layers; l++) {
    if (d->layers[l].n_in < 1 || d->layers[l].n_out < 1) return "layer sizes must be positive";
    if (l > 0 && d->layers[l].n_in != d->layers[l - 1].n_out) return "layer sizes do not chain";
    if (d->layers[l].n_in > 4096 || d->layers[l].n_out > 4096) return "layer wider than 4096";
  }
  if (d->layers[d->n_layers - 1].n_out < d->n_M) return "network has fewer outputs than n_M";
  if (d->n_cor_ends_P < 0 || d->n_cor_ends_P > MAXP || d->n_cor_ends_M < 0 || d->n_cor_ends_M > MAXP)
    return "n_cor_ends out of range";
  if (d->n_P > 0 && d->n_cor_ends_P > 0 && d->cor_ends_P[d->n_cor_ends_P - 1] != d->n_P) return "cor_ends_P must end at n_P";
  if (d->n_cor_ends_M > 0 && d->cor_ends_M[d->n_cor_ends_M - 1] != d->n_M) return "cor_ends_M must end at n_M";
  for (int b = 1; b < d->n_cor_ends_P; b++) if (d->cor_ends_P[b] <= d->cor_ends_P[b - 1]) return "cor_ends_P not increasing";
  for (int b = 1; b < d->n_cor_ends_M; b++) if (d->cor_ends_M[b] <= d->cor_ends_M[b - 1]) return "cor_ends_M not increasing";
  for (int s = 0; s < 4; s++) {
    if (d->slot_src[s] == HVI_SRC_P && (d->slot_idx[s] < 0 || d->slot_idx[s] >= d->n_P)) return "slot_idx (P) out of range";
    if (d->slot_src[s] == HVI_SRC_M && (d->slot_idx[s] < 0 || d->slot_idx[s] >= d->n_M)) return "slot_idx (M) out of range";
    if (d->slot_src[s] < 0 || d->slot_src[s] > 2) return "slot_src invalid";
  }
  for (int i = 0; i < d->n_P; i++) {
    if (d->trans_P[i] < 0 || d->trans_P[i] > 2) return "trans_P invalid";
    if (d->prior_P[i].kind < 0 || d->prior_P[i].kind > 2) return "prior_P kind invalid";
    if (d->prior_P[i].kind && !(d->prior_P[i].b > 0)) return "prior_P scale must be positive";
  }
  for (int i = 0; i < d->n_M; i++) {
    if (d->trans_M[i] < 0 || d->trans_M[i] > 2) return "trans_M invalid";
    if (d->prior_M[i].kind < 0 || d->prior_M[i].kind > 2) return "prior_M kind invalid";
    if (d->prior_M[i].kind && !(d->prior_M[i].b > 0)) return "prior_M scale must be positive";
  }
  if (d->approx != HVI_APPROX_MEAN && d->approx != HVI_APPROX_MEANVAR_SEP) return "approx invalid";
  if (d->approx == HVI_APPROX_MEANVAR_SEP && (d->n_site_total < 1 || d->n_site_total * d->n_M > 1000000000LL))
    return "MeanVarSep needs 1 <= n_site_total (and n_M*n_site_total <= 1e9)";
  if (!stream_supported(d->n_P, d->n_M, d->n_obs)) return "(n_P, n_M, n_obs) combination has no kernel instantiation";
  if (d->n_pbm_covar > 0 && !mlp_supports_pbm(d) && !is_wide(d))
    return "pbm_covars need one of the register-tiled applicator shapes (5-20-20-2, 6-24-24-2) or a wide network";
  return nullptr;
}

extern "C" int hvi_version(void) { return HVI_ABI_VERSION; }

extern "C" const char* hvi_last_error(const hvi_plan* plan) { return plan ? plan->err.c_str() : g_create_err.c_str(); }

extern "C" int hvi_plan_create(const hvi_desc* desc, hvi_plan** out) {
  if (!out) { g_create_err = "out is NULL"; return HVI_EINVAL; }
  *out = nullptr;
  if (const char* msg = validate(desc)) { g_create_err = msg; return HVI_EINVAL; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this path has no CPU fallback)";
    return HVI_ECUDA;
  }
  if (desc->device < 0 || desc->device >= ndev) { g_create_err = "device ordinal out of range"; return HVI_EINVAL; }
  hvi_plan* p = new (std::nothrow) hvi_plan();
  if (!p) { g_create_err = "out of host memory"; return HVI_ENOMEM; }
  p->desc = *desc;
  p->dtype = desc->dtype;
  p->es = desc->dtype == HVI_F32 ? 4 : 8;
  fill_cfg<float>(*desc, p->cf);
  fill_cfg<double>(*desc, p->cd);
  p->nb = 0; p->launches = 0;
  p->d_ls = nullptr; p->cap_ls = 0;
  p->d_phi_opt = nullptr; p->d_adam_m = p->d_adam_v = p->d_trace = nullptr; p->cap_trace = 0;
  p->d_ctr = nullptr; p->adam_graph = nullptr; p->graph_valid = 0; p->graph_nodes = 0; p->use_graphs = 1; p->use_fused = 1; p->d_clocks = nullptr;
  p->own_stream = nullptr; p->ev_user = p->ev_plan = nullptr; p->user_pending = 0;
  p->d_bconst = nullptr; p->h_bconst = nullptr;
  p->ds_n = 0; p->ds_has_obs = 0; p->ds_xM = p->ds_xP = p->ds_yo = p->ds_yu = nullptr; p->d_idx = nullptr; p->cap_idx = 0; p->n_isites = 0;
  p->gpbm = nullptr; p->n_x = 2 * desc->n_obs; p->d_yo = p->d_yu = nullptr; p->d_pen = nullptr;
  for (double& v : p->pbm_fix) v = 0.0;
  p->copy_stream = nullptr; p->ev_copied = nullptr; p->d_xM_next = p->d_xP_next = nullptr;
  p->cap_cur = p->cap_next = p->nb_next = 0; p->has_obs_next = 0;
  p->d_xM = p->d_rec = p->d_mu = p->d_mubar = nullptr; p->d_xP = nullptr; p->has_obs = 0; p->d_isites = nullptr;
  p->d_phi = p->d_zP = p->d_zMs = p->d_ec = p->d_pdraw = nullptr;
  p->cap_zP = p->cap_zMs = p->cap_pdraw = 0;
  p->d_MU = p->d_MUBAR = nullptr; p->d_part_x = p->d_xred = nullptr;
  p->cap_MU = p->cap_MUBAR = p->cap_part_x = p->cap_xred = 0;
  p->d_stage = nullptr; p->cap_stage = 0; p->h_pre = nullptr;
  p->wide = is_wide(desc) ? 1 : 0; p->d_wide = nullptr; p->cap_wide = 0;
  p->d_wide_bwd = nullptr; p->cap_wide_bwd = 0; p->d_gacc = nullptr;
  p->d_wide_cache = nullptr; p->cap_wide_cache = 0;
  p->d_part_stream = p->d_part_mlp = p->d_out = p->d_red = nullptr;
  p->d_grad = nullptr; p->h_out = nullptr; p->h_grad = nullptr;
  p->anomalies = 0;
  p->profiling = 0; p->n_ev = 0;
  for (auto& e_ : p->ev) e_ = nullptr;
#define CRE(call)                                                                   \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      g_create_err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
      hvi_plan_destroy(p);                                                          \
      return HVI_ECUDA;                                                             \
    }                                                                               \
  } while (0)
  CRE(cudaSetDevice(desc->device));
  cudaDeviceProp prop;
  CRE(cudaGetDeviceProperties(&prop, desc->device));
  p->sm_count = prop.multiProcessorCount;
  p->stream = nullptr;
  CRE(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
  p->stream = p->own_stream;
  CRE(cudaEventCreateWithFlags(&p->ev_user, cudaEventDisableTiming));
  CRE(cudaEventCreateWithFlags(&p->ev_plan, cudaEventDisableTiming));
  const int nphi = p->cf.nphi;
  p->max_rows = stream_max_rows(p->sm_count);
  p->max_blk = mlp_bwd_max_blocks(p->sm_count);
  CRE(cudaMalloc(&p->d_phi, p->es * (size_t)nphi));
  CRE(cudaMalloc(&p->d_ec, sizeof(EvalConsts<double>)));
  CRE(cudaMalloc(&p->d_part_stream, sizeof(double) * (size_t)p->max_rows * nacc(p->cf.nP, p->cf.nM)));
  if (!p->wide) CRE(cudaMalloc(&p->d_part_mlp, sizeof(double) * (size_t)2 * p->max_blk * p->cf.nphig));
  else CRE(cudaMalloc((void**)&p->d_gacc, sizeof(double) * (size_t)p->cf.nphig));
  CRE(cudaMalloc(&p->d_out, sizeof(double) * (size_t)(nphi + 8)));
  CRE(cudaMalloc(&p->d_red, sizeof(double) * (size_t)(128 + (p->cf.nphig + 31) / 32 + 2)));
  CRE(cudaMalloc(&p->d_grad, p->es * (size_t)nphi));
  CRE(cudaMallocHost(&p->h_out, sizeof(double) * (size_t)(nphi + 8)));
  CRE(cudaMallocHost(&p->h_grad, p->es * (size_t)nphi));
  CRE(cudaMallocHost(&p->h_pre, sizeof(double) * (size_t)2 * p->sm_count * 8 * 8));
  CRE(cudaMalloc((void**)&p->d_bconst, sizeof(double) * 2));
  CRE(cudaMallocHost((void**)&p->h_bconst, sizeof(double) * 2));
  CRE(cudaMalloc((void**)&p->d_ctr, sizeof(unsigned long long) * 4));
  CRE(cudaMemset(p->d_ctr, 0, sizeof(unsigned long long) * 4));
#undef CRE
  *out = p;
  return HVI_OK;
}

extern "C" int hvi_plan_destroy(hvi_plan* p) {
  if (!p) return HVI_OK;
  cudaSetDevice(p->desc.device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->own_stream && p->own_stream != p->stream) cudaStreamSynchronize(p->own_stream);
  if (p->user_pending && p->ev_user) cudaEventSynchronize(p->ev_user);
  if (p->nccl_comm) hvi_plan_comm_destroy(p);
  if (p->adam_graph) cudaGraphExecDestroy(p->adam_graph);
  if (p->gpbm) generic_pbm_destroy(p->gpbm);
  if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
  if (p->aux_stream) { cudaStreamSynchronize(p->aux_stream); cudaStreamDestroy(p->aux_stream); }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  if (p->ev_copied) cudaEventDestroy(p->ev_copied);
  void* dev[] = {p->d_clocks, p->d_yo, p->d_yu, p->d_pen, p->d_bconst, p->d_ctr, p->ds_xM, p->ds_xP, p->ds_yo, p->ds_yu, p->d_idx, p->d_ls, p->d_phi_opt, p->d_adam_m, p->d_adam_v, p->d_trace, p->d_xM_next, p->d_xP_next, p->d_isites, p->d_xP, p->d_xM, p->d_rec, p->d_mu, p->d_mubar, p->d_phi, p->d_zP, p->d_zMs, p->d_ec, p->d_pdraw,
                 p->d_part_stream, p->d_part_mlp, p->d_out, p->d_red, p->d_grad, p->d_MU, p->d_MUBAR, p->d_part_x, p->d_xred, p->d_stage, p->d_wide, p->d_wide_bwd, p->d_gacc, p->d_wide_cache};
  for (void* q : dev) if (q) cudaFree(q);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->h_grad) cudaFreeHost(p->h_grad);
  if (p->h_pre) cudaFreeHost(p->h_pre);
  if (p->h_bconst) cudaFreeHost(p->h_bconst);
  for (auto& e_ : p->ev) if (e_) cudaEventDestroy(e_);
  if (p->ev_user) cudaEventDestroy(p->ev_user);
  if (p->ev_plan) cudaEventDestroy(p->ev_plan);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  delete p;
  return HVI_OK;
}

extern "C" int64_t hvi_phi_len(const hvi_plan* p) { return p ? p->cf.nphi : -1; }

extern "C" int hvi_phi_positions(const hvi_plan* p, int64_t off[6], int64_t len[6]) {
  if (!p || !off || !len) return HVI_EINVAL;
  const Cfg<float>& c = p->cf;
  const int o[7] = {0, c.o_ls2P, c.o_coef, c.o_rhoP, c.o_rhoM, c.o_muP, c.nphi};
  for (int i = 0; i < 6; i++) { off[i] = o[i]; len[i] = o[i + 1] - o[i]; }
  return HVI_OK;
}

// src/cholesky.jl:30-37 (one-based)
extern "C" int hvi_vec2utri_pos(int64_t s, int64_t* row, int64_t* col) {
  if (s < 1 || !row || !col) return HVI_EINVAL;
  int64_t c = (int64_t)ceil(-0.5 + sqrt(0.25 + 2.0 * (double)s));
  while (c * (c + 1) / 2 < s) c++;          // guard the float inversion
  while ((c - 1) * c / 2 >= s) c--;
  const int64_t cl = c - 1;
  *row = s - cl * (cl + 1) / 2;
  *col = c;
  return HVI_OK;
}
// src/cholesky.jl:42-50
extern "C" int64_t hvi_utri2vec_pos(int64_t row, int64_t col) {
  if (row < 1 || row > col) return -1;
  return (col - 1) * col / 2 + row;
}
// src/cholesky.jl:233-247
extern "C" int64_t hvi_cor_count(const int32_t* cor_ends, int32_t n) {
  if (n <= 0 || !cor_ends) return 0;
  int64_t tot = 0;
  for (int i = 0; i < n; i++) {
    const int nb = i == 0 ? cor_ends[0] : cor_ends[i] - cor_ends[i - 1];
    tot += (int64_t)(nb - 1) * nb / 2;
  }
  return tot;
}

extern "C" int64_t hvi_launch_count(const hvi_plan* p) { return p ? p->launches : -1; }

// ---- stream ordering ---------------------------------------------------------------------
// Work that an _async / _apply_dev call left on a caller's stream is awaited (on the device, not the host) before
// anything else touches the plan's buffers on `stream`.
static int join_user(hvi_plan* p) {
  if (p->user_pending) {
    CU(p, cudaStreamWaitEvent(p->stream, p->ev_user, 0));
    p->user_pending = 0;
  }
  return HVI_OK;
}
// a caller's stream `s` is about to read/write plan buffers: order it behind everything queued on `stream`
static int fork_to(hvi_plan* p, cudaStream_t s) {
  if (s != p->stream) {
    CU(p, cudaEventRecord(p->ev_plan, p->stream));
    CU(p, cudaStreamWaitEvent(s, p->ev_plan, 0));
  }
  return HVI_OK;
}
static int mark_user(hvi_plan* p, cudaStream_t s) {
  if (s != p->stream) {
    CU(p, cudaEventRecord(p->ev_user, s));
    p->user_pending = 1;
  }
  return HVI_OK;
}

extern "C" int hvi_plan_set_stream(hvi_plan* p, void* cuda_stream, int use_own) {
  if (!p) return HVI_EINVAL;
  CU(p, cudaSetDevice(p->desc.device));
  cudaStream_t ns = use_own ? p->own_stream : (cudaStream_t)cuda_stream;
  if (ns == p->stream) return HVI_OK;
  int rc = join_user(p);
  if (rc) return rc;
  // the new stream continues after everything queued on the old one
  CU(p, cudaEventRecord(p->ev_plan, p->stream));
  CU(p, cudaStreamWaitEvent(ns, p->ev_plan, 0));
  p->stream = ns;
  p->graph_valid = 0;   // a captured optimiser iteration belongs to the stream it was captured on
  return HVI_OK;
}

extern "C" int hvi_plan_set_pbm(hvi_plan* p, const char* pbm, int32_t n_x, int32_t n_fix, const double* fix) {
  if (!p) return HVI_EINVAL;
  if (!pbm) PLAN_FAIL(p, HVI_EINVAL, "set_pbm: pbm is NULL");
  if (n_fix < 0 || n_fix > MAXP || (n_fix > 0 && !fix)) PLAN_FAIL(p, HVI_EINVAL, "set_pbm: 0 <= n_fix <= %d fixed parameters", MAXP);
  CU(p, cudaSetDevice(p->desc.device));
  int rc = join_user(p);
  if (rc) return rc;
  CU(p, cudaStreamSynchronize(p->stream));
  // a different process model invalidates the registered batch (its layout depends on the model)
  for (void** q : {&p->d_rec, &p->d_mu, &p->d_mubar, &p->d_yo, &p->d_yu, &p->d_xM, &p->d_xP}) { if (*q) cudaFree(*q); *q = nullptr; }
  p->nb = 0; p->cap_cur = 0; p->graph_valid = 0;
  if (p->gpbm) { generic_pbm_destroy(p->gpbm); p->gpbm = nullptr; }
  for (int i = 0; i < MAXP; i++) p->pbm_fix[i] = i < n_fix ? fix[i] : 0.0;
  p->pbm_spec = pbm;
  if (p->pbm_spec == "doubleMM") {   // the hand-written kernel (default)
    p->n_x = 2 * p->desc.n_obs;
    return HVI_OK;
  }
  if (n_x < 1) PLAN_FAIL(p, HVI_EINVAL, "set_pbm: n_x (driver rows per site) must be >= 1");
  if (p->wide) PLAN_FAIL(p, HVI_EINVAL, "set_pbm: a generic process model is not available with the wide applicator");
  if (p->pbm_spec == "builtin:doubleMM")   // θFix of the slot map comes from the description
    for (int s_ = 0; s_ < 4; s_++) p->pbm_fix[s_] = p->desc.slot_src[s_] == HVI_SRC_FIX ? p->desc.slot_fix[s_] : 0.0;
  std::string err;
  rc = generic_pbm_create(pbm, p->dtype, p->desc.n_P, p->desc.n_M, p->desc.n_obs, n_x, p->cf.npc > 0, err, &p->gpbm);
  if (rc) PLAN_FAIL(p, rc, "%s", err.c_str());
  p->n_x = n_x;
  if (!p->d_pen) CU(p, cudaMalloc((void**)&p->d_pen, sizeof(double) * (size_t)p->max_rows));
  return HVI_OK;
}

// NVRTC compilation of a process model for the given problem shape without touching a device (log: the compiler's
// message on failure).  Lets a host check a model's source before a GPU is involved.
extern "C" int hvi_pbm_compile_check(const char* pbm, int32_t dtype, int32_t n_P, int32_t n_M, int32_t n_obs, int32_t n_x,
                                     char* log, int32_t log_len) {
  std::string err;
  const int rc = generic_pbm_compile_check(pbm, dtype, n_P, n_M, n_obs, n_x, err);
  if (log && log_len > 0) { strncpy(log, err.c_str(), (size_t)log_len - 1); log[log_len - 1] = 0; }
  return rc;
}

extern "C" int hvi_plan_set_sites(hvi_plan* p, const int64_t* i_sites, int64_t n) {
  if (!p) return HVI_EINVAL;
  if (p->nb <= 0) PLAN_FAIL(p, HVI_ENODATA, "hvi_plan_set_data has not been called");
  if (!i_sites || n != p->nb) PLAN_FAIL(p, HVI_EINVAL, "set_sites: need one index per site of the registered batch");
  if (p->cf.sepvar)
    for (int64_t i = 0; i < n; i++)
      if (i_sites[i] < 0 || i_sites[i] >= p->cf.n_site_total) PLAN_FAIL(p, HVI_EINVAL, "set_sites: index out of range");
  CU(p, cudaSetDevice(p->desc.device));
  int rc = join_user(p);
  if (rc) return rc;
  CU(p, cudaStreamSynchronize(p->stream));   // nothing in flight reads the old index array
  if (p->d_isites) { cudaFree(p->d_isites); p->d_isites = nullptr; p->n_isites = 0; }
  CU(p, cudaMalloc((void**)&p->d_isites, sizeof(int64_t) * (size_t)n));
  p->n_isites = n;
  CU(p, cudaMemcpyAsync(p->d_isites, i_sites, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, p->stream));
  CU(p, cudaStreamSynchronize(p->stream));
  return HVI_OK;
}

extern "C" int hvi_plan_set_profiling(hvi_plan* p, int on) {
  if (!p) return HVI_EINVAL;
  CU(p, cudaSetDevice(p->desc.device));
  if (on)
    for (auto& evt : p->ev) if (!evt) CU(p, cudaEventCreate(&evt));
  p->profiling = on ? 1 : 0;
  p->n_ev = 0;
  return HVI_OK;
}

extern "C" int hvi_plan_kernel_times(hvi_plan* p, float ms[5]) {
  if (!p || !ms) return HVI_EINVAL;
  if (!p->profiling || p->n_ev < 6) PLAN_FAIL(p, HVI_EINVAL, "no profiled evaluation recorded");
  CU(p, cudaEventSynchronize(p->ev[5]));
  for (int i = 0; i < 5; i++) CU(p, cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
  return HVI_OK;
}

static cudaError_t copy_in(void* dst, const void* src, size_t bytes, int mem_kind, cudaStream_t s) {
  return cudaMemcpyAsync(dst, src, bytes, mem_kind == HVI_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s);
}

static int ensure(hvi_plan* p, void** buf, int64_t* cap, int64_t need_bytes);

// (re)allocate the per-batch buffers for nb sites; keep_raw: d_xM/d_xP already hold the batch (commit of a prefetch)
template <typename T>
static int alloc_batch(hvi_plan* p, int64_t nb, bool keep_raw) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar, nM = p->desc.n_M;
  if (nb == p->nb && (keep_raw || p->cap_cur >= nb)) return HVI_OK;
  for (void** q : {&p->d_rec, &p->d_mu, &p->d_mubar, &p->d_yo, &p->d_yu}) { if (*q) cudaFree(*q); *q = nullptr; }
  p->nb = 0;
  const int64_t ntiles = (nb + 31) / 32;
  if (!keep_raw) {
    for (void** q : {&p->d_xM, &p->d_xP}) { if (*q) cudaFree(*q); *q = nullptr; }
    p->cap_cur = 0;
    CU(p, cudaMalloc(&p->d_xM, sizeof(T) * (size_t)nC * nb));
    CU(p, cudaMalloc(&p->d_xP, sizeof(T) * (size_t)p->n_x * nb));
    p->cap_cur = nb;
  }
  if (p->gpbm) {   // the generic kernel reads the raw observations instead of DoubleMM records
    CU(p, cudaMalloc(&p->d_yo, sizeof(T) * (size_t)nO * nb));
    CU(p, cudaMalloc(&p->d_yu, sizeof(T) * (size_t)nO * nb));
  }
  CU(p, cudaMalloc(&p->d_rec, sizeof(T) * (size_t)ntiles * 32 * 4 * nO));
  CU(p, cudaMalloc(&p->d_mu, sizeof(T) * (size_t)nM * nb));
  CU(p, cudaMalloc(&p->d_mubar, sizeof(T) * (size_t)nM * nb));
  p->nb = nb;
  return HVI_OK;
}

// after the partial sums of a (gather-)preprocess launch: reduce them into the device-resident batch constants
static int finish_consts(hvi_plan* p, int nrows, bool sync_fetch) {
  Launch L{p->stream, p->sm_count, &p->launches};
  CU(p, launch_batch_consts(L, p->d_part_stream, nrows, p->d_bconst));
  p->anomalies = -1;   // fetched with the next synchronous result (or here)
  if (sync_fetch) {
    CU(p, cudaMemcpyAsync(p->h_bconst, p->d_bconst, sizeof(double) * 2, cudaMemcpyDeviceToHost, p->stream));
    CU(p, cudaStreamSynchronize(p->stream));
    p->anomalies = (int64_t)p->h_bconst[1];
  }
  return HVI_OK;
}

// re-tile the raw batch on the device (k_preprocess) and fetch the two batch constants
template <typename T>
static int finish_batch(hvi_plan* p, int64_t nb, const T* dxP, const T* dyo, const T* dyu) {
  const int nO = p->desc.n_obs;
  Launch L{p->stream, p->sm_count, &p->launches};
  int nrows = 0;
  // the stream partial buffer doubles as scratch for the 2·nwarps preprocess partials
  const int NACC = nacc(p->cf.nP, p->cf.nM);
  if ((size_t)2 * p->sm_count * 8 * 8 > (size_t)p->max_rows * NACC) PLAN_FAIL(p, HVI_EINVAL, "internal: scratch too small");
  if (p->gpbm) {
    if (!dyo || !dyu) PLAN_FAIL(p, HVI_ENODATA, "a generic process model needs observations (y_o, y_unc)");
    const size_t n2 = (size_t)nO * nb;
    CU(p, cudaMemcpyAsync(p->d_yo, dyo, sizeof(T) * n2, cudaMemcpyDeviceToDevice, p->stream));
    CU(p, cudaMemcpyAsync(p->d_yu, dyu, sizeof(T) * n2, cudaMemcpyDeviceToDevice, p->stream));
    CU(p, launch_preprocess_generic<T>(L, (int64_t)n2, (const T*)p->d_yo, (const T*)p->d_yu, p->d_part_stream, &nrows));
    return finish_consts(p, nrows, true);
  }
  CU(p, launch_preprocess<T>(L, nO, nb, dxP, dyo, dyu, (T*)p->d_rec, p->d_part_stream, &nrows));
  // synchronous: the caller's host arrays are free again at return, and the anomaly count is known
  return finish_consts(p, nrows, true);
}

template <typename T>
static int set_data_t(hvi_plan* p, int64_t nb, const void* xM, const void* xP, const void* y_o, const void* y_unc,
                      int mem_kind) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar;
  if (p->nb_next > 0) {   // a pending prefetch shares the staging buffer: drop it
    CU(p, cudaStreamSynchronize(p->copy_stream));
    p->nb_next = 0;
  }
  int rc = join_user(p);
  if (rc) return rc;
  CU(p, cudaStreamSynchronize(p->stream));   // evaluations in flight still read the old batch
  rc = alloc_batch<T>(p, nb, false);
  if (rc) return rc;
  CU(p, copy_in(p->d_xM, xM, sizeof(T) * (size_t)nC * nb, mem_kind, p->stream));
  const T *dxP = (const T*)xP, *dyo = (const T*)y_o, *dyu = (const T*)y_unc;
  {
    const size_t n1 = (size_t)p->n_x * nb, n2 = (size_t)nO * nb;
    CU(p, copy_in(p->d_xP, xP, sizeof(T) * n1, mem_kind, p->stream));
    dxP = (const T*)p->d_xP;
    p->has_obs = (y_o && y_unc) ? 1 : 0;
    if (mem_kind == HVI_MEM_HOST && p->has_obs) {
      rc = ensure(p, &p->d_stage, &p->cap_stage, (int64_t)(sizeof(T) * 2 * n2));
      if (rc) return rc;
      T* t = (T*)p->d_stage;
      CU(p, copy_in(t, y_o, sizeof(T) * n2, mem_kind, p->stream));
      CU(p, copy_in(t + n2, y_unc, sizeof(T) * n2, mem_kind, p->stream));
      dyo = t; dyu = t + n2;
    }
    if (!p->has_obs) { dyo = nullptr; dyu = nullptr; }
  }
  return finish_batch<T>(p, nb, dxP, dyo, dyu);
}

// dst (rows x nb, column-major) = the one column src (rows) repeated: arrays that are identical for every site are
// uploaded once (PBMPopulationGlobalApplicator keeps shared drivers out of the per-site data, src/PBMApplicator.jl:305-375)
template <typename T>
__global__ void __launch_bounds__(256) k_expand_cols(T* __restrict__ dst, const T* __restrict__ src, int rows, int64_t nb) {
  const int64_t n = (int64_t)rows * nb;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) dst[e] = src[e % rows];
}

template <typename T>
static int expand_cols(hvi_plan* p, T* dst, const T* src_dev, int rows, int64_t nb) {
  int64_t blocks = ((int64_t)rows * nb + 255) / 256;
  if (blocks > (int64_t)p->sm_count * 16) blocks = (int64_t)p->sm_count * 16;
  k_expand_cols<T><<<(int)blocks, 256, 0, p->stream>>>(dst, src_dev, rows, nb);
  CU(p, cudaGetLastError());
  ++p->launches;
  return HVI_OK;
}

// set_data with arrays that are the same for every site given as ONE column
template <typename T>
static int set_data_bcast_t(hvi_plan* p, int64_t nb, const void* xM, const void* xP, int xP_shared, const void* y_o,
                            const void* y_unc, int y_unc_shared, int mem_kind) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar, nX = p->n_x;
  if (p->nb_next > 0) { CU(p, cudaStreamSynchronize(p->copy_stream)); p->nb_next = 0; }
  int rc = join_user(p);
  if (rc) return rc;
  CU(p, cudaStreamSynchronize(p->stream));
  rc = alloc_batch<T>(p, nb, false);
  if (rc) return rc;
  const size_t n2 = (size_t)nO * nb;
  // staging: y_o | y_unc | one column of xP | one column of y_unc
  rc = ensure(p, &p->d_stage, &p->cap_stage, (int64_t)(sizeof(T) * (2 * n2 + nX + nO)));
  if (rc) return rc;
  T* t = (T*)p->d_stage;
  T* col_x = t + 2 * n2;
  T* col_u = col_x + nX;
  CU(p, copy_in(p->d_xM, xM, sizeof(T) * (size_t)nC * nb, mem_kind, p->stream));
  if (xP_shared) {
    CU(p, copy_in(col_x, xP, sizeof(T) * (size_t)nX, mem_kind, p->stream));
    rc = expand_cols<T>(p, (T*)p->d_xP, col_x, nX, nb);
    if (rc) return rc;
  } else {
    CU(p, copy_in(p->d_xP, xP, sizeof(T) * (size_t)nX * nb, mem_kind, p->stream));
  }
  p->has_obs = 1;
  CU(p, copy_in(t, y_o, sizeof(T) * n2, mem_kind, p->stream));
  if (y_unc_shared) {
    CU(p, copy_in(col_u, y_unc, sizeof(T) * (size_t)nO, mem_kind, p->stream));
    rc = expand_cols<T>(p, t + n2, col_u, nO, nb);
    if (rc) return rc;
  } else {
    CU(p, copy_in(t + n2, y_unc, sizeof(T) * n2, mem_kind, p->stream));
  }
  return finish_batch<T>(p, nb, (const T*)p->d_xP, t, t + n2);
}

// upload of the NEXT batch on the plan's copy stream; nothing of the current batch is touched
template <typename T>
static int prefetch_t(hvi_plan* p, int64_t nb, const void* xM, const void* xP, const void* y_o, const void* y_unc) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar;
  if (!p->copy_stream) CU(p, cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
  if (!p->ev_copied) CU(p, cudaEventCreateWithFlags(&p->ev_copied, cudaEventDisableTiming));
  if (p->nb_next > 0) CU(p, cudaStreamSynchronize(p->copy_stream));   // replaces a batch that was never committed
  p->nb_next = 0;
  if (p->cap_next < nb) {
    for (void** q : {&p->d_xM_next, &p->d_xP_next}) { if (*q) cudaFree(*q); *q = nullptr; }
    p->cap_next = 0;
    CU(p, cudaMalloc(&p->d_xM_next, sizeof(T) * (size_t)nC * nb));
    CU(p, cudaMalloc(&p->d_xP_next, sizeof(T) * (size_t)2 * nO * nb));
    p->cap_next = nb;
  }
  const size_t n1 = (size_t)2 * nO * nb, n2 = (size_t)nO * nb;
  p->has_obs_next = (y_o && y_unc) ? 1 : 0;
  cudaStream_t cs = p->copy_stream;
  CU(p, cudaMemcpyAsync(p->d_xM_next, xM, sizeof(T) * (size_t)nC * nb, cudaMemcpyHostToDevice, cs));
  CU(p, cudaMemcpyAsync(p->d_xP_next, xP, sizeof(T) * n1, cudaMemcpyHostToDevice, cs));
  if (p->has_obs_next) {
    // the staging buffer is free: every k_preprocess that read it has completed (set_data / commit_data synchronise)
    int rc = ensure(p, &p->d_stage, &p->cap_stage, (int64_t)(sizeof(T) * 2 * n2));
    if (rc) return rc;
    T* t = (T*)p->d_stage;
    CU(p, cudaMemcpyAsync(t, y_o, sizeof(T) * n2, cudaMemcpyHostToDevice, cs));
    CU(p, cudaMemcpyAsync(t + n2, y_unc, sizeof(T) * n2, cudaMemcpyHostToDevice, cs));
  }
  CU(p, cudaEventRecord(p->ev_copied, cs));
  p->nb_next = nb;
  return HVI_OK;
}

template <typename T>
static int commit_t(hvi_plan* p) {
  const int nO = p->desc.n_obs;
  const int64_t nb = p->nb_next;
  int rcj = join_user(p);
  if (rcj) return rcj;
  CU(p, cudaStreamSynchronize(p->stream));   // evaluations in flight still read the old batch
  CU(p, cudaStreamWaitEvent(p->stream, p->ev_copied, 0));
  std::swap(p->d_xM, p->d_xM_next);
  std::swap(p->d_xP, p->d_xP_next);
  std::swap(p->cap_cur, p->cap_next);
  p->nb_next = 0;
  int rc = alloc_batch<T>(p, nb, true);
  if (rc) return rc;
  p->has_obs = p->has_obs_next;
  const T* t = (const T*)p->d_stage;
  const size_t n2 = (size_t)nO * nb;
  return finish_batch<T>(p, nb, (const T*)p->d_xP, p->has_obs ? t : nullptr, p->has_obs ? t + n2 : nullptr);
}

extern "C" int hvi_plan_set_data(hvi_plan* p, int64_t n_site, const void* xM, const void* xP, const void* y_o,
                                 const void* y_unc, int mem_kind) {
  if (!p) return HVI_EINVAL;
  if (n_site < 1 || !xM || !xP) PLAN_FAIL(p, HVI_EINVAL, "set_data: n_site must be >= 1 and xM, xP non-NULL");
  if ((y_o == nullptr) != (y_unc == nullptr)) PLAN_FAIL(p, HVI_EINVAL, "set_data: y_o and y_unc must both be given or both be NULL");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "set_data: bad mem_kind");
  CU(p, cudaSetDevice(p->desc.device));
  if (p->cf.sepvar && n_site > p->cf.n_site_total) PLAN_FAIL(p, HVI_EINVAL, "set_data: batch has more sites than n_site_total");
  if (p->d_isites) { cudaStreamSynchronize(p->stream); cudaFree(p->d_isites); p->d_isites = nullptr; p->n_isites = 0; }   // back to the identity map
  return p->dtype == HVI_F32 ? set_data_t<float>(p, n_site, xM, xP, y_o, y_unc, mem_kind)
                             : set_data_t<double>(p, n_site, xM, xP, y_o, y_unc, mem_kind);
}

extern "C" int hvi_plan_set_data_bcast(hvi_plan* p, int64_t n_site, const void* xM, const void* xP, int xP_shared,
                                       const void* y_o, const void* y_unc, int y_unc_shared, int mem_kind) {
  if (!p) return HVI_EINVAL;
  if (n_site < 1 || !xM || !xP || !y_o || !y_unc) PLAN_FAIL(p, HVI_EINVAL, "set_data_bcast: n_site must be >= 1 and all arrays non-NULL");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "set_data_bcast: bad mem_kind");
  CU(p, cudaSetDevice(p->desc.device));
  if (p->cf.sepvar && n_site > p->cf.n_site_total) PLAN_FAIL(p, HVI_EINVAL, "set_data_bcast: batch has more sites than n_site_total");
  if (p->d_isites) { cudaStreamSynchronize(p->stream); cudaFree(p->d_isites); p->d_isites = nullptr; p->n_isites = 0; }
  return p->dtype == HVI_F32 ? set_data_bcast_t<float>(p, n_site, xM, xP, xP_shared, y_o, y_unc, y_unc_shared, mem_kind)
                             : set_data_bcast_t<double>(p, n_site, xM, xP, xP_shared, y_o, y_unc, y_unc_shared, mem_kind);
}

extern "C" int hvi_plan_prefetch_data(hvi_plan* p, int64_t n_site, const void* xM, const void* xP, const void* y_o,
                                      const void* y_unc) {
  if (!p) return HVI_EINVAL;
  if (n_site < 1 || !xM || !xP) PLAN_FAIL(p, HVI_EINVAL, "prefetch_data: n_site must be >= 1 and xM, xP non-NULL");
  if ((y_o == nullptr) != (y_unc == nullptr)) PLAN_FAIL(p, HVI_EINVAL, "prefetch_data: y_o and y_unc must both be given or both be NULL");
  if (p->gpbm) PLAN_FAIL(p, HVI_EINVAL, "prefetch_data: not available with a generic process model (use hvi_plan_set_data)");
  CU(p, cudaSetDevice(p->desc.device));
  if (p->cf.sepvar && n_site > p->cf.n_site_total) PLAN_FAIL(p, HVI_EINVAL, "prefetch_data: batch has more sites than n_site_total");
  return p->dtype == HVI_F32 ? prefetch_t<float>(p, n_site, xM, xP, y_o, y_unc) : prefetch_t<double>(p, n_site, xM, xP, y_o, y_unc);
}

extern "C" int hvi_plan_commit_data(hvi_plan* p) {
  if (!p) return HVI_EINVAL;
  if (p->nb_next <= 0) PLAN_FAIL(p, HVI_ENODATA, "commit_data: no prefetched batch (call hvi_plan_prefetch_data first)");
  CU(p, cudaSetDevice(p->desc.device));
  if (p->d_isites) { cudaStreamSynchronize(p->stream); cudaFree(p->d_isites); p->d_isites = nullptr; p->n_isites = 0; }   // back to the identity map
  return p->dtype == HVI_F32 ? commit_t<float>(p) : commit_t<double>(p);
}

// ---- device-resident dataset and minibatch selection -----------------------------------------
// gdev_hybridproblem_dataloader (src/AbstractHybridProblem.jl:220-250) moves the whole training set to the device once;
// solve walks its minibatches (src/HybridSolver.jl:259-260; consecutive column ranges, last partial batch dropped,
// src/AbstractHybridProblem.jl:206-207).  set_dataset is the one upload; select_range / select_batch make a minibatch
// the registered batch with one gather kernel and no host round trip.
template <typename T>
static int set_dataset_t(hvi_plan* p, int64_t N, const void* xM, const void* xP, const void* y_o, const void* y_unc,
                         int mem_kind) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar;
  int rc = join_user(p);
  if (rc) return rc;
  CU(p, cudaStreamSynchronize(p->stream));
  for (void** q : {&p->ds_xM, &p->ds_xP, &p->ds_yo, &p->ds_yu}) { if (*q) cudaFree(*q); *q = nullptr; }
  p->ds_n = 0;
  p->ds_has_obs = (y_o && y_unc) ? 1 : 0;
  CU(p, cudaMalloc(&p->ds_xM, sizeof(T) * (size_t)nC * N));
  CU(p, cudaMalloc(&p->ds_xP, sizeof(T) * (size_t)2 * nO * N));
  CU(p, copy_in(p->ds_xM, xM, sizeof(T) * (size_t)nC * N, mem_kind, p->stream));
  CU(p, copy_in(p->ds_xP, xP, sizeof(T) * (size_t)2 * nO * N, mem_kind, p->stream));
  if (p->ds_has_obs) {
    CU(p, cudaMalloc(&p->ds_yo, sizeof(T) * (size_t)nO * N));
    CU(p, cudaMalloc(&p->ds_yu, sizeof(T) * (size_t)nO * N));
    CU(p, copy_in(p->ds_yo, y_o, sizeof(T) * (size_t)nO * N, mem_kind, p->stream));
    CU(p, copy_in(p->ds_yu, y_unc, sizeof(T) * (size_t)nO * N, mem_kind, p->stream));
  }
  CU(p, cudaStreamSynchronize(p->stream));   // the caller's arrays are free again at return
  p->ds_n = N;
  return HVI_OK;
}

// MeanVarSep reads its per-site parameters through d_isites: the gather kernel writes the selected sites there
static int select_sites(hvi_plan* p, int64_t n);

// buffers and bookkeeping of a minibatch of n sites of the device-resident dataset (no kernel)
template <typename T>
static int select_prepare(hvi_plan* p, int64_t n) {
  if (p->nb_next > 0) { CU(p, cudaStreamSynchronize(p->copy_stream)); p->nb_next = 0; }
  if (n != p->nb || p->cap_cur < n) {
    CU(p, cudaStreamSynchronize(p->stream));   // buffers of the old size may still be in use
    int rc = alloc_batch<T>(p, n, false);
    if (rc) return rc;
  }
  int rc = select_sites(p, n);
  if (rc) return rc;
  p->has_obs = p->ds_has_obs;
  return HVI_OK;
}

template <typename T>
static int select_t(hvi_plan* p, const int64_t* idx_dev, int64_t first, int64_t n,
                    const unsigned long long* ctr = nullptr, int64_t n_batches = 1) {
  const int nO = p->desc.n_obs, nC = p->desc.n_covar;
  int rc = select_prepare<T>(p, n);
  if (rc) return rc;
  Launch L{p->stream, p->sm_count, &p->launches};
  int nrows = 0;
  CU(p, launch_gather_preprocess<T>(L, nC, nO, n, idx_dev, first, (const T*)p->ds_xM, (const T*)p->ds_xP,
                                    (const T*)p->ds_yo, (const T*)p->ds_yu, (T*)p->d_xM, (T*)p->d_xP, (T*)p->d_rec,
                                    p->d_part_stream, &nrows, ctr, n_batches, p->d_isites));
  return finish_consts(p, nrows, false);
}

extern "C" int hvi_plan_set_dataset(hvi_plan* p, int64_t n_site_total, const void* xM, const void* xP, const void* y_o,
                                    const void* y_unc, int mem_kind) {
  if (!p) return HVI_EINVAL;
  if (n_site_total < 1 || !xM || !xP) PLAN_FAIL(p, HVI_EINVAL, "set_dataset: n_site_total must be >= 1 and xM, xP non-NULL");
  if ((y_o == nullptr) != (y_unc == nullptr)) PLAN_FAIL(p, HVI_EINVAL, "set_dataset: y_o and y_unc must both be given or both be NULL");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "set_dataset: bad mem_kind");
  if (p->gpbm) PLAN_FAIL(p, HVI_EINVAL, "set_dataset: not available with a generic process model (use hvi_plan_set_data)");
  if (p->cf.sepvar && n_site_total != p->cf.n_site_total)
    PLAN_FAIL(p, HVI_EINVAL, "set_dataset: MeanVarSep needs the dataset to hold exactly n_site_total sites");
  CU(p, cudaSetDevice(p->desc.device));
  return p->dtype == HVI_F32 ? set_dataset_t<float>(p, n_site_total, xM, xP, y_o, y_unc, mem_kind)
                             : set_dataset_t<double>(p, n_site_total, xM, xP, y_o, y_unc, mem_kind);
}

extern "C" int hvi_plan_select_range(hvi_plan* p, int64_t first, int64_t n) {
  if (!p) return HVI_EINVAL;
  if (p->ds_n <= 0) PLAN_FAIL(p, HVI_ENODATA, "select_range: hvi_plan_set_dataset has not been called");
  if (first < 0 || n < 1 || first + n > p->ds_n) PLAN_FAIL(p, HVI_EINVAL, "select_range: [first, first + n) is not inside the dataset");
  CU(p, cudaSetDevice(p->desc.device));
  int rc = join_user(p);
  if (rc) return rc;
  return p->dtype == HVI_F32 ? select_t<float>(p, nullptr, first, n) : select_t<double>(p, nullptr, first, n);
}

extern "C" int hvi_plan_select_batch(hvi_plan* p, const int64_t* i_sites, int64_t n, int mem_kind) {
  if (!p) return HVI_EINVAL;
  if (p->ds_n <= 0) PLAN_FAIL(p, HVI_ENODATA, "select_batch: hvi_plan_set_dataset has not been called");
  if (!i_sites || n < 1) PLAN_FAIL(p, HVI_EINVAL, "select_batch: need n >= 1 indices");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "select_batch: bad mem_kind");
  if (mem_kind == HVI_MEM_HOST)
    for (int64_t i = 0; i < n; i++)
      if (i_sites[i] < 0 || i_sites[i] >= p->ds_n) PLAN_FAIL(p, HVI_EINVAL, "select_batch: index out of range");
  CU(p, cudaSetDevice(p->desc.device));
  int rc = join_user(p);
  if (rc) return rc;
  if (p->cap_idx < n) {
    CU(p, cudaStreamSynchronize(p->stream));
    if (p->d_idx) cudaFree(p->d_idx);
    p->d_idx = nullptr; p->cap_idx = 0;
    CU(p, cudaMalloc((void**)&p->d_idx, sizeof(int64_t) * (size_t)n));
    p->cap_idx = n;
  }
  CU(p, copy_in(p->d_idx, i_sites, sizeof(int64_t) * (size_t)n, mem_kind, p->stream));
  return p->dtype == HVI_F32 ? select_t<float>(p, p->d_idx, 0, n) : select_t<double>(p, p->d_idx, 0, n);
}

static int select_sites(hvi_plan* p, int64_t n) {
  if (!p->cf.sepvar) {
    if (p->d_isites) { CU(p, cudaStreamSynchronize(p->stream)); cudaFree(p->d_isites); p->d_isites = nullptr; p->n_isites = 0; }
    return HVI_OK;
  }
  // (re)allocate only when the batch size changes: the array is rewritten in stream order by the gather kernel
  if (!p->d_isites || n != p->n_isites) {
    CU(p, cudaStreamSynchronize(p->stream));
    if (p->d_isites) cudaFree(p->d_isites);
    p->d_isites = nullptr; p->n_isites = 0;
    CU(p, cudaMalloc((void**)&p->d_isites, sizeof(int64_t) * (size_t)n));
    p->n_isites = n;
  }
  return HVI_OK;
}

static int ensure(hvi_plan* p, void** buf, int64_t* cap, int64_t need_bytes) {
  if (*cap >= need_bytes) return HVI_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr; *cap = 0;
  CU(p, cudaMalloc(buf, (size_t)need_bytes));
  *cap = need_bytes;
  return HVI_OK;
}

template <typename T> static const Cfg<T>& cfg_of(const hvi_plan* p);
template <> const Cfg<float>& cfg_of<float>(const hvi_plan* p) { return p->cf; }
template <> const Cfg<double>& cfg_of<double>(const hvi_plan* p) { return p->cd; }

// g forward for all paths: register-tiled / generic CUDA-core kernels, or the wide tensor-core chain
// (xM_sub / nb_sub: a range of the batch's sites instead of all of them)
template <typename T>
static int run_g_fwd(hvi_plan* p, const Launch& L, const Cfg<T>& cfg, const T* phi, T* out, int M_units, const T* zetaP,
                     WideCache cache = WideCache{nullptr, 0, 0}, const T* xM_sub = nullptr, int64_t nb_sub = -1) {
  const T* xM = xM_sub ? xM_sub : (const T*)p->d_xM;
  const int64_t nb = nb_sub >= 0 ? nb_sub : p->nb;
  if (p->wide) {
    int rc = ensure(p, &p->d_wide, &p->cap_wide, (int64_t)(sizeof(T) * wide_work_elems(cfg.max_width, cfg.nphig)));
    if (rc) return rc;
    CU(p, launch_mlp_fwd_wide<T>(L, cfg, phi, xM, nb, out, M_units, zetaP, (T*)p->d_wide, cache));
    return HVI_OK;
  }
  CU(p, launch_mlp_fwd<T>(L, cfg, phi, xM, nb, out, M_units, zetaP));
  return HVI_OK;
}

// Site blocks of the wide applicator with pbm covariates (BASELINE config 5).  Pass 2 runs g on n_MC columns per site and
// its activations ((n_layers − 1)·width·4 bytes per column as bf16 hi/lo planes) exceed HBM by far, so an evaluation that
// does forward(all) → stream(all) → backward(all) must recompute most of them in the backward pass.  The only coupling
// between sites is the sum over sites at the very end, so the evaluation walks blocks of sites instead and does
// forward → stream → backward per block while the block's activations are still in the cache: nothing is recomputed.
// Returns the sites per block (0: one block = the unblocked schedule) and the rows of a cache slot.
// HVI_WIDE_BLOCK_SITES forces a block size (tests), HVI_WIDE_CACHE_MB caps the cache.
static int64_t wide_block_sites(hvi_plan* p, int nL, int maxw, int M, int64_t* slot_rows, int shrink = 0) {
  *slot_rows = 0;
  const int64_t col_bytes = (int64_t)(nL - 1) * maxw * 4;
  const int64_t nch = bf3_chunks(p->nb), WIDE_CHUNK = bf3_chunk_rows();
  if (const char* e = getenv("HVI_WIDE_BLOCK_SITES")) {
    int64_t b = (atoll(e) + 127) / 128 * 128;
    if (b <= 0 || b >= p->nb) return 0;
    if (b < WIDE_CHUNK) *slot_rows = b; else b = b / WIDE_CHUNK * WIDE_CHUNK;
    return b;
  }
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 0;
  // what the cache may take: everything that is free now (plus what it holds already) minus head room for the buffers
  // allocated after it (per-draw means of a block, caller's own allocations)
  int64_t avail = (int64_t)fr + p->cap_wide_cache - ((int64_t)8 << 30);
  if (const char* e = getenv("HVI_WIDE_CACHE_MB")) { const int64_t cap = atoll(e) << 20; if (cap < avail) avail = cap; }
  avail >>= shrink;   // after a failed allocation (fragmented free memory): half, a quarter, ...
  const int64_t rows1 = nch == 1 ? (p->nb + 127) / 128 * 128 : 0;
  if (avail >= nch * (1 + (int64_t)M) * (int64_t)bf3_cache_slot_bytes(nL, maxw, rows1)) return 0;   // everything fits: one block
  const int64_t sites_max = avail / col_bytes / M;
  if (sites_max < 4096) return 0;            // too little memory for useful blocks: unblocked with recomputation
  if (sites_max >= WIDE_CHUNK) return sites_max / WIDE_CHUNK * WIDE_CHUNK;
  const int64_t nblocks = (p->nb + sites_max - 1) / sites_max;
  int64_t b = ((p->nb + nblocks - 1) / nblocks + 127) / 128 * 128;
  if (b > sites_max) b = sites_max / 128 * 128;
  *slot_rows = b;
  return b;
}

// enqueue the whole evaluation on `s`; all pointers are device pointers
template <typename T>
static int enqueue_eval(hvi_plan* p, int M, const T* phi, const T* zP, const T* zMs, double* out, T* grad_T,
                        cudaStream_t s, bool rng = false, uint64_t seed = 0, int64_t site_offset = 0,
                        const unsigned long long* seed_dev = nullptr) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  int rc = ensure(p, &p->d_pdraw, &p->cap_pdraw, (int64_t)(sizeof(T) * pdraw_elems(cfg.nP, M)));
  if (rc) return rc;
  Launch L{s, p->sm_count, &p->launches};
  EvalConsts<T>* ec = (EvalConsts<T>*)p->d_ec;
  T* pdraw = (T*)p->d_pdraw;
  int iev = 0;
#define MARK() do { if (p->profiling) CU(p, cudaEventRecord(p->ev[iev++], s)); } while (0)
  MARK();
  // ϕq → factors and per-draw θP is a single block of latency (≈ 11 µs) that only the streaming kernel waits for: it runs
  // on an auxiliary stream beside the applicator's forward pass (HVI_NO_OVERLAP=1 or profiling: in line)
  static const bool no_overlap = getenv("HVI_NO_OVERLAP") != nullptr;
  const bool overlap = !p->profiling && !no_overlap;
  if (overlap) {
    if (!p->aux_stream) {
      CU(p, cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking));
      CU(p, cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
      CU(p, cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    CU(p, cudaEventRecord(p->ev_fork, s));
    CU(p, cudaStreamWaitEvent(p->aux_stream, p->ev_fork, 0));
    Launch La{p->aux_stream, p->sm_count, &p->launches};
    CU(p, launch_prologue<T>(La, cfg, phi, zP, M, ec, pdraw));
    CU(p, cudaEventRecord(p->ev_join, p->aux_stream));
  } else {
    CU(p, launch_prologue<T>(L, cfg, phi, zP, M, ec, pdraw));
  }
  MARK();
  const bool pbm = cfg.npc > 0;
  const T* zetaP = pdraw + (size_t)cfg.nP * M;
  // wide Float32 applicator: the forward pass keeps the activations for the backward pass -- of every chunk where they fit,
  // of one block of sites at a time where they do not (pbm covariates), of as many chunks as fit otherwise
  WideCache cache_s{nullptr, 0, 0}, cache_u{nullptr, 0, 0};
  int64_t blk = 0, slot_rows = 0;   // blk > 0: site-blocked schedule
  if constexpr (sizeof(T) == 4) {
    if (p->wide && bf3_eligible(cfg)) {
      if (pbm && !p->gpbm) {
        // the work spaces first, so that the free memory seen by the block size is what is really left
        rc = ensure(p, &p->d_wide, &p->cap_wide, (int64_t)(sizeof(T) * wide_work_elems(cfg.max_width, cfg.nphig)));
        if (rc) return rc;
        rc = ensure(p, &p->d_wide_bwd, &p->cap_wide_bwd, (int64_t)wide_bwd_work_bytes(cfg.nL, cfg.max_width, cfg.nphig));
        if (rc) return rc;
        blk = wide_block_sites(p, cfg.nL, cfg.max_width, M, &slot_rows);
      }
      // the cache of one block; a failed allocation (free memory in pieces) is retried with smaller blocks
      for (int shrink = 1; blk > 0; shrink++) {
        const int64_t need = (int64_t)M * bf3_chunks(blk) * (int64_t)bf3_cache_slot_bytes(cfg.nL, cfg.max_width, slot_rows);
        rc = ensure(p, &p->d_wide_cache, &p->cap_wide_cache, need);
        if (rc == HVI_OK) break;
        (void)cudaGetLastError();
        if (shrink > 3) return rc;
        blk = wide_block_sites(p, cfg.nL, cfg.max_width, M, &slot_rows, shrink);   // 0: the unblocked schedule below
      }
      if (blk > 0) {
        cache_u = WideCache{p->d_wide_cache, (int64_t)M * bf3_chunks(blk), 0, slot_rows};
      } else {
        const int64_t nch = bf3_chunks(p->nb), need = nch * (1 + (pbm ? M : 0));
        // batches below one chunk: slots of the batch's own size (a 1 500-site test problem would take 2 GB per slot otherwise)
        const int64_t rows1 = nch == 1 ? (p->nb + 127) / 128 * 128 : 0;
        const int64_t slot = (int64_t)bf3_cache_slot_bytes(cfg.nL, cfg.max_width, rows1);
        if (p->cap_wide_cache < need * slot) {
          size_t fr = 0, tot = 0;
          CU(p, cudaMemGetInfo(&fr, &tot));
          int64_t can = ((int64_t)fr + p->cap_wide_cache) / 2 / slot;
          if (can > need) can = need;
          if (can * slot > p->cap_wide_cache) {
            rc = ensure(p, &p->d_wide_cache, &p->cap_wide_cache, can * slot);
            if (rc) return rc;
          }
        }
        const int64_t have = p->cap_wide_cache / slot;
        cache_s = WideCache{p->d_wide_cache, have, 0, rows1};
        cache_u = WideCache{p->d_wide_cache, have, nch, rows1};
      }
    }
  }
  const int64_t nbB_max = blk > 0 ? blk : p->nb;    // sites whose per-draw means are alive at once
  rc = run_g_fwd<T>(p, L, cfg, phi, (T*)p->d_mu, 0, nullptr, cache_s);
  if (rc) return rc;
  if (overlap) CU(p, cudaStreamWaitEvent(s, p->ev_join, 0));   // pass 2 of g and the streaming kernel read the prologue's output
  if (pbm) {  // pass 2: g([xM; ζP_m[idx]]) for every draw (src/elbo.jl:508-515)
    const int64_t nmu = (int64_t)sizeof(T) * M * cfg.nM * nbB_max;
    rc = ensure(p, &p->d_MU, &p->cap_MU, nmu);
    if (rc) return rc;
    rc = ensure(p, &p->d_MUBAR, &p->cap_MUBAR, nmu);
    if (rc) return rc;
    rc = ensure(p, (void**)&p->d_part_x, &p->cap_part_x, (int64_t)sizeof(double) * 2 * p->max_blk * (M + 1) * cfg.npc);
    if (rc) return rc;
    rc = ensure(p, (void**)&p->d_xred, &p->cap_xred, (int64_t)sizeof(double) * (M + 1) * cfg.npc);
    if (rc) return rc;
    if (blk == 0) {
      rc = run_g_fwd<T>(p, L, cfg, phi, (T*)p->d_MU, M, zetaP, cache_u);
      if (rc) return rc;
    }
  }
  MARK();
  // the streaming kernel on sites [off, off + n) of the batch; per-draw means and cotangents are local to the range
  auto stream_args = [&](int64_t off, int64_t n) {
    StreamArgs<T> a;
    a.rec = (const T*)p->d_rec + (size_t)off * 4 * cfg.nO; a.mu = (const T*)p->d_mu + (size_t)off * cfg.nM;
    a.zMs = zMs ? zMs + (size_t)off * cfg.nM * M : nullptr;
    a.pd = pdraw + (((size_t)4 * (cfg.nP > 0 ? cfg.nP : 1) * M + 3) & ~(size_t)3); a.pd_in_smem = 0;
    a.ec = ec; a.mubar = (T*)p->d_mubar + (size_t)off * cfg.nM; a.partials = p->d_part_stream; a.nb = n; a.M = M;
    a.MU = pbm ? (const T*)p->d_MU : nullptr; a.MUBAR = pbm ? (T*)p->d_MUBAR : nullptr;
    a.staged = !rng && (M % (16 / (int)sizeof(T)) == 0) && (((uintptr_t)a.zMs & 15) == 0);
    a.rng = rng ? 1 : 0; a.seed = seed; a.seed_dev = seed_dev; a.col_offset = (site_offset + off) * cfg.nM;
    a.phi = phi; a.i_sites = p->d_isites ? p->d_isites + off : nullptr; a.out = out; a.grad_T = grad_T;
    return a;
  };
  if (cfg.sepvar) {   // sites outside the batch have zero gradient
    const size_t big = (size_t)cfg.nM * cfg.n_site_total;
    CU(p, cudaMemsetAsync(out + cfg.o_coef, 0, sizeof(double) * big, s));
    if (grad_T) CU(p, cudaMemsetAsync(grad_T + cfg.o_coef, 0, sizeof(T) * big, s));
  }
  int nrows = 0, nblk = 0, nblk2 = 0;
  const bool penalty = generic_pbm_has_penalty(p->gpbm);
  double* part_xs = pbm ? p->d_part_x : nullptr;
  double* part_xu = pbm ? p->d_part_x + (size_t)p->max_blk * cfg.npc : nullptr;
  const double* part_g = p->d_part_mlp;
  if (blk > 0) {
    // site-blocked wide evaluation: pass-2 forward, stream and pass-2 backward per block (the time of the whole loop is
    // reported in the "stream" slot of hvi_plan_kernel_times); the pass-1 backward follows once μ̄ of every site is known
    CU(p, cudaMemsetAsync(p->d_gacc, 0, sizeof(double) * (size_t)cfg.nphig, s));
    CU(p, cudaMemsetAsync(p->d_part_x, 0, sizeof(double) * (size_t)2 * p->max_blk * (M + 1) * cfg.npc, s));
    for (int64_t off = 0; off < p->nb; off += blk) {
      const int64_t n = p->nb - off < blk ? p->nb - off : blk;
      const T* xM_b = (const T*)p->d_xM + (size_t)off * cfg.nC;
      rc = run_g_fwd<T>(p, L, cfg, phi, (T*)p->d_MU, M, zetaP, cache_u, xM_b, n);
      if (rc) return rc;
      StreamArgs<T> a = stream_args(off, n);
      a.partials = p->d_part_stream + (size_t)nrows * nacc(cfg.nP, cfg.nM);
      int nr = 0;
      CU(p, launch_stream<T>(L, cfg, a, &nr, p->max_rows - nrows));
      nrows += nr;
      CU(p, launch_mlp_bwd_wide<T>(L, cfg, phi, xM_b, n, (const T*)p->d_MUBAR, M, zetaP, p->d_wide_bwd, p->d_gacc, part_xu,
                                   cache_u));
      if (p->max_rows - nrows < 1) { p->err = "hvi: too many site blocks for the partial-sum rows"; return HVI_EINVAL; }
    }
    MARK();
    CU(p, launch_mlp_bwd_wide<T>(L, cfg, phi, (const T*)p->d_xM, p->nb, (const T*)p->d_mubar, 0, nullptr, p->d_wide_bwd,
                                 p->d_gacc, part_xs, cache_s));
    part_g = p->d_gacc;
    nblk = 1; nblk2 = 0;
  } else {
    StreamArgs<T> a = stream_args(0, p->nb);
    if (p->gpbm)
      CU(p, launch_generic_stream<T>(L, p->gpbm, cfg, a, (const T*)p->d_xP, (const T*)p->d_yo, (const T*)p->d_yu, p->pbm_fix,
                                     penalty ? p->d_pen : nullptr, &nrows, p->max_rows));
    else
      CU(p, launch_stream<T>(L, cfg, a, &nrows, p->max_rows));
    MARK();
    if (p->wide) {
      // wide applicator: tensor-core backward into one double accumulator (a single "partial row")
      rc = ensure(p, &p->d_wide_bwd, &p->cap_wide_bwd, (int64_t)wide_bwd_work_bytes(cfg.nL, cfg.max_width, cfg.nphig));
      if (rc) return rc;
      CU(p, cudaMemsetAsync(p->d_gacc, 0, sizeof(double) * (size_t)cfg.nphig, s));
      if (pbm) CU(p, cudaMemsetAsync(p->d_part_x, 0, sizeof(double) * (size_t)2 * p->max_blk * (M + 1) * cfg.npc, s));
      CU(p, launch_mlp_bwd_wide<T>(L, cfg, phi, (const T*)p->d_xM, p->nb, (const T*)p->d_mubar, 0, nullptr, p->d_wide_bwd,
                                   p->d_gacc, part_xs, cache_s));
      if (pbm)
        CU(p, launch_mlp_bwd_wide<T>(L, cfg, phi, (const T*)p->d_xM, p->nb, (const T*)p->d_MUBAR, M, zetaP, p->d_wide_bwd,
                                     p->d_gacc, part_xu, cache_u));
      part_g = p->d_gacc;
      nblk = 1; nblk2 = 0;
    } else {
      // μ_ζ of the forward passes of this evaluation is still in d_mu / d_MU: the backward pass reads Φ⁻¹(p) off it
      CU(p, launch_mlp_bwd<T>(L, cfg, phi, (const T*)p->d_xM, (const T*)p->d_mubar, p->nb, p->d_part_mlp, &nblk, p->max_blk, 0,
                              nullptr, part_xs, (const T*)p->d_mu));
      if (pbm)
        CU(p, launch_mlp_bwd<T>(L, cfg, phi, (const T*)p->d_xM, (const T*)p->d_MUBAR, p->nb,
                                p->d_part_mlp + (size_t)nblk * cfg.nphig, &nblk2, p->max_blk, M, zetaP, part_xu, (const T*)p->d_MU));
    }
  }
  MARK();
  CU(p, launch_finalize<T>(L, cfg, phi, zP, pdraw, ec, p->d_part_stream, nrows, part_g, nblk + nblk2, p->d_bconst,
                           p->nb, M, out, grad_T, p->d_red, part_xs, p->wide ? 1 : nblk, part_xu, p->wide ? 1 : nblk2, p->d_xred));
  if (penalty) CU(p, launch_add_penalty(L, p->d_pen, nrows, M, cfg.nphi, out));
  MARK();
#undef MARK
  p->n_ev = iev;
  return HVI_OK;
}

// small batches with device draws: the whole evaluation (and optionally the minibatch gather before it and the Adam update
// after it) as ONE launch of one block
template <typename T>
static bool fused_ok(const hvi_plan* p, int64_t nb, int M, bool looped = false) {
  // one block does the work of the whole GPU: the single launch wins only while its ≈ 25 µs + 0.3 µs per site stay below the
  // ≈ 45 µs of the graph-replayed latency-bound launches (profiles/r02s_small_batch_latency_loop_in_kernel.txt, Float32, µs
  // per iteration with the optimiser loop inside the launch: 20 x 3: 29 against 43; 20 x 12: 33 against 47; 50 x 12: 41
  // against 47; 200 x 12: 87 against 47; a single evaluation without the loop: 37 against 43 at 20 x 3) -- use_fused = 1
  // applies that break-even, 2 forces the path wherever it is supported
  // looped: the iterations of an optimiser call run inside one launch (no launch, warm instruction cache): the single block
  // stays ahead up to ≈ 72 sites (50 x 12: 32 against 43 µs) instead of 48 for a lone evaluation
  const int64_t nb_max = looped ? 72 : 48;
  const bool fits = p->use_fused >= 2 ? (nb <= 1024 && nb * (int64_t)M <= 16384) : (nb <= nb_max && nb * (int64_t)M <= 1024);
  return p->use_fused && !p->wide && !p->gpbm && !p->profiling && fits &&
         small_iter_supported<T>(cfg_of<T>(const_cast<hvi_plan*>(p)), nb, M);
}

template <typename T>
static int enqueue_fused(hvi_plan* p, int M, T* phi, double* out, T* grad_T, cudaStream_t s, uint64_t seed, int64_t site_offset,
                         bool adam, int64_t batch_size, int64_t first_batch, int n_inner = 1) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  int rc = ensure(p, &p->d_pdraw, &p->cap_pdraw, (int64_t)(sizeof(T) * pdraw_elems(cfg.nP, M)));
  if (rc) return rc;
  rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)sizeof(T) * M * (cfg.nP > 0 ? cfg.nP : 1));
  if (rc) return rc;
  SmallArgs<T> a;
  memset(&a, 0, sizeof a);
  T* pdraw = (T*)p->d_pdraw;
  a.st.rec = (const T*)p->d_rec; a.st.mu = (const T*)p->d_mu; a.st.zMs = nullptr;
  a.st.pd = pdraw + (((size_t)4 * (cfg.nP > 0 ? cfg.nP : 1) * M + 3) & ~(size_t)3);
  a.st.ec = (const EvalConsts<T>*)p->d_ec; a.st.mubar = (T*)p->d_mubar; a.st.partials = p->d_part_stream;
  a.st.nb = p->nb; a.st.M = M; a.st.rng = 1; a.st.seed = seed; a.st.col_offset = site_offset * cfg.nM;
  a.st.phi = phi; a.st.i_sites = nullptr; a.st.out = out; a.st.grad_T = grad_T;
  a.phi = phi; a.zP = (T*)p->d_zP; a.pdraw = pdraw; a.xM = (const T*)p->d_xM;
  a.part_mlp = p->d_part_mlp; a.red = p->d_red; a.out = out; a.grad_T = grad_T; a.bconst = p->d_bconst;
  if (batch_size > 0) {
    a.ds_xM = (const T*)p->ds_xM; a.ds_xP = (const T*)p->ds_xP; a.ds_yo = (const T*)p->ds_yo; a.ds_yu = (const T*)p->ds_yu;
    a.xM_b = (T*)p->d_xM; a.xP_b = (T*)p->d_xP; a.rec_w = (T*)p->d_rec;
    a.first = first_batch; a.n_batches = p->ds_n / batch_size;
  }
  if (adam) {
    a.adam_m = p->d_adam_m; a.adam_v = p->d_adam_v; a.trace = p->d_trace;
    a.eta = p->adam_eta; a.b1 = p->adam_b1; a.b2 = p->adam_b2; a.eps = p->adam_eps; a.ctr = p->d_ctr;
  }
  a.clocks = p->d_clocks;
  a.n_inner = n_inner;
  Launch L{s, p->sm_count, &p->launches};
  CU(p, launch_small_iter<T>(L, cfg, a));
  return HVI_OK;
}

// ---- NCCL through dlopen (the library does not link against it): one all-reduce per evaluation of a site-sharded job ----
struct HviNcclId { char internal[128]; };   // ncclUniqueId, passed by value to ncclCommInitRank
namespace {
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, HviNcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("HVI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      *(void**)(&api.GetUniqueId) = dlsym(api.h, "ncclGetUniqueId");
      *(void**)(&api.CommInitRank) = dlsym(api.h, "ncclCommInitRank");
      *(void**)(&api.CommDestroy) = dlsym(api.h, "ncclCommDestroy");
      *(void**)(&api.AllReduce) = dlsym(api.h, "ncclAllReduce");
      *(void**)(&api.GetErrorString) = dlsym(api.h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) api.h = nullptr;
    }
  }
  return api.h ? &api : nullptr;
}
#define NCCL_FAIL(p, rc_, what)                                                                                       \
  PLAN_FAIL(p, HVI_ENCCL, "%s: %s", what, (nccl_api() && nccl_api()->GetErrorString) ? nccl_api()->GetErrorString(rc_) : "NCCL error")

// Σ over the ranks of out[0 .. n_phi + 8) (double, in place) on stream s
static int nccl_allreduce_out(hvi_plan* p, double* out_dev, cudaStream_t s) {
  if (!p->nccl_comm) PLAN_FAIL(p, HVI_EINVAL, "hvi_plan_comm_init has not been called");
  NcclApi* api = nccl_api();
  if (!api) PLAN_FAIL(p, HVI_ENCCL, "libnccl.so.2 could not be loaded");
  const int nphi = p->dtype == HVI_F32 ? cfg_of<float>(p).nphi : cfg_of<double>(p).nphi;
  const int rc = api->AllReduce(out_dev, out_dev, (size_t)(nphi + 8), /* ncclFloat64 */ 8, /* ncclSum */ 0, p->nccl_comm, s);
  if (rc != 0) NCCL_FAIL(p, rc, "ncclAllReduce");
  return HVI_OK;
}

template <typename T>
static int eval_sync_t(hvi_plan* p, int M, const void* phi, const void* zP, const void* zMs, bool use_rng, uint64_t seed,
                       int64_t site_offset, int mem_kind, double comps[6], double* nL, void* grad, bool reduce = false) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  const int nphi = cfg.nphi;
  cudaStream_t s = p->stream;
  int rcj = join_user(p);
  if (rcj) return rcj;
  const T* dphi = (const T*)phi;
  if (mem_kind == HVI_MEM_HOST) {
    CU(p, cudaMemcpyAsync(p->d_phi, phi, sizeof(T) * nphi, cudaMemcpyHostToDevice, s));
    dphi = (const T*)p->d_phi;
  }
  const int64_t nzP = (int64_t)M * (cfg.nP > 0 ? cfg.nP : 1), nzM = (int64_t)M * cfg.nM * p->nb;
  const T *dzP = (const T*)zP, *dzM = (const T*)zMs;
  const bool fused = use_rng && fused_ok<T>(p, p->nb, M);   // small batch, device draws: one launch
  if (use_rng || mem_kind == HVI_MEM_HOST) {
    int rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)sizeof(T) * nzP);
    if (rc) return rc;
    if (use_rng) {
      // global-block draws: columns keyed far above any site column; the site draws are generated
      // inside the streaming kernel from the same (seed, column, draw) keys
      Launch L{s, p->sm_count, &p->launches};
      if (!fused) CU(p, launch_gen_normal<T>(L, (T*)p->d_zP, cfg.nP, M, seed, (int64_t)1 << 62));
      dzM = nullptr;
    } else {
      rc = ensure(p, &p->d_zMs, &p->cap_zMs, (int64_t)sizeof(T) * nzM);
      if (rc) return rc;
      if (cfg.nP > 0) CU(p, cudaMemcpyAsync(p->d_zP, zP, sizeof(T) * (size_t)M * cfg.nP, cudaMemcpyHostToDevice, s));
      CU(p, cudaMemcpyAsync(p->d_zMs, zMs, sizeof(T) * (size_t)nzM, cudaMemcpyHostToDevice, s));
      dzM = (const T*)p->d_zMs;
    }
    dzP = (const T*)p->d_zP;
  }
  const bool want_grad = grad != nullptr;
  int rc = fused ? enqueue_fused<T>(p, M, const_cast<T*>(dphi), p->d_out, want_grad ? (T*)p->d_grad : nullptr, s, seed,
                                    site_offset, false, 0, 0)
                 : enqueue_eval<T>(p, M, dphi, dzP, dzM, p->d_out, want_grad ? (T*)p->d_grad : nullptr, s, use_rng, seed,
                                   site_offset);
  if (rc) return rc;
  if (reduce) {
    // site-sharded job: Σ over the ranks of [∇ϕ; components; nL; non-finite] in double, on the same stream; the
    // gradient is then taken from the reduced vector (d_grad holds this shard's part only)
    rc = nccl_allreduce_out(p, p->d_out, s);
    if (rc) return rc;
    CU(p, cudaMemcpyAsync(p->h_out, p->d_out, sizeof(double) * (size_t)(nphi + 8), cudaMemcpyDeviceToHost, s));
    if (p->anomalies < 0) CU(p, cudaMemcpyAsync(p->h_bconst, p->d_bconst, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
    CU(p, cudaStreamSynchronize(s));
    if (p->anomalies < 0) p->anomalies = (int64_t)p->h_bconst[1];
    if (want_grad) for (int i = 0; i < nphi; i++) ((T*)grad)[i] = (T)p->h_out[i];
    const double* Cr = p->h_out + nphi;
    if (comps) for (int i = 0; i < 6; i++) comps[i] = Cr[i];
    if (nL) *nL = Cr[6];
    if (Cr[7] != 0.0) PLAN_FAIL(p, HVI_ENONFINITE, "non-finite value or gradient, or an observation that is finite where a driver is not, on some rank (%g entries)", Cr[7]);
    return HVI_OK;
  }
  CU(p, cudaMemcpyAsync(p->h_out + nphi, p->d_out + nphi, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
  if (p->anomalies < 0) CU(p, cudaMemcpyAsync(p->h_bconst, p->d_bconst, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
  if (want_grad) {
    if (mem_kind == HVI_MEM_HOST) CU(p, cudaMemcpyAsync(p->h_grad, p->d_grad, sizeof(T) * nphi, cudaMemcpyDeviceToHost, s));
    else CU(p, cudaMemcpyAsync(grad, p->d_grad, sizeof(T) * nphi, cudaMemcpyDeviceToDevice, s));
  }
  CU(p, cudaStreamSynchronize(s));
  if (p->anomalies < 0) p->anomalies = (int64_t)p->h_bconst[1];
  if (want_grad && mem_kind == HVI_MEM_HOST) memcpy(grad, p->h_grad, sizeof(T) * nphi);
  const double* C = p->h_out + nphi;
  if (comps) for (int i = 0; i < 6; i++) comps[i] = C[i];
  if (nL) *nL = C[6];
  if (p->anomalies > 0) {
    // reference behaviour: NaN loss; we evaluate the masked value and flag it
    PLAN_FAIL(p, HVI_ENONFINITE, "%lld observation(s) are finite where a driver is not: the reference returns NaN here",
              (long long)p->anomalies);
  }
  if (C[7] != 0.0) PLAN_FAIL(p, HVI_ENONFINITE, "non-finite value or gradient (%g entries)", C[7]);
  return HVI_OK;
}

// the device-resident entry points cannot return the reference's NaN for a finite observation under a missing driver
// (SURVEY Appendix C-2): refuse the batch up front when the count is known, and in any case the count is folded into
// out[n_phi + 7] by k_finalize
static int refuse_anomalies(hvi_plan* p) {
  if (p->anomalies > 0)
    PLAN_FAIL(p, HVI_ENONFINITE, "%lld observation(s) are finite where a driver is not: the reference returns NaN here",
              (long long)p->anomalies);
  return HVI_OK;
}

static int check_eval_args(hvi_plan* p, int32_t n_MC, const void* phi, int mem_kind) {
  if (!p) return HVI_EINVAL;
  if (p->nb <= 0) PLAN_FAIL(p, HVI_ENODATA, "hvi_plan_set_data has not been called");
  if (n_MC < 1) PLAN_FAIL(p, HVI_EINVAL, "n_MC must be >= 1");
  if (!phi) PLAN_FAIL(p, HVI_EINVAL, "phi is NULL");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "bad mem_kind");
  CU(p, cudaSetDevice(p->desc.device));
  return join_user(p);   // the plan's stream continues behind work left on a caller's stream
}

extern "C" int hvi_neg_elbo_grad(hvi_plan* p, int32_t n_MC, const void* phi, const void* zP, const void* zMs, int mem_kind,
                                 double comps[6], double* nL, void* grad) {
  int rc = check_eval_args(p, n_MC, phi, mem_kind);
  if (rc) return rc;
  if (!zMs || (p->desc.n_P > 0 && !zP)) PLAN_FAIL(p, HVI_EINVAL, "zP/zMs are NULL (use hvi_neg_elbo_grad_rng for device draws)");
  return p->dtype == HVI_F32 ? eval_sync_t<float>(p, n_MC, phi, zP, zMs, false, 0, 0, mem_kind, comps, nL, grad)
                             : eval_sync_t<double>(p, n_MC, phi, zP, zMs, false, 0, 0, mem_kind, comps, nL, grad);
}

extern "C" int hvi_neg_elbo_grad_rng(hvi_plan* p, int32_t n_MC, const void* phi, uint64_t seed, int64_t site_offset,
                                     int mem_kind, double comps[6], double* nL, void* grad) {
  int rc = check_eval_args(p, n_MC, phi, mem_kind);
  if (rc) return rc;
  return p->dtype == HVI_F32
             ? eval_sync_t<float>(p, n_MC, phi, nullptr, nullptr, true, seed, site_offset, mem_kind, comps, nL, grad)
             : eval_sync_t<double>(p, n_MC, phi, nullptr, nullptr, true, seed, site_offset, mem_kind, comps, nL, grad);
}

// ---- site-sharded jobs: the collective behind the C ABI ------------------------------------------
extern "C" int hvi_nccl_unique_id(void* id128) {
  if (!id128) return HVI_EINVAL;
  NcclApi* api = nccl_api();
  if (!api) return HVI_ENCCL;
  return api->GetUniqueId(id128) == 0 ? HVI_OK : HVI_ENCCL;
}

extern "C" int hvi_plan_comm_init(hvi_plan* p, const void* id128, int32_t rank, int32_t world) {
  if (!p) return HVI_EINVAL;
  if (!id128 || world < 1 || rank < 0 || rank >= world) PLAN_FAIL(p, HVI_EINVAL, "bad communicator arguments (rank %d of %d)", rank, world);
  if (p->nccl_comm) PLAN_FAIL(p, HVI_EINVAL, "the plan already has a communicator");
  NcclApi* api = nccl_api();
  if (!api) PLAN_FAIL(p, HVI_ENCCL, "libnccl.so.2 could not be loaded (set HVI_NCCL_LIB)");
  CU(p, cudaSetDevice(p->desc.device));
  HviNcclId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  const int rc = api->CommInitRank(&comm, world, id, rank);
  if (rc != 0) NCCL_FAIL(p, rc, "ncclCommInitRank");
  p->nccl_comm = comm; p->nccl_rank = rank; p->nccl_world = world;
  return HVI_OK;
}

extern "C" int hvi_plan_comm_destroy(hvi_plan* p) {
  if (!p) return HVI_EINVAL;
  if (p->nccl_comm) {
    NcclApi* api = nccl_api();
    cudaSetDevice(p->desc.device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (api) api->CommDestroy(p->nccl_comm);
    p->nccl_comm = nullptr; p->nccl_world = 1; p->nccl_rank = 0;
  }
  return HVI_OK;
}

extern "C" int hvi_allreduce_result(hvi_plan* p, double* out_dev, void* cuda_stream) {
  if (!p) return HVI_EINVAL;
  if (!out_dev) PLAN_FAIL(p, HVI_EINVAL, "out_dev is NULL");
  CU(p, cudaSetDevice(p->desc.device));
  return nccl_allreduce_out(p, out_dev, cuda_stream ? (cudaStream_t)cuda_stream : p->stream);
}

extern "C" int hvi_neg_elbo_grad_sharded(hvi_plan* p, int32_t n_MC, const void* phi, const void* zP, const void* zMs,
                                         uint64_t seed, int64_t site_offset, int mem_kind, double comps[6], double* nL,
                                         void* grad_host) {
  int rc = check_eval_args(p, n_MC, phi, mem_kind);
  if (rc) return rc;
  if (!p->nccl_comm) PLAN_FAIL(p, HVI_EINVAL, "hvi_plan_comm_init has not been called");
  if (zMs && p->desc.n_P > 0 && !zP) PLAN_FAIL(p, HVI_EINVAL, "zP is NULL");
  const bool rng = zMs == nullptr;
  return p->dtype == HVI_F32
             ? eval_sync_t<float>(p, n_MC, phi, zP, zMs, rng, seed, site_offset, mem_kind, comps, nL, grad_host, true)
             : eval_sync_t<double>(p, n_MC, phi, zP, zMs, rng, seed, site_offset, mem_kind, comps, nL, grad_host, true);
}

template <typename T>
static int eval_async_t(hvi_plan* p, int M, const void* phi, const void* zP, const void* zMs, uint64_t seed,
                        int64_t site_offset, double* out, cudaStream_t s) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  const T *dzP = (const T*)zP, *dzM = (const T*)zMs;
  const bool rng = (zMs == nullptr);
  int rc = refuse_anomalies(p);
  if (rc) return rc;
  // the caller's stream is ordered behind the plan's own work (batch registration, earlier calls) ...
  rc = fork_to(p, s);
  if (rc) return rc;
  if (rng) {
    const int64_t nzP = (int64_t)M * (cfg.nP > 0 ? cfg.nP : 1);
    rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)sizeof(T) * nzP);
    if (rc) return rc;
    Launch L{s, p->sm_count, &p->launches};
    CU(p, launch_gen_normal<T>(L, (T*)p->d_zP, cfg.nP, M, seed, (int64_t)1 << 62));
    dzP = (const T*)p->d_zP;
  }
  rc = enqueue_eval<T>(p, M, (const T*)phi, dzP, dzM, out, nullptr, s, rng, seed, site_offset);
  if (rc) return rc;
  // ... and later calls on the plan's stream are ordered behind this evaluation
  return mark_user(p, s);
}

extern "C" int hvi_neg_elbo_grad_async(hvi_plan* p, int32_t n_MC, const void* phi_dev, const void* zP_dev,
                                       const void* zMs_dev, uint64_t seed, int64_t site_offset, double* out_dev,
                                       void* cuda_stream) {
  int rc = check_eval_args(p, n_MC, phi_dev, HVI_MEM_DEVICE);
  if (rc) return rc;
  if (!out_dev) PLAN_FAIL(p, HVI_EINVAL, "out_dev is NULL");
  if (zMs_dev && p->desc.n_P > 0 && !zP_dev) PLAN_FAIL(p, HVI_EINVAL, "zP_dev is NULL");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  return p->dtype == HVI_F32 ? eval_async_t<float>(p, n_MC, phi_dev, zP_dev, zMs_dev, seed, site_offset, out_dev, s)
                             : eval_async_t<double>(p, n_MC, phi_dev, zP_dev, zMs_dev, seed, site_offset, out_dev, s);
}

template <typename T>
static int generate_zeta_t(hvi_plan* p, int M, const void* phi, const void* zP, const void* zMs, int mem_kind, void* zetaP,
                           void* zetaMs_tr, void* sigma) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  cudaStream_t s = p->stream;
  const int nphi = cfg.nphi, nP = cfg.nP, nM = cfg.nM;
  const int64_t nb = p->nb;
  const size_t nzP = (size_t)M * (nP > 0 ? nP : 1), nzM = (size_t)M * nM * nb;
  const T *dphi = (const T*)phi, *dzP = (const T*)zP, *dzM = (const T*)zMs;
  void *o1 = zetaP, *o2 = zetaMs_tr, *o3 = sigma, *tmp = nullptr;
  if (mem_kind == HVI_MEM_HOST) {
    CU(p, cudaMemcpyAsync(p->d_phi, phi, sizeof(T) * nphi, cudaMemcpyHostToDevice, s));
    int rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)(sizeof(T) * nzP));
    if (rc) return rc;
    rc = ensure(p, &p->d_zMs, &p->cap_zMs, (int64_t)(sizeof(T) * nzM));
    if (rc) return rc;
    if (nP > 0) CU(p, cudaMemcpyAsync(p->d_zP, zP, sizeof(T) * (size_t)M * nP, cudaMemcpyHostToDevice, s));
    CU(p, cudaMemcpyAsync(p->d_zMs, zMs, sizeof(T) * nzM, cudaMemcpyHostToDevice, s));
    dphi = (const T*)p->d_phi; dzP = (const T*)p->d_zP; dzM = (const T*)p->d_zMs;
    CU(p, cudaMalloc(&tmp, sizeof(T) * ((size_t)nP * M + nzM + nP + (size_t)nM * nb)));
    o1 = tmp; o2 = (T*)tmp + (size_t)nP * M; o3 = (T*)o2 + nzM;
  }
  int rc = ensure(p, &p->d_pdraw, &p->cap_pdraw, (int64_t)(sizeof(T) * pdraw_elems(nP, M)));
  if (rc) { if (tmp) cudaFree(tmp); return rc; }
  Launch L{s, p->sm_count, &p->launches};
  EvalConsts<T>* ec = (EvalConsts<T>*)p->d_ec;
  T* pdraw = (T*)p->d_pdraw;
  StreamArgs<T> a;
  memset(&a, 0, sizeof a);
  a.mu = (const T*)p->d_mu; a.zMs = dzM; a.ec = ec; a.nb = nb; a.M = M;
  a.phi = dphi; a.i_sites = p->d_isites;
  if (cfg.npc > 0) {
    rc = ensure(p, &p->d_MU, &p->cap_MU, (int64_t)sizeof(T) * M * nM * nb);
    if (rc) { if (tmp) cudaFree(tmp); return rc; }
    a.MU = (const T*)p->d_MU;
  }
  cudaError_t e = launch_prologue<T>(L, cfg, dphi, dzP, M, ec, pdraw);
  if (e == cudaSuccess && run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_mu, 0, nullptr)) e = cudaErrorUnknown;
  if (e == cudaSuccess && cfg.npc > 0 && run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_MU, M, pdraw + (size_t)nP * M)) e = cudaErrorUnknown;
  if (e == cudaSuccess) e = launch_generate_zeta<T>(L, cfg, a, pdraw, (T*)o1, (T*)o2, (T*)o3);
  if (e == cudaSuccess && mem_kind == HVI_MEM_HOST) {
    if (nP > 0) e = cudaMemcpyAsync(zetaP, o1, sizeof(T) * (size_t)nP * M, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(zetaMs_tr, o2, sizeof(T) * nzM, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sigma, o3, sizeof(T) * (nP + (size_t)nM * nb), cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) PLAN_FAIL(p, HVI_ECUDA, "generate_zeta: %s", cudaGetErrorString(e));
  return HVI_OK;
}

extern "C" int hvi_generate_zeta(hvi_plan* p, int32_t n_MC, const void* phi, const void* zP, const void* zMs, int mem_kind,
                                 void* zetaP, void* zetaMs_tr, void* sigma) {
  int rc = check_eval_args(p, n_MC, phi, mem_kind);
  if (rc) return rc;
  if (!zMs || !zetaMs_tr || !sigma || (p->desc.n_P > 0 && (!zP || !zetaP))) PLAN_FAIL(p, HVI_EINVAL, "NULL array argument");
  return p->dtype == HVI_F32 ? generate_zeta_t<float>(p, n_MC, phi, zP, zMs, mem_kind, zetaP, zetaMs_tr, sigma)
                             : generate_zeta_t<double>(p, n_MC, phi, zP, zMs, mem_kind, zetaP, zetaMs_tr, sigma);
}

template <typename T>
static int apply_g_t(hvi_plan* p, const void* phi, int mem_kind, void* mu) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  cudaStream_t s = p->stream;
  const T* dphi = (const T*)phi;
  if (mem_kind == HVI_MEM_HOST) {
    CU(p, cudaMemcpyAsync(p->d_phi, phi, sizeof(T) * cfg.nphi, cudaMemcpyHostToDevice, s));
    dphi = (const T*)p->d_phi;
  }
  Launch L{s, p->sm_count, &p->launches};
  int rcg = run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_mu, 0, nullptr);
  if (rcg) return rcg;
  CU(p, cudaMemcpyAsync(mu, p->d_mu, sizeof(T) * (size_t)cfg.nM * p->nb,
                        mem_kind == HVI_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  CU(p, cudaStreamSynchronize(s));
  return HVI_OK;
}

extern "C" int hvi_apply_g(hvi_plan* p, const void* phi, int mem_kind, void* mu_zeta) {
  int rc = check_eval_args(p, 1, phi, mem_kind);
  if (rc) return rc;
  if (!mu_zeta) PLAN_FAIL(p, HVI_EINVAL, "mu_zeta is NULL");
  return p->dtype == HVI_F32 ? apply_g_t<float>(p, phi, mem_kind, mu_zeta) : apply_g_t<double>(p, phi, mem_kind, mu_zeta);
}

template <typename T>
static int randn_t(hvi_plan* p, int M, int64_t n_cols, uint64_t seed, int64_t col_offset, int mem_kind, void* out) {
  cudaStream_t s = p->stream;
  Launch L{s, p->sm_count, &p->launches};
  const size_t bytes = sizeof(T) * (size_t)M * n_cols;
  void* tmp = nullptr;
  T* dst = (T*)out;
  if (mem_kind == HVI_MEM_HOST) { CU(p, cudaMalloc(&tmp, bytes)); dst = (T*)tmp; }
  cudaError_t e = launch_gen_normal<T>(L, dst, n_cols, M, seed, col_offset);
  if (e == cudaSuccess && tmp) e = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) PLAN_FAIL(p, HVI_ECUDA, "hvi_randn: %s", cudaGetErrorString(e));
  return HVI_OK;
}

extern "C" int hvi_randn(hvi_plan* p, int32_t n_MC, int64_t n_cols, uint64_t seed, int64_t col_offset, int mem_kind,
                         void* out) {
  if (!p) return HVI_EINVAL;
  if (n_MC < 1 || n_cols < 1 || !out) PLAN_FAIL(p, HVI_EINVAL, "hvi_randn: bad arguments");
  CU(p, cudaSetDevice(p->desc.device));
  return p->dtype == HVI_F32 ? randn_t<float>(p, n_MC, n_cols, seed, col_offset, mem_kind, out)
                             : randn_t<double>(p, n_MC, n_cols, seed, col_offset, mem_kind, out);
}

// predict_hvi (src/elbo.jl:320-352) = sample_posterior (:381-458) + the 3-D PBM applicator
template <typename T>
static int predict_t(hvi_plan* p, int M, const void* phi, const void* zP, const void* zMs, uint64_t seed,
                     int64_t site_offset, int mem_kind, void* thetaP, void* thetaMs_tr, void* y, double* entropy) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  cudaStream_t s = p->stream;
  const int nphi = cfg.nphi, nP = cfg.nP, nM = cfg.nM, nO = cfg.nO;
  const int64_t nb = p->nb;
  const bool rng = (zMs == nullptr);
  const size_t nzP = (size_t)M * (nP > 0 ? nP : 1), nzM = (size_t)M * nM * nb, ny = (size_t)nO * nb * M;
  const T *dphi = (const T*)phi, *dzP = (const T*)zP, *dzM = (const T*)zMs;
  Launch L{s, p->sm_count, &p->launches};
  int rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)(sizeof(T) * nzP));
  if (rc) return rc;
  if (mem_kind == HVI_MEM_HOST) {
    CU(p, cudaMemcpyAsync(p->d_phi, phi, sizeof(T) * nphi, cudaMemcpyHostToDevice, s));
    dphi = (const T*)p->d_phi;
  }
  if (rng) {
    CU(p, launch_gen_normal<T>(L, (T*)p->d_zP, nP, M, seed, (int64_t)1 << 62));
    dzP = (const T*)p->d_zP;
  } else if (mem_kind == HVI_MEM_HOST) {
    rc = ensure(p, &p->d_zMs, &p->cap_zMs, (int64_t)(sizeof(T) * nzM));
    if (rc) return rc;
    if (nP > 0) CU(p, cudaMemcpyAsync(p->d_zP, zP, sizeof(T) * (size_t)M * nP, cudaMemcpyHostToDevice, s));
    CU(p, cudaMemcpyAsync(p->d_zMs, zMs, sizeof(T) * nzM, cudaMemcpyHostToDevice, s));
    dzP = (const T*)p->d_zP; dzM = (const T*)p->d_zMs;
  }
  rc = ensure(p, &p->d_pdraw, &p->cap_pdraw, (int64_t)(sizeof(T) * pdraw_elems(nP, M)));
  if (rc) return rc;
  void *o1 = thetaP, *o2 = thetaMs_tr, *o3 = y, *tmp = nullptr;
  if (mem_kind == HVI_MEM_HOST) {
    // the three result arrays in one allocation, each starting on a 256-byte boundary (the prediction kernel
    // writes y with 256-bit stores)
    const size_t oM = (((size_t)nP * M + 63) / 64) * 64, oY = oM + ((nzM + 63) / 64) * 64;
    CU(p, cudaMalloc(&tmp, sizeof(T) * (oY + ny)));
    o1 = tmp; o2 = (T*)tmp + oM; o3 = (T*)tmp + oY;
  }
  const int64_t nth = nb > (int64_t)nP * M ? nb : (int64_t)nP * M;
  const size_t nls = (size_t)((nth + 255) / 256) * 8;
  rc = ensure(p, (void**)&p->d_ls, &p->cap_ls, (int64_t)(sizeof(double) * nls));
  if (rc) { if (tmp) cudaFree(tmp); return rc; }
  double* d_ls = p->d_ls;
  cudaError_t e = cudaSuccess;
  EvalConsts<T>* ec = (EvalConsts<T>*)p->d_ec;
  T* pdraw = (T*)p->d_pdraw;
  StreamArgs<T> a;
  memset(&a, 0, sizeof a);
  a.mu = (const T*)p->d_mu; a.zMs = dzM; a.ec = ec; a.nb = nb; a.M = M;
  a.rng = rng ? 1 : 0; a.seed = seed; a.col_offset = site_offset * nM;
  a.phi = dphi; a.i_sites = p->d_isites;
  if (e == cudaSuccess && cfg.npc > 0) {
    rc = ensure(p, &p->d_MU, &p->cap_MU, (int64_t)sizeof(T) * M * nM * nb);
    if (rc) { if (tmp) cudaFree(tmp); return rc; }
    a.MU = (const T*)p->d_MU;
  }
  int nparts = 0;
  if (e == cudaSuccess) e = launch_prologue<T>(L, cfg, dphi, dzP, M, ec, pdraw);
  if (e == cudaSuccess && run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_mu, 0, nullptr)) e = cudaErrorUnknown;
  if (e == cudaSuccess && cfg.npc > 0 && run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_MU, M, pdraw + (size_t)nP * M)) e = cudaErrorUnknown;
  if (e == cudaSuccess) e = launch_predict<T>(L, cfg, a, pdraw, (const T*)p->d_xP, (T*)o1, (T*)o2, (T*)o3, d_ls, &nparts);
  std::vector<double> hls(nparts > 0 ? nparts : 1);
  std::vector<T> hq(nphi);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hls.data(), d_ls, sizeof(double) * nparts, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hq.data(), dphi, sizeof(T) * nphi, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && mem_kind == HVI_MEM_HOST) {
    if (nP > 0) e = cudaMemcpyAsync(thetaP, o1, sizeof(T) * (size_t)nP * M, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(thetaMs_tr, o2, sizeof(T) * nzM, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y, o3, sizeof(T) * ny, cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) PLAN_FAIL(p, HVI_ECUDA, "hvi_predict: %s", cudaGetErrorString(e));
  if (entropy) {
    // entropy_MvNormal(length(σ), 2 Σ log σ)  (src/elbo.jl:449-450, src/logden_normal.jl:74)
    double logsig = 0;
    for (int i = 0; i < nparts; i++) logsig += hls[i];
    for (int j = 0; j < nP; j++) logsig += 0.5 * (double)hq[cfg.o_ls2P + j];
    *entropy = 0.5 * ((double)nP + (double)nM * (double)nb) * 2.8378770664093454836 + logsig;
  }
  return HVI_OK;
}

extern "C" int hvi_predict(hvi_plan* p, int32_t n_sample, const void* phi, const void* zP, const void* zMs, uint64_t seed,
                           int64_t site_offset, int mem_kind, void* thetaP, void* thetaMs_tr, void* y, double* entropy) {
  int rc = check_eval_args(p, n_sample, phi, mem_kind);
  if (rc) return rc;
  if (!thetaMs_tr || !y || (p->desc.n_P > 0 && !thetaP)) PLAN_FAIL(p, HVI_EINVAL, "hvi_predict: NULL output");
  if (p->gpbm) PLAN_FAIL(p, HVI_EINVAL, "hvi_predict: not available with a generic process model");
  if (zMs && p->desc.n_P > 0 && !zP) PLAN_FAIL(p, HVI_EINVAL, "hvi_predict: zP is NULL");
  return p->dtype == HVI_F32
             ? predict_t<float>(p, n_sample, phi, zP, zMs, seed, site_offset, mem_kind, thetaP, thetaMs_tr, y, entropy)
             : predict_t<double>(p, n_sample, phi, zP, zMs, seed, site_offset, mem_kind, thetaP, thetaMs_tr, y, entropy);
}

namespace hvi {
cudaError_t launch_dense_tc(cudaStream_t stream, const float* A, const float* W, const float* bias, float* C, int64_t Mc,
                            int N, int K, int act, const float* Hd);
}

// One dense layer on the tensor cores (tcgen05, 3xTF32): C (Mc x N, row-major) = act(A (Mc x K) W(N x K)^T + b).
// Host pointers; building block of the wide applicator, exported for its parity test.
extern "C" int hvi_dense_tc(int32_t device, int64_t Mc, int32_t N, int32_t K, const float* A, const float* W,
                            const float* bias, int32_t act, float* C) {
  if (!A || !W || !C || Mc < 1 || N < 1 || K < 1 || K % 32 != 0 || N % 128 != 0) { g_create_err = "hvi_dense_tc: bad arguments (K % 32 == 0, N % 128 == 0)"; return HVI_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_create_err = "hvi_dense_tc: no CUDA device (no CPU fallback)"; return HVI_ECUDA; }
  cudaSetDevice(device);
  float *dA = nullptr, *dW = nullptr, *db = nullptr, *dC = nullptr;
  cudaError_t e = cudaMalloc((void**)&dA, sizeof(float) * (size_t)Mc * K);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dW, sizeof(float) * (size_t)N * K);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dC, sizeof(float) * (size_t)Mc * N);
  if (e == cudaSuccess && bias) e = cudaMalloc((void**)&db, sizeof(float) * (size_t)N);
  if (e == cudaSuccess) e = cudaMemcpy(dA, A, sizeof(float) * (size_t)Mc * K, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dW, W, sizeof(float) * (size_t)N * K, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && bias) e = cudaMemcpy(db, bias, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_dense_tc(nullptr, dA, dW, db, dC, Mc, N, K, act, nullptr);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(C, dC, sizeof(float) * (size_t)Mc * N, cudaMemcpyDeviceToHost);
  for (float* q : {dA, dW, db, dC}) if (q) cudaFree(q);
  if (e != cudaSuccess) { g_create_err = std::string("hvi_dense_tc: ") + cudaGetErrorString(e); return HVI_ECUDA; }
  return HVI_OK;
}

// One GEMM of the TMA-fed 3xBF16 chain (hvi_dense_tma.cu) on host arrays; exported for its parity test.
//   mode 0: C = act(A·Bᵀ + bias), A (M x K), B (N x K);  mode 1: C = (A·Bᵀ) ⊙ act'(Hd);
//   mode 2: C = AᵀB through the split-K, MN-major path, A (K x M), B (K x N).  Hd / C (M x N), all row-major.
extern "C" int hvi_gemm_bf3(int32_t device, int32_t mode, int64_t M, int32_t N, int64_t K, const float* A, const float* B,
                            const float* bias, int32_t act, const float* Hd, float* C) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1 || N % 128 != 0 || mode < 0 || mode > 2 || (mode == 1 && !Hd) ||
      (mode == 2 ? M % 32 != 0 : K % 32 != 0)) {
    g_create_err = "hvi_gemm_bf3: bad arguments (N % 128 == 0; K % 32 == 0, or M % 32 == 0 in mode 2)";
    return HVI_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_create_err = "hvi_gemm_bf3: no CUDA device (no CPU fallback)"; return HVI_ECUDA; }
  cudaSetDevice(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  float *dA = nullptr, *dB = nullptr, *db = nullptr, *dH = nullptr, *dC = nullptr;
  cudaError_t e = cudaMalloc((void**)&dA, sizeof(float) * (size_t)M * K);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dB, sizeof(float) * (size_t)N * K);
  if (e == cudaSuccess) e = cudaMalloc((void**)&dC, sizeof(float) * (size_t)M * N);
  if (e == cudaSuccess && bias) e = cudaMalloc((void**)&db, sizeof(float) * (size_t)N);
  if (e == cudaSuccess && Hd) e = cudaMalloc((void**)&dH, sizeof(float) * (size_t)M * N);
  if (e == cudaSuccess) e = cudaMemcpy(dA, A, sizeof(float) * (size_t)M * K, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dB, B, sizeof(float) * (size_t)N * K, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && bias) e = cudaMemcpy(db, bias, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && Hd) e = cudaMemcpy(dH, Hd, sizeof(float) * (size_t)M * N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = gemm_bf3_device(nullptr, prop.multiProcessorCount, mode, M, N, K, dA, dB, db, act, dH, dC);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(C, dC, sizeof(float) * (size_t)M * N, cudaMemcpyDeviceToHost);
  for (float* q : {dA, dB, db, dH, dC}) if (q) cudaFree(q);
  if (e != cudaSuccess) { g_create_err = std::string("hvi_gemm_bf3: ") + cudaGetErrorString(e); return HVI_ECUDA; }
  return HVI_OK;
}


// ---------------------------------------------------------------------------------------
// Point-solver loss loss_gf (src/gf.jl:227-295) and its gradient w.r.t. ϕ = [ϕg; ϕP]
// ---------------------------------------------------------------------------------------
template <typename T>
static int loss_gf_t(hvi_plan* p, const void* phi_gf, int mem_kind, double comps[3], void* grad) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  if (p->wide) PLAN_FAIL(p, HVI_EINVAL, "loss_gf: not available with the wide applicator");
  if (p->gpbm) PLAN_FAIL(p, HVI_EINVAL, "loss_gf: not available with a generic process model");
  if (!p->has_obs) PLAN_FAIL(p, HVI_ENODATA, "loss_gf: the registered batch has no observations");
  cudaStream_t s = p->stream;
  int rcj = join_user(p);
  if (rcj) return rcj;
  const int ng = cfg.nphig, nP = cfg.nP;
  const cudaMemcpyKind kind = mem_kind == HVI_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  // full-layout ϕ: ϕg in place, ϕP in the μP slot (what g's pbm covariates and θP = transP(ϕP) read), the rest zero
  T* dphi = (T*)p->d_phi;
  CU(p, cudaMemsetAsync(dphi, 0, sizeof(T) * (size_t)cfg.nphi, s));
  CU(p, cudaMemcpyAsync(dphi, phi_gf, sizeof(T) * (size_t)ng, kind, s));
  if (nP > 0) CU(p, cudaMemcpyAsync(dphi + cfg.o_muP, (const T*)phi_gf + ng, sizeof(T) * (size_t)nP, kind, s));
  Launch L{s, p->sm_count, &p->launches};
  int rc = run_g_fwd<T>(p, L, cfg, dphi, (T*)p->d_mu, 0, nullptr);
  if (rc) return rc;
  if ((size_t)point_max_rows(p->sm_count) * point_row_len() > (size_t)p->max_rows * nacc(p->cf.nP, p->cf.nM))
    PLAN_FAIL(p, HVI_EINVAL, "internal: scratch too small");
  int nrows = 0, nblk = 0;
  CU(p, launch_point<T>(L, cfg, dphi, (const T*)p->d_rec, (const T*)p->d_mu, (T*)p->d_mubar, p->nb, p->d_part_stream, &nrows));
  double* part_xs = nullptr;
  if (cfg.npc > 0) {
    rc = ensure(p, (void**)&p->d_part_x, &p->cap_part_x, (int64_t)sizeof(double) * 2 * p->max_blk * cfg.npc);
    if (rc) return rc;
    part_xs = p->d_part_x;
  }
  CU(p, launch_mlp_bwd<T>(L, cfg, dphi, (const T*)p->d_xM, (const T*)p->d_mubar, p->nb, p->d_part_mlp, &nblk, p->max_blk, 0,
                          nullptr, part_xs));
  const bool want_grad = grad != nullptr;
  CU(p, launch_point_finalize<T>(L, cfg, dphi, p->d_part_stream, nrows, p->d_part_mlp, nblk, part_xs, p->d_bconst, p->d_out,
                                 want_grad ? (T*)p->d_grad : nullptr));
  CU(p, cudaMemcpyAsync(p->h_out, p->d_out + ng + nP, sizeof(double) * 4, cudaMemcpyDeviceToHost, s));
  if (p->anomalies < 0) CU(p, cudaMemcpyAsync(p->h_bconst, p->d_bconst, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
  if (want_grad) {
    if (mem_kind == HVI_MEM_HOST) CU(p, cudaMemcpyAsync(p->h_grad, p->d_grad, sizeof(T) * (size_t)(ng + nP), cudaMemcpyDeviceToHost, s));
    else CU(p, cudaMemcpyAsync(grad, p->d_grad, sizeof(T) * (size_t)(ng + nP), cudaMemcpyDeviceToDevice, s));
  }
  CU(p, cudaStreamSynchronize(s));
  if (p->anomalies < 0) p->anomalies = (int64_t)p->h_bconst[1];
  if (want_grad && mem_kind == HVI_MEM_HOST) memcpy(grad, p->h_grad, sizeof(T) * (size_t)(ng + nP));
  if (comps) for (int i = 0; i < 3; i++) comps[i] = p->h_out[i];
  if (p->anomalies > 0)
    PLAN_FAIL(p, HVI_ENONFINITE, "%lld observation(s) are finite where a driver is not: the reference returns NaN here",
              (long long)p->anomalies);
  if (p->h_out[3] != 0.0) PLAN_FAIL(p, HVI_ENONFINITE, "non-finite value or gradient (%g entries)", p->h_out[3]);
  return HVI_OK;
}

extern "C" int hvi_loss_gf_grad(hvi_plan* p, const void* phi_gf, int mem_kind, double comps[3], void* grad) {
  if (!p) return HVI_EINVAL;
  if (p->nb <= 0) PLAN_FAIL(p, HVI_ENODATA, "hvi_plan_set_data has not been called");
  if (!phi_gf) PLAN_FAIL(p, HVI_EINVAL, "phi_gf is NULL");
  if (mem_kind != HVI_MEM_HOST && mem_kind != HVI_MEM_DEVICE) PLAN_FAIL(p, HVI_EINVAL, "bad mem_kind");
  CU(p, cudaSetDevice(p->desc.device));
  return p->dtype == HVI_F32 ? loss_gf_t<float>(p, phi_gf, mem_kind, comps, grad) : loss_gf_t<double>(p, phi_gf, mem_kind, comps, grad);
}

// ---------------------------------------------------------------------------------------
// Device-resident optimiser steps: Optimization.solve(optprob, Adam(η); …) of
// src/HybridSolver.jl:259-260 without the host round trip per iteration.  The update is the
// Optimisers.jl Adam rule: m ← β1 m + (1−β1) g, v ← β2 v + (1−β2) g², ϕ ← ϕ − η·m̂/(√v̂ + ε).
// An iteration whose value or gradient is not finite leaves ϕ, m, v untouched (the reference
// raises there) and is visible as a non-finite entry of the loss trace.
// ---------------------------------------------------------------------------------------
// Device-side counters (plan->d_ctr) make an iteration independent of host state, so that it can be captured once as
// a CUDA graph and replayed: [0] Adam steps applied so far (advanced only when an update lands: the bias correction
// never counts a skipped iteration), [1] iteration index (trace slot, minibatch choice), [2] seed of the iteration's
// draws, [3] block counter.  All blocks read the counters before the last block to finish advances them.
template <typename T>
__global__ void __launch_bounds__(256) k_adam(int64_t n, const double* __restrict__ out, T* __restrict__ phi,
                                              double* __restrict__ m, double* __restrict__ v, double eta, double b1,
                                              double b2, double eps, double* __restrict__ trace,
                                              unsigned long long* __restrict__ ctr) {
  const unsigned long long t_applied = ctr[0], it = ctr[1];
  const bool bad = out[n + 7] != 0.0 || !isfinite(out[n + 6]);
  if (!bad) {
    // β^t by repeated squaring, as the Adam phase of k_small_iter: ≈ 2 log2(t) multiplications instead of two pow()
    double p1 = 1.0, p2 = 1.0, q1 = b1, q2 = b2;
    _Pragma("unroll 1") for (unsigned long long e = t_applied + 1; e > 0; e >>= 1) {
      if (e & 1) { p1 *= q1; p2 *= q2; }
      q1 *= q1; q2 *= q2;
    }
    const double c1 = 1.0 / (1.0 - p1), c2 = 1.0 / (1.0 - p2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double g = out[i];
      const double mi = b1 * m[i] + (1.0 - b1) * g, vi = b2 * v[i] + (1.0 - b2) * g * g;
      m[i] = mi; v[i] = vi;
      phi[i] = (T)((double)phi[i] - eta * (mi * c1) / (sqrt(vi * c2) + eps));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0 && trace) trace[it] = bad ? nan("") : out[n + 6];
    __threadfence();
    if (atomicAdd(&ctr[3], 1ULL) == (unsigned long long)gridDim.x - 1) {
      ctr[3] = 0;
      if (!bad) ctr[0] = t_applied + 1;
      ctr[1] = it + 1;
      ctr[2] += 1;
    }
  }
}

__global__ void k_set_ctr(unsigned long long* ctr, unsigned long long it0, unsigned long long seed0, int reset_steps) {
  if (reset_steps) ctr[0] = 0;
  ctr[1] = it0; ctr[2] = seed0; ctr[3] = 0;
}

template <typename T>
static int adam_apply_t(hvi_plan* p, const double* out_dev, cudaStream_t s, bool trace) {
  const int64_t n = cfg_of<T>(p).nphi;
  int blocks = (int)((n + 255) / 256);
  if (blocks > p->sm_count * 8) blocks = p->sm_count * 8;
  k_adam<T><<<blocks, 256, 0, s>>>(n, out_dev, (T*)p->d_phi_opt, p->d_adam_m, p->d_adam_v, p->adam_eta, p->adam_b1, p->adam_b2,
                                  p->adam_eps, trace ? p->d_trace : nullptr, p->d_ctr);
  CU(p, cudaGetLastError());
  ++p->launches;
  return HVI_OK;
}

extern "C" int hvi_adam_init(hvi_plan* p, const void* phi0, double eta, double beta1, double beta2, double eps) {
  if (!p) return HVI_EINVAL;
  if (!phi0 || !(eta > 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(eps > 0))
    PLAN_FAIL(p, HVI_EINVAL, "adam_init: need phi0, eta > 0, 0 <= beta < 1, eps > 0");
  CU(p, cudaSetDevice(p->desc.device));
  int rcj = join_user(p);
  if (rcj) return rcj;
  const size_t n = (size_t)p->cf.nphi;
  if (!p->d_phi_opt) {
    CU(p, cudaMalloc(&p->d_phi_opt, p->es * n));
    CU(p, cudaMalloc((void**)&p->d_adam_m, sizeof(double) * n));
    CU(p, cudaMalloc((void**)&p->d_adam_v, sizeof(double) * n));
  }
  CU(p, cudaMemcpyAsync(p->d_phi_opt, phi0, p->es * n, cudaMemcpyHostToDevice, p->stream));
  CU(p, cudaMemsetAsync(p->d_adam_m, 0, sizeof(double) * n, p->stream));
  CU(p, cudaMemsetAsync(p->d_adam_v, 0, sizeof(double) * n, p->stream));
  CU(p, cudaMemsetAsync(p->d_ctr, 0, sizeof(unsigned long long) * 4, p->stream));
  CU(p, cudaStreamSynchronize(p->stream));
  p->adam_eta = eta; p->adam_b1 = beta1; p->adam_b2 = beta2; p->adam_eps = eps;
  p->graph_valid = 0;   // the hyper-parameters are baked into a captured iteration
  return HVI_OK;
}

// one optimiser iteration on `s`: [minibatch selection] → draws of the global block → evaluation → Adam update.
// Everything that varies between iterations (seed, minibatch, step count, trace slot) is read from p->d_ctr.
template <typename T>
static int adam_iteration(hvi_plan* p, int M, int64_t site_offset, int64_t batch_size, int64_t first_batch, cudaStream_t s,
                          bool looped = false) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  Launch L{s, p->sm_count, &p->launches};
  if (fused_ok<T>(p, batch_size > 0 ? batch_size : p->nb, M, looped)) {   // the reference's operating point: one launch per iteration
    if (batch_size > 0) {
      int rc = select_prepare<T>(p, batch_size);
      if (rc) return rc;
      p->anomalies = -1;
    }
    return enqueue_fused<T>(p, M, (T*)p->d_phi_opt, p->d_out, nullptr, s, 0, site_offset, true, batch_size, first_batch);
  }
  if (batch_size > 0) {
    const int64_t n_batches = p->ds_n / batch_size;   // partial = false (src/AbstractHybridProblem.jl:206-207)
    int rc = select_t<T>(p, nullptr, first_batch, batch_size, p->d_ctr, n_batches);
    if (rc) return rc;
  }
  CU(p, launch_gen_normal<T>(L, (T*)p->d_zP, cfg.nP, M, 0, (int64_t)1 << 62, p->d_ctr + 2));
  int rc = enqueue_eval<T>(p, M, (const T*)p->d_phi_opt, (const T*)p->d_zP, nullptr, p->d_out, nullptr, s, true, 0, site_offset,
                           p->d_ctr + 2);
  if (rc) return rc;
  return adam_apply_t<T>(p, p->d_out, s, true);
}

// every device pointer and size an iteration bakes into its kernel arguments; a captured graph is replayed only while
// this is unchanged
static void graph_key(const hvi_plan* p, int M, int64_t site_offset, int64_t batch_size, int64_t first_batch, GraphKey* k) {
  memset(k, 0, sizeof *k);
  const void* ptrs[] = {p->d_xM, p->d_xP, p->d_rec, p->d_mu, p->d_mubar, p->d_isites, p->d_phi_opt, p->d_adam_m, p->d_adam_v,
                        p->d_trace, p->d_zP, p->d_pdraw, p->d_ec, p->d_part_stream, p->d_part_mlp, p->d_out, p->d_red, p->d_MU,
                        p->d_MUBAR, p->d_part_x, p->d_xred, p->ds_xM, p->ds_xP, p->ds_yo, p->ds_yu, p->d_wide, p->d_wide_bwd,
                        p->d_gacc, p->d_wide_cache, (const void*)p->stream};
  static_assert(sizeof(ptrs) / sizeof(ptrs[0]) <= sizeof(k->ptrs) / sizeof(k->ptrs[0]), "GraphKey too small");
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) k->ptrs[i] = ptrs[i];
  k->v[0] = p->nb; k->v[1] = M; k->v[2] = site_offset; k->v[3] = batch_size; k->v[4] = first_batch; k->v[5] = p->ds_n;
  k->v[6] = p->has_obs; k->v[7] = p->use_fused;
}

template <typename T>
static int adam_run_t(hvi_plan* p, int n_iter, int M, uint64_t seed0, int64_t site_offset, int64_t batch_size,
                      int64_t first_batch, double* trace_host) {
  const Cfg<T>& cfg = cfg_of<T>(p);
  cudaStream_t s = p->stream;
  int rc = join_user(p);
  if (rc) return rc;
  if (p->cap_trace < (int64_t)sizeof(double) * n_iter) {
    CU(p, cudaStreamSynchronize(s));
    rc = ensure(p, (void**)&p->d_trace, &p->cap_trace, (int64_t)sizeof(double) * n_iter);
    if (rc) return rc;
  }
  const int64_t nzP = (int64_t)M * (cfg.nP > 0 ? cfg.nP : 1);
  rc = ensure(p, &p->d_zP, &p->cap_zP, (int64_t)sizeof(T) * nzP);
  if (rc) return rc;
  k_set_ctr<<<1, 1, 0, s>>>(p->d_ctr, 0ULL, (unsigned long long)seed0, 0);
  CU(p, cudaGetLastError());
  ++p->launches;
  // the first iteration runs eagerly (it sizes every scratch buffer); the rest replay one captured iteration.  The wide
  // applicator issues thousands of launches per evaluation through its own chunk loop and is not captured.
  int it = 0;
  const bool looped = n_iter > 1 && !p->profiling && !getenv("HVI_SMALL_NO_LOOP") &&
                      fused_ok<T>(p, batch_size > 0 ? batch_size : p->nb, M, true);
  rc = adam_iteration<T>(p, M, site_offset, batch_size, first_batch, s, looped);
  if (rc) return rc;
  it = 1;
  // single-launch iterations: the remaining ones as a loop inside ONE launch (chunks of 4096 iterations)
  if (looped) {
    while (it < n_iter) {
      const int n = n_iter - it < 4096 ? n_iter - it : 4096;
      rc = enqueue_fused<T>(p, M, (T*)p->d_phi_opt, p->d_out, nullptr, s, 0, site_offset, true, batch_size, first_batch, n);
      if (rc) return rc;
      it += n;
    }
  }
  const bool use_graph = !p->wide && !p->profiling && n_iter > 1 && p->use_graphs && it < n_iter;
  if (use_graph) {
    GraphKey key;
    graph_key(p, M, site_offset, batch_size, first_batch, &key);
    if (!p->graph_valid || memcmp(&key, &p->graph_key, sizeof key) != 0) {
      if (p->adam_graph) { cudaGraphExecDestroy(p->adam_graph); p->adam_graph = nullptr; }
      p->graph_valid = 0;
      cudaGraph_t g = nullptr;
      const int64_t launches_before = p->launches;
      CU(p, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      rc = adam_iteration<T>(p, M, site_offset, batch_size, first_batch, s);
      cudaError_t e = cudaStreamEndCapture(s, &g);
      p->launches = launches_before;   // nothing ran during the capture
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) PLAN_FAIL(p, HVI_ECUDA, "adam_run: graph capture failed: %s", cudaGetErrorString(e));
      size_t nn = 0;
      cudaGraphGetNodes(g, nullptr, &nn);
      p->graph_nodes = (int)nn;
      e = cudaGraphInstantiate(&p->adam_graph, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) PLAN_FAIL(p, HVI_ECUDA, "adam_run: graph instantiation failed: %s", cudaGetErrorString(e));
      graph_key(p, M, site_offset, batch_size, first_batch, &p->graph_key);
      p->graph_valid = 1;
    }
    for (; it < n_iter; it++) {
      CU(p, cudaGraphLaunch(p->adam_graph, s));
      p->launches += p->graph_nodes;
    }
  } else {
    for (; it < n_iter; it++) {
      rc = adam_iteration<T>(p, M, site_offset, batch_size, first_batch, s);
      if (rc) return rc;
    }
  }
  if (trace_host) {
    CU(p, cudaMemcpyAsync(trace_host, p->d_trace, sizeof(double) * n_iter, cudaMemcpyDeviceToHost, s));
    CU(p, cudaStreamSynchronize(s));
  }
  return HVI_OK;
}

extern "C" int hvi_adam_run(hvi_plan* p, int32_t n_iter, int32_t n_MC, uint64_t seed0, int64_t site_offset,
                            double* nL_trace) {
  if (!p) return HVI_EINVAL;
  if (!p->d_phi_opt) PLAN_FAIL(p, HVI_EINVAL, "adam_run: call hvi_adam_init first");
  if (p->nb <= 0) PLAN_FAIL(p, HVI_ENODATA, "hvi_plan_set_data has not been called");
  if (n_iter < 1 || n_MC < 1) PLAN_FAIL(p, HVI_EINVAL, "adam_run: n_iter and n_MC must be >= 1");
  CU(p, cudaSetDevice(p->desc.device));
  int rc = refuse_anomalies(p);
  if (rc) return rc;
  return p->dtype == HVI_F32 ? adam_run_t<float>(p, n_iter, n_MC, seed0, site_offset, 0, 0, nL_trace)
                             : adam_run_t<double>(p, n_iter, n_MC, seed0, site_offset, 0, 0, nL_trace);
}

// The minibatch loop of solve (src/HybridSolver.jl:259-260 over the DataLoader of src/AbstractHybridProblem.jl:206-207:
// consecutive column ranges of the training set, last partial batch dropped, no shuffling) on the device-resident
// dataset: iteration i evaluates minibatch (first_batch + i) mod n_batches and applies the Adam update.
extern "C" int hvi_adam_run_minibatches(hvi_plan* p, int32_t n_iter, int32_t n_MC, uint64_t seed0, int64_t batch_size,
                                        int64_t first_batch, double* nL_trace) {
  if (!p) return HVI_EINVAL;
  if (!p->d_phi_opt) PLAN_FAIL(p, HVI_EINVAL, "adam_run_minibatches: call hvi_adam_init first");
  if (p->ds_n <= 0) PLAN_FAIL(p, HVI_ENODATA, "adam_run_minibatches: hvi_plan_set_dataset has not been called");
  if (n_iter < 1 || n_MC < 1) PLAN_FAIL(p, HVI_EINVAL, "adam_run_minibatches: n_iter and n_MC must be >= 1");
  if (batch_size < 1 || batch_size > p->ds_n) PLAN_FAIL(p, HVI_EINVAL, "adam_run_minibatches: batch_size must be in 1 .. n_site_total");
  if (first_batch < 0 || first_batch >= p->ds_n / batch_size) PLAN_FAIL(p, HVI_EINVAL, "adam_run_minibatches: first_batch out of range");
  if (!p->ds_has_obs) PLAN_FAIL(p, HVI_ENODATA, "adam_run_minibatches: the dataset has no observations");
  CU(p, cudaSetDevice(p->desc.device));
  return p->dtype == HVI_F32 ? adam_run_t<float>(p, n_iter, n_MC, seed0, 0, batch_size, first_batch, nL_trace)
                             : adam_run_t<double>(p, n_iter, n_MC, seed0, 0, batch_size, first_batch, nL_trace);
}

extern "C" int hvi_plan_use_fused(hvi_plan* p, int on) {
  if (!p) return HVI_EINVAL;
  p->use_fused = on < 0 ? 0 : (on > 2 ? 2 : on);
  p->graph_valid = 0;
  return HVI_OK;
}

// Profiling aid of the single-launch path: from the next fused launch on, thread 0 stamps clock64() at the phase
// boundaries; clocks[0..10] = start, after gather, zP draws, prologue, g forward, streaming phase, g backward, reduce,
// finalize, Adam (SM clock ticks).  The reference has no tracing.
extern "C" int hvi_plan_fused_phase_clocks(hvi_plan* p, int64_t clocks[10]) {
  if (!p) return HVI_EINVAL;
  CU(p, cudaSetDevice(p->desc.device));
  if (!p->d_clocks) {
    CU(p, cudaMalloc((void**)&p->d_clocks, sizeof(long long) * 16));
    CU(p, cudaMemset(p->d_clocks, 0, sizeof(long long) * 16));
    p->graph_valid = 0;
  }
  if (clocks) {
    CU(p, cudaStreamSynchronize(p->stream));
    long long h[16];
    CU(p, cudaMemcpy(h, p->d_clocks, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 10; i++) clocks[i] = (int64_t)h[i];
  }
  return HVI_OK;
}

extern "C" int hvi_plan_use_graphs(hvi_plan* p, int on) {
  if (!p) return HVI_EINVAL;
  p->use_graphs = on ? 1 : 0;
  return HVI_OK;
}

extern "C" int hvi_adam_apply_dev(hvi_plan* p, const double* out_dev, void* cuda_stream) {
  if (!p) return HVI_EINVAL;
  if (!p->d_phi_opt) PLAN_FAIL(p, HVI_EINVAL, "adam_apply_dev: call hvi_adam_init first");
  if (!out_dev) PLAN_FAIL(p, HVI_EINVAL, "adam_apply_dev: out_dev is NULL");
  CU(p, cudaSetDevice(p->desc.device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int rc = fork_to(p, s);
  if (rc) return rc;
  rc = p->dtype == HVI_F32 ? adam_apply_t<float>(p, out_dev, s, false) : adam_apply_t<double>(p, out_dev, s, false);
  if (rc) return rc;
  return mark_user(p, s);
}

extern "C" int hvi_adam_phi_dev(hvi_plan* p, void** phi_dev) {
  if (!p || !phi_dev) return HVI_EINVAL;
  if (!p->d_phi_opt) PLAN_FAIL(p, HVI_EINVAL, "adam_phi_dev: call hvi_adam_init first");
  *phi_dev = p->d_phi_opt;
  return HVI_OK;
}

extern "C" int hvi_adam_get_phi(hvi_plan* p, void* phi_host) {
  if (!p || !phi_host) return HVI_EINVAL;
  if (!p->d_phi_opt) PLAN_FAIL(p, HVI_EINVAL, "adam_get_phi: call hvi_adam_init first");
  CU(p, cudaSetDevice(p->desc.device));
  int rcj = join_user(p);   // an update applied on a caller's stream (hvi_adam_apply_dev) lands first
  if (rcj) return rcj;
  CU(p, cudaMemcpyAsync(phi_host, p->d_phi_opt, p->es * (size_t)p->cf.nphi, cudaMemcpyDeviceToHost, p->stream));
  CU(p, cudaStreamSynchronize(p->stream));
  return HVI_OK;
}
