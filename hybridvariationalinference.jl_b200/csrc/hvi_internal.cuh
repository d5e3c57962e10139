// hvi_internal.cuh -- shared definitions of the B200 (sm_100a) neg-ELBO kernels.
// Reference citations (file:line) are relative to /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/hvi_b200.h"

namespace hvi {

constexpr int MAXP = HVI_MAX_PAR;
constexpr int MAXL = HVI_MAX_LAYERS;

// Static description of the problem in the working type T, passed to kernels by value
// (__grid_constant__), i.e. through the constant bank.
template <typename T>
struct Cfg {
  int nP, nM, nO, nC;
  int transP[MAXP], transM[MAXP];
  int prP_kind[MAXP], prM_kind[MAXP];
  T prP_a[MAXP], prP_ib[MAXP];  // prior location, 1/scale
  T prM_a[MAXP], prM_ib[MAXP];
  T prM_nab[MAXP];                    // −a/b, so that u = fma(x, 1/b, −a/b)
  double prP_c0[MAXP], prM_c0[MAXP];  // log(scale) + log(2π)/2 (0 if no prior)
  int slot_code[4];                   // -1: fixed; j<nP: θP[j]; nP+k: θM[k]
  T slot_fix[4];
  int omit_priors, count_globals;
  T clamp_lo, clamp_hi;               // sqrt(floatmin), sqrt(floatmax) (src/elbo.jl:797-798)
  T ln_clamp_lo, ln_clamp_hi;         // their logarithms
  int blkP_start[MAXP], rhoP_off[MAXP], blkM_start[MAXP], rhoM_off[MAXP];
  int o_ls2P, o_coef, o_rhoP, o_rhoM, o_muP, nphig, nphi;
  int nL, l_in[MAXL], l_out[MAXL], l_act[MAXL], l_bias[MAXL], woff[MAXL], boff[MAXL];
  int sum_width;   // Σ_l n_in(l) + n_out(last): rows of the activation stack
  int sum_out;     // Σ_l n_out(l): rows of the delta stack
  int max_width;
  int scale_kind;
  T scale_a[MAXP], scale_b[MAXP];
  int npc, pbm_covar_idx[MAXP];
  // MeanVarSepHVIApproximation (src/elbo_sepvec.jl): o_coef is then the offset of logσ2_ζMs (nM × n_site_total)
  int sepvar, n_site_total;
};

// Per-evaluation constants derived from ϕq by the prologue kernel.
template <typename T>
struct EvalConsts {
  T UM[MAXP][MAXP], UP[MAXP][MAXP];  // column-normalised unit-upper factors (src/cholesky.jl:156-174)
  T nrmM[MAXP], nrmP[MAXP];
  T coefA[MAXP], coefB[MAXP];        // coef_logσ2_ζMs[1,k], [2,k] (src/elbo.jl:652-653)
  T sigP[MAXP], muP[MAXP];
};

// number of per-warp accumulators of the streaming kernel
__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__host__ __device__ constexpr int nacc(int nP, int nM) { return 4 + 2 * nM + tri(nM) + nP + tri(nP); }
// accumulator slots
enum { ACC_NLY = 0, ACC_PRIOR = 1, ACC_LJ = 2, ACC_LOGSIG = 3, ACC_FIRST_VEC = 4 };

template <typename T>
struct StreamArgs {
  const T* rec;      // tiled site records [tile][4*nO][32]: S1, S2, y_o, w = mask·exp(−y_unc)
  const T* mu;       // μ_ζ (nM × nb) column-major
  const T* zMs;      // (M × nM·nb) column-major
  const T* pd;       // per-draw records of the global block [M][pd_stride(nP)]: θP, dθP/dζP, zP
  const EvalConsts<T>* ec;
  T* mubar;          // cotangent of μ_ζ (nM × nb)
  const T* MU;       // pbm covariates only: per-draw means [M][nM][nb] (NULL otherwise)
  T* MUBAR;          // and their cotangents, scaled by 1/n_MC
  double* partials;  // [gridDim.x * warps][nacc]
  int64_t nb;
  int M;
  int pd_in_smem;    // 1: the kernel copies pd into shared memory (it fits)
  const T* phi;      // MeanVarSep: ϕ (per-site log-variances are read from it) …
  const int64_t* i_sites;   // … at the batch's global site indices (NULL: 0..nb-1)
  double* out;       // … and their gradient goes straight into the result vector
  T* grad_T;
  int rng;           // 1: draw ξ inside the kernel from (seed, col_offset + site·n_θM + k, draw); zMs unused
  uint64_t seed;
  const unsigned long long* seed_dev;   // non-NULL: the seed is read from device memory (optimiser loops replayed as CUDA graphs)
  int64_t col_offset;
  int staged;        // 1: M % (16/sizeof(T)) == 0 and zMs 16-byte aligned → cp.async staging
};

// arguments of the single-launch small-batch iteration (k_small_iter)
template <typename T>
struct SmallArgs {
  StreamArgs<T> st;     // as for k_stream: rec, mu, pd, ec, mubar, partials, nb, M, seed, col_offset
  T* phi;               // ϕ on the device (updated in place by the Adam phase)
  T* zP;                // [n_P][M] draws of the global block
  T* pdraw;
  const T* xM;
  double *part_mlp, *red, *out;
  T* grad_T;
  double* bconst;       // batch constants {½Σy_unc, anomalies}: read, and written by the gather phase
  // minibatch of the device-resident dataset (ds_xM != NULL)
  const T *ds_xM, *ds_xP, *ds_yo, *ds_yu;
  T *xM_b, *xP_b, *rec_w;
  int64_t first, n_batches;
  int64_t* isites_out;
  // Adam phase (adam_m != NULL) and the device-side iteration counters (k_adam)
  double *adam_m, *adam_v, *trace;
  double eta, b1, b2, eps;
  unsigned long long* ctr;
  long long* clocks;    // optional: clock64() at the phase boundaries (profiling aid), NULL otherwise
  int n_inner;          // optimiser iterations run by ONE launch (the loop is inside the kernel); 0 or 1: one
  int units_mode;       // > 0: shared-memory bytes the (site, draw)-parallel streaming phase may use; 0: thread per site
};

struct Launch {
  cudaStream_t stream;
  int sm_count;
  int64_t* counter;
};

// activations of the wide applicator kept from the forward pass for the backward pass of the same evaluation:
// slot (slot0 + m·n_chunks + chunk) holds every layer output of one chunk of columns; chunks beyond n_slots
// are recomputed in the backward pass
struct WideCache {
  void* base;
  int64_t n_slots;
  int64_t slot0;
  int64_t rows;   // columns per slot; 0: WIDE_CHUNK (site-blocked evaluations use slots of one block's chunk)
};

// ---- kernel launchers (hvi_kernels.cu) -------------------------------------------------
template <typename T>
cudaError_t launch_preprocess(const Launch& L, int nO, int64_t nb, const T* xP, const T* y_o, const T* y_unc, T* rec,
                              double* part /*[2*nwarps]*/, int* nrows_out);
// select a minibatch of a device-resident dataset (hvi_plan_select_batch / _range): gather the sites idx[0..nb) (NULL:
// first .. first+nb) of the raw dataset arrays into the batch arrays xM_b, xP_b and the tiled records; same partial sums.
// ctr non-NULL: `first` counts minibatches and the device-side iteration counter ctr[1] advances it (modulo n_batches).
// isites_out non-NULL: the dataset index of every batch site is written there (MeanVarSep)
template <typename T>
cudaError_t launch_gather_preprocess(const Launch& L, int nC, int nO, int64_t nb, const int64_t* idx, int64_t first,
                                     const T* ds_xM, const T* ds_xP, const T* ds_yo, const T* ds_yu, T* xM_b, T* xP_b,
                                     T* rec, double* part /*[2*nwarps]*/, int* nrows_out,
                                     const unsigned long long* ctr = nullptr, int64_t n_batches = 1,
                                     int64_t* isites_out = nullptr);
// consts[0] = Σ part[2r] (the ϕ-independent ½Σy_unc of the batch), consts[1] = Σ part[2r+1] (anomalies); fixed order
cudaError_t launch_batch_consts(const Launch& L, const double* part, int nrows, double* consts);
template <typename T>
cudaError_t launch_prologue(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* zP, int M, EvalConsts<T>* ec,
                            T* pdraw /*4 arrays [nP][M]: rP, zetaP, thP, gP; then [M][pd_stride] records*/);
template <typename T>
cudaError_t launch_mlp_fwd(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, T* mu,
                           int M_units = 0, const T* zetaP = nullptr);
template <typename T>
cudaError_t launch_mlp_bwd(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, const T* mubar, int64_t nb,
                           double* part /*[nblk][nphig]*/, int* nblk_out, int nblk_max, int M_units = 0,
                           const T* zetaP = nullptr, double* part_x = nullptr /*[nblk][max(M_units,1)*npc]*/,
                           const T* mu_fwd = nullptr /*μ_ζ written by launch_mlp_fwd for the same columns, if still valid*/);
template <typename T>
cudaError_t launch_stream(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, int* nrows_out, int nrows_max);
template <typename T>
cudaError_t launch_finalize(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* zP, const T* pdraw,
                            const EvalConsts<T>* ec, const double* part_stream, int nrows, const double* part_mlp,
                            int nblk, const double* batch_consts /*device: ½Σy_unc, anomalies*/, int64_t nb, int M,
                            double* out, T* grad_T,
                            double* red /*[128 + ceil(nphig/32) + 1] scratch*/, const double* part_xs = nullptr,
                            int nblk_xs = 0, const double* part_xu = nullptr, int nblk_xu = 0,
                            double* xred = nullptr /*[(1 + M)·npc] scratch*/);
template <typename T>
cudaError_t launch_gen_normal(const Launch& L, T* z, int64_t n_cols, int M, uint64_t seed, int64_t col_offset,
                              const unsigned long long* seed_dev = nullptr);
template <typename T>
cudaError_t launch_generate_zeta(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, const T* pdraw, T* zetaP,
                                 T* zetaMs_tr, T* sigma);
__host__ __device__ constexpr int pd_stride(int nP) { return ((3 * nP + 3) / 4) * 4; }
inline size_t pdraw_elems(int nP, int M) { return (size_t)4 * (nP > 0 ? nP : 1) * M + (size_t)M * pd_stride(nP) + 4; }
template <typename T>
cudaError_t launch_predict(const Launch& L, const Cfg<T>& cfg, const StreamArgs<T>& a, const T* pdraw, const T* xP,
                           T* thetaP, T* thetaMs_tr, T* y, double* part_logsig /*[sm_count*16*8]*/, int* nparts);
// wide applicator (any layer wider than HVI_MAX_WIDTH): forward only, tensor cores for the hidden layers
template <typename T>
cudaError_t launch_mlp_fwd_wide(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, T* mu,
                                int M_units, const T* zetaP, T* work, WideCache cache = WideCache{nullptr, 0, 0});
size_t wide_work_elems(int max_width, int nphig);
template <typename T>
cudaError_t launch_mlp_bwd_wide(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* xM, int64_t nb, const T* mubar,
                                int M_units, const T* zetaP, void* work, double* grad_acc /*[nphig], zeroed*/,
                                double* xacc /*[max(M_units,1)*npc], zeroed*/, WideCache cache = WideCache{nullptr, 0, 0});
size_t wide_bwd_work_bytes(int nL, int max_width, int nphig);
// TMA-fed 3xBF16 chain of the wide applicator (hvi_dense_tma.cu), Float32 and hidden widths % 128 == 0
bool bf3_eligible(const Cfg<float>& cfg);
size_t bf3_fwd_work_bytes(int maxw, int nphig);
size_t bf3_bwd_work_bytes(int nL, int maxw, int nphig);
size_t bf3_cache_slot_bytes(int nL, int maxw, int64_t rows = 0);
int64_t bf3_chunks(int64_t nb);
int64_t bf3_chunk_rows();   // columns per chunk of the wide chain (WIDE_CHUNK)
cudaError_t launch_mlp_fwd_bf3(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb, float* mu,
                               int M_units, const float* zetaP, void* work, WideCache cache);
cudaError_t launch_mlp_bwd_bf3(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb, const float* mubar,
                               int M_units, const float* zetaP, void* work, double* grad_acc, double* xacc, WideCache cache);
cudaError_t gemm_bf3_device(cudaStream_t stream, int sm_count, int mode, int64_t M, int N, int64_t K, const float* A, const float* B,
                            const float* bias, int act, const float* Hd, float* C);
// tensor-core (mma.sync, 3xBF16) small applicator, Float32 (hvi_mlp_mma.cu); cudaErrorNotSupported: not one of its shapes
cudaError_t launch_mlp3_mma_fwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb,
                                float* mu, int M_units, const float* zetaP);
cudaError_t launch_mlp3_mma_bwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM,
                                const float* mubar, int64_t nb, double* part, int* nblk_out, int nblk_max, int M_units,
                                const float* zetaP, double* part_x);
// tcgen05 small applicator, Float32 (hvi_mlp_tc.cu); cudaErrorNotSupported: not one of its shapes or HVI_MLP_TC=0
cudaError_t launch_mlp3_tc_fwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM, int64_t nb,
                               float* mu, int M_units, const float* zetaP);
cudaError_t launch_mlp3_tc_bwd(const Launch& L, const Cfg<float>& cfg, const float* phi, const float* xM,
                               const float* mubar, int64_t nb, double* part, int* nblk_out, int nblk_max, int M_units,
                               const float* zetaP, double* part_x, const float* mufwd /*μ_ζ of the forward pass or NULL*/);
// point-solver loss (hvi_point.cu)
template <typename T>
cudaError_t launch_point(const Launch& L, const Cfg<T>& cfg, const T* phi, const T* rec, const T* mu, T* mubar, int64_t nb,
                         double* part /*[point_max_rows][point_row_len]*/, int* nrows_out);
template <typename T>
cudaError_t launch_point_finalize(const Launch& L, const Cfg<T>& cfg, const T* phi, const double* part, int nrows,
                                  const double* part_mlp, int nblk, const double* part_x, const double* batch_consts,
                                  double* out,
                                  T* grad_T);
template <typename T>
bool small_iter_supported(const Cfg<T>& cfg, int64_t nb, int M);
template <typename T>
cudaError_t launch_small_iter(const Launch& L, const Cfg<T>& cfg, SmallArgs<T> a);
int point_max_rows(int sm_count);
int point_row_len();
int stream_max_rows(int sm_count);
// generic process-based-model seam (hvi_generic.cu, hvi_pbm.cuh)
struct GenericPbm;
int generic_pbm_create(const char* spec, int dtype, int NP, int NM, int NO, int NX, bool want_mu, std::string& err, GenericPbm** out);
int generic_pbm_compile_check(const char* spec, int dtype, int NP, int NM, int NO, int NX, std::string& err);
void generic_pbm_destroy(GenericPbm* g);
bool generic_pbm_has_penalty(const GenericPbm* g);
template <typename T>
cudaError_t launch_generic_stream(const Launch& L, GenericPbm* g, const Cfg<T>& cfg, const StreamArgs<T>& s, const T* xP,
                                  const T* y_o, const T* y_unc, const double* fix, double* pen_partials, int* nrows_out,
                                  int nrows_max);
template <typename T>
cudaError_t launch_preprocess_generic(const Launch& L, int64_t n, const T* y_o, const T* y_unc, double* part, int* nrows_out);
cudaError_t launch_add_penalty(const Launch& L, const double* pen, int nrows, int M, int nphi, double* out);
int mlp_bwd_max_blocks(int sm_count);
bool stream_supported(int nP, int nM, int nO);
bool mlp_supports_pbm(const hvi_desc* d);

}  // namespace hvi
