// swr_launch.h -- host-side launch descriptors shared by the kernel files and the
// program executor (swr_exec.cu).  Everything here is plain pointers + sizes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "swr_common.cuh"

namespace swr {

struct GatherLaunch {
  const float* const* tables; const int64_t* vocab; const void* const* idx; const int32_t* idx_dtype;
  const int32_t* E;                 // per-field embed dim
  const void* const* dense; const int32_t* dense_dtype;
  float* out; int64_t ld; int64_t B; int n_sparse; int n_dense; int32_t* oob;
  const int32_t* world = nullptr;                 // per field: > 0 = row-sharded over that many ranks (vocab = full vocabulary)
  const float* const* const* peers = nullptr;     // per field: device array of `world` shard base pointers
};
int launch_gather(const GatherLaunch& g, cudaStream_t st);

struct ScatterLaunch {
  const float* g; int64_t ld; int64_t B;
  const void* const* idx; const int32_t* idx_dtype; float* const* gtables; const int64_t* vocab;
  const int32_t* E; const int32_t* col;   // col may be null (fields packed in order)
  int n_sparse;
  const int32_t* world = nullptr;          // per field: > 0 = row-sharded gradient (see GatherLaunch)
  float* const* const* peers = nullptr;
};
int launch_scatter(const ScatterLaunch& s, cudaStream_t st);

int launch_colstats(const float* x, int64_t B, int n, int64_t ld, double* stats, cudaStream_t st);

// ---- fully-connected stack ------------------------------------------------------------
struct FcGroup {
  ActDev A;             // input activation (lazy)
  ActDev Y;             // output activation: raw = Linear output, norm/act = what follows it
  const float* W; const float* W2;        // effective weight = W (.) W2
  const float* bias; const float* bias2;  // effective bias = bias + bias2
  float* dW; float* dW2; float* dbias; float* dbias2;
  double* stats_out;    // [N][2] column moments of Y (train-mode BatchNorm), or null
  const float* img_f;   // presplit weight images (swr_fc_tc.cu): [2][N][K32] forward, [2][K][N32] data gradient; null = none
  const float* img_d;
  int k_full;           // input width of the layer's weight (A.n may be narrowed to the columns that receive a data gradient)
  int w_layout; int ldw;
  int e_act; float e_scale;   // epilogue activation for layers without a norm (GateNU)
  int flags;            // bit0: A needs a gradient, bit1: accumulate into A.dz
};
enum { FC_A_NEEDS_GRAD = 1, FC_A_ACCUMULATE = 2 };

int launch_fc_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);
int launch_fc_dgrad(const FcGroup* groups, const int* dst_of, int n_groups, int64_t B, cudaStream_t st);
int launch_fc_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);
// tcgen05 path (swr_fc_tc.cu); the launchers above route to it when fc_tc_wanted() says so
bool fc_tc_wanted(const FcGroup* groups, int n_groups, int64_t B);
int fc_mode_get();
int fc_mode_set(int mode);
// TMA-fed presplit weights, persistent warp-specialised kernels.
// pass: 0 forward, 1 data gradient, 2 weight gradient
bool fc_tc2_usable(const FcGroup* groups, int n_groups, int pass);
int launch_fc_presplit(const FcGroup* groups, int n_groups, cudaStream_t st);
int launch_fc_tc2_fwd(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);
int launch_fc_tc2_dgrad(const FcGroup* groups, const int* dst_group, int n_dst, int n_groups, int64_t B, cudaStream_t st);
int launch_fc_tc2_wgrad(const FcGroup* groups, int n_groups, int64_t B, cudaStream_t st);

// ---- row-local ops ----------------------------------------------------------------------
constexpr int kMaxPoolExperts = 16;
struct PoolGate {
  ActDev gate;          // logits [B, nE] (BatchNorm, no activation; softmax applied here)
  float* pooled;        // [B, ldp] plain output
  const float* dpooled; // [B, ldp] gradient of the output
  float* probs;         // [B, nE] saved softmax
  int ldp; int nE;
  int expert[kMaxPoolExperts];   // indices into the unique expert list
};
struct PoolLaunch {
  const PoolGate* gates; int n_gates;
  const ActDev* experts; int n_experts;
  int64_t B; int H;
};
int launch_pool_fwd(const PoolLaunch& p, cudaStream_t st);
int launch_pool_bwd(const PoolLaunch& p, cudaStream_t st);

struct HeadDomain {
  ActDev A;             // [B, H] lazy tower hidden (or [B,1] when w == null)
  const float* w; const float* bias; float* dw; float* dbias;
};
struct HeadLaunch {
  const HeadDomain* dom; int n_domains;
  const void* domain_id; int dom_dtype;
  float* out; const float* gout;           // [B]
  const float* add; float* dadd;           // optional plain [B, ld_add] column added before the sigmoid (STAR aux)
  int ld_add;
  int sig_before_select;                   // 1: select(sigmoid(v_d))   0: sigmoid(select(v_d) + add)   2: sigmoid(v_0), no select
  int64_t B;
};
int launch_head_fwd(const HeadLaunch& h, cudaStream_t st);
int launch_head_bwd(const HeadLaunch& h, cudaStream_t st);

struct BnLayer {
  ActDev A;
  int repeat;                                          // BN_UPDATE: apply the update this many times (HAMUR shared hyper-net)
  float* rmean; float* rvar; int64_t* nbt;            // BN_UPDATE
  float* dgamma; float* dgamma2; float* dbeta; float* dbeta2;   // BN_PGRAD
};
int launch_bn_update(const BnLayer* layers, int n, int64_t B, float momentum, cudaStream_t st);
int launch_bn_pgrad(const BnLayer* layers, int n, int64_t B, cudaStream_t st);

// ---- element-wise / row-local glue (swr_glue.cu) ------------------------------------------------
struct EwGroup {
  ActDev A, C;              // C unused for SWR_EW_COPY
  float* out; const float* dout; int ld_out;   // plain [B, n] result and its gradient
  int mode; float scale;
  int flags;                // bit0 A needs grad, bit1 C needs grad, bit2 accumulate into A.dz, bit3 accumulate into C.dz
};
enum { EW_A_GRAD = 1, EW_C_GRAD = 2, EW_A_ACC = 4, EW_C_ACC = 8 };
int launch_ew_fwd(const EwGroup* g, int n_groups, int64_t B, cudaStream_t st);
int launch_ew_bwd(const EwGroup* g, int n_groups, int64_t B, cudaStream_t st);

constexpr int kMaxViews = 16;
struct SumGradLaunch { ActDev dst; ActDev views[kMaxViews]; int n_views; int accumulate; int64_t B; };
int launch_sumgrad(const SumGradLaunch& s, cudaStream_t st);

struct SelectLaunch {
  ActDev Y[16]; int n_domains;
  const void* domain_id; int dom_dtype;
  float* out; const float* dout; int ld_out; int n; int64_t B;
};
int launch_select_fwd(const SelectLaunch& s, cudaStream_t st);
int launch_select_bwd(const SelectLaunch& s, cudaStream_t st);

struct LnGroup {
  const float* y; float* dy; int ld_y;         // plain Linear output and its gradient
  const float* gamma; const float* beta; float* dgamma; float* dbeta;
  float* out; const float* dout; int ld_out;   // act(LayerNorm(y))
  float* rowstats;                             // [B][2] mean, rstd
  int n; int act; float eps;
};
int launch_ln_fwd(const LnGroup* g, int n_groups, int64_t B, cudaStream_t st);
int launch_ln_bwd(const LnGroup* g, int n_groups, int64_t B, cudaStream_t st);

struct MixLaunch {
  const float* X[16]; float* dX[16]; int ldx;          // domain-expert outputs (plain) and their gradients
  float* out[16]; const float* dout[16]; int ldo;      // pooled outputs (accumulated into) and their gradients
  const float* w_exp; const float* w_bal; float* dw_exp; float* dw_bal;
  double* red;                                         // [2] scratch, zeroed by the caller
  int D; int H; int64_t B;
};
int launch_mix_fwd(const MixLaunch& m, cudaStream_t st);
int launch_mix_bwd(const MixLaunch& m, cudaStream_t st);

struct BmvGroup { const float* p; float* dp; int ldp; float* q; const float* dq; int ldq; };
struct BmvLaunch {
  BmvGroup g[16]; int n_groups;
  const float* H; float* dH; int ldh;                  // plain [B, k*k] and its gradient
  int k; int accumulate_dH; int64_t B;
};
int launch_bmv_fwd(const BmvLaunch& m, cudaStream_t st);
int launch_bmv_bwd(const BmvLaunch& m, cudaStream_t st);

// ---- trainer-side ops (swr_train.cu) ---------------------------------------------------------------
int launch_bce(const float* pred, const void* label, int label_dtype, float* gout, float* loss_ring, const int32_t* ctrl,
               int ring, int64_t B, float gscale, cudaStream_t st);
int launch_adam(float* p, float* g, float* m, float* v, const float* hyper, int64_t n, int zero_grad, cudaStream_t st);

// row-lazy Adam for embedding tables (swr_train.cu): catch-up before the gather, update after the scatter, flush
struct LazyField {
  float* p; float* g; float* m; float* v;     // [vocab, E] views into the flat arenas
  int* last; int* claim;                      // [vocab]
  const void* idx; int idx_dtype;             // index column [B]
  int64_t vocab; int E;
  int world, rank;                            // world > 1: the table is this rank's shard (row r lives on rank r % world at local row
                                              // r / world): lookups of other ranks' rows are skipped, idx holds global rows
};
int launch_adam_rows(const LazyField* fields, int n_fields, int64_t B, const float* hyper, const int32_t* ctrl, float4* hist, int phase,
                     cudaStream_t st);
int launch_adam_flush(const LazyField* fields, int n_fields, const float* hyper, const int32_t* ctrl, const float4* hist, cudaStream_t st);

}  // namespace swr
