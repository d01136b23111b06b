/*
 * swr_b200.h -- C ABI of the B200-native Scenario-Wise-Rec hot path.
 *
 * The reference (Xiaopengli1/Scenario-Wise-Rec) is pure Python on PyTorch and defines
 * no FFI; its boundary for this path is the nn.Module API.  This
 * header is the plugin boundary a reference maintainer would bind (ctypes stub in
 * INTEGRATION.md): plain pointers and sizes, no torch types.  Every device pointer is
 * BORROWED for the duration of the call; the library allocates nothing persistent.
 * All entry points are asynchronous on `stream` (a cudaStream_t passed as void*),
 * re-entrant across devices/streams (no global mutable state except the per-thread
 * error string), and return 0 on success or a negative swr_status.
 *
 * Reference interfaces replaced (paths relative to /root/reference/scenario_wise_rec):
 *   swr_embedding_gather_fwd   basic/layers.py:64-114   EmbeddingLayer.forward(squeeze_dim=True)
 *   swr_embedding_scatter_bwd  autograd of basic/layers.py:70 (aten::embedding_dense_backward)
 *   swr_program_run            the expert / gate / domain-tower stack executed by
 *                              basic/layers.py:231-264 (MLP), :307-320 (GateNU) and the
 *                              model forwards models/multi_domain/{sharebottom.py:28-50,
 *                              mmoe.py:33-56, ple.py:41-136, star.py:78-118, ppnet.py:21-67,
 *                              epnet.py:25-32, hamur.py:101-378, m3oe.py:135-198} plus
 *                              their autograd backward (trainers/ctr_trainer.py:72).
 */
#ifndef SWR_B200_H_
#define SWR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWR_ABI_VERSION 2

#if defined(__GNUC__)
#define SWR_API __attribute__((visibility("default")))
#else
#define SWR_API
#endif

typedef enum swr_status {
  SWR_OK = 0,
  SWR_ERR_INVALID = -1,     /* bad argument / malformed program record            */
  SWR_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed               */
  SWR_ERR_UNSUPPORTED = -3, /* shape / dtype outside what the kernels implement   */
  SWR_ERR_NO_DEVICE = -4    /* no sm_100 device available                         */
} swr_status;

/* element types of caller-provided feature columns (the reference casts with
 * .long() / .float(), basic/layers.py:70,89; reduce_mem_usage yields int8..int64 and
 * float16..float64 columns, utils/data.py:109-124) */
typedef enum swr_dtype {
  SWR_I8 = 0, SWR_I16 = 1, SWR_I32 = 2, SWR_I64 = 3, SWR_U8 = 4,
  SWR_F16 = 8, SWR_BF16 = 9, SWR_F32 = 10, SWR_F64 = 11
} swr_dtype;

SWR_API int swr_abi_version(void);
/* last error message of the calling thread ("" if none) */
SWR_API const char* swr_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's
 * gpu_launches claim is counted here, not estimated) */
SWR_API int64_t swr_launch_count(void);
/* 0 if the current device is sm_100-class, else SWR_ERR_NO_DEVICE */
SWR_API int swr_device_check(void);
/* Arithmetic of the grouped FC ops (the reference's nn.Linear contractions, basic/layers.py:254):
 *   SWR_FC_SIMT  fp32 FFMA kernels (round-to-nearest accumulate, the numerics of a stock fp32 GPU run)
 *   SWR_FC_TC    tcgen05 3xTF32 kernels for every layer whose shape they accept
 *   SWR_FC_AUTO  tcgen05 for the wide layers, FFMA for the narrow ones (default; env SWR_FC_TC=0|1|2 presets it)
 * Process-wide; takes effect for programs run (or CUDA graphs captured) afterwards.  Returns the previous mode. */
typedef enum { SWR_FC_SIMT = 0, SWR_FC_TC = 1, SWR_FC_AUTO = 2 } swr_fc_mode;
SWR_API int swr_set_fc_mode(int mode);
SWR_API int swr_get_fc_mode(void);

/* Per-op device timing of swr_program_run (bench.py's live roofline measurement).
 * swr_profile_begin(): every op of the program runs issued by this process (any thread) from now
 * on is bracketed by CUDA events on the launch stream (no synchronisation is added).
 * swr_profile_end(): synchronises, writes for each recorded op its op kind, the index of
 * its header record inside its program and its device time in ms (up to `cap` entries),
 * stops profiling and returns the number of ops recorded (or a negative swr_status). */
SWR_API int swr_profile_begin(void);
SWR_API int swr_profile_end(int32_t* kinds, int32_t* rec_index, float* ms, int32_t cap);

/* ------------------------------------------------------------------------------------
 * K1: fused multi-field embedding gather (+ dense append)
 *   out[b, f*E .. f*E+E)          = tables[f][ idx[f][b] , : ]      f < n_sparse
 *   out[b, n_sparse*E + j]         = (float) dense[j][b]            j < n_dense
 * Sparse fields first in list order, then dense scalars (basic/layers.py:93-104).
 * An index outside [0, vocab[f]) writes a zero row and sets oob_flag[0] = 1,
 * oob_flag[1] = f (the reference raises IndexError / device-asserts; the host wrapper
 * raises IndexError from the flag).  oob_flag may live in mapped pinned host memory.
 * ---------------------------------------------------------------------------------- */
SWR_API int swr_embedding_gather_fwd(const float* const* tables, const int64_t* vocab,
                             const void* const* idx, const int32_t* idx_dtype,
                             const void* const* dense, const int32_t* dense_dtype,
                             float* out, int64_t ld_out, int64_t batch,
                             int32_t n_sparse, int32_t embed_dim, int32_t n_dense,
                             int32_t* oob_flag, void* stream);

/* ------------------------------------------------------------------------------------
 * K2: embedding-gradient scatter-add (dense-gradient semantics of nn.Embedding)
 *   grad_tables[f][ idx[f][b], : ] += grad_out[b, f*E .. f*E+E)
 * grad_tables must be zero-filled by the caller (the dense [V_f, E] gradient the
 * reference materialises); accumulation uses fp32 red.global atomics (vectorised),
 * duplicate rows inside a warp are pre-combined.  Out-of-range indices are skipped.
 * ---------------------------------------------------------------------------------- */
SWR_API int swr_embedding_scatter_bwd(const float* grad_out, int64_t ld_grad, int64_t batch,
                              const void* const* idx, const int32_t* idx_dtype,
                              float* const* grad_tables, const int64_t* vocab,
                              int32_t n_sparse, int32_t embed_dim, void* stream);

/* ------------------------------------------------------------------------------------
 * Device program: the expert / gate / domain-tower stack and its backward as a list of
 * fused ops over a pointer ("slot") table.  The host mirror (python package) lowers a
 * model to records once; per call it patches the per-call slots (feature columns,
 * gradient arenas) and issues ONE swr_program_run, which launches every kernel of the
 * pass on `stream` (CUDA-graph capturable: no syncs, no allocations).
 *
 * A record is either an op header (kind = SWR_OP_*, n_sub = number of group records
 * that follow) or a group record.  Field use per op is documented in DESIGN.md
 * "Program records" and mirrored by oracle/ops_ref.py (the CPU checker).
 * ---------------------------------------------------------------------------------- */
#define SWR_REC_INTS 32
#define SWR_REC_FLOATS 8
#define SWR_REC_SLOTS 32

typedef struct swr_rec {
  int32_t kind;
  int32_t n_sub;
  int32_t i[SWR_REC_INTS];
  float f[SWR_REC_FLOATS];
  int32_t s[SWR_REC_SLOTS]; /* indices into the slot table, -1 = null */
} swr_rec_t;

typedef enum swr_op_kind {
  SWR_OP_ZERO = 1,        /* memset a slot                                             */
  SWR_OP_GATHER = 2,      /* K1 via records                                            */
  SWR_OP_SCATTER = 3,     /* K2 via records                                            */
  SWR_OP_COLSTATS = 4,    /* column sum / sum-of-squares of a plain [B, n] buffer      */
  SWR_OP_FC_FWD = 5,      /* grouped Y = act_in(norm(A)) * W^T + b (+ column stats)    */
  SWR_OP_FC_DGRAD = 6,    /* fan-in dA = sum_g dY_g * W_g, fused norm/act backward     */
  SWR_OP_FC_WGRAD = 7,    /* grouped dW = dY^T * act_in(norm(A)), db                   */
  SWR_OP_POOL_FWD = 8,    /* gate softmax + expert pooling                             */
  SWR_OP_POOL_BWD = 9,
  SWR_OP_HEAD_FWD = 10,   /* per-domain Linear(H,1) + sigmoid + domain mask-select     */
  SWR_OP_HEAD_BWD = 11,
  SWR_OP_BN_UPDATE = 12,  /* BatchNorm running-stat update for a list of layers        */
  SWR_OP_BN_PGRAD = 13,   /* d gamma / d beta from the reduced backward statistics     */
  SWR_OP_EW_FWD = 14,     /* out = value(A) * value(C) * scale | value(A) + value(C) | value(A) */
  SWR_OP_EW_BWD = 15,
  SWR_OP_SUMGRAD = 16,    /* dst.dz (+)= sum_i dY(view_i): backward of normalised views of one
                             raw tensor (STAR partitioned norm, star.py:95-100)          */
  SWR_OP_SELECT_FWD = 17, /* out[b,:] = value(Y_{domain[b]})[b,:]  (M3oE star layer)     */
  SWR_OP_SELECT_BWD = 18,
  SWR_OP_LN_FWD = 19,     /* out = act(LayerNorm(Y)) row-wise (m3oe.py:45-68)            */
  SWR_OP_LN_BWD = 20,
  SWR_OP_MIX_FWD = 21,    /* M3oE domain-expert mixing with sigmoid scalar weights       */
  SWR_OP_MIX_BWD = 22,
  SWR_OP_BMV_FWD = 23,    /* HAMUR per-sample q[b,:] = p[b,:] * H_b (hamur.py:177-189)   */
  SWR_OP_BMV_BWD = 24,
  SWR_OP_BCE = 25,        /* BCELoss forward (mean) + d loss / d prediction (ctr_trainer.py:70) */
  SWR_OP_ADAM = 26,       /* torch.optim.Adam over flat param/grad/moment arenas (ctr_trainer.py:73) */
  SWR_OP_FC_PRESPLIT = 27,/* effective weights W (.) W2 -> hi / lo TF32 images (both orientations) that the
                             tcgen05 FC kernels read through TMA; sub-records = FC group records whose
                             s[10] / s[11] name the forward / data-gradient image buffers           */
  SWR_OP_ADAM_ROWS = 28,  /* row-lazy Adam on embedding tables: i[1] = 0 catch-up of the batch's rows before the
                             gather, 1 update after the scatter (same trajectory as the dense sweep)    */
  SWR_OP_ADAM_FLUSH = 29, /* replay every postponed row update of the listed tables up to the current step    */
  SWR_OP_GROUP = 100      /* a group record belonging to the preceding header          */
} swr_op_kind;

/* norm modes of a lazily-normalised activation */
enum { SWR_NORM_NONE = 0, SWR_NORM_BATCH = 1, SWR_NORM_RUNNING = 2 };
/* activations */
enum { SWR_ACT_NONE = 0, SWR_ACT_RELU = 1, SWR_ACT_SIGMOID = 2, SWR_ACT_LEAKY = 3 };
/* weight layouts */
enum { SWR_W_NK = 0 /* nn.Linear [N,K] */, SWR_W_KN = 1 /* STAR [K,N] */ };
/* element-wise modes of SWR_OP_EW_* */
enum { SWR_EW_MUL = 0, SWR_EW_ADD = 1, SWR_EW_COPY = 2 };
/* head modes: select(sigmoid(v_d)) | sigmoid(select(v_d) + add) | sigmoid(v_0) for every row */
enum { SWR_HEAD_SELECT_SIG = 1, SWR_HEAD_SIG_SELECT_ADD = 0, SWR_HEAD_NO_SELECT = 2 };

SWR_API int swr_program_run(const swr_rec_t* recs, int32_t n_recs, void* const* slots,
                    int32_t n_slots, void* stream);

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): the trainer's single H2D copy of a
 * packed batch (pinned host staging -> device staging) and the D2H read-back of the loss ring. */
SWR_API int swr_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Peer memory for row-sharded embedding tables (one process per GPU on one NVSwitch node; no reference counterpart:
 * the reference keeps every table whole on one device, basic/features.py:76-79).  A rank allocates its shard arenas
 * with swr_peer_alloc, publishes the 64-byte handle (torch.distributed all_gather), and maps every peer's arena with
 * swr_peer_open; K1 then reads looked-up rows and K2 adds gradient rows directly in the owner's memory over NVLink.
 * ---------------------------------------------------------------------------------- */
#define SWR_PEER_HANDLE_BYTES 64
SWR_API int swr_peer_alloc(int64_t bytes, void** ptr);                 /* cudaMalloc + zero fill                   */
SWR_API int swr_peer_free(void* ptr);
SWR_API int swr_peer_handle(void* ptr, unsigned char* handle64);       /* cudaIpcGetMemHandle                      */
SWR_API int swr_peer_open(const unsigned char* handle64, void** ptr);  /* cudaIpcOpenMemHandle (lazy peer access)  */
SWR_API int swr_peer_close(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* SWR_B200_H_ */
