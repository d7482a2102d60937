/*
 * matx_b200.h — C ABI of libmatx_b200.so, the B200 (sm_100a) engine behind the MatX operator API.
 *
 * This is the drop-in boundary.  Everything above it (the MatX expression templates,
 * `(out = sum(a*b+c, {1})).run(exec)`) stays the reference's own code; the header shim
 * `include/matx_b200/executor.h` lowers a statement to the plain-C descriptors declared here and
 * calls one of the entry points below.  No C++ types, no torch types, no exceptions cross this line.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference tree):
 *   mxb_elementwise  <- cudaExecutor::Exec                    include/matx/executors/cuda.h:84-230
 *                       + matxOpT{0..4,D}Kernel               include/matx/executors/kernel.h:41-223
 *   mxb_reduce       <- sum_impl / mean_impl / var_impl / stdd_impl / max_impl / min_impl / argmax_impl /
 *                       argmin_impl / any_impl / all_impl / prod_impl (cudaExecutor overloads)
 *                                                             include/matx/transforms/reduce.h:266-290,646-655,
 *                                                             712-722,786-795,856-869,934-943,1003-1017,
 *                                                             1170-1178,1243-1251,1406-1444,1474-1479
 *                       + matxCubPlan_t::Exec{Sum,Min,Max,Reduce,ArgReduce}
 *                                                             include/matx/transforms/cub.h:647-894,1281-1328
 *   mxb_softmax      <- softmax_impl (both overloads)         include/matx/transforms/reduce.h:362-445
 *   mxb_cumsum       <- cumsum_impl + ExecPrefixScanEx        include/matx/transforms/cub.h:2367-2395,375-408
 *   mxb_find         <- find_impl / find_idx_impl + ExecSelect / ExecSelectIndex
 *                                                             include/matx/transforms/cub.h:2609-2625,2705-2721,912-1010
 *   mxb_hist         <- hist_impl + ExecHistEven              include/matx/transforms/cub.h:2464-2503,320-359
 *   mxb_sort         <- sort_impl + ExecSort                  include/matx/transforms/cub.h:2145-2190,428-560
 *   mxb_unique       <- unique_impl + ExecUnique              include/matx/transforms/cub.h:2796-2842,1052-1110
 *   mxb_argminmax    <- argminmax_impl + ExecDualArgReduce    include/matx/transforms/reduce.h:1090-1109, cub.h:1439-1491
 *   mxb_create / mxb_destroy / mxb_set_stream
 *                    <- cudaExecutor ctor / getStream         include/matx/executors/cuda.h:60-82
 *   mxb_sync         <- CudaExecutorBase::sync                include/matx/executors/cuda_executor_common.h:137
 *
 * Conventions
 *   - every call is asynchronous and ordered on the handle's stream (like the reference);
 *   - every call returns an mxb_status_t; mxb_last_error() gives the text for the calling thread;
 *   - sizes, strides and indices are int64 end to end (the reference's CUB path is limited to int);
 *   - strides are in ELEMENTS, stride 0 = broadcast (MatX clone / scalar / lower-rank operand);
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     MXB_ERR_NO_DEVICE.
 */
#ifndef MATX_B200_H
#define MATX_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MXB_VERSION_MAJOR 0
#define MXB_VERSION_MINOR 1

#define MXB_MAX_RANK   8   /* rank of an expression's index space */
#define MXB_MAX_LEAVES 12  /* distinct tensor views in one expression */
#define MXB_MAX_NODES  96  /* SSA nodes in one expression */
#define MXB_MAX_CONSTS 24  /* scalar constants in one expression */

typedef enum {
  MXB_OK = 0,
  MXB_ERR_INVALID = 1,       /* bad argument / malformed program     -> matxInvalidParameter */
  MXB_ERR_NOT_SUPPORTED = 2, /* legal MatX, not lowered by this path -> matxNotSupported (shim falls back) */
  MXB_ERR_CUDA = 3,          /* CUDA runtime error                   -> matxCudaError */
  MXB_ERR_NO_DEVICE = 4,     /* no CUDA device visible */
  MXB_ERR_JIT = 5,           /* NVRTC missing or generated kernel failed to build */
  MXB_ERR_SIZE = 6           /* output shape does not match          -> matxInvalidSize */
} mxb_status_t;

/* element types (reference: value_type of tensor_t; core/half.h for the 16-bit floats) */
typedef enum {
  MXB_F32 = 0,
  MXB_F64 = 1,
  MXB_BF16 = 2,
  MXB_F16 = 3,
  MXB_C64 = 4,  /* cuda::std::complex<float>, interleaved re,im */
  MXB_I32 = 5,
  MXB_I64 = 6,  /* matx::index_t */
  MXB_U8 = 7,   /* bool / uint8 */
  MXB_DTYPE_COUNT
} mxb_dtype_t;

/* reductions (reference: transforms/reduce.h) */
typedef enum {
  MXB_RED_SUM = 0,
  MXB_RED_MEAN = 1,
  MXB_RED_VAR = 2,   /* two-pass arithmetic, divisor N - ddof */
  MXB_RED_STDD = 3,
  MXB_RED_MAX = 4,
  MXB_RED_MIN = 5,
  MXB_RED_ARGMAX = 6, /* value + absolute flat index, lowest index wins ties */
  MXB_RED_ARGMIN = 7,
  MXB_RED_ANY = 8,
  MXB_RED_ALL = 9,
  MXB_RED_PROD = 10,
  MXB_RED_COUNT
} mxb_reduce_op_t;

/* node opcodes of the expression program (reference functors: operators/scalar_ops.h:434-503) */
typedef enum {
  MXB_OP_LEAF = 0,  /* src[0] = leaf index */
  MXB_OP_CONST = 1, /* src[0] = constant index */
  /* binary */
  MXB_OP_ADD = 10, MXB_OP_SUB, MXB_OP_MUL, MXB_OP_DIV, MXB_OP_MOD, MXB_OP_POW, MXB_OP_MAX, MXB_OP_MIN,
  MXB_OP_LT, MXB_OP_GT, MXB_OP_LE, MXB_OP_GE, MXB_OP_EQ, MXB_OP_NE, MXB_OP_AND, MXB_OP_OR, MXB_OP_ATAN2,
  /* unary */
  MXB_OP_NEG = 40, MXB_OP_SQRT, MXB_OP_RSQRT, MXB_OP_EXP, MXB_OP_LOG, MXB_OP_LOG2, MXB_OP_LOG10, MXB_OP_ABS,
  MXB_OP_ABS2, MXB_OP_CONJ, MXB_OP_REAL, MXB_OP_IMAG, MXB_OP_SIN, MXB_OP_COS, MXB_OP_TAN, MXB_OP_TANH,
  MXB_OP_NORMCDF, MXB_OP_NOT, MXB_OP_ISNAN, MXB_OP_ISINF, MXB_OP_FLOOR, MXB_OP_CEIL, MXB_OP_ROUND,
  MXB_OP_SINH, MXB_OP_COSH, MXB_OP_ASIN, MXB_OP_ACOS, MXB_OP_ATAN, MXB_OP_EXPJ, MXB_OP_CSQRT_UNUSED,
  /* casts: src[0] = value, result dtype = aux */
  MXB_OP_CAST = 80
} mxb_opcode_t;

/* one SSA node: value id == its position in nodes[] */
typedef struct {
  int32_t opcode;
  int32_t src[2]; /* operand value ids (or leaf / constant index) */
  int32_t aux;    /* MXB_OP_CAST: target mxb_dtype_t */
} mxb_node_t;

/* a tensor view inside an expression: pointer + strides over the EXPRESSION's dims
 * (reference: tensor_impl_t = ldata_ + tensor_desc_t, core/tensor_impl.h, core/tensor_desc.h:221-233).
 * Permute / clone / slice / lower-rank broadcast are all expressed by the strides. */
typedef struct {
  const void *data; /* device pointer to element (0,0,...,0) of the view */
  int32_t dtype;    /* mxb_dtype_t */
  int32_t _pad;
  int64_t stride[MXB_MAX_RANK];
} mxb_leaf_t;

typedef struct {
  double re, im;  /* im used only for MXB_C64 */
  int32_t dtype;  /* mxb_dtype_t the literal has in the C++ expression (0.5f -> F32) */
  int32_t _pad;
} mxb_const_t;

/* expression program: SSA nodes over an N-D index space */
typedef struct {
  int32_t rank; /* 0 = scalar */
  int32_t n_nodes;
  int32_t n_leaves;
  int32_t n_consts;
  int32_t root; /* value id the expression evaluates to */
  int32_t _pad;
  int64_t size[MXB_MAX_RANK];
  mxb_node_t nodes[MXB_MAX_NODES];
  mxb_leaf_t leaves[MXB_MAX_LEAVES];
  mxb_const_t consts[MXB_MAX_CONSTS];
} mxb_expr_t;

/* destination view (reference: the LHS tensor of set<T,Op>, operators/set.h:121-553) */
typedef struct {
  void *data;
  int32_t dtype;
  int32_t rank;
  int64_t size[MXB_MAX_RANK];
  int64_t stride[MXB_MAX_RANK];
} mxb_out_t;

typedef struct mxb_context *mxb_handle_t;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Creates a handle bound to the calling thread's current device and to `stream` (a cudaStream_t,
 * NULL = legacy default stream).  The handle owns the scratch its launches use (grid-combine partials, tile exchange
 * slots); every entry point locks the handle, so it may be shared by several host threads — their statements are
 * serialised onto the handle's stream.  Threads that want to overlap use one handle each. */
int mxb_create(mxb_handle_t *out_handle, void *stream);
int mxb_destroy(mxb_handle_t h);
int mxb_set_stream(mxb_handle_t h, void *stream);
int mxb_sync(mxb_handle_t h);

/* ---- the hot path --------------------------------------------------------------------------- */
/* out(idx...) = expr(idx...)   — one fused kernel, each distinct leaf read once. */
int mxb_elementwise(mxb_handle_t h, const mxb_expr_t *expr, const mxb_out_t *out);

/* Reduce the trailing `n_reduce_dims` dims of `expr` (the caller has already moved the reduced dims
 * innermost, as the reference's sum(x, dims) does through permute: operators/sum.h:373-405).
 * `out` has rank expr->rank - n_reduce_dims.  `idx_out` (MXB_I64) is required for ARGMAX/ARGMIN and
 * receives b * R + r, the absolute flat offset in the collapsed input (cub.h:1289-1326 convention),
 * R = product of the reduced sizes.  `ddof` is used by VAR/STDD only.
 * One launch per call (VAR/STDD: one launch when a reduced row fits in shared memory). */
int mxb_reduce(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int n_reduce_dims,
               const mxb_out_t *out, const mxb_out_t *idx_out, int ddof);

/* min and max of the trailing `n_reduce_dims` dims with their absolute flat indices, ONE read of the operand (reference:
 * argminmax_impl, transforms/reduce.h:1090-1109 -> cub_dualargreduce, cub.h:1439-1491).  Four outputs of rank expr->rank -
 * n_reduce_dims; the two value outputs share a dtype, the index outputs are MXB_I64; lowest index wins ties on both sides. */
int mxb_argminmax(mxb_handle_t h, const mxb_expr_t *expr, int n_reduce_dims, const mxb_out_t *out_min, const mxb_out_t *idx_min,
                  const mxb_out_t *out_max, const mxb_out_t *idx_max);

/* out(b..., r...) = exp(x(b..., r...) - max_r x(b..., :)) / sum_r exp(x(b..., :) - max_r x(b..., :)) over the trailing
 * `n_reduce_dims` dims of `expr` (1 <= n_reduce_dims <= rank; the caller permutes the softmax axes innermost in BOTH
 * `expr` and `out`, which have the same rank and sizes) — the arithmetic of the reference's softmax_impl
 * (transforms/reduce.h:362-445: max_impl, sum_impl(exp(in - max)), then the divide; three passes, two temporaries and
 * five launches there).  Real floating expressions only.  One launch with one read and one write when a row fits in
 * the registers of a CTA (<= 64 KB of fp32), otherwise a one-pass statistics launch + one elementwise launch. */
int mxb_softmax(mxb_handle_t h, const mxb_expr_t *expr, int n_reduce_dims, const mxb_out_t *out);

/* out(b..., j) = sum_{i <= j} expr(b..., i): inclusive prefix sum along the LAST dim of `expr` (reference: cumsum_impl,
 * transforms/cub.h:2367-2395 -> matxCubPlan_t::ExecPrefixScanEx :375-408, cub::DeviceScan::InclusiveSum launched once
 * per row).  `out` has the rank and sizes of `expr`; accumulation in the expression's arithmetic type (fp32 for 16-bit
 * floats).  One launch, fixed summation order (run-to-run deterministic): many rows -> a CTA per row with a running
 * carry (one read, one write); few long rows -> 4 KB warp tiles whose totals are exchanged through L2 inside the launch;
 * a tile is read twice, the second time out of L2 for the most part (one write). */
int mxb_cumsum(mxb_handle_t h, const mxb_expr_t *expr, const mxb_out_t *out);

/* Stream compaction: out[0 .. n) = the elements x of `expr` (row-major flat order, stable) with `x <select_op> threshold`
 * (want_indices == 0) or their flat indices (want_indices != 0), and *count_out = n.  Reference: find_impl / find_idx_impl
 * (cudaExecutor overloads, transforms/cub.h:2609-2625,2705-2721 -> matxCubPlan_t::ExecSelect / ExecSelectIndex :912-1010,
 * cub::DeviceSelect::If) with the selection functors LT / GT / EQ / NEQ / LTE / GTE (:2521-2588); the host path
 * (:2656-2675,2752-2770) defines the order and the int count.  `out` is rank 1 and contiguous; values are converted to
 * its dtype, indices need MXB_I32 (the reference's static_cast<int>) or MXB_I64.  Elements beyond out->size[0] are counted
 * but not written (the reference would write past the end).  `count_out` is rank 0, MXB_I32.  Real value types only.
 * Views that collapse to one dim: ONE cooperative launch (every element read once; counts of 2048-element warp tiles
 * exchanged through epoch-tagged device slots, nothing to clear between calls); N-D views that do not collapse: two launches
 * (count + in-launch scan, scatter).  Deterministic and stable either way. */
typedef enum {
  MXB_SEL_LT = 0, MXB_SEL_GT = 1, MXB_SEL_EQ = 2, MXB_SEL_NEQ = 3, MXB_SEL_LTE = 4, MXB_SEL_GTE = 5, MXB_SEL_COUNT
} mxb_select_op_t;
int mxb_find(mxb_handle_t h, const mxb_expr_t *expr, int select_op, double threshold, const mxb_out_t *out,
             const mxb_out_t *count_out, int want_indices);

/* out(b..., k) = number of elements x of row b of `expr` (its last dim) with lower <= x < upper that fall in bin k of
 * `bins = out->size[last]` even-width bins (reference: hist_impl, transforms/cub.h:2464-2503 -> ExecHistEven :320-359,
 * cub::DeviceHistogram::HistogramEven with num_levels = bins + 1; the bin arithmetic is CUB's: int((x - lower) * (T(bins) /
 * T(upper - lower))) for floating types, ((x - lower) * bins) / (upper - lower) in 64-bit for integers).  `out` is MXB_I32,
 * same leading sizes as `expr`, bins of a row contiguous.  One memset + one launch. */
int mxb_hist(mxb_handle_t h, const mxb_expr_t *expr, double lower, double upper, const mxb_out_t *out);

/* Sort every row (last dim) of `expr` into `out`, ascending (descending != 0: descending) — keys only (reference:
 * sort_impl, transforms/cub.h:2145-2190 -> cub::DeviceRadixSort::SortKeys, per row for rank > 1; the HostExecutor
 * overload is std::sort per row).  fp32 / fp64 / int32 / int64; `out` contiguous, the operand's value type.  A
 * non-contiguous view or an expression is evaluated into `out` first (the reference copies as well). */
int mxb_sort(mxb_handle_t h, const mxb_expr_t *expr, const mxb_out_t *out, int descending);

/* out[0 .. n) = the distinct values of the rank-1 `expr` in ascending order, *count_out = n (reference: unique_impl,
 * transforms/cub.h:2796-2842: sort + cub::DeviceSelect::Unique; the HostExecutor overload is std::sort + std::unique).
 * Elements beyond out->size[0] are counted but not written. */
int mxb_unique(mxb_handle_t h, const mxb_expr_t *expr, const mxb_out_t *out, const mxb_out_t *count_out);

/* ---- multi-GPU (no counterpart in the reference) -------------------------------- */
/* Slab-sharded full-tensor reductions: each rank reduces its slab with mxb_reduce_partial into a
 * 32-byte device record, the host exchanges the records with ONE collective (NCCL all-gather of
 * world*32 bytes), and mxb_reduce_finalize folds them in rank order (deterministic; lowest global
 * index wins ties).  `slab_offset` is the global flat index of the slab's first element and
 * `global_count` the total element count (MEAN divisor).  VAR / STDD records hold the slab's (mean, sum |x-mean|^2, n)
 * in fp64 and are combined with Chan's formula in rank order, divisor N - ddof.  Rank r's record is read at
 * gathered_records + r * record_stride_bytes, so the records of several statements can share one
 * exchange (record_stride_bytes = 32 * statements per step; 0 means 32). */
#define MXB_PARTIAL_BYTES 32
int mxb_reduce_partial(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int64_t slab_offset,
                       void *partial_record);
int mxb_reduce_finalize(mxb_handle_t h, int reduce_op, int32_t value_dtype, const void *gathered_records,
                        int world, int64_t record_stride_bytes, int64_t global_count, int ddof,
                        const mxb_out_t *out, const mxb_out_t *idx_out);

/* Fused exchange over NVLink peer memory (one box, <= 8 ranks): instead of handing the records to a collective,
 * the reduction kernel's last block stores its record directly into EVERY rank's exchange buffer through peer
 * mappings and bumps an arrival counter there; mxb_exchange_finalize is one small kernel per step that waits for the
 * counters of all ranks and folds all statements' records in rank order.  No NCCL call, no host round trip, graph
 * capturable.  The host maps the buffers once (CUDA IPC): rec[r] / flag[r] point at rank r's buffers as seen from
 * THIS device; buffers are zero-initialised and sized MXB_EXCHANGE_REC_BYTES(world) / MXB_EXCHANGE_FLAG_BYTES(world);
 * epoch is a zeroed LOCAL control block of MXB_EXCHANGE_CTL_BYTES (completed exchanges, a sticky error word, and how many
 * records of every source rank have been consumed: steps with different n_items may share one exchange).  Records are
 * double-buffered by step parity, so a fast rank can run one step ahead.  A peer that does not deliver within 5 s makes
 * the fold write NaN / -1 into that step's outputs and set the error word: mxb_exchange_check (which synchronises the
 * handle's stream) then returns MXB_ERR_CUDA. */
#define MXB_EXCHANGE_CTL_BYTES 128
#define MXB_MAX_PEERS 8
#define MXB_MAX_ITEMS 8
#define MXB_EXCHANGE_REC_BYTES(world) (2 * (world) * MXB_MAX_ITEMS * MXB_PARTIAL_BYTES)
#define MXB_EXCHANGE_FLAG_BYTES(world) ((world) * 4)
typedef struct {
  void *rec[MXB_MAX_PEERS];
  void *flag[MXB_MAX_PEERS];
  void *epoch;
  int32_t world, rank;
} mxb_peers_t;
typedef struct {
  int32_t reduce_op;   /* SUM, MEAN, PROD, MAX, MIN, ARGMAX, ARGMIN, ANY, ALL, VAR, STDD */
  int32_t value_dtype; /* arithmetic type of the reduced expression (F32, F64, C64, I32, I64) */
  void *out;           /* one element of value_dtype (VAR / STDD: the real type, fp32 for F32 / C64, fp64 for F64) */
  void *idx_out;       /* one int64 (ARGMAX / ARGMIN), else NULL */
  int32_t ddof;        /* VAR / STDD divisor N - ddof */
  int32_t _pad;
} mxb_fold_item_t;
int mxb_reduce_partial_push(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int64_t slab_offset,
                            const mxb_peers_t *peers, int item, int n_items);
int mxb_exchange_finalize(mxb_handle_t h, const mxb_peers_t *peers, const mxb_fold_item_t *items, int n_items,
                          int64_t global_count);
int mxb_exchange_check(mxb_handle_t h, const mxb_peers_t *peers);
/* Exchange-buffer plumbing (CUDA IPC, one box).  mxb_exchange_alloc: cudaMalloc + zero `bytes` on the handle's
 * device and export a 64-byte IPC handle; the host passes the handles around (any transport) and every other rank
 * maps them with mxb_exchange_open from ITS device (peer access is enabled as part of the mapping).  The owner
 * releases with mxb_exchange_free, importers with mxb_exchange_close. */
#define MXB_IPC_HANDLE_BYTES 64
int mxb_exchange_alloc(mxb_handle_t h, size_t bytes, void **out_ptr, unsigned char ipc_handle_out[MXB_IPC_HANDLE_BYTES]);
int mxb_exchange_open(mxb_handle_t h, const unsigned char ipc_handle[MXB_IPC_HANDLE_BYTES], void **out_peer_ptr);
int mxb_exchange_close(mxb_handle_t h, void *peer_ptr);
int mxb_exchange_free(mxb_handle_t h, void *ptr);

/* ---- introspection ---------------------------------------------------------------------------- */
int mxb_version(void);                /* major*1000 + minor */
void mxb_reload_env(void);            /* the MXB_* tuning knobs are read once per thread: re-read them after changing one */
const char *mxb_last_error(void);     /* thread-local text of the last non-OK status */
int mxb_device_count(void);           /* 0 when no CUDA device / driver */
/* Name of the kernel family the last mxb_elementwise / mxb_reduce on this handle launched
 * (e.g. "reduce_inner<f32,sum,aot:fma3>"), and how many kernels the handle has launched. */
const char *mxb_last_kernel(mxb_handle_t h);
int64_t mxb_launch_count(mxb_handle_t h);
/* 1 if (expr, op) is served by an ahead-of-time kernel, 0 if it would be JIT-compiled. */
int mxb_is_aot(const mxb_expr_t *expr, int reduce_op_or_minus1);

#ifdef __cplusplus
}
#endif
#endif /* MATX_B200_H */
