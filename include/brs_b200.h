/*
 * brs_b200.h -- C ABI of libbrs_b200.so, the sm_100a hot path under
 * beta_rec's MatrixFactorization.train() / NeuCF.train() / LightGCN.train().
 *
 * The reference (beta-team/beta-recsys) is pure Python and has no FFI: its seam
 * is the duck-typed engine contract consumed by TrainEngine._train
 * (beta_rec/core/train_engine.py:225-240).  The Python engines in
 * beta_recsys_b200/engines mirror that contract and call ONLY the functions
 * declared here, through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every function returns BRS_OK (0) or a negative brs_status; no exceptions
 *     cross the ABI; brs_strerror() names a status.
 *   - all pointers inside the descriptor structs are DEVICE pointers owned by
 *     the caller (PyTorch allocates them); nothing is allocated, freed or
 *     retained by the library.  Tables are updated IN PLACE.
 *   - stream-ordered, no hidden synchronisation; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).
 *   - floats are fp32, indices are int64 (torch.LongTensor) exactly as the
 *     reference passes them (beta_rec/data/data_loaders.py:43-49).
 *   - batch-synchronous semantics: every sample of a batch reads PRE-step
 *     weights, duplicate rows' gradients are summed, then one optimizer update
 *     is applied (what autograd + torch.optim do in the reference).
 */
#ifndef BRS_B200_H
#define BRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRS_ABI_VERSION 2

typedef enum brs_status {
    BRS_OK = 0,
    BRS_ERR_INVALID_ARG = -1,   /* NULL pointer, negative size, misaligned table */
    BRS_ERR_UNSUPPORTED = -2,   /* dim / optimizer / layer shape the kernels do not cover */
    BRS_ERR_CUDA = -3,          /* a CUDA runtime call failed; see brs_last_cuda_error() */
    BRS_ERR_INDEX_RANGE = -4,   /* reserved: index outside its table (debug check) */
    BRS_ERR_NO_DEVICE = -5      /* no sm_100 device / driver */
} brs_status;

/* ---- optimizers: ModelEngine.set_optimizer (beta_rec/models/torch_engine.py:23-39) ---- */
typedef enum brs_opt_kind { BRS_SGD = 0, BRS_ADAM = 1, BRS_RMSPROP = 2 } brs_opt_kind;

/* BRS_DENSE reproduces the reference exactly: nn.Embedding(sparse=False) gives dense
 * gradients, so torch.optim.Adam/RMSprop update EVERY row every step (rows with zero
 * gradient keep moving through the decaying first moment).  BRS_TOUCHED_ROWS updates only
 * rows present in the batch ("lazy" Adam): exact for SGD, NOT equivalent for Adam/RMSprop
 * after the first step -- offered as a named, non-default mode. */
typedef enum brs_opt_mode { BRS_DENSE = 0, BRS_TOUCHED_ROWS = 1 } brs_opt_mode;

/* hyper-parameters are doubles: torch.optim keeps them as Python floats and derives the
 * per-step scalars (lr / (1 - beta1^t), sqrt(1 - beta2^t)) in double before casting to fp32 */
typedef struct brs_opt {
    int32_t kind;         /* brs_opt_kind */
    int32_t mode;         /* brs_opt_mode (ignored for SGD: always touched rows, which is exact) */
    double lr;
    double beta1, beta2;  /* Adam (torch defaults 0.9, 0.999) */
    double eps;           /* Adam / RMSprop (1e-8) */
    double alpha;         /* RMSprop (0.99) */
} brs_opt;

/* ---- one embedding table + its per-step gradient scratch and optimizer state ---- */
typedef struct brs_table {
    float *weight;   /* [n_rows, dim] row-major; 16-byte aligned when dim % 4 == 0 */
    float *grad;     /* capacity*dim floats of COMPACT gradient scratch: slot s holds the summed gradient of
                        table row rowset.list[s]; all-zero between steps.  The same few tens of MB are
                        reused every step, so the scatter-add stays in L2 instead of touching a
                        table-sized dense gradient (what autograd materialises in the reference).
                        Layout: row-major [capacity][dim], EXCEPT when dim % 8 == 0, where it is
                        sector-blocked [dim/8][capacity][8]: element (s, c) at ((c>>3)*capacity + s)*8 + (c&7),
                        so one row's 32-byte sectors land on different L2 slices (RED throughput). */
    float *m;        /* Adam exp_avg            (NULL for SGD / RMSprop) */
    float *v;        /* Adam exp_avg_sq / RMSprop square_avg (NULL for SGD) */
    int64_t n_rows;
    int32_t dim;
    int32_t pad_;
} brs_table;

/* ---- rows of one entity (users or items) touched by the current step ---- */
typedef struct brs_rowset {
    int32_t *slot_map; /* [n_rows] slot of the row in list/grad scratch, -1 when untouched; all -1 between steps */
    int32_t *list;     /* [capacity] touched row ids in slot order */
    int32_t *count;    /* [1] number of valid entries in list; zero between steps */
    int64_t n_rows;
    int32_t capacity;  /* >= min(n_rows, occurrences per batch) */
    int32_t pad_;
} brs_rowset;

#define BRS_MAX_ENTITY_TABLES 4
/* an entity = one id space (users / items) + the tables indexed by it */
typedef struct brs_entity {
    brs_rowset rows;
    int32_t n_tables;
    int32_t pad_;
    brs_table table[BRS_MAX_ENTITY_TABLES];
} brs_entity;

/* dense (non-embedding) parameters: global bias, Linear weights/biases.  Updated densely. */
typedef struct brs_dense_param {
    float *weight, *grad, *m, *v;
    int64_t numel;
} brs_dense_param;

/* device scratch shared by the kernels of one engine: 256 bytes, zero-initialised once */
#define BRS_STEP_WS_BYTES 256

/* what a step publishes (device memory, 4 floats): the two floats
 * MFEngine.train_single_batch returns (beta_rec/models/mf.py:119) + a status word */
typedef struct brs_step_out {
    float loss;         /* batch loss (mean over the batch) */
    float regularizer;  /* MF regularizer (mf.py:49-54); 0 for other models */
    int32_t status;     /* bit mask: 0 ok | 1 an index was outside its table (the reference raises IndexError)
                           | 2 touched-row list capacity exceeded */
    float reserved;
} brs_step_out;

/* ---- MF: beta_rec/models/mf.py (MF module + MFEngine) ---- */
/* Device scratch for the index plan of ONE batch (row-owner step): the batch's samples grouped by user
 * row (the user stream) and its (sample, item) entries grouped by item row (the item stream), one
 * contiguous segment per unique row.  Built by brs_mf_plan_build from the index arrays alone (no table
 * access), consumed by brs_mf_step_planned.  buf: brs_mf_plan_bytes(...) bytes, zero-filled once by the
 * caller. */
typedef struct brs_mf_plan {
    void *buf;
    int64_t bytes;
    int64_t batch_capacity;               /* samples per batch the plan can hold */
    int32_t user_capacity, item_capacity; /* = the rowset capacities it was sized for */
} brs_mf_plan;

typedef struct brs_mf_model {
    brs_entity user;              /* table[0] = user_emb [U,D], table[1] = user_bias [U,1] */
    brs_entity item;              /* table[0] = item_emb [I,D], table[1] = item_bias [I,1] */
    brs_dense_param global_bias;  /* numel 1 */
    void *ws;                     /* BRS_STEP_WS_BYTES of device scratch */
    /* optional second set of slot maps / lists / counters (same capacities; slot_map NULL = absent).  With
     * them brs_mf_train_batches runs the slot pre-pass of batch b+1 inside the apply launch of batch b. */
    brs_rowset user_rows_alt, item_rows_alt;
    /* row-owner step (brs_mf_step / brs_mf_train_batches): per-batch index plans (one per
     * rowset parity; plan[1] may be absent) and the staging copy of the batch's pre-step user rows
     * ([user rowset capacity, dim] fp32).  plan[0].buf == NULL or user_stage == NULL selects the
     * gradient-scratch path (brs_mf_*_fwd_bwd + brs_mf_apply) everywhere. */
    brs_mf_plan plan[2];
    float *user_stage;
} brs_mf_model;

/* one rank's MF shard as seen from the calling process (pointers mapped through CUDA IPC / NVLink peer
 * access); a device-resident array of these, indexed by rank, drives the row-sharded kernels */
typedef struct brs_mf_peer_tables {
    const float *user_emb, *item_emb, *user_bias, *item_bias;    /* shard weights [local_rows, dim] / [local_rows] */
    float *g_user_emb, *g_item_emb, *g_user_bias, *g_item_bias;  /* DENSE per-shard gradient tables, same shapes,
                                                                    row-major, all-zero between steps */
    uint32_t *user_bits, *item_bits;                             /* touched bitmaps [(local_rows+31)/32] */
} brs_mf_peer_tables;

/* library / device */
int brs_abi_version(void);
const char *brs_strerror(int status);
const char *brs_last_cuda_error(void);
int brs_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int64_t *l2_bytes);

/*
 * MF forward + backward for one BPR batch (replaces MF.forward x2 + ModelEngine.bpr_loss +
 * loss.backward(): beta_rec/models/mf.py:32-55,102-107,116-117; torch_engine.py:92-106).
 *   score = sigmoid(u.i + b_u + b_i + b_g)            (sigmoid BEFORE the BPR difference)
 *   loss  = -mean(logsigmoid(score_pos - score_neg));  regularizer as mf.py:49-54 for both calls
 * Accumulates d(loss + reg_weight*regularizer)/d(row) into table.grad, records touched rows
 * in the rowsets, and adds the batch's loss / regularizer sums into ws.  One fused kernel.
 */
int brs_mf_bpr_fwd_bwd(const brs_mf_model *model, const int64_t *users, const int64_t *pos_items,
                       const int64_t *neg_items, int64_t batch, float reg_weight, void *stream);

/* brs_mf_bpr_fwd_bwd = brs_mf_bpr_prepare (slot pre-pass: range check + one slot of the compact gradient
 * scratch per touched row) followed by brs_mf_bpr_fwd_bwd_prepared (the fused kernel); exposed separately
 * so that each launch can be timed / profiled on its own */
int brs_mf_bpr_prepare(const brs_mf_model *model, const int64_t *users, const int64_t *pos_items,
                       const int64_t *neg_items, int64_t batch, void *stream);
int brs_mf_bpr_fwd_bwd_prepared(const brs_mf_model *model, const int64_t *users, const int64_t *pos_items,
                                const int64_t *neg_items, int64_t batch, float reg_weight, void *stream);

/* diagnostics: L2 eviction priority (0 normal, 1 evict_first, 2 evict_last) used for the embedding-row
 * gathers, the compact gradient scratch, and the weight-row updates of the apply kernels */
int brs_debug_set_l2_policy(int gather, int scratch, int weight);

/* Same for loss == "bce" (beta_rec/models/mf.py:108-111; torch_engine.py:108-121, nn.BCELoss). */
int brs_mf_bce_fwd_bwd(const brs_mf_model *model, const int64_t *users, const int64_t *items,
                       const float *ratings, int64_t batch, float reg_weight, void *stream);

/*
 * optimizer.step() + the two .item() results (beta_rec/models/mf.py:118-119): applies `opt`
 * to the rows recorded by the preceding *_fwd_bwd call (and, in BRS_DENSE mode with
 * Adam/RMSprop, to every other row with g = 0), clears the gradient scratch / rowsets,
 * writes the step's brs_step_out to `out` (device) and resets ws.
 */
int brs_mf_apply(const brs_mf_model *model, const brs_opt *opt, int64_t batch, float *out /* brs_step_out */,
                 void *stream);

/*
 * MFEngine.train_an_epoch's inner loop (beta_rec/models/mf.py:132-136) over index arrays
 * already resident in HBM: for b in range(ceil(n/batch)): fwd_bwd + apply on
 * [b*batch, min((b+1)*batch, n)).  loss_kind 0 = bpr (third = neg item ids, int64),
 * 1 = bce (third = ratings, float).  out is device brs_step_out[n_batches].
 */
int brs_mf_train_batches(const brs_mf_model *model, const brs_opt *opt, int32_t loss_kind,
                         const int64_t *users, const int64_t *items, const void *third, int64_t n,
                         int64_t batch, float reg_weight, float *out /* brs_step_out[] */, void *stream);

/*
 * The same loop fed from HOST memory (the loader's side of train_an_epoch, mf.py:132-136, where the
 * reference moves each batch to the device inside train_single_batch): h_users / h_items / h_third
 * are host arrays (pinned for full speed; pageable works), h_out is host brs_step_out[n_batches].
 * Batch b+2 is copied into a device ring on a private copy stream while batch b computes, every step's
 * record is DMA'd back as soon as it is published.  Returns after the last step completed
 * (synchronizes `stream`).  One such loop at a time per process.
 */
int brs_mf_train_batches_host(const brs_mf_model *model, const brs_opt *opt, int32_t loss_kind,
                              const int64_t *h_users, const int64_t *h_items, const void *h_third, int64_t n,
                              int64_t batch, float reg_weight, float *h_out /* brs_step_out[] */, void *stream);

/*
 * Row-owner MF step (round 2; replaces the fwd_bwd + apply pair on the training path).  The same maths and
 * the same batch-synchronous semantics as brs_mf_*_fwd_bwd + brs_mf_apply (beta_rec/models/mf.py:92-119),
 * organised so that every touched table row is read from HBM once and written once, with no gradient
 * scratch and no atomics on the common path:
 *   plan   (index-only, brs_mf_plan_build): slot per unique row, samples grouped by user row and
 *          (sample, item) entries grouped by item row, one contiguous segment per row;
 *   users  (equal ranges of the user stream per lane group): gathers the three rows of each sample,
 *          evaluates score / loss / d loss once per block of 4 samples, keeps the user-row gradient in
 *          registers while the user stays the same, hands (coefficient, user slot) to the item stream,
 *          stages the PRE-step user row and writes the updated row in place;
 *   items  (equal ranges of the item stream): sums coefficient * staged user row in registers and writes
 *          the updated item row in place; the last block applies the global bias step and publishes
 *          brs_step_out.
 * Rows whose segment crosses a range boundary (Zipf head) combine their partial sums through table.grad
 * (row-major [capacity][dim] here) and the last part to arrive applies the update.
 * Adam / RMSprop in BRS_DENSE mode additionally sweep the untouched rows with g = 0 (reference-exact).
 * On an out-of-range index the step publishes status 1 and leaves every parameter untouched.
 */
int64_t brs_mf_plan_bytes(int64_t batch_capacity, int32_t user_capacity, int32_t item_capacity);
/* which = 0/1 selects {user.rows,item.rows,plan[0]} or {user_rows_alt,item_rows_alt,plan[1]};
 * loss_kind 0 = bpr (third = neg item ids, int64), 1 = bce (third = ratings, float) */
int brs_mf_plan_build(const brs_mf_model *model, int32_t which, int32_t loss_kind, const int64_t *users,
                      const int64_t *items, const void *third, int64_t batch, void *stream);
int brs_mf_step_planned(const brs_mf_model *model, int32_t which, const brs_opt *opt, int32_t loss_kind,
                        int64_t batch, float reg_weight, float *out /* brs_step_out */, void *stream);
/* plan_build(which = 0) + step_planned on one stream: MFEngine.train_single_batch */
int brs_mf_step(const brs_mf_model *model, const brs_opt *opt, int32_t loss_kind, const int64_t *users,
                const int64_t *items, const void *third, int64_t batch, float reg_weight,
                float *out /* brs_step_out */, void *stream);

/* diagnostics (per-kernel timing on a FIXED plan; leaves the step incomplete): 0 = both row kernels,
 * 1 = users kernel only, 2 = items kernel only */
int brs_debug_set_mf_rows_only(int which);

/* MF.predict / MF.forward under no_grad (beta_rec/models/mf.py:57-70): scores[k] = sigmoid(...) */
int brs_mf_predict(const brs_mf_model *model, const int64_t *users, const int64_t *items, int64_t n,
                   float *scores, void *stream);

/* ---- NCF family: beta_rec/models/{gmf,mlp,ncf}.py ---- */
#define BRS_NCF_GMF 0    /* GMF  : sigmoid(w.(u*i)+b)                                   gmf.py:29-36  */
#define BRS_NCF_MLP 1    /* MLP  : cat -> [Linear,ReLU]xn -> Linear -> sigmoid          mlp.py:40-51  */
#define BRS_NCF_NEUMF 2  /* NeuMF: relu(cat) -> [Linear,ReLU]xn ; cat with u_mf*i_mf -> Linear -> sigmoid  ncf.py:52-71 */
#define BRS_NCF_MAX_LAYERS 6

typedef struct brs_ncf_model {
    int32_t kind;      /* BRS_NCF_* */
    int32_t n_layers;  /* config["model"]["mlp_config"]["n_layers"]; 0 for GMF */
    int32_t emb_dim;   /* config["model"]["emb_dim"]: GMF / MF-part width and tower output width */
    int32_t mlp_dim;   /* emb_dim * 2^(n_layers-1): width of the MLP-side embedding rows; 0 for GMF */
    brs_entity user;   /* NeuMF: table[0] = embedding_user_mlp [U,mlp_dim], table[1] = embedding_user_mf [U,emb_dim];
                          MLP / GMF: table[0] = embedding_user */
    brs_entity item;   /* same for items */
    brs_dense_param fc_weight[BRS_NCF_MAX_LAYERS]; /* fc_layers.{3l+1}.weight [in_l/2, in_l], in_l = 2*mlp_dim >> l */
    brs_dense_param fc_bias[BRS_NCF_MAX_LAYERS];
    float *fc_weight_t[BRS_NCF_MAX_LAYERS];        /* scratch [in_l, in_l/2] for W^T (dgrad on tensor cores) or NULL */
    brs_dense_param out_weight;                    /* affine_output.weight [1, emb_dim (GMF, MLP) | 2*emb_dim (NeuMF)] */
    brs_dense_param out_bias;                      /* affine_output.bias [1] */
    float *act[BRS_NCF_MAX_LAYERS + 1];   /* act[0] = tower input [max_batch, 2*mlp_dim], act[l] = output of layer l */
    float *dact[BRS_NCF_MAX_LAYERS + 1];  /* gradients of the same shapes */
    float *mfv;                           /* [max_batch, emb_dim] u_mf * i_mf (NeuMF) */
    float *dz;                            /* [max_batch] d loss / d logit */
    int64_t max_batch;
    void *ws;                             /* BRS_STEP_WS_BYTES */
} brs_ncf_model;

/* forward + BCELoss + backward of one batch: gradients of the embedding rows go to table.grad
 * (touched rows recorded), gradients of the Linear layers to their brs_dense_param.grad
 * (replaces NeuMF.forward / GMF.forward / MLP.forward + nn.BCELoss + loss.backward():
 * ncf.py:109-117, gmf.py:69-77, mlp.py:87-95) */
int brs_ncf_fwd_bwd(const brs_ncf_model *model, const int64_t *users, const int64_t *items, const float *ratings,
                    int64_t batch, void *stream);
/* optimizer.step() over the embedding tables and the Linear parameters + loss.item() (ncf.py:118-119) */
int brs_ncf_apply(const brs_ncf_model *model, const brs_opt *opt, int64_t batch, float *out /* brs_step_out */,
                  void *stream);
/* NeuMF.predict / GMF.predict / MLP.predict (ncf.py:73-78): sigmoid scores, chunked by max_batch */
int brs_ncf_predict(const brs_ncf_model *model, const int64_t *users, const int64_t *items, int64_t n, float *scores,
                    void *stream);
/* train_an_epoch's inner loop (ncf.py:132-138) over (user,item,rating) arrays resident in HBM */
int brs_ncf_train_batches(const brs_ncf_model *model, const brs_opt *opt, const int64_t *users, const int64_t *items,
                          const float *ratings, int64_t n, int64_t batch, float *out /* brs_step_out[] */,
                          void *stream);

/* the Linear building blocks of the tower (nn.Linear forward / backward), row-major fp32:
 *   fwd: y[m,n] = (relu)(x[m,k] . w[n,k]^T + b[n])
 *   bwd: dw[n,k] += dy^T x ; db[n] += colsum(dy) (db may be NULL);
 *        dx[m,k] = (dy . w) * (relu_mask_src[m,k] > 0)   (dx may be NULL; mask source may be NULL) */
/* same forward on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with 3xTF32 error compensation
 * (hi.hi + hi.lo + lo.hi accumulated in TMEM), fused bias / ReLU / keep-where-positive mask.
 * Shapes: k % 32 == 0, n % 64 == 0; BRS_ERR_UNSUPPORTED otherwise (callers fall back to brs_mlp_fwd). */
int brs_mlp_fwd_tc(const float *x, const float *w, const float *b, float *y, const float *keep_where_positive,
                   int64_t m, int32_t n, int32_t k, int32_t relu, void *stream);
/* 0 = exact-fp32 FFMA Linear kernels, 1 = tcgen05 3xTF32 where the shape allows (default) */
int brs_set_gemm_backend(int backend);
int brs_mlp_fwd(const float *x, const float *w, const float *b, float *y, int64_t m, int32_t n, int32_t k,
                int32_t relu, void *stream);
int brs_mlp_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *db,
                const float *relu_mask_src, int64_t m, int32_t n, int32_t k, void *stream);

/* ---- LightGCN: beta_rec/models/lightgcn.py ---- */
#define BRS_LGCN_MAX_LAYERS 6

/* CSR sparse matrix; nnz < 2^31.  edge_id (optional) maps each stored non-zero back to its position in the
 * coalesced (row-major) edge list of the ORIGINAL matrix -- used by the transpose so that it consumes the
 * same edge-dropout keep mask as the forward matrix */
typedef struct brs_csr {
    const int32_t *row_ptr;  /* [n_rows + 1] */
    const int32_t *col;      /* [nnz] */
    const float *val;        /* [nnz] */
    const int32_t *edge_id;  /* [nnz] or NULL (identity) */
    int64_t n_rows, nnz;
} brs_csr;

/* y[r, :] += sum over kept non-zeros e of row r of (val[e] / keep_prob) * x[col[e], :]
 * (torch.sparse.mm at lightgcn.py:73 with LightGCN.dropout's mask/rescale at :32-36 folded in).
 * keep_mask: uint8 per ORIGINAL edge, or NULL for no dropout.  y must be pre-initialised. */
int brs_spmm_csr(const brs_csr *a, const uint8_t *keep_mask, float keep_prob, const float *x, float *y, int32_t dim,
                 void *stream);

typedef struct brs_lightgcn_model {
    int64_t n_users, n_items;
    int32_t dim, n_layers;
    float decay;  /* regs[0] (lightgcn.py:113) */
    float pad_;
    brs_csr adj;    /* norm_adj = D^-1 (A + I), [N, N], N = n_users + n_items (row-normalised => asymmetric) */
    brs_csr adj_t;  /* its transpose in CSR, edge_id -> edge of adj */
    float *emb[BRS_LGCN_MAX_LAYERS + 1]; /* emb[0] = the parameters [N, dim] (users first, then items:
                                            torch.cat at lightgcn.py:55-57); emb[l] = layer-l buffer */
    float *d;         /* [N, dim] scratch: d loss / d E^(l) */
    float *g[2];      /* [N, dim] x2 scratch for the backward chain */
    brs_dense_param param;  /* weight = emb[0]; grad [N, dim] receives d loss / d E^(0); m, v for Adam */
    void *ws;
} brs_lightgcn_model;

/* LightGCN.forward's propagate (lightgcn.py:46-74): fills emb[1..L] */
int brs_lightgcn_propagate(const brs_lightgcn_model *model, const uint8_t *keep_mask, float keep_prob, void *stream);
/* train_single_batch minus the optimizer (lightgcn.py:119-149): propagate, softplus-BPR + L2 tail, backward
 * through the propagate; the dense gradient of the embeddings lands in param.grad */
int brs_lightgcn_fwd_bwd(const brs_lightgcn_model *model, const uint8_t *keep_mask, float keep_prob,
                         const int64_t *users, const int64_t *pos_items, const int64_t *neg_items, int64_t batch,
                         void *stream);
/* The two non-SpMM pieces of brs_lightgcn_fwd_bwd, for the row-partitioned multi-GPU step (the propagate and its
 * backward then run as per-rank block products with brs_spmm_csr and collectives in between):
 * brs_lightgcn_tail: softplus-BPR + L2 tail of `batch` samples on the layer buffers emb[0..L] (lightgcn.py:171-191),
 *   scaled for a mean over global_batch; clears d [N, dim] and scatters d loss / d E^(l) into it; loss sum -> ws;
 * brs_lightgcn_reg_grad: the L2 term's gradient on the layer-0 rows of the batch, restricted to node rows
 *   [row_lo, row_hi) and added into grad_rows [(row_hi - row_lo), dim] (row 0 = node row_lo). */
int brs_lightgcn_tail(const brs_lightgcn_model *model, const int64_t *users, const int64_t *pos_items,
                      const int64_t *neg_items, int64_t batch, int64_t global_batch, void *stream);
int brs_lightgcn_reg_grad(const brs_lightgcn_model *model, const int64_t *users, const int64_t *pos_items,
                          const int64_t *neg_items, int64_t batch, int64_t global_batch, int64_t row_lo, int64_t row_hi,
                          float *grad_rows, void *stream);
/* optimizer.step() over all rows + loss.item() (lightgcn.py:150-151) */
int brs_lightgcn_apply(const brs_lightgcn_model *model, const brs_opt *opt, int64_t batch,
                       float *out /* brs_step_out */, void *stream);
/* sigmoid(ebar_u . ebar_i) from the current layer buffers (LightGCN.predict, lightgcn.py:80-101) */
int brs_lightgcn_scores(const brs_lightgcn_model *model, const int64_t *users, const int64_t *items, int64_t n,
                        float *scores, void *stream);

/* ---- generic sparse-embedding-gradient building blocks (K5/K6 in SURVEY.md section 2b) ---- */
/* range-check idx[0..n) against rows->n_rows and give every distinct row a slot in the entity's
 * compact gradient scratch (ws: BRS_STEP_WS_BYTES of zeroed device scratch; ws.err_flag collects errors) */
int brs_rows_assign(const brs_rowset *rows, const int64_t *idx, int64_t n, void *ws, void *stream);
/* grad_scratch[slot(idx[k])] += scale * src[k, :]   (src is [n, dim] of entity->table[table]) */
int brs_rows_scatter_grad(const brs_entity *entity, int32_t table, const int64_t *idx, int64_t n, const float *src,
                          float scale, void *stream);
/* out[k, :] = grad_scratch[slot(idx[k])] (0 for rows without a slot): reads the accumulated gradient rows back
 * without applying them -- the row-sharded NeuMF step ships them to the rows' owners, and the parity tests
 * compare gradients, not only updated parameters */
int brs_rows_read_grad(const brs_entity *entity, int32_t table, const int64_t *idx, int64_t n, float *out,
                       void *stream);
/* touched rows only: p -= lr*g (exactly what torch.optim.SGD does, g = 0 elsewhere) */
int brs_rows_sgd(const brs_entity *entities, int32_t n_entities, double lr, void *stream);
/* touched rows only, Adam with explicit step number t (1-based) */
int brs_rows_adam(const brs_entity *entities, int32_t n_entities, const brs_opt *opt, int64_t t, void *stream);
/* every row of every table: reference-exact Adam / RMSprop (g = 0 for untouched rows) */
int brs_dense_adam_sweep(const brs_entity *entities, int32_t n_entities, const brs_opt *opt, int64_t t,
                         void *stream);
/* dense parameters (Linear layers, global bias); clears their grads */
int brs_dense_params_step(const brs_dense_param *params, int32_t n_params, const brs_opt *opt, int64_t t,
                          void *stream);

/* ---- multi-GPU: row-sharded tables over NVLink peer memory (new work, the
 *      reference has no distributed path) ----
 * owner(row) = row mod world, local row = row div world (world a power of two <= 8).  Every rank maps
 * every other rank's shard (CUDA IPC) and the kernels address it directly:
 *   - the fused kernel gathers embedding rows with peer loads (or, mode 2, from local staging tables that a
 *     pull kernel filled with every UNIQUE row of the batch once) and accumulates the batch's gradients
 *     in a LOCAL compact scratch (slots over GLOBAL ids) -- no remote atomics per sample;
 *   - a push kernel then sends ONE coalesced row per unique touched row to its OWNER with 128-bit peer REDs:
 *     SGD adds -lr*g into the owner's weights directly; Adam/RMSprop add g into the owner's dense per-shard
 *     gradient table and set the row's bit in the owner's touched bitmap (red.or), and after a flag barrier
 *     every owner applies the optimizer to the rows whose bit is set.
 * Two flag barriers per step, no NCCL on the data path. */
#define BRS_MAX_RANKS 8
#define BRS_IPC_HANDLE_BYTES 64

/* device memory that can be exported to the other ranks of the node (cudaMalloc'ed, zero-filled) */
int brs_shm_alloc(int64_t bytes, void **ptr);
int brs_shm_free(void *ptr);
int brs_ipc_get_handle(void *ptr, uint8_t handle[BRS_IPC_HANDLE_BYTES]);
int brs_ipc_open_handle(const uint8_t handle[BRS_IPC_HANDLE_BYTES], void **ptr); /* enables peer access */
int brs_ipc_close_handle(void *ptr);

typedef struct brs_peer_sync {
    int32_t world, rank;
    uint64_t *flags[BRS_MAX_RANKS];   /* flags[r]    = rank r's uint64[BRS_MAX_RANKS] arrival flags (zeroed) */
    double *partials[BRS_MAX_RANKS];  /* partials[r] = rank r's double[BRS_MAX_RANKS][4] step-sum mailbox   */
} brs_peer_sync;

/* All ranks call this with the same, strictly increasing `epoch` (1, 2, ...).  Returns (stream-ordered)
 * once every rank has arrived; writes issued by earlier kernels of every rank (incl. peer REDs) are
 * visible afterwards.  With ws != NULL the ranks also exchange their step sums {loss, regularizer,
 * d global_bias}: each rank's ws ends up holding the sums over ALL ranks, added in rank order (so
 * replicated parameters stay bit-identical). */
int brs_peer_barrier(const brs_peer_sync *sync, uint64_t epoch, void *ws, void *stream);

typedef struct brs_mf_sharded {
    int32_t world, rank;
    int64_t n_users, n_items;            /* GLOBAL row counts */
    int64_t local_users, local_items;    /* rows reserved per shard: ceil(N / world) */
    /* this rank's staging + shard state: rowsets (slot maps) are indexed by GLOBAL ids, table[k].grad is the
     * local compact scratch; table[k].weight / m / v are this rank's SHARD tables (owner-side optimizer) */
    brs_mf_model stage;
    const brs_mf_peer_tables *peers;     /* DEVICE array [world]; peers[rank] = own shard */
    brs_mf_peer_tables own;              /* host copy of peers[rank] (dense gradients + bitmaps of this shard) */
    /* staging tables for the batch's unique rows, one row per slot of stage.{user,item}.rows
     * ([capacity, dim] row-major / [capacity]); filled by the pull kernel each step */
    float *pull_user_emb, *pull_item_emb, *pull_user_bias, *pull_item_bias;
} brs_mf_sharded;

/* forward + backward of this rank's `batch` BPR triples (GLOBAL ids) against the sharded tables: the
 * pre-pass gives every unique row a slot; rows are read either per sample with peer loads inside the fused
 * kernel, or (mode 2) a pull kernel copies each unique row ONCE from its owner into the staging tables and
 * the fused kernel runs on those; the batch's gradients are aggregated in the LOCAL compact scratch; the loss is the mean over `global_batch` = sum of all ranks'
 * batches. */
int brs_mf_sharded_bpr_fwd_bwd(const brs_mf_sharded *model, const int64_t *users, const int64_t *pos_items,
                               const int64_t *neg_items, int64_t batch, int64_t global_batch, float reg_weight,
                               void *stream);
/* how brs_mf_sharded_bpr_fwd_bwd reads remote rows: 1 = per-sample peer loads inside the fused kernel,
 * 2 = pull every unique row once into the staging tables, then compute locally, 0 = by world size
 * (default: 2 from 8 ranks up, else 1).  Also BRS_SHARD_MODE in the environment. */
int brs_debug_set_shard_mode(int mode);
/* push ONE coalesced row per unique touched row to its owner (128-bit peer REDs) and release the local slots.
 *   opt->kind == BRS_SGD: adds -lr * g straight into the owner's weight rows (the update is linear in g, so
 *     no owner-side pass is needed) -- every rank must have finished gathering: barrier BEFORE this call;
 *   otherwise: adds g into the owner's dense per-shard gradient table and sets the touched bit -- barrier
 *     AFTER this call, then brs_mf_sharded_apply. */
int brs_mf_sharded_push(const brs_mf_sharded *model, const brs_opt *opt, void *stream);
/* owner side: Adam / RMSprop on this rank's shard for the rows whose touched bit is set (BRS_DENSE: every
 * row, g = 0 where the bit is clear); for every optimizer the replicated global_bias step and the
 * publication of {loss, regularizer, status}.  Needs the step sums exchanged (brs_peer_barrier with ws). */
int brs_mf_sharded_apply(const brs_mf_sharded *model, const brs_opt *opt, int64_t global_batch,
                         float *out /* brs_step_out */, void *stream);
/* one whole step in the right order for opt->kind, with two flag barriers (epochs `epoch`, `epoch`+1):
 *   SGD : fwd_bwd, barrier(+sums), push (in-place), apply (bias/record), barrier
 *   else: fwd_bwd, push, barrier(+sums), apply, barrier */
int brs_mf_sharded_step(const brs_mf_sharded *model, const brs_peer_sync *sync, const brs_opt *opt,
                        const int64_t *users, const int64_t *pos_items, const int64_t *neg_items, int64_t batch,
                        int64_t global_batch, float reg_weight, uint64_t epoch, float *out /* brs_step_out */,
                        void *stream);

/* the sharded epoch inner loop over this rank's index arrays resident in HBM (same n and batch on every rank):
 * brs_mf_sharded_step per batch; barrier epochs first_epoch, first_epoch+1, ... (2 per batch); out = brs_step_out[ceil(n/batch)] */
int brs_mf_sharded_train_batches(const brs_mf_sharded *model, const brs_peer_sync *sync, const brs_opt *opt,
                                 const int64_t *users, const int64_t *pos_items, const int64_t *neg_items, int64_t n,
                                 int64_t batch, int64_t global_batch, float reg_weight, uint64_t first_epoch,
                                 float *out /* brs_step_out[] */, void *stream);

/* the same loop with this rank's index arrays in HOST memory (pinned for full speed) and h_out on the host:
 * batches are streamed through the device ring of brs_mf_train_batches_host; returns after the last step
 * completed.  Every rank must call it with the same n / batch / first_epoch. */
int brs_mf_sharded_train_batches_host(const brs_mf_sharded *model, const brs_peer_sync *sync, const brs_opt *opt,
                                      const int64_t *h_users, const int64_t *h_pos_items, const int64_t *h_neg_items,
                                      int64_t n, int64_t batch, int64_t global_batch, float reg_weight,
                                      uint64_t first_epoch, float *h_out /* brs_step_out[] */, void *stream);

/* stable bucket of (u,i,j) triples by owner(u) = u mod world (the send buffer of the NCCL all-to-all that
 * routes triples to the user-row owner): out_* hold the triples grouped by destination, original order
 * kept inside a group; counts[world] (device int64) receives the group sizes */
int brs_route_triples(const int64_t *users, const int64_t *pos_items, const int64_t *neg_items, int64_t n,
                      int32_t world, int64_t *out_users, int64_t *out_pos, int64_t *out_neg, int64_t *counts,
                      void *stream);

/* ---- embedding gather / scatter-add micro-ops (BASELINE.json config 5) ---- */
int brs_gather(const float *table, int64_t n_rows, int32_t dim, const int64_t *idx, int64_t n, float *out,
               void *stream);
int brs_scatter_add(float *table, int64_t n_rows, int32_t dim, const int64_t *idx, int64_t n, const float *src,
                    float scale, void *stream);
/* table[idx[k]] -= lr * src[k] after gathering: gather + SGD update micro-op */
int brs_gather_sgd_update(float *table, int64_t n_rows, int32_t dim, const int64_t *idx, int64_t n, float lr,
                          float *out, void *stream);

/* ---- ranking evaluation: beta_rec/core/eval_engine.py:49-87 (evaluate), beta_rec/utils/evaluation.py:459-785 ----
 * ndcg / map / precision / recall at k of `n_pred` prediction rows (user, item, score) against `n_true`
 * true rows (user, item, rating; only rating >= 1 counts, evaluation.py:492), as the reference's pandas code
 * defines them: per user the k highest scores (equal scores keep the frame's row order: nlargest
 * keep="first" + rank(method="first")), hits = top-k rows whose (user, item) is a true row, users = users with
 * at least one true row and one prediction row.  User ids lie in [0, n_user_ids), item ids in [0, 2^32).
 * out[8] (double, device): {sum_u ndcg_u, sum_u ap_u, sum_u hit_u / k, sum_u hit_u / actual_u, users, hits,
 * status (1 = an id was out of range), 0}; the caller divides the sums by `users` (0.0 when hits == 0,
 * evaluation.py:627).  workspace: brs_rank_metrics_workspace_bytes(...) bytes of device memory, any content.
 * 1 <= k <= 1024.  Stream-ordered, no host synchronisation. */
int64_t brs_rank_metrics_workspace_bytes(int64_t n_true, int64_t n_pred, int64_t n_user_ids);
int brs_rank_metrics(const int64_t *true_users, const int64_t *true_items, const float *true_ratings, int64_t n_true,
                     const int64_t *pred_users, const int64_t *pred_items, const float *pred_scores, int64_t n_pred,
                     int64_t n_user_ids, int32_t k, void *workspace, int64_t workspace_bytes, double *out,
                     void *stream);

/* ---- negative sampling: BaseData.instance_bpr_loader / instance_bce_loader (beta_rec/data/base_data.py:182-253) ----
 * brs_pairset_build puts the training interactions (user, item) into a device hash set (`set`:
 * brs_pairset_bytes(n) bytes).  brs_sample_negatives then writes, for every row r of `users`, `num_negative`
 * pairwise distinct items drawn uniformly from the items users[r] has NOT interacted with (the reference's
 * `random.sample(set(item_id_pool) - positive_items, num_negative)`), by rejection on the counter-based stream
 *     candidate(r, t, attempt) = mix64(seed + r*C1 + t*C2 + attempt*C3) mod n_items
 * (splitmix64 finaliser; constants in csrc/sample_kernels.cu), a pure function of its arguments.
 * neg_out: int64 [n, num_negative], 1 <= num_negative <= 64.  brs_pairset_status (synchronises) reports
 * bit 0 = an interaction outside [0, n_users) x [0, n_items), bit 1 = a user had no item left. */
int64_t brs_pairset_bytes(int64_t n_pairs);
int brs_pairset_build(const int64_t *users, const int64_t *items, int64_t n, int64_t n_users, int64_t n_items, void *set,
                      int64_t set_bytes, void *stream);
int brs_sample_negatives(const void *set, int64_t n_pairs_in_set, const int64_t *users, int64_t n, int64_t n_items,
                         int32_t num_negative, uint64_t seed, int64_t *neg_out, void *stream);
int brs_pairset_status(const void *set, uint32_t *status_out, void *stream);

/* ---- normalised adjacency: BaseData.create_adj_mat (beta_rec/data/base_data.py:337-360) + normalized_adj_single
 * (beta_rec/utils/common_util.py:24-41) ----
 * Interactions (users[e], items[e]) -> CSR of norm_adj = D^-1 (A + I) (self_loops = 1) or mean_adj = D^-1 A
 * (self_loops = 0) over the n_users + n_items nodes, A = [[0, R], [R^T, 0]], R[u, i] = 1 (duplicates collapse);
 * rows and columns in coalesced order, values fp32(1.0 / rowsum) with the division in float64 like the reference.
 * The pattern is symmetric, so the transpose shares row_ptr / col: val_t holds its values and edge_id_t maps each
 * transposed non-zero to its forward edge (what brs_spmm_csr's backward and the dropout mask need).
 * Outputs (device): row_ptr int32 [n + 1]; col, val, val_t, edge_id_t sized for the bound 2 * n_interactions + n;
 * nnz_out int64 [1] = entries actually written.  2 * n_interactions + n < 2^31.  Sorting is cub::DeviceRadixSort
 * (library plumbing; this runs once per training run).  brs_adj_status (synchronises): 1 = an id out of range. */
int64_t brs_adj_workspace_bytes(int64_t n_interactions, int64_t n_users, int64_t n_items);
int brs_adj_build(const int64_t *users, const int64_t *items, int64_t n_interactions, int64_t n_users, int64_t n_items,
                  int32_t self_loops, void *workspace, int64_t workspace_bytes, int32_t *row_ptr, int32_t *col, float *val,
                  float *val_t, int32_t *edge_id_t, int64_t *nnz_out, void *stream);
int brs_adj_status(const void *workspace, int64_t n_interactions, int64_t n_users, int64_t n_items, uint32_t *status_out,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BRS_B200_H */
