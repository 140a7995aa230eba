// Library-level C ABI: status strings, device info, MF step composition, the
// epoch loop over HBM-resident batches, and the gather / scatter-add micro-ops
// of BASELINE.json config 5.  See include/brs_b200.h for the contract.
#include <mutex>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

int brs_mf_fwd_bwd_impl(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                        const void* third, int64_t batch, float reg_weight, void* stream);
int brs_apply_impl(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                   int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                   long long batch, long long max_rows_hint, void* stream);
int brs_apply_impl_next(const brs_entity* ents, int n_ent, const brs_dense_param* dense, int n_dense,
                        int dense_grad_from_ws, const brs_opt* opt, void* ws, long long t_explicit, float* out,
                        long long batch, long long max_rows_hint, void* stream, const brs_rowset* next_rs,
                        const long long* const* next_idx, const long long* next_n, int next_arrays, int parity);
int brs_mf_fwd_bwd_phases(const brs_mf_model* model, int loss_kind, const int64_t* users, const int64_t* items,
                          const void* third, int64_t batch, float reg_weight, void* stream, int phases);

namespace {
char g_cuda_err[512] = "";
std::mutex g_err_mu;
constexpr int kMaxDevices = 64;
int g_sm_count[kMaxDevices] = {};
}  // namespace

void brs_set_cuda_error(cudaError_t e, const char* what, int line) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), what, line);
    (void)cudaGetLastError();
}

int brs_sm_count() {  // of the CURRENT device (cached per device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        (void)cudaGetLastError();
        return 148;  // B200
    }
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count[dev] = n;
        else
            g_sm_count[dev] = 148;
    }
    return g_sm_count[dev];
}

namespace {
// defaults from the round-1 sweep (tools/sweep_mf.py l2, profiles/r01_sweeps.md): streamed weight rows
// are demoted, the compact gradient scratch is pinned -> fused kernel 58 -> 50 us at config 2
brs_l2_policy_cfg g_l2 = {BRS_L2_EVICT_FIRST, BRS_L2_EVICT_LAST, BRS_L2_EVICT_FIRST};
}
const brs_l2_policy_cfg& brs_l2_cfg() { return g_l2; }

extern "C" int brs_debug_set_l2_policy(int gather, int scratch, int weight) {
    if (gather < 0 || gather > 2 || scratch < 0 || scratch > 2 || weight < 0 || weight > 2) return BRS_ERR_INVALID_ARG;
    g_l2.gather = gather;
    g_l2.scratch = scratch;
    g_l2.weight = weight;
    return BRS_OK;
}

extern "C" int brs_abi_version(void) { return BRS_ABI_VERSION; }

extern "C" const char* brs_strerror(int status) {
    switch (status) {
        case BRS_OK: return "ok";
        case BRS_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size or misaligned table)";
        case BRS_ERR_UNSUPPORTED: return "unsupported dimension / optimizer / shape";
        case BRS_ERR_CUDA: return "CUDA runtime error (see brs_last_cuda_error)";
        case BRS_ERR_INDEX_RANGE: return "index out of range";
        case BRS_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

extern "C" const char* brs_last_cuda_error(void) { return g_cuda_err; }

extern "C" int brs_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        (void)cudaGetLastError();
        return BRS_ERR_NO_DEVICE;
    }
    int v = 0;
    if (sm_count) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        *sm_count = v;
    }
    if (cc_major) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
        *cc_major = v;
    }
    if (cc_minor) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
        *cc_minor = v;
    }
    if (l2_bytes) {
        BRS_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
        *l2_bytes = v;
    }
    return BRS_OK;
}

// ---------------------------------------------------------------------------
// MF step composition
// ---------------------------------------------------------------------------
extern "C" int brs_mf_apply(const brs_mf_model* model, const brs_opt* opt, int64_t batch, float* out_loss_reg,
                            void* stream) {
    if (!model || !opt || !model->ws || batch < 0) return BRS_ERR_INVALID_ARG;
    brs_entity ents[2] = {model->user, model->item};
    // a batch touches at most `batch` user rows and 2*batch item rows
    const long long hint = 3 * (long long)batch;
    return brs_apply_impl(ents, 2, &model->global_bias, 1, /*dense_grad_from_ws=*/1, opt, model->ws, 0, out_loss_reg,
                          batch, hint, stream);
}

namespace {
// where the epoch loop finds batch b's index arrays on the device, and what has to happen around it
struct BatchFeed {
    virtual ~BatchFeed() {}
    virtual void ptrs(int64_t b, const int64_t** users, const int64_t** items, const void** third) = 0;
    virtual int before_read(int64_t b, cudaStream_t st) { return BRS_OK; }  // batch b is about to be read on st
    virtual int after_step(int64_t b, float* d_rec, cudaStream_t st) { return BRS_OK; }  // step b fully enqueued
};

struct ResidentFeed : BatchFeed {  // the whole epoch's arrays already live in HBM
    const int64_t *users, *items;
    const char* third;
    int64_t batch;
    size_t third_sz;
    void ptrs(int64_t b, const int64_t** u, const int64_t** i, const void** t) override {
        *u = users + b * batch;
        *i = items + b * batch;
        *t = third + (size_t)(b * batch) * third_sz;
    }
};

// Per batch: fused fwd/bwd + apply, 2 launches when the slot pre-pass of batch b+1 can ride in the apply
// launch of batch b (alternate rowsets + touched-rows optimizer), else the plain 3-launch sequence.
// side stream + events of the row-owner epoch loop, one set per device
struct PlanPipe {
    cudaStream_t side = nullptr;
    cudaEvent_t planned[2] = {}, freed[2] = {};
    int init() {
        if (side) return BRS_OK;
        BRS_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            BRS_CUDA_CHECK(cudaEventCreateWithFlags(&planned[k], cudaEventDisableTiming));
            BRS_CUDA_CHECK(cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming));
        }
        return BRS_OK;
    }
};
PlanPipe g_plan_pipe[kMaxDevices];
std::mutex g_plan_pipe_mu;

// Row-owner loop (mf_rowwise.cu): the index plan of batch b+1 (claim / scan / fill: reads only the index
// arrays) is built on a side stream while batch b's two row kernels run on `stream`; plans and rowsets
// alternate, events order the two streams.  With a single plan everything is issued on `stream`.
int mf_epoch_loop_rowwise(const brs_mf_model* model, const brs_opt* opt, int loss_kind, BatchFeed& feed, int64_t n,
                          int64_t batch, float reg_weight, float* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb = (n + batch - 1) / batch;
    auto size_of = [&](int64_t b) { return (b == nb - 1) ? (n - b * batch) : batch; };
    const int64_t *u, *i;
    const void* t;
    int rc;
    const bool two = model->plan[1].buf && model->user_rows_alt.slot_map && model->item_rows_alt.slot_map;
    if (!two) {
        for (int64_t b = 0; b < nb; ++b) {
            if ((rc = feed.before_read(b, st)) != BRS_OK) return rc;
            feed.ptrs(b, &u, &i, &t);
            if ((rc = brs_mf_step(model, opt, loss_kind, u, i, t, size_of(b), reg_weight, d_out + 4 * b, stream)) != BRS_OK)
                return rc;
            if ((rc = feed.after_step(b, d_out + 4 * b, st)) != BRS_OK) return rc;
        }
        return BRS_OK;
    }
    int dev = 0;
    BRS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return BRS_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_plan_pipe_mu);
    PlanPipe& pp = g_plan_pipe[dev];
    if ((rc = pp.init()) != BRS_OK) return rc;
    // whatever `stream` did before (previous steps released the rowsets) precedes the first plan
    BRS_CUDA_CHECK(cudaEventRecord(pp.freed[0], st));
    BRS_CUDA_CHECK(cudaEventRecord(pp.freed[1], st));
    auto plan = [&](int64_t b) -> int {
        const int k = (int)(b & 1);
        BRS_CUDA_CHECK(cudaStreamWaitEvent(pp.side, pp.freed[k], 0));
        int r = feed.before_read(b, pp.side);
        if (r != BRS_OK) return r;
        feed.ptrs(b, &u, &i, &t);
        r = brs_mf_plan_build(model, k, loss_kind, u, i, t, size_of(b), pp.side);
        if (r != BRS_OK) return r;
        BRS_CUDA_CHECK(cudaEventRecord(pp.planned[k], pp.side));
        return BRS_OK;
    };
    if (nb > 0 && (rc = plan(0)) != BRS_OK) return rc;
    for (int64_t b = 0; b < nb; ++b) {
        const int k = (int)(b & 1);
        if (b + 1 < nb && (rc = plan(b + 1)) != BRS_OK) break;
        BRS_CUDA_CHECK(cudaStreamWaitEvent(st, pp.planned[k], 0));
        if ((rc = brs_mf_step_planned(model, k, opt, loss_kind, size_of(b), reg_weight, d_out + 4 * b, stream)) != BRS_OK) break;
        BRS_CUDA_CHECK(cudaEventRecord(pp.freed[k], st));
        if ((rc = feed.after_step(b, d_out + 4 * b, st)) != BRS_OK) break;
    }
    // the side stream's work is ordered before anything issued on `stream` afterwards
    cudaEvent_t& last = pp.planned[0];
    if (cudaEventRecord(last, pp.side) == cudaSuccess) (void)cudaStreamWaitEvent(st, last, 0);
    return rc;
}

int mf_epoch_loop(const brs_mf_model* model, const brs_opt* opt, int loss_kind, BatchFeed& feed, int64_t n, int64_t batch,
                  float reg_weight, float* d_out, void* stream) {
    if (model->plan[0].buf && model->user_stage)
        return mf_epoch_loop_rowwise(model, opt, loss_kind, feed, n, batch, reg_weight, d_out, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb = (n + batch - 1) / batch;
    auto size_of = [&](int64_t b) { return (b == nb - 1) ? (n - b * batch) : batch; };
    const bool overlap = model->user_rows_alt.slot_map && model->item_rows_alt.slot_map &&
                         (opt->kind == BRS_SGD || opt->mode == BRS_TOUCHED_ROWS);
    const int64_t *u, *i;
    const void* t;
    int rc;
    if (!overlap) {
        for (int64_t b = 0; b < nb; ++b) {
            if ((rc = feed.before_read(b, st)) != BRS_OK) return rc;
            feed.ptrs(b, &u, &i, &t);
            if ((rc = brs_mf_fwd_bwd_impl(model, loss_kind, u, i, t, size_of(b), reg_weight, stream)) != BRS_OK) return rc;
            if ((rc = brs_mf_apply(model, opt, size_of(b), d_out + 4 * b, stream)) != BRS_OK) return rc;
            if ((rc = feed.after_step(b, d_out + 4 * b, st)) != BRS_OK) return rc;
        }
        return BRS_OK;
    }
    brs_mf_model m[2] = {*model, *model};  // m[1] works on the alternate rowsets
    m[1].user.rows = model->user_rows_alt;
    m[1].item.rows = model->item_rows_alt;
    m[1].user_rows_alt = model->user.rows;
    m[1].item_rows_alt = model->item.rows;
    if (nb > 0) {  // pre-pass of batch 0 on its own
        if ((rc = feed.before_read(0, st)) != BRS_OK) return rc;
        feed.ptrs(0, &u, &i, &t);
        if ((rc = brs_mf_fwd_bwd_phases(&m[0], loss_kind, u, i, t, size_of(0), reg_weight, stream, 1)) != BRS_OK) return rc;
    }
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t cur = size_of(b);
        const brs_mf_model& cm = m[b & 1];
        feed.ptrs(b, &u, &i, &t);
        if ((rc = brs_mf_fwd_bwd_phases(&cm, loss_kind, u, i, t, cur, reg_weight, stream, 2)) != BRS_OK) return rc;
        brs_entity ents[2] = {cm.user, cm.item};
        if (b + 1 < nb) {
            if ((rc = feed.before_read(b + 1, st)) != BRS_OK) return rc;
            const int64_t ncur = size_of(b + 1);
            const brs_mf_model& nm = m[(b + 1) & 1];
            const int64_t *nu, *ni;
            const void* nt;
            feed.ptrs(b + 1, &nu, &ni, &nt);
            const brs_rowset rs[3] = {nm.user.rows, nm.item.rows, nm.item.rows};
            const long long* idx[3] = {(const long long*)nu, (const long long*)ni, (const long long*)nt};
            const long long nn[3] = {ncur, ncur, ncur};
            rc = brs_apply_impl_next(ents, 2, &cm.global_bias, 1, 1, opt, cm.ws, 0, d_out + 4 * b, cur, 3 * cur, stream,
                                     rs, idx, nn, loss_kind == 0 ? 3 : 2, (int)(b & 1));
        } else {
            rc = brs_apply_impl_next(ents, 2, &cm.global_bias, 1, 1, opt, cm.ws, 0, d_out + 4 * b, cur, 3 * cur, stream,
                                     nullptr, nullptr, nullptr, 0, (int)(b & 1));
        }
        if (rc != BRS_OK) return rc;
        if ((rc = feed.after_step(b, d_out + 4 * b, st)) != BRS_OK) return rc;
    }
    return BRS_OK;
}

// Epoch arrays in HOST memory: batch b+2 is copied to a 4-slot device ring on a private copy stream while
// batch b computes; each step's record is DMA'd back into pinned memory as soon as it is published.
struct HostFeed : BatchFeed {
    static constexpr int R = 4, AHEAD = 2;
    int device = -1;
    cudaStream_t copy = nullptr;
    cudaEvent_t ready[R] = {}, done[R] = {};
    char* ring = nullptr;       // R slots x 3 arrays x slot_elems x 8 bytes
    int64_t slot_elems = 0;
    float* d_out = nullptr;     // device records
    float* h_rec = nullptr;     // pinned host records
    int64_t rec_cap = 0;
    // per call
    const int64_t *h_users = nullptr, *h_items = nullptr;
    const char* h_third = nullptr;
    int64_t n = 0, batch = 0, nb = 0, issued = 0;
    size_t third_sz = 8;

    int prepare(int64_t batch_, int64_t nb_) {
        int dev = 0;
        BRS_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev != device) {  // first use on this device (the feed object itself is per device)
            device = dev;
            BRS_CUDA_CHECK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
            for (int k = 0; k < R; ++k) {
                BRS_CUDA_CHECK(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
                BRS_CUDA_CHECK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
            }
        }
        if (batch_ > slot_elems) {
            if (ring) BRS_CUDA_CHECK(cudaFree(ring));
            ring = nullptr;
            BRS_CUDA_CHECK(cudaMalloc(&ring, (size_t)R * 3 * (size_t)batch_ * 8));
            slot_elems = batch_;
        }
        if (nb_ > rec_cap) {
            if (d_out) BRS_CUDA_CHECK(cudaFree(d_out));
            if (h_rec) BRS_CUDA_CHECK(cudaFreeHost(h_rec));
            d_out = h_rec = nullptr;
            const int64_t cap = nb_ < 1024 ? 1024 : nb_;
            BRS_CUDA_CHECK(cudaMalloc(&d_out, (size_t)cap * 16));
            BRS_CUDA_CHECK(cudaMallocHost(&h_rec, (size_t)cap * 16));
            rec_cap = cap;
        }
        return BRS_OK;
    }
    char* slot(int64_t b, int arr) const { return ring + ((size_t)(b % R) * 3 + arr) * (size_t)slot_elems * 8; }
    int issue(int64_t b) {  // H2D of batch b on the copy stream
        const int k = (int)(b % R);
        if (b >= R) BRS_CUDA_CHECK(cudaStreamWaitEvent(copy, done[k], 0));  // batch b-R is done with the slot
        const int64_t off = b * batch, cur = (b == nb - 1) ? (n - off) : batch;
        BRS_CUDA_CHECK(cudaMemcpyAsync(slot(b, 0), h_users + off, (size_t)cur * 8, cudaMemcpyHostToDevice, copy));
        BRS_CUDA_CHECK(cudaMemcpyAsync(slot(b, 1), h_items + off, (size_t)cur * 8, cudaMemcpyHostToDevice, copy));
        BRS_CUDA_CHECK(cudaMemcpyAsync(slot(b, 2), h_third + (size_t)off * third_sz, (size_t)cur * third_sz,
                                       cudaMemcpyHostToDevice, copy));
        BRS_CUDA_CHECK(cudaEventRecord(ready[k], copy));
        return BRS_OK;
    }
    void ptrs(int64_t b, const int64_t** u, const int64_t** i, const void** t) override {
        *u = (const int64_t*)slot(b, 0);
        *i = (const int64_t*)slot(b, 1);
        *t = slot(b, 2);
    }
    int before_read(int64_t b, cudaStream_t st) override {
        while (issued < nb && issued <= b + AHEAD) {
            int rc = issue(issued);
            if (rc != BRS_OK) return rc;
            ++issued;
        }
        BRS_CUDA_CHECK(cudaStreamWaitEvent(st, ready[b % R], 0));
        return BRS_OK;
    }
    int after_step(int64_t b, float* d_rec, cudaStream_t st) override {
        BRS_CUDA_CHECK(cudaEventRecord(done[b % R], st));
        // the record goes home on the copy stream, so the compute stream never waits for a DMA
        BRS_CUDA_CHECK(cudaStreamWaitEvent(copy, done[b % R], 0));
        BRS_CUDA_CHECK(cudaMemcpyAsync(h_rec + 4 * b, d_rec, 16, cudaMemcpyDeviceToHost, copy));
        return BRS_OK;
    }
};
HostFeed g_host_feed[kMaxDevices];  // one ring / copy stream / record buffer per device
std::mutex g_host_feed_mu;
HostFeed* host_feed_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return &g_host_feed[dev];
}
}  // namespace

extern "C" int brs_mf_train_batches(const brs_mf_model* model, const brs_opt* opt, int32_t loss_kind,
                                    const int64_t* users, const int64_t* items, const void* third, int64_t n,
                                    int64_t batch, float reg_weight, float* out_loss_reg, void* stream) {
    if (!model || !opt || !users || !items || !third || n < 0 || batch <= 0 || !out_loss_reg) return BRS_ERR_INVALID_ARG;
    ResidentFeed feed;
    feed.users = users;
    feed.items = items;
    feed.third = (const char*)third;
    feed.batch = batch;
    feed.third_sz = loss_kind == 0 ? 8 : 4;
    return mf_epoch_loop(model, opt, loss_kind, feed, n, batch, reg_weight, out_loss_reg, stream);
}

extern "C" int brs_mf_train_batches_host(const brs_mf_model* model, const brs_opt* opt, int32_t loss_kind,
                                         const int64_t* h_users, const int64_t* h_items, const void* h_third,
                                         int64_t n, int64_t batch, float reg_weight, float* h_out_loss_reg,
                                         void* stream) {
    if (!model || !opt || !h_users || !h_items || !h_third || n < 0 || batch <= 0 || !h_out_loss_reg)
        return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    std::lock_guard<std::mutex> lk(g_host_feed_mu);
    HostFeed* fp = host_feed_for_current_device();
    if (!fp) return BRS_ERR_NO_DEVICE;
    HostFeed& f = *fp;
    const int64_t nb = (n + batch - 1) / batch;
    int rc = f.prepare(batch < n ? batch : n, nb);
    if (rc != BRS_OK) return rc;
    f.h_users = h_users;
    f.h_items = h_items;
    f.h_third = (const char*)h_third;
    f.n = n;
    f.batch = batch;
    f.nb = nb;
    f.issued = 0;
    f.third_sz = loss_kind == 0 ? 8 : 4;
    rc = mf_epoch_loop(model, opt, loss_kind, f, n, batch, reg_weight, f.d_out, stream);
    // drain both streams whatever happened: the ring and the pinned records are reused by the next call
    cudaError_t e1 = cudaStreamSynchronize((cudaStream_t)stream);
    cudaError_t e2 = cudaStreamSynchronize(f.copy);
    if (rc != BRS_OK) return rc;
    BRS_CUDA_CHECK(e1);
    BRS_CUDA_CHECK(e2);
    memcpy(h_out_loss_reg, f.h_rec, (size_t)nb * 16);
    return BRS_OK;
}

// the sharded epoch inner loop: every rank calls it with ITS OWN index arrays (same n / batch on all ranks)
extern "C" int brs_mf_sharded_step(const brs_mf_sharded* model, const brs_peer_sync* sync, const brs_opt* opt,
                                  const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                                  int64_t batch, int64_t global_batch, float reg_weight, uint64_t epoch, float* out,
                                  void* stream) {
    if (!model || !sync || !opt || !out || epoch == 0) return BRS_ERR_INVALID_ARG;
    int rc = brs_mf_sharded_bpr_fwd_bwd(model, users, pos_items, neg_items, batch, global_batch, reg_weight, stream);
    if (rc != BRS_OK) return rc;
    if (opt->kind == BRS_SGD) {
        // the push updates the owners' weights in place: every rank must be done gathering first
        rc = brs_peer_barrier(sync, epoch, model->stage.ws, stream);
        if (rc != BRS_OK) return rc;
        rc = brs_mf_sharded_push(model, opt, stream);
        if (rc != BRS_OK) return rc;
        rc = brs_mf_sharded_apply(model, opt, global_batch, out, stream);
        if (rc != BRS_OK) return rc;
        return brs_peer_barrier(sync, epoch + 1, nullptr, stream);
    }
    rc = brs_mf_sharded_push(model, opt, stream);
    if (rc != BRS_OK) return rc;
    rc = brs_peer_barrier(sync, epoch, model->stage.ws, stream);
    if (rc != BRS_OK) return rc;
    rc = brs_mf_sharded_apply(model, opt, global_batch, out, stream);
    if (rc != BRS_OK) return rc;
    return brs_peer_barrier(sync, epoch + 1, nullptr, stream);
}

extern "C" int brs_mf_sharded_train_batches(const brs_mf_sharded* model, const brs_peer_sync* sync, const brs_opt* opt,
                                            const int64_t* users, const int64_t* pos_items, const int64_t* neg_items,
                                            int64_t n, int64_t batch, int64_t global_batch, float reg_weight,
                                            uint64_t first_epoch, float* out, void* stream) {
    if (!model || !sync || !opt || !users || !pos_items || !neg_items || !out || n < 0 || batch <= 0 || first_epoch == 0)
        return BRS_ERR_INVALID_ARG;
    uint64_t epoch = first_epoch;
    int64_t b = 0;
    for (int64_t off = 0; off < n; off += batch, ++b, epoch += 2) {
        const int64_t cur = (n - off < batch) ? (n - off) : batch;
        const int64_t gb = (cur == batch) ? global_batch : cur * model->world;
        int rc = brs_mf_sharded_step(model, sync, opt, users + off, pos_items + off, neg_items + off, cur, gb, reg_weight,
                                     epoch, out + 4 * b, stream);
        if (rc != BRS_OK) return rc;
    }
    return BRS_OK;
}

// the sharded epoch loop fed from HOST memory: every rank streams ITS OWN index arrays through the ring
extern "C" int brs_mf_sharded_train_batches_host(const brs_mf_sharded* model, const brs_peer_sync* sync,
                                                 const brs_opt* opt, const int64_t* h_users, const int64_t* h_pos_items,
                                                 const int64_t* h_neg_items, int64_t n, int64_t batch,
                                                 int64_t global_batch, float reg_weight, uint64_t first_epoch,
                                                 float* h_out, void* stream) {
    if (!model || !sync || !opt || !h_users || !h_pos_items || !h_neg_items || !h_out || n < 0 || batch <= 0 ||
        first_epoch == 0)
        return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    std::lock_guard<std::mutex> lk(g_host_feed_mu);
    HostFeed* fp = host_feed_for_current_device();
    if (!fp) return BRS_ERR_NO_DEVICE;
    HostFeed& f = *fp;
    const int64_t nb = (n + batch - 1) / batch;
    int rc = f.prepare(batch < n ? batch : n, nb);
    if (rc != BRS_OK) return rc;
    f.h_users = h_users;
    f.h_items = h_pos_items;
    f.h_third = (const char*)h_neg_items;
    f.n = n;
    f.batch = batch;
    f.nb = nb;
    f.issued = 0;
    f.third_sz = 8;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t epoch = first_epoch;
    for (int64_t b = 0; b < nb && rc == BRS_OK; ++b, epoch += 2) {
        const int64_t cur = (b == nb - 1) ? (n - b * batch) : batch;
        const int64_t gb = (cur == batch) ? global_batch : cur * model->world;
        const int64_t *u, *i;
        const void* t;
        if ((rc = f.before_read(b, st)) != BRS_OK) break;
        f.ptrs(b, &u, &i, &t);
        rc = brs_mf_sharded_step(model, sync, opt, u, i, (const int64_t*)t, cur, gb, reg_weight, epoch, f.d_out + 4 * b,
                                 stream);
        if (rc == BRS_OK) rc = f.after_step(b, f.d_out + 4 * b, st);
    }
    cudaError_t e1 = cudaStreamSynchronize(st);
    cudaError_t e2 = cudaStreamSynchronize(f.copy);
    if (rc != BRS_OK) return rc;
    BRS_CUDA_CHECK(e1);
    BRS_CUDA_CHECK(e2);
    memcpy(h_out, f.h_rec, (size_t)nb * 16);
    return BRS_OK;
}

// ---------------------------------------------------------------------------
// gather / scatter-add micro-ops (config 5): warp-group per index, 128-bit accesses
// ---------------------------------------------------------------------------
namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

enum { OP_GATHER = 0, OP_SCATTER_ADD = 1, OP_GATHER_SGD = 2 };

template <int LPR, int VPL, int OP>
__global__ void __launch_bounds__(kThreads) rows_op_kernel(float* __restrict__ table, long long n_rows, int D,
                                                           const long long* __restrict__ idx, long long n,
                                                           float* __restrict__ buf, float scale) {
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPR, grp = lane / LPR;
    const long long warp_global = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * kWarps;
    for (long long base = warp_global * SPW; base < n; base += n_warps * SPW) {
        const long long s = base + grp;
        if (s >= n) continue;
        const long long r = idx[s];
        if ((unsigned long long)r >= (unsigned long long)n_rows) continue;
        float* row = table + r * D;
        float* b = buf + s * D;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int col = (v * LPR + gl) * 4;
            if (col < D) {
                if (OP == OP_GATHER) {
                    *(float4*)(b + col) = ld_row4(row + col);
                } else if (OP == OP_SCATTER_ADD) {
                    float4 x = *(const float4*)(b + col);
                    red_add4(row + col, make_float4(scale * x.x, scale * x.y, scale * x.z, scale * x.w));
                } else {  // gather, hand the row out, apply an SGD-style update in place
                    float4 x = ld_row4(row + col);
                    *(float4*)(b + col) = x;
                    red_add4(row + col, make_float4(-scale * x.x, -scale * x.y, -scale * x.z, -scale * x.w));
                }
            }
        }
    }
}

template <int OP>
int launch_rows_op(float* table, int64_t n_rows, int32_t D, const int64_t* idx, int64_t n, float* buf, float scale,
                   void* stream) {
    if (!table || !idx || !buf || n_rows < 0 || n < 0) return BRS_ERR_INVALID_ARG;
    if (D <= 0 || D % 4 != 0 || D > 512) return BRS_ERR_UNSUPPORTED;
    if ((((uintptr_t)table | (uintptr_t)buf) & 15) != 0) return BRS_ERR_INVALID_ARG;
    if (n == 0) return BRS_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define BRS_ROWS(LPR, VPL)                                                                                       \
    do {                                                                                                         \
        auto k = rows_op_kernel<LPR, VPL, OP>;                                                                   \
        int per_sm = 1;                                                                                          \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);                                  \
        long long blocks = (n + kWarps * (32 / LPR) - 1) / (kWarps * (32 / LPR));                                \
        long long grid = (long long)brs_sm_count() * (per_sm < 1 ? 1 : per_sm);                                  \
        if (grid > blocks) grid = blocks;                                                                        \
        k<<<(int)grid, kThreads, 0, st>>>(table, n_rows, D, (const long long*)idx, n, buf, scale);               \
    } while (0)
    if (D <= 4) BRS_ROWS(1, 1);
    else if (D <= 8) BRS_ROWS(2, 1);
    else if (D <= 16) BRS_ROWS(4, 1);
    else if (D <= 32) BRS_ROWS(8, 1);
    else if (D <= 64) BRS_ROWS(16, 1);
    else if (D <= 128) BRS_ROWS(32, 1);
    else if (D <= 256) BRS_ROWS(32, 2);
    else if (D <= 384) BRS_ROWS(32, 3);
    else BRS_ROWS(32, 4);
#undef BRS_ROWS
    BRS_CUDA_CHECK(cudaGetLastError());
    return BRS_OK;
}
}  // namespace

extern "C" int brs_gather(const float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n, float* out,
                          void* stream) {
    return launch_rows_op<OP_GATHER>(const_cast<float*>(table), n_rows, dim, idx, n, out, 0.f, stream);
}

extern "C" int brs_scatter_add(float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n,
                               const float* src, float scale, void* stream) {
    return launch_rows_op<OP_SCATTER_ADD>(table, n_rows, dim, idx, n, const_cast<float*>(src), scale, stream);
}

extern "C" int brs_gather_sgd_update(float* table, int64_t n_rows, int32_t dim, const int64_t* idx, int64_t n,
                                     float lr, float* out, void* stream) {
    return launch_rows_op<OP_GATHER_SGD>(table, n_rows, dim, idx, n, out, lr, stream);
}
