// multi.cu -- one host process driving N GPUs behind the drop-in boundary (rla_set_devices).
//
// The reference call sites (src/matrix/mat_mul.rs:57-67, src/matrix/decomposition/lu.rs:163-195) can only ever reach
// rla_dgemm / rla_dgetrf; after rla_set_devices(N) those two calls shard over the first N GPUs of the box:
//
//   GEMM  row panels: GPU g owns rows [g*mg, (g+1)*mg) of A and C.  B is the path's one exchange step.  Every column
//         chunk of B is uploaded from the host by ONE GPU (chunks dealt round-robin, so the N PCIe links share the
//         upload of B) and fanned out to the other N-1 over NVLink/NVSwitch by peer copies the receivers pull as soon
//         as the owner's H2D has landed.  Uploads, pulls, DMMA kernels and downloads run as a 2-D wavefront per GPU
//         (A row strips x B column chunks), all asynchronous, issued by the calling thread; every C element is one
//         full-k product of the same kernel, so the result is bit-identical to the single-GPU call.
//   LU    1D block-cyclic column blocks (block 256, owner(J) = J mod N); the factored panel travels by peer copies.
//
// There is no NCCL in this file: inside one process the exchange is peer memory (cudaMemcpyPeer semantics under UVA);
// the one-process-per-GPU driver (rulinalg_b200/sharded*.py, bench.py under torchrun) uses NCCL for the same step.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "context.cuh"

namespace rla {
namespace {

std::mutex g_multi_mu;                          // one multi-device operation (or reconfiguration) at a time
std::atomic<int> g_ndev{1};
std::unique_ptr<Context> g_mctx[RLA_MAX_DEVICES];

struct DeviceRestore {
    int dev = -1;
    DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = -1; } }
    ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

int device_ctx(int d, Context **out) {
    RLA_CUDA(cudaSetDevice(d));
    if (!g_mctx[d]) g_mctx[d].reset(new Context());
    RLA_TRY(g_mctx[d]->init(d));
    *out = g_mctx[d].get();
    return RLA_OK;
}

}  // namespace

int multi_device_count() { return g_ndev.load(std::memory_order_relaxed); }

void multi_release() {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    for (auto &c : g_mctx) c.reset();
}

int multi_set_devices(int n) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    const int cnt = probe_device_count();
    if (cnt == 0) return RLA_ERR_NO_DEVICE;
    if (n < 1 || n > cnt) return RLA_ERR_INVALID;
    DeviceRestore restore;
    if (n > 1) {
        for (int i = 0; i < n; ++i) {
            if (!device_usable(i)) return RLA_ERR_NO_DEVICE;
            RLA_CUDA(cudaSetDevice(i));
            for (int j = 0; j < n; ++j) {
                if (i == j) continue;
                int can = 0;
                RLA_CUDA(cudaDeviceCanAccessPeer(&can, i, j));
                if (!can) return RLA_ERR_INVALID;             // the exchange step needs peer memory
                const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { note_cuda_error(e); return RLA_ERR_CUDA; }
                (void)cudaGetLastError();
            }
        }
    }
    g_ndev.store(n, std::memory_order_relaxed);
    return RLA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// GEMM: C = alpha * A * B over N GPUs (beta == 0, unit column strides; the caller packed anything else).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
int gemm_host_multi(size_t m, size_t k, size_t n, T alpha, const T *ha, size_t hrsa, const T *hb, size_t hrsb, T *hc,
                    size_t hrsc, Stager &stg) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    DeviceRestore restore;
    const int G = multi_device_count();
    const size_t mg = ((m + G - 1) / G + 127) / 128 * 128;         // rows per GPU
    const int Gu = int((m + mg - 1) / mg);                          // GPUs that get rows
    const bool pin_a = host_is_pinned(ha), pin_b = host_is_pinned(hb), pin_c = host_is_pinned(hc);
    const size_t lda = pad_ld(k, sizeof(T)), ldb = pad_ld(n, sizeof(T)), ldc = pad_ld(n, sizeof(T));

    Context *cx[RLA_MAX_DEVICES];
    T *dA[RLA_MAX_DEVICES], *dB[RLA_MAX_DEVICES], *dC[RLA_MAX_DEVICES];
    size_t rows_of[RLA_MAX_DEVICES];
    for (int g = 0; g < Gu; ++g) {
        RLA_TRY(device_ctx(g, &cx[g]));
        rows_of[g] = (size_t(g) + 1) * mg <= m ? mg : m - size_t(g) * mg;
        RLA_TRY(cx[g]->dA.ensure(rows_of[g] * lda * sizeof(T)));
        RLA_TRY(cx[g]->dB.ensure(k * ldb * sizeof(T)));
        RLA_TRY(cx[g]->dC.ensure(rows_of[g] * ldc * sizeof(T)));
        dA[g] = static_cast<T *>(cx[g]->dA.p);
        dB[g] = static_cast<T *>(cx[g]->dB.p);
        dC[g] = static_cast<T *>(cx[g]->dC.p);
    }

    // strips: ~512 rows / columns (the single-GPU pipeline's measured optimum), at least one B chunk per GPU
    size_t S = (mg > n ? mg : n) / 512;
    S = S < 4 ? 4 : (S > 32 ? 32 : S);
    if (S < size_t(Gu)) S = size_t(Gu);
    const size_t pm = ((mg + S - 1) / S + 127) / 128 * 128, pn = ((n + S - 1) / S + 127) / 128 * 128;
    const size_t sm = (mg + pm - 1) / pm, sn = (n + pn - 1) / pn;
    const size_t steps = sm > sn ? sm : sn;
    enum { EV_IN = 0, EV_ROW, EV_COL, EV_P2P, EV_B, EV_PER_STEP };
    auto ev = [&](int g, size_t st, int which, cudaEvent_t *e) { return cx[g]->event(EV_PER_STEP * st + which, e); };

    for (size_t st = 0; st < steps; ++st) {
        const size_t c0 = st * pn;
        const size_t cols = st < sn ? (c0 + pn <= n ? pn : n - c0) : 0;
        const int owner = int(st % size_t(Gu));
        // ---- phase 1: host -> device.  The chunk's owner uploads it first (the others are waiting for it) ----
        for (int g = 0; g < Gu; ++g) {
            RLA_CUDA(cudaSetDevice(g));
            cudaEvent_t e;
            if (cols && g == owner) {
                RLA_TRY(stg.upload2d(dB[g] + c0, ldb * sizeof(T), hb + c0, hrsb * sizeof(T), cols * sizeof(T), k, pin_b, g, cx[g]->copy_in));
                RLA_TRY(ev(g, st, EV_B, &e));
                RLA_CUDA(cudaEventRecord(e, cx[g]->copy_in));
            }
            const size_t r0 = st * pm;
            const size_t rows = r0 < rows_of[g] ? (r0 + pm <= rows_of[g] ? pm : rows_of[g] - r0) : 0;
            if (rows)
                RLA_TRY(stg.upload2d(dA[g] + r0 * lda, lda * sizeof(T), ha + (size_t(g) * mg + r0) * hrsa, hrsa * sizeof(T),
                                     k * sizeof(T), rows, pin_a, g, cx[g]->copy_in));
            RLA_TRY(ev(g, st, EV_IN, &e));
            RLA_CUDA(cudaEventRecord(e, cx[g]->copy_in));
        }
        // ---- phase 2: the exchange step.  Every other GPU pulls the chunk from its owner over NVLink ----
        if (cols) {
            cudaEvent_t eb;
            RLA_TRY(ev(owner, st, EV_B, &eb));
            for (int g = 0; g < Gu; ++g) {
                if (g == owner) continue;
                RLA_CUDA(cudaSetDevice(g));
                RLA_CUDA(cudaStreamWaitEvent(cx[g]->p2p, eb, 0));
                RLA_CUDA(cudaMemcpy2DAsync(dB[g] + c0, ldb * sizeof(T), dB[owner] + c0, ldb * sizeof(T), cols * sizeof(T), k,
                                           cudaMemcpyDefault, cx[g]->p2p));
                cudaEvent_t e;
                RLA_TRY(ev(g, st, EV_P2P, &e));
                RLA_CUDA(cudaEventRecord(e, cx[g]->p2p));
            }
        }
        // ---- phase 3: every C tile that just became computable, then its download ----
        for (int g = 0; g < Gu; ++g) {
            RLA_CUDA(cudaSetDevice(g));
            Context &c = *cx[g];
            const size_t r0 = st * pm;
            const size_t rows = r0 < rows_of[g] ? (r0 + pm <= rows_of[g] ? pm : rows_of[g] - r0) : 0;
            const size_t ncols_avail = (st + 1 < sn ? (st + 1) * pn : n);
            const size_t nrows_prev = r0 < rows_of[g] ? r0 : rows_of[g];
            cudaEvent_t e_in, e_row, e_col, e_p2p = nullptr;
            RLA_TRY(ev(g, st, EV_IN, &e_in));
            RLA_TRY(ev(g, st, EV_ROW, &e_row));
            RLA_TRY(ev(g, st, EV_COL, &e_col));
            if (cols && g != owner) RLA_TRY(ev(g, st, EV_P2P, &e_p2p));
            T *hcg = hc + size_t(g) * mg * hrsc;
            if (rows) {
                // row strip st against every chunk that has arrived (earlier chunks were waited for in earlier steps)
                RLA_CUDA(cudaStreamWaitEvent(c.stream, e_in, 0));
                if (e_p2p) RLA_CUDA(cudaStreamWaitEvent(c.stream, e_p2p, 0));
                RLA_TRY(gemm_dev<T>(rows, k, ncols_avail, alpha, dA[g] + r0 * lda, lda, dB[g], ldb, T(0), dC[g] + r0 * ldc, ldc, c.stream));
                RLA_CUDA(cudaEventRecord(e_row, c.stream));
                RLA_CUDA(cudaStreamWaitEvent(c.copy_out, e_row, 0));
                RLA_TRY(stg.download2d(hcg + r0 * hrsc, hrsc * sizeof(T), dC[g] + r0 * ldc, ldc * sizeof(T), ncols_avail * sizeof(T),
                                       rows, pin_c, g, c.copy_out));
            }
            if (cols && nrows_prev) {
                // column strip: the row strips before st against this step's chunk (disjoint C tiles, second stream)
                RLA_CUDA(cudaStreamWaitEvent(c.stream2, e_in, 0));
                if (e_p2p) RLA_CUDA(cudaStreamWaitEvent(c.stream2, e_p2p, 0));
                RLA_TRY(gemm_dev<T>(nrows_prev, k, cols, alpha, dA[g], lda, dB[g] + c0, ldb, T(0), dC[g] + c0, ldc, c.stream2));
                RLA_CUDA(cudaEventRecord(e_col, c.stream2));
                RLA_CUDA(cudaStreamWaitEvent(c.copy_out, e_col, 0));
                RLA_TRY(stg.download2d(hcg + c0, hrsc * sizeof(T), dC[g] + c0, ldc * sizeof(T), cols * sizeof(T), nrows_prev, pin_c,
                                       g, c.copy_out));
            }
        }
    }
    for (int g = 0; g < Gu; ++g) {
        RLA_CUDA(cudaSetDevice(g));
        RLA_CUDA(cudaStreamSynchronize(cx[g]->copy_out));
        RLA_CUDA(cudaStreamSynchronize(cx[g]->stream));
        RLA_CUDA(cudaStreamSynchronize(cx[g]->stream2));
        RLA_CUDA(cudaStreamSynchronize(cx[g]->p2p));      // nobody may still be reading this GPU's copy of B
    }
    return stg.finish();
}

// ---------------------------------------------------------------------------------------------------------------
// LU: PartialPivLu::decompose over N GPUs, 1D block-cyclic column blocks.  Global column block J (256
// wide) lives on GPU J mod N at local columns [(J/N)*256, ...), every GPU stores its blocks side by side as an
// n x ncl row-major matrix, so a panel is entirely local to its owner and the pivot search needs no communication.
// Per block:  owner: factor in place (lu.cu panel kernels), pack [net row permutation | info | L11 over L21];
//             others: pull that buffer over NVLink;  all: interchanges on the local columns outside the block,
//             U12 = L11^-1 A12, A22 -= L21 U12 on the local blocks right of J.
// Look-ahead 1: the owner of block J+1 updates that block's columns first and factors it on a high-priority side
// stream while every GPU (itself included) runs the rest of block J's update; the pulls of panel J+1 hide under it.
// The per-element operation order equals the single-GPU factorisation's, so the result is bit-identical to it.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t LU_BLOCK = 256;
constexpr size_t LU_HEADER = 16384;      // >= lu_plan_bytes() + 8; a multiple of 256 so the panel stays 16-byte aligned

struct LuDev {
    Context *cx = nullptr;
    cudaStream_t side = nullptr;         // look-ahead factorisation
    size_t ncl = 0, ld = 0;              // local columns, row stride
    void *a = nullptr;                   // local matrix
    unsigned char *buf[2] = {nullptr, nullptr};
};
}  // namespace

template <typename T>
int getrf_host_multi(size_t n_, T *lu, size_t *perm, Stager &stg) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    DeviceRestore restore;
    if (n_ > 0x3fffffffull || n_ > lu_max_n(sizeof(T))) return RLA_ERR_INVALID;
    const int n = int(n_);
    const int nb = int((n_ + LU_BLOCK - 1) / LU_BLOCK);
    const int G = multi_device_count() < nb ? multi_device_count() : nb;
    const bool pinned = host_is_pinned(lu);
    const size_t plan_bytes = lu_plan_bytes();
    const size_t info_off = (plan_bytes + 7) / 8 * 8;
    if (info_off + 8 > LU_HEADER) return RLA_ERR_INVALID;
    auto width = [&](int J) { return size_t(J) * LU_BLOCK + LU_BLOCK <= n_ ? LU_BLOCK : n_ - size_t(J) * LU_BLOCK; };
    auto lcol0 = [&](int J) { return size_t(J / G) * LU_BLOCK; };

    // RLA_MULTI_TRACE=1: phase times on stderr (the phases are hard dependencies anyway, so the extra syncs cost nothing)
    const bool trace = getenv("RLA_MULTI_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    LuDev dv[RLA_MAX_DEVICES];
    enum { EV_UP = 0, EV_HEAD, EV_READY, EV_DONE, EV_COUNT };      // + per block J: [arrived] EV_COUNT + 2J, [packed] EV_COUNT + 2J + 1
    for (int g = 0; g < G; ++g) {
        LuDev &d = dv[g];
        RLA_TRY(device_ctx(g, &d.cx));
        for (int J = g; J < nb; J += G) d.ncl += width(J);
        d.ld = pad_ld(d.ncl, sizeof(T));
        RLA_TRY(d.cx->dA.ensure(n_ * d.ld * sizeof(T)));
        d.a = d.cx->dA.p;
        for (int i = 0; i < 2; ++i) {
            RLA_TRY(d.cx->dPanel[i].ensure(LU_HEADER + n_ * LU_BLOCK * sizeof(T)));
            d.buf[i] = static_cast<unsigned char *>(d.cx->dPanel[i].p);
        }
        if (!d.cx->lu_ws.side) {
            int lo = 0, hi = 0;
            RLA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            RLA_CUDA(cudaStreamCreateWithPriority(&d.cx->lu_ws.side, cudaStreamNonBlocking, hi));
            RLA_CUDA(cudaEventCreateWithFlags(&d.cx->lu_ws.ev_head, cudaEventDisableTiming));
            RLA_CUDA(cudaEventCreateWithFlags(&d.cx->lu_ws.ev_fact, cudaEventDisableTiming));
        }
        d.side = d.cx->lu_ws.side;
        for (int which = 0; which < EV_COUNT + 2 * nb; ++which) {   // all events exist before the worker threads start sharing them
            cudaEvent_t e;
            RLA_TRY(d.cx->event(size_t(which), &e));
        }
    }
    auto evt = [&](int g, int which, cudaEvent_t *e) { return dv[g].cx->event(size_t(which), e); };
    auto info_ptr = [&](int g, int b) { return reinterpret_cast<int32_t *>(dv[g].buf[b] + info_off); };
    auto panel_ptr = [&](int g, int b) { return reinterpret_cast<T *>(dv[g].buf[b] + LU_HEADER); };
    auto A = [&](int g) { return static_cast<T *>(dv[g].a); };

    // ---- upload: block columns in ascending order, every GPU over its own PCIe link ----
    for (int J = 0; J < nb; ++J) {
        const int g = J % G;
        RLA_CUDA(cudaSetDevice(g));
        RLA_TRY(stg.upload2d(A(g) + lcol0(J), dv[g].ld * sizeof(T), lu + size_t(J) * LU_BLOCK, n_ * sizeof(T), width(J) * sizeof(T), n_,
                             pinned, g, dv[g].cx->copy_in));
    }
    for (int g = 0; g < G; ++g) {
        RLA_CUDA(cudaSetDevice(g));
        cudaEvent_t e;
        RLA_TRY(evt(g, EV_UP, &e));
        RLA_CUDA(cudaEventRecord(e, dv[g].cx->copy_in));
        RLA_CUDA(cudaStreamWaitEvent(dv[g].cx->stream, e, 0));
    }
    double t_up = 0, t_fact = 0;
    if (trace) {
        for (int g = 0; g < G; ++g) { RLA_CUDA(cudaSetDevice(g)); RLA_CUDA(cudaStreamSynchronize(dv[g].cx->copy_in)); }
        t_up = ms_since(t_begin);
    }
    // row-origin vector (-> perm) lives on GPU 0
    RLA_CUDA(cudaSetDevice(0));
    RLA_TRY(dv[0].cx->dRowid.ensure(n_ * sizeof(int32_t)));
    RLA_TRY(dv[0].cx->dPerm.ensure(n_ * sizeof(int64_t)));
    RLA_TRY(dv[0].cx->hSmall.ensure(64));
    int32_t *rowid = static_cast<int32_t *>(dv[0].cx->dRowid.p);
    RLA_TRY(lu_rowid_init_dev(rowid, n, dv[0].cx->stream));

    // owner side: factor block J in place and pack [plan | info | panel] into buf[b], all on `st`
    auto factor_and_pack = [&](int J, int b, const int32_t *prev_info, cudaStream_t st) -> int {
        const int g = J % G;
        const size_t row0 = size_t(J) * LU_BLOCK, w = width(J), lc0 = lcol0(J);
        int32_t *hinfo = info_ptr(g, b);
        if (prev_info) RLA_CUDA(cudaMemcpyAsync(hinfo, prev_info, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        else RLA_CUDA(cudaMemsetAsync(hinfo, 0, sizeof(int32_t), st));
        RLA_TRY(lu_factor_block_dev<T>(n, A(g), dv[g].ld, int(row0), int(lc0), int(w), hinfo, dv[g].buf[b], dv[g].cx->lu_ws, st));
        RLA_CUDA(cudaMemcpy2DAsync(panel_ptr(g, b), w * sizeof(T), A(g) + row0 * dv[g].ld + lc0, dv[g].ld * sizeof(T), w * sizeof(T),
                                   n_ - row0, cudaMemcpyDeviceToDevice, st));
        return RLA_OK;
    };
    auto update = [&](int g, int J, int b, size_t c0, size_t c1) -> int {
        if (c1 <= c0) return RLA_OK;
        return lu_update_dev<T>(n, A(g), dv[g].ld, int(size_t(J) * LU_BLOCK), int(width(J)), panel_ptr(g, b), width(J), int(c0), int(c1),
                                info_ptr(g, b), dv[g].cx->stream);
    };
    // first local column (on g) that belongs to a global block > J
    auto first_after = [&](int g, int J) -> size_t {
        const int Jn = J < g ? g : J + 1 + ((g - (J + 1)) % G + G) % G;     // smallest block index > J owned by g
        return Jn < nb ? lcol0(Jn) : dv[g].ncl;
    };

    // ---- the block loop, SPMD: one host thread per GPU issues that GPU's work.  GPU-side ordering is by events (one
    //      [packed] and one [arrived] event per block and GPU, never re-recorded); what the host threads must agree on is
    //      only "the event I am about to wait on has already been recorded", which two monotonic sequence numbers give:
    //        packed_seq    highest block whose [factored + packed] event has been recorded by its owner
    //        pulled[g]     highest block whose [arrived] event GPU g has recorded (owners count as arrived)
    //      Fan-out of panel J: it is on the critical path of exactly ONE GPU, the owner of block J+1, which therefore pulls
    //      first and alone (full NVLink rate: 67 MB in ~0.1 ms); the others follow down a binary tree (each holder serves
    //      its two children one after the other), ~4 transfer times in all at 8 GPUs.  The first version let all N-1 GPUs
    //      pull from the owner at once: 7 x 67 MB through one GPU's egress put ~0.7 ms on EVERY block's critical path
    //      (8 GPUs, n = 32768: 357 ms against the NCCL driver's 195 ms).
    // finished block rows stream back to a pinned host matrix during the factorisation when the block-cyclic layout is
    // regular (n a multiple of 256 * G: each GPU's share of a row is nblk chunks of 2 KiB at a constant host stride)
    const bool stream_rows = pinned && n_ % (LU_BLOCK * size_t(G)) == 0;
    std::atomic<int> packed_seq{-1}, failed{RLA_OK};
    std::atomic<int> pulled[RLA_MAX_DEVICES];
    for (int g = 0; g < G; ++g) pulled[g].store(-1);
    std::atomic<uint64_t> launches{0};
    auto wait_for = [&](std::atomic<int> &x, int v) -> bool {
        while (x.load(std::memory_order_acquire) < v) {
            if (failed.load(std::memory_order_relaxed) != RLA_OK) return false;
            std::this_thread::yield();
        }
        return true;
    };
    auto got_event = [&](int dev, int J, cudaEvent_t *e) { return evt(dev, EV_COUNT + 2 * J, e); };
    auto packed_event = [&](int dev, int J, cudaEvent_t *e) { return evt(dev, EV_COUNT + 2 * J + 1, e); };
    auto worker = [&](int g) -> int {
        RLA_CUDA(cudaSetDevice(g));
        cudaStream_t st = dv[g].cx->stream;
        // buf[J & 1] of this GPU is about to be overwritten with block J: whoever read block J-2 out of it (its children in
        // that block's tree) must be done.  Conservative and cheap: wait for every receiver's [arrived] event of block J-2.
        auto buffer_reusable = [&](int J, cudaStream_t on) -> int {
            if (J < 2) return RLA_OK;
            for (int o = 0; o < G; ++o) {
                if (o == g || o == (J - 2) % G) continue;           // (owners do not pull)
                if (!wait_for(pulled[o], J - 2)) return RLA_ERR_CUDA;
                cudaEvent_t got;
                RLA_TRY(got_event(o, J - 2, &got));
                RLA_CUDA(cudaStreamWaitEvent(on, got, 0));
            }
            return RLA_OK;
        };
        auto do_pull = [&](int J) -> int {
            const int b = J & 1, O = J % G, P = (J + 1) % G;
            // tree: node 0 = owner, node 1 = the next block's owner, then the other GPUs in ascending order
            int node = 0, order[RLA_MAX_DEVICES];
            order[0] = O;
            int cnt = 1;
            if (P != O) order[cnt++] = P;
            for (int x = 0; x < G; ++x)
                if (x != O && x != P) order[cnt++] = x;
            for (int i = 0; i < cnt; ++i)
                if (order[i] == g) node = i;
            const int parent = (node - 1) / 2, src = order[parent];
            cudaStream_t ps = dv[g].cx->p2p;
            RLA_TRY(buffer_reusable(J, ps));
            cudaEvent_t e;
            if (parent == 0) {
                if (!wait_for(packed_seq, J)) return RLA_ERR_CUDA;
                RLA_TRY(packed_event(O, J, &e));
            } else {
                if (!wait_for(pulled[src], J)) return RLA_ERR_CUDA;
                RLA_TRY(got_event(src, J, &e));
            }
            RLA_CUDA(cudaStreamWaitEvent(ps, e, 0));
            if ((node & 1) == 0) {                                  // second child: after its sibling, so each gets the full link
                const int sib = order[node - 1];
                if (!wait_for(pulled[sib], J)) return RLA_ERR_CUDA;
                RLA_TRY(got_event(sib, J, &e));
                RLA_CUDA(cudaStreamWaitEvent(ps, e, 0));
            }
            const size_t bytes = LU_HEADER + (n_ - size_t(J) * LU_BLOCK) * width(J) * sizeof(T);
            cudaEvent_t ready, got;
            RLA_TRY(evt(g, EV_READY, &ready));                      // my own readers of buf[b] (update J-2) are queued on `st`
            RLA_CUDA(cudaEventRecord(ready, st));
            RLA_CUDA(cudaStreamWaitEvent(ps, ready, 0));
            RLA_CUDA(cudaMemcpyPeerAsync(dv[g].buf[b], g, dv[src].buf[b], src, bytes, ps));
            RLA_TRY(got_event(g, J, &got));
            RLA_CUDA(cudaEventRecord(got, ps));
            pulled[g].store(J, std::memory_order_release);
            return RLA_OK;
        };
        auto do_factor = [&](int J, const int32_t *prev_info, cudaStream_t fs) -> int {
            const int b = J & 1;
            RLA_TRY(buffer_reusable(J, fs));
            RLA_TRY(factor_and_pack(J, b, prev_info, fs));
            cudaEvent_t packed;
            RLA_TRY(packed_event(g, J, &packed));
            RLA_CUDA(cudaEventRecord(packed, fs));
            pulled[g].store(J, std::memory_order_release);
            packed_seq.store(J, std::memory_order_release);
            return RLA_OK;
        };
        // block 0 has no predecessor
        if (g == 0) RLA_TRY(do_factor(0, nullptr, st)); else RLA_TRY(do_pull(0));
        for (int J = 0; J < nb; ++J) {
            const int b = J & 1, owner = J % G;
            const bool have_next = J + 1 < nb;
            const int ow2 = (J + 1) % G;
            const size_t w = width(J);
            // ---- panel J has arrived -> interchanges on the local columns outside the block ----
            if (g != owner) {
                cudaEvent_t got;
                RLA_TRY(got_event(g, J, &got));
                RLA_CUDA(cudaStreamWaitEvent(st, got, 0));
            } else if (J > 0) {
                cudaEvent_t packed;                       // factored on my side stream
                RLA_TRY(packed_event(g, J, &packed));
                RLA_CUDA(cudaStreamWaitEvent(st, packed, 0));
            }
            if (g == 0) RLA_TRY(lu_rowid_apply_dev(dv[g].buf[b], rowid, info_ptr(g, b), st));
            if (g == owner) {
                const size_t lc0 = lcol0(J);
                RLA_TRY(lu_laswp_dev<T>(A(g), dv[g].ld, int(w), dv[g].buf[b], info_ptr(g, b), 0, int(lc0), int(lc0 + w), int(dv[g].ncl), st));
            } else {
                RLA_TRY(lu_laswp_dev<T>(A(g), dv[g].ld, int(w), dv[g].buf[b], info_ptr(g, b), 0, int(dv[g].ncl), 0, 0, st));
            }
            const size_t lo = first_after(g, J);
            if (have_next && g == ow2) {
                // ---- the next block's owner: its columns first, then factor + pack on the side stream under the tail update ----
                const size_t wn = width(J + 1);               // lo = block J+1's first local column
                RLA_TRY(update(g, J, b, lo, lo + wn));
                cudaEvent_t head;
                RLA_TRY(evt(g, EV_HEAD, &head));
                RLA_CUDA(cudaEventRecord(head, st));
                RLA_CUDA(cudaStreamWaitEvent(dv[g].side, head, 0));    // also orders buf[1-b] after its local readers
                RLA_TRY(do_factor(J + 1, info_ptr(g, b), dv[g].side));
                RLA_TRY(update(g, J, b, lo + wn, dv[g].ncl));
            } else {
                if (have_next) RLA_TRY(do_pull(J + 1));          // the pull of panel J+1 hides under the update
                RLA_TRY(update(g, J, b, lo, dv[g].ncl));
            }
            if (stream_rows) {
                // rows [J*256, J*256 + w) of every local column are final from here on (later interchanges only touch rows
                // below): ONE strided 3-D copy takes this GPU's 256-column chunks of those rows to their places in the host
                // matrix while the factorisation goes on
                cudaEvent_t e;
                RLA_TRY(evt(g, EV_DONE, &e));
                RLA_CUDA(cudaEventRecord(e, st));
                RLA_CUDA(cudaStreamWaitEvent(dv[g].cx->copy_out, e, 0));
                const size_t chunk = LU_BLOCK * sizeof(T), nblk = dv[g].ncl / LU_BLOCK, row0 = size_t(J) * LU_BLOCK;
                cudaMemcpy3DParms p3 = {};
                p3.srcPtr = make_cudaPitchedPtr(A(g) + row0 * dv[g].ld, chunk, chunk, nblk);
                p3.dstPtr = make_cudaPitchedPtr(lu + row0 * n_ + size_t(g) * LU_BLOCK, size_t(G) * chunk, chunk, nblk);
                p3.extent = make_cudaExtent(chunk, nblk, w);
                p3.kind = cudaMemcpyDeviceToHost;
                RLA_CUDA(cudaMemcpy3DAsync(&p3, dv[g].cx->copy_out));
            }
        }
        launches.fetch_add(launch_count_take(), std::memory_order_relaxed);
        return RLA_OK;
    };
    {
        std::vector<std::thread> threads;
        for (int g = 1; g < G; ++g)
            threads.emplace_back([&, g] {
                const int st = worker(g);
                if (st != RLA_OK) failed.store(st, std::memory_order_relaxed);
            });
        const int st0 = worker(0);
        if (st0 != RLA_OK) failed.store(st0, std::memory_order_relaxed);
        for (auto &t : threads) t.join();
        note_launch(unsigned(launches.load()));
        if (failed.load() != RLA_OK) return failed.load();
    }
    if (trace) {
        for (int g = 0; g < G; ++g) {
            RLA_CUDA(cudaSetDevice(g));
            RLA_CUDA(cudaStreamSynchronize(dv[g].cx->stream));
            RLA_CUDA(cudaStreamSynchronize(dv[g].side));
            RLA_CUDA(cudaStreamSynchronize(dv[g].cx->p2p));
        }
        t_fact = ms_since(t_begin);
    }
    // ---- perm, info, download ----
    const int lastb = (nb - 1) & 1;
    RLA_CUDA(cudaSetDevice(0));
    {
        Context &c0 = *dv[0].cx;
        int64_t *dP = static_cast<int64_t *>(c0.dPerm.p);
        int32_t *hInfo = static_cast<int32_t *>(c0.hSmall.p);
        RLA_TRY(lu_perm_from_rowid_dev(rowid, dP, n, info_ptr(0, lastb), c0.stream));
        RLA_CUDA(cudaMemcpyAsync(hInfo, info_ptr(0, lastb), sizeof(int32_t), cudaMemcpyDeviceToHost, c0.stream));
        static_assert(sizeof(size_t) == sizeof(int64_t), "LP64 expected");
        RLA_CUDA(cudaMemcpyAsync(perm, dP, n_ * sizeof(int64_t), cudaMemcpyDeviceToHost, c0.stream));
    }
    for (int g = 0; g < G; ++g) {
        RLA_CUDA(cudaSetDevice(g));
        cudaEvent_t done;
        RLA_TRY(evt(g, EV_DONE, &done));
        RLA_CUDA(cudaEventRecord(done, dv[g].cx->stream));
        RLA_CUDA(cudaStreamWaitEvent(dv[g].cx->copy_out, done, 0));
    }
    for (int J = 0; J < nb && !stream_rows; ++J) {
        const int g = J % G;
        RLA_CUDA(cudaSetDevice(g));
        RLA_TRY(stg.download2d(lu + size_t(J) * LU_BLOCK, n_ * sizeof(T), A(g) + lcol0(J), dv[g].ld * sizeof(T), width(J) * sizeof(T), n_,
                               pinned, g, dv[g].cx->copy_out));
    }
    for (int g = 0; g < G; ++g) {
        RLA_CUDA(cudaSetDevice(g));
        RLA_CUDA(cudaStreamSynchronize(dv[g].cx->copy_out));
        RLA_CUDA(cudaStreamSynchronize(dv[g].cx->stream));
        RLA_CUDA(cudaStreamSynchronize(dv[g].side));
        RLA_CUDA(cudaStreamSynchronize(dv[g].cx->p2p));
    }
    RLA_TRY(stg.finish());
    if (trace)
        fprintf(stderr, "[rla multi getrf] n=%d gpus=%d pinned=%d: upload %.1f ms, factor %.1f ms, download %.1f ms, total %.1f ms\n", n, G,
                int(pinned), t_up, t_fact - t_up, ms_since(t_begin) - t_fact, ms_since(t_begin));
    return *static_cast<int32_t *>(dv[0].cx->hSmall.p) != 0 ? RLA_ERR_SINGULAR : RLA_OK;
}

template int getrf_host_multi<double>(size_t, double *, size_t *, Stager &);
template int getrf_host_multi<float>(size_t, float *, size_t *, Stager &);

template int gemm_host_multi<double>(size_t, size_t, size_t, double, const double *, size_t, const double *, size_t, double *,
                                     size_t, Stager &);
template int gemm_host_multi<float>(size_t, size_t, size_t, float, const float *, size_t, const float *, size_t, float *, size_t,
                                    Stager &);

}  // namespace rla
