// api.cu -- the extern "C" boundary of librla_b200.so (see include/rla_b200.h).
//
// Host-pointer entry points = stage operands into HBM, run the device twin, copy the result back.
// There is no CPU compute path here: if no sm_100 device is usable every call returns
// RLA_ERR_NO_DEVICE.  One context per host thread (stream + grow-only device/pinned buffers), so the
// library is re-entrant like the single-threaded reference is (Matrix<T>: Send + Sync).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "context.cuh"

namespace rla {

extern int g_dgemm_cfg, g_dgemm_streamk;   // dgemm.cu
extern int g_sgemm_cfg;   // sgemm.cu
extern int g_lu_gmax, g_lu_dbg, g_lu_cluster, g_lu_slab_rows, g_lu_k3e_rows;   // lu.cu
int g_host_gemm_2d = 1;           // rla_set_tuning("host_gemm_2d", 0/1): 2-D wavefront host pipeline on/off
int g_host_gemm_s = 0;            // rla_set_tuning("host_gemm_s", S): panels/chunks per dimension of that pipeline; 0 = auto (~512-row strips)
int g_host_gemm_kprefix = -1;     // rla_set_tuning("host_gemm_kprefix", v): 2-D pipeline, fraction of k (in 1/16) uploaded and multiplied as
                                  // rank-kc updates of ALL of C before the wavefront starts; -1 = auto (k/4 when k >= 4096), 0 = off
int g_host_gemm_kchunk = 256;     // rla_set_tuning("host_gemm_kchunk", kc): depth of those rank-kc updates (measured 256 / 512 / 1024: 35.7 / 36.1 / 37.2 ms)
int g_host_gemm_grade = 0;        // rla_set_tuning("host_gemm_grade", 0/1): graded first / last strips of that pipeline (off: measured 37.9 vs 38.2 ms at 8192^3 pinned -- the start-up loss is the quadratic growth of computable work, not the strip size)
int g_host_stage = 1;             // rla_set_tuning("host_stage", 0/1): pageable operands through the pinned staging ring (host.cu)

namespace {
thread_local cudaError_t tl_last_cuda = cudaSuccess;
thread_local uint64_t tl_launches = 0;
}  // namespace

int Buffer::ensure(size_t bytes) {
    if (bytes <= cap) return RLA_OK;
    release();
    // grow geometrically to avoid realloc churn on size sweeps
    size_t want = bytes + bytes / 8;
    cudaError_t e = pinned_host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = pinned_host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        note_cuda_error(e);
        cudaGetLastError();
        return RLA_ERR_NOMEM;
    }
    cap = want;
    return RLA_OK;
}
void Buffer::release() {
    if (p) {
        if (pinned_host) cudaFreeHost(p); else cudaFree(p);
        (void)cudaGetLastError();
    }
    p = nullptr;
    cap = 0;
}

int Context::init(int dev) {
    if (ready && device == dev) return RLA_OK;
    destroy();
    RLA_CUDA(cudaSetDevice(dev));
    device = dev;
    RLA_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    RLA_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    RLA_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
    RLA_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
    RLA_CUDA(cudaStreamCreateWithFlags(&p2p, cudaStreamNonBlocking));
    ready = true;
    return RLA_OK;
}

// Everything this context allocated goes back: streams, events, device and pinned buffers, the LU workspace and the
// staging ring.  Runs at host-thread exit (thread_local destructor), from rla_shutdown, and is harmless when the CUDA
// runtime is already unloading (every call then just fails).
void Context::destroy() {
    if (device < 0) return;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { (void)cudaGetLastError(); cur = -1; }
    if (cudaSetDevice(device) == cudaSuccess) {
        for (cudaStream_t *s : {&stream, &stream2, &copy_in, &copy_out, &p2p}) {
            if (*s) { cudaStreamSynchronize(*s); cudaStreamDestroy(*s); }
            *s = nullptr;
        }
        stager.release();
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
        for (Buffer *b : {&dA, &dB, &dC, &dPerm, &dInfo, &dVec, &dVec2, &dSync, &dTrsv, &dChol, &dPanel[0], &dPanel[1], &dRowid, &hSmall, &hPack, &dPack})
            b->release();
        lu_workspace_release(lu_ws);
    }
    (void)cudaGetLastError();
    if (cur >= 0) cudaSetDevice(cur);
    (void)cudaGetLastError();
    ready = false;
    device = -1;
}

int Context::event(size_t i, cudaEvent_t *out) {
    while (events.size() <= i) {
        cudaEvent_t e;
        RLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events.push_back(e);
    }
    *out = events[i];
    return RLA_OK;
}

namespace {

// One context per (host thread, device): a thread that moves between devices (rla_init(d), or the caller's own
// cudaSetDevice for the *_dev twins) gets separate streams, buffers and workspaces on each of them.
struct ThreadContexts {
    std::unique_ptr<Context> per_dev[RLA_MAX_DEVICES];
};
thread_local ThreadContexts tl_ctxs;

std::once_flag g_dev_once;
int g_dev_count = 0;
bool g_dev_ok[RLA_MAX_DEVICES] = {};

void probe_devices() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        g_dev_count = 0;
        return;
    }
    if (cnt > RLA_MAX_DEVICES) cnt = RLA_MAX_DEVICES;
    for (int d = 0; d < cnt; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) != cudaSuccess) { cudaGetLastError(); major = 0; }
        g_dev_ok[d] = (major == 10);       // kernels are sm_100a only; no fallback
    }
    g_dev_count = cnt;
}

}  // namespace

int probe_device_count() {
    std::call_once(g_dev_once, probe_devices);
    return g_dev_count;
}

bool device_usable(int d) { return d >= 0 && d < probe_device_count() && g_dev_ok[d]; }

Context &thread_ctx() {
    const int d = current_device();
    if (!tl_ctxs.per_dev[d]) tl_ctxs.per_dev[d].reset(new Context());
    return *tl_ctxs.per_dev[d];
}

// Binds the calling thread to `device` (or to its current CUDA device when device < 0) and makes sure the
// (thread, device) context exists.  Every entry point starts here, so the device is always the one the context's
// streams and buffers live on.
int ensure_ctx(int device) {
    if (probe_device_count() == 0) return RLA_ERR_NO_DEVICE;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); return RLA_ERR_NO_DEVICE; }
    } else {
        if (device >= g_dev_count) return RLA_ERR_NO_DEVICE;
        RLA_CUDA(cudaSetDevice(device));
    }
    if (device >= g_dev_count || !g_dev_ok[device]) return RLA_ERR_NO_DEVICE;
    Context &c = thread_ctx();
    if (c.ready) return RLA_OK;
    return c.init(device);
}

template <>
int gemm_dev<double>(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda, const double *b,
                     size_t ldb, double beta, double *c, size_t ldc, cudaStream_t st) {
    return dgemm_launch(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
}
template <>
int gemm_dev<float>(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b,
                    size_t ldb, float beta, float *c, size_t ldc, cudaStream_t st) {
    return sgemm_launch(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
}
// C <- alpha * (C + A*B), bit-identical continuation of a k-split product (dgemm.cu / sgemm.cu ACC_C)
template <typename T> constexpr bool gemm_has_acc() { return true; }
template <typename T>
int gemm_dev_acc(size_t, size_t, size_t, T, const T *, size_t, const T *, size_t, T *, size_t, cudaStream_t);
template <>
int gemm_dev_acc<float>(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b, size_t ldb,
                        float *c, size_t ldc, cudaStream_t st) {
    return sgemm_launch(m, k, n, alpha, a, lda, b, ldb, 0.f, c, ldc, st, true);
}
template <>
int gemm_dev_acc<double>(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda, const double *b, size_t ldb,
                         double *c, size_t ldc, cudaStream_t st) {
    return dgemm_launch(m, k, n, alpha, a, lda, b, ldb, 0.0, c, ldc, st, true);
}

// ---- held operands (device-resident operands across host-API calls) --------------------------
// An operand may stay resident in HBM across calls only if the library KNOWS it cannot have changed; behind
// matrixmultiply's signature nothing says so, so the caller says it: rla_operand_hold(ptr, bytes) declares the host range
// immutable until rla_operand_release(ptr) (in Rust: a guard that holds `&Matrix` -- the borrow checker then enforces the
// promise).  While a range is held, the first product that reads an operand inside it keeps that operand's device copy, and
// later calls with the same (pointer, shape, row stride) skip its H2D.
namespace {
struct HeldRegion { const unsigned char *base; size_t bytes; int refs; };
struct HeldView { const void *p; size_t rows, cols, rs, elem; int device; void *dptr; };
std::mutex g_held_mu;
std::vector<HeldRegion> g_held_regions;
std::vector<HeldView> g_held_views;

bool held_covers(const void *p, size_t extent) {
    std::lock_guard<std::mutex> lk(g_held_mu);
    const unsigned char *q = static_cast<const unsigned char *>(p);
    for (const HeldRegion &r : g_held_regions)
        if (q >= r.base && q + extent <= r.base + r.bytes) return true;
    return false;
}
void *held_lookup(const void *p, size_t rows, size_t cols, size_t rs, size_t elem, int device) {
    std::lock_guard<std::mutex> lk(g_held_mu);
    for (const HeldView &v : g_held_views)
        if (v.p == p && v.rows == rows && v.cols == cols && v.rs == rs && v.elem == elem && v.device == device) return v.dptr;
    return nullptr;
}
// registers a freshly filled device copy; returns false (caller frees its copy) when the range was released meanwhile or
// another thread registered the same view first
bool held_insert(const void *p, size_t extent, size_t rows, size_t cols, size_t rs, size_t elem, int device, void *dptr) {
    std::lock_guard<std::mutex> lk(g_held_mu);
    const unsigned char *q = static_cast<const unsigned char *>(p);
    bool covered = false;
    for (const HeldRegion &r : g_held_regions)
        if (q >= r.base && q + extent <= r.base + r.bytes) covered = true;
    if (!covered) return false;
    for (const HeldView &v : g_held_views)
        if (v.p == p && v.rows == rows && v.cols == cols && v.rs == rs && v.elem == elem && v.device == device) return false;
    g_held_views.push_back(HeldView{p, rows, cols, rs, elem, device, dptr});
    return true;
}
void held_free_views(const unsigned char *base, size_t bytes, bool all, const std::vector<HeldView> *keep = nullptr) {
    std::vector<HeldView> drop;
    {
        std::lock_guard<std::mutex> lk(g_held_mu);
        for (size_t i = 0; i < g_held_views.size();) {
            const unsigned char *q = static_cast<const unsigned char *>(g_held_views[i].p);
            bool kept = false;
            if (keep)
                for (const HeldView &kv : *keep) kept = kept || kv.dptr == g_held_views[i].dptr;
            if (!kept && (all || (q >= base && q < base + bytes))) {
                drop.push_back(g_held_views[i]);
                g_held_views.erase(g_held_views.begin() + long(i));
            } else ++i;
        }
    }
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { (void)cudaGetLastError(); cur = -1; }
    for (const HeldView &v : drop)
        if (cudaSetDevice(v.device) == cudaSuccess) cudaFree(v.dptr);
    (void)cudaGetLastError();
    if (cur >= 0) cudaSetDevice(cur);
}

// One operand of a host-API call: resident (a held view exists), to be kept (held range, first use) or transient.
template <typename T>
struct Operand {
    T *dev = nullptr;          // where the kernels read it
    bool resident = false;     // device copy already valid: skip the H2D
    bool fresh = false;        // allocated by this call for a held range: register on success, free otherwise
    const void *host = nullptr;
    size_t extent = 0, rows = 0, cols = 0, rs = 0;
    int resolve(const T *h, bool packed, size_t rows_, size_t cols_, size_t rs_, size_t ld, Buffer &fallback, int device) {
        host = h; rows = rows_; cols = cols_; rs = rs_;
        extent = rows ? ((rows - 1) * rs + cols) * sizeof(T) : 0;
        if (!packed && rows && cols && held_covers(h, extent)) {
            if (void *v = held_lookup(h, rows, cols, rs, sizeof(T), device)) {
                dev = static_cast<T *>(v);
                resident = true;
                return RLA_OK;
            }
            void *pnew = nullptr;
            if (cudaMalloc(&pnew, rows * ld * sizeof(T)) == cudaSuccess) {
                dev = static_cast<T *>(pnew);
                fresh = true;
                return RLA_OK;
            }
            (void)cudaGetLastError();          // no room for a resident copy: behave as if the range were not held
        }
        RLA_TRY(fallback.ensure((rows ? rows : 1) * ld * sizeof(T)));
        dev = static_cast<T *>(fallback.p);
        return RLA_OK;
    }
    void finish(bool ok, int device) {
        if (!fresh) return;
        if (!ok || !held_insert(host, extent, rows, cols, rs, sizeof(T), device, dev)) cudaFree(dev);
        fresh = false;
    }
    bool owns(const T *dst, size_t ld) const { return dst >= dev && dst < dev + (rows ? rows : 1) * ld; }
};
}  // namespace

namespace {

cudaStream_t pick_stream(void *s) { return static_cast<cudaStream_t>(s); }   // NULL = CUDA legacy default stream

// Row-strided host matrix -> contiguous-ish device matrix (ld elements per row).  Large pageable operands travel through
// the calling thread's staging ring (host.cu); callers of download_matrix finish with cx.stager.finish().
inline bool treat_as_pinned(const void *p, size_t bytes) {
    return !g_host_stage || bytes < (size_t(2) << 20) || host_is_pinned(p);
}
template <typename T>
int upload_matrix(T *dst, size_t ld, const T *src, size_t rs, size_t rows, size_t cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return RLA_OK;
    Context &cx = thread_ctx();
    return cx.stager.upload2d(dst, ld * sizeof(T), src, rs * sizeof(T), cols * sizeof(T), rows,
                              treat_as_pinned(src, rows * cols * sizeof(T)), cx.device, st);
}
template <typename T>
int download_matrix(T *dst, size_t rs, const T *src, size_t ld, size_t rows, size_t cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return RLA_OK;
    Context &cx = thread_ctx();
    return cx.stager.download2d(dst, rs * sizeof(T), src, ld * sizeof(T), cols * sizeof(T), rows,
                                treat_as_pinned(dst, rows * cols * sizeof(T)), cx.device, st);
}

// The matrix operand of a solve / matrix-vector call: resident if it lies in a held range (rla_operand_hold) and was seen
// before, otherwise uploaded (and kept when the range is held).  For these O(n^2)-flop calls the upload IS the cost:
// n = 4096 f64 solve 2.8 ms with the factors re-uploaded, 0.4 ms with them held.
template <typename T>
struct MatOperand {
    Operand<T> op;
    int device = 0;
    int prepare(const T *h, size_t rows, size_t cols, size_t rs, size_t ld, Buffer &fallback, cudaStream_t st) {
        Context &cx = thread_ctx();
        device = cx.device;
        RLA_TRY(op.resolve(h, false, rows, cols, rs, ld, fallback, device));
        if (!op.resident) {
            const int s = upload_matrix(op.dev, ld, h, rs, rows, cols, st);
            if (s != RLA_OK) {
                cudaStreamSynchronize(st);
                (void)cudaGetLastError();
                op.finish(false, device);
                return s;
            }
        }
        return RLA_OK;
    }
    // `uploaded`: the stream has been synchronised since prepare() and the copy went through
    void done(bool uploaded) { op.finish(uploaded, device); }
};

// Host operand with arbitrary (possibly negative / non-unit) strides -> packed row-major copy.
template <typename T>
void pack_host(std::vector<T> &out, const T *src, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols) {
    out.resize(rows * cols);
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) out[i * cols + j] = src[ptrdiff_t(i) * rs + ptrdiff_t(j) * cs];
}

// Small calls (the only shapes the reference itself benchmarks: benches/linalg/matrix.rs:40-65, lu.rs:52-139) are bound by
// API calls, not by bytes: the general path costs ~45 us in three pageable copies, events and three stream syncs.  Here
// the operands are packed into ONE pinned block by the CPU (nanoseconds at these sizes), travel in ONE H2D copy, and the
// result comes back in ONE D2H copy behind a single synchronisation: 2 copies + the launches + 1 sync.
constexpr size_t SMALL_CALL_BYTES = size_t(256) << 10;
inline size_t round16(size_t b) { return (b + 15) / 16 * 16; }

template <typename T>
int gemm_host_small(size_t m, size_t k, size_t n, T alpha, const T *a, ptrdiff_t rsa, ptrdiff_t csa, const T *b,
                    ptrdiff_t rsb, ptrdiff_t csb, T beta, T *c, ptrdiff_t rsc, ptrdiff_t csc) {
    Context &cx = thread_ctx();
    const size_t lda = pad_ld(k ? k : 1, sizeof(T)), ldb = pad_ld(n, sizeof(T)), ldc = pad_ld(n, sizeof(T));
    const size_t offA = 0, offB = round16(m * lda * sizeof(T)), offC = offB + round16((k ? k : 1) * ldb * sizeof(T));
    const size_t total = offC + round16(m * ldc * sizeof(T));
    RLA_TRY(cx.hPack.ensure(total));
    RLA_TRY(cx.dPack.ensure(total));
    unsigned char *hp = static_cast<unsigned char *>(cx.hPack.p), *dp = static_cast<unsigned char *>(cx.dPack.p);
    T *ha = reinterpret_cast<T *>(hp + offA), *hb = reinterpret_cast<T *>(hp + offB), *hc = reinterpret_cast<T *>(hp + offC);
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < k; ++j) ha[i * lda + j] = a[ptrdiff_t(i) * rsa + ptrdiff_t(j) * csa];
    for (size_t i = 0; i < k; ++i)
        for (size_t j = 0; j < n; ++j) hb[i * ldb + j] = b[ptrdiff_t(i) * rsb + ptrdiff_t(j) * csb];
    size_t up = offC;
    if (beta != T(0)) {
        for (size_t i = 0; i < m; ++i)
            for (size_t j = 0; j < n; ++j) hc[i * ldc + j] = c[ptrdiff_t(i) * rsc + ptrdiff_t(j) * csc];
        up = total;
    }
    RLA_CUDA(cudaMemcpyAsync(dp, hp, up, cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(gemm_dev<T>(m, k, n, alpha, reinterpret_cast<T *>(dp + offA), lda, reinterpret_cast<T *>(dp + offB), ldb, beta,
                        reinterpret_cast<T *>(dp + offC), ldc, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hc, dp + offC, m * ldc * sizeof(T), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) c[ptrdiff_t(i) * rsc + ptrdiff_t(j) * csc] = hc[i * ldc + j];
    return RLA_OK;
}

// The host-pointer GEMM.  Pipeline: B is uploaded whole (it is reused by every row panel); A is
// uploaded and C downloaded in row panels so PCIe transfers overlap the kernel of the previous /
// next panel (three streams, events).
template <typename T>
int gemm_host(size_t m, size_t k, size_t n, T alpha, const T *a, ptrdiff_t rsa, ptrdiff_t csa, const T *b,
              ptrdiff_t rsb, ptrdiff_t csb, T beta, T *c, ptrdiff_t rsc, ptrdiff_t csc) {
    if (m == 0 || n == 0) return RLA_OK;
    if ((k > 0 && (!a || !b)) || !c) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    if ((m * k + k * n + m * n) * sizeof(T) <= SMALL_CALL_BYTES)
        return gemm_host_small<T>(m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc);

    std::vector<T> pa, pb, pc;
    const T *ha = a, *hb = b;
    size_t hrsa = size_t(rsa), hrsb = size_t(rsb);
    if (k > 0 && (csa != 1 || rsa < ptrdiff_t(k))) { pack_host(pa, a, rsa, csa, m, k); ha = pa.data(); hrsa = k; }
    if (k > 0 && (csb != 1 || rsb < ptrdiff_t(n))) { pack_host(pb, b, rsb, csb, k, n); hb = pb.data(); hrsb = n; }
    const bool c_direct = (csc == 1 && rsc >= ptrdiff_t(n));
    T *hc = c;
    size_t hrsc = size_t(rsc);
    if (!c_direct) {
        if (beta != T(0)) pack_host(pc, c, rsc, csc, m, n); else pc.resize(m * n);
        hc = pc.data();
        hrsc = n;
    }

    // Pageable operands (what a Rust Vec<T> is) travel through the pinned staging ring; pinned ones are DMA'd in place.
    Stager &stg = cx.stager;
    const bool stage = g_host_stage != 0;
    const bool pin_a = !stage || host_is_pinned(ha), pin_b = !stage || host_is_pinned(hb), pin_c = !stage || host_is_pinned(hc);
    const double flops = 2.0 * double(m) * double(k) * double(n);
    const int ndev = multi_device_count();
    if (ndev > 1 && beta == T(0) && k > 0 && m >= size_t(256) * ndev && n >= 1024 && flops >= 4e11) {
        // rla_set_devices(N): row panels of A and C over N GPUs, B chunks fanned out over NVLink (multi.cu)
        RLA_TRY(gemm_host_multi<T>(m, k, n, alpha, ha, hrsa, hb, hrsb, hc, hrsc, stg));
    } else {
    const size_t lda = pad_ld(k ? k : 1, sizeof(T)), ldb = pad_ld(n, sizeof(T)), ldc = pad_ld(n, sizeof(T));
    Operand<T> opA, opB;                       // held operands stay resident across calls (rla_operand_hold)
    RLA_TRY(opA.resolve(ha, ha != a, k ? m : 0, k, hrsa, lda, cx.dA, cx.device));
    RLA_TRY(opB.resolve(hb, hb != b, k, k ? n : 0, hrsb, ldb, cx.dB, cx.device));
    RLA_TRY(cx.dC.ensure(m * ldc * sizeof(T)));
    T *dA = opA.dev, *dB = opB.dev, *dC = static_cast<T *>(cx.dC.p);
    auto up = [&](T *dst, size_t ld, const T *src, size_t rs, size_t rows, size_t cols, bool pinned) -> int {
        if ((opA.resident && opA.owns(dst, lda)) || (opB.resident && opB.owns(dst, ldb))) return RLA_OK;
        return stg.upload2d(dst, ld * sizeof(T), src, rs * sizeof(T), cols * sizeof(T), rows, pinned, cx.device, cx.copy_in);
    };
    auto pipeline = [&]() -> int {
    auto down = [&](T *dst, size_t rs, const T *src, size_t ld, size_t rows, size_t cols) -> int {
        return stg.download2d(dst, rs * sizeof(T), src, ld * sizeof(T), cols * sizeof(T), rows, pin_c, cx.device, cx.copy_out);
    };

    const size_t bytes_per_row = (k + n) * sizeof(T);
    const bool big = m * bytes_per_row > (size_t(96) << 20);
    if (big && beta == T(0) && k > 0 && m >= 1024 && n >= 1024 && g_host_gemm_2d) {
        // 2-D wavefront pipeline.  A is cut into S row panels, B into S column chunks; step s uploads panel s
        // and chunk s, then computes every C tile that just became computable (row strip s x [0..s], column strip
        // [0..s-1] x s) and downloads it.  Work availability grows quadratically while uploads proceed linearly,
        // so the kernel starts after 2/S of the H2D traffic instead of after all of B, and PCIe in both
        // directions stays busy under the DMMA kernel.
        size_t S = size_t(g_host_gemm_s);
        if (S == 0) {                    // measured at n = 8192 (tools/e2e_probe.py): 45 / 40 / 38.4 / 38.8 ms for S = 4 / 8 / 16 / 32
            S = (m > n ? m : n) / 512;
            S = S < 4 ? 4 : (S > 32 ? 32 : S);
        }
        // Strip boundaries (multiples of 128): uniform strips of ~512.  Optional grading ("host_gemm_grade", off by default):
        // first strip cut 128 / 128 / rest, last strip rest / 128 / 128, so the first kernel starts after two 8 MiB uploads and
        // the tail after the final upload is a quarter strip.  Measured at 8192^3: 37.9 vs 38.2 ms -- the uploads deliver
        // computable work quadratically (W(t) = (t / 19.5 ms)^2 x 31.9 ms), which starves the GPU for the first ~6 ms whatever
        // the strip size; that, not the strip granularity, is the gap to the kernel time.
        const size_t pm = ((m + S - 1) / S + 127) / 128 * 128, pn = ((n + S - 1) / S + 127) / 128 * 128;
        auto boundaries = [&](size_t total, size_t p, std::vector<size_t> &bd) {
            bd.clear();
            bd.push_back(0);
            const size_t nstrips = (total + p - 1) / p;
            for (size_t i = 0; i < nstrips; ++i) {
                const size_t lo = i * p, hi = (lo + p < total ? lo + p : total);
                const bool grade = g_host_gemm_grade && p >= 512 && hi - lo == p && nstrips >= 4;
                if (grade && i == 0) { bd.push_back(lo + 128); bd.push_back(lo + 256); }
                if (grade && i == nstrips - 1) { bd.push_back(hi - 256); bd.push_back(hi - 128); }
                bd.push_back(hi);
            }
        };
        std::vector<size_t> rb, cb;
        boundaries(m, pm, rb);
        boundaries(n, pn, cb);
        const size_t sm = rb.size() - 1, sn = cb.size() - 1;
        const size_t steps = sm > sn ? sm : sn;
        // k-prefix: the wavefront makes work available QUADRATICALLY in the uploaded bytes (a C tile needs a whole row
        // panel and a whole column chunk), which starves the GPU for the first ~6 ms at 8192^3.  Rank-kc updates of ALL of C
        // need only kc columns of A and kc rows of B each -- work LINEAR in the bytes -- so the first k1 = k/4 of the k
        // range is uploaded and multiplied that way (chunk 0 plain, later chunks continuing the accumulators: ACC_C), and the
        // wavefront then runs on the remaining k range while the GPU still has the prefix to chew on.  Every C element is
        // still ONE accumulation chain in k order => bit-identical to the plain pipelines (asserted in tests).
        // Measured at 8192^3: 38.3 -> 35.7 ms pinned, 44.1 -> 41.6 ms pageable (profiles/r02_e2e_kprefix_probe.jsonl).
        size_t k1 = 0, kc = size_t(g_host_gemm_kchunk > 0 ? g_host_gemm_kchunk : 256) / 32 * 32;
        if (kc == 0) kc = 32;
        if (gemm_has_acc<T>() && g_host_gemm_kprefix != 0) {
            const size_t sixteenths = g_host_gemm_kprefix < 0 ? 4 : size_t(g_host_gemm_kprefix > 15 ? 15 : g_host_gemm_kprefix);
            if (g_host_gemm_kprefix > 0 || k >= 4096) k1 = (k * sixteenths / 16) / kc * kc;
            if (k1 + 1024 > k) k1 = 0;
        }
        const size_t nchunk = k1 / kc;
        while (cx.events.size() < 3 * steps + nchunk + 2) {
            cudaEvent_t e;
            RLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            cx.events.push_back(e);
        }
        for (size_t ch = 0; ch < nchunk; ++ch) {
            const size_t kk = ch * kc;
            RLA_TRY(up(dA + kk, lda, ha + kk, hrsa, m, kc, pin_a));
            RLA_TRY(up(dB + kk * ldb, ldb, hb + kk * hrsb, hrsb, kc, n, pin_b));
            cudaEvent_t ev = cx.events[3 * steps + 1 + ch];
            RLA_CUDA(cudaEventRecord(ev, cx.copy_in));
            RLA_CUDA(cudaStreamWaitEvent(cx.stream, ev, 0));
            if (ch == 0) RLA_TRY(gemm_dev<T>(m, kc, n, T(1), dA, lda, dB, ldb, T(0), dC, ldc, cx.stream));
            else RLA_TRY(gemm_dev_acc<T>(m, kc, n, T(1), dA + kk, lda, dB + kk * ldb, ldb, dC, ldc, cx.stream));
        }
        cudaEvent_t ev_prefix = cx.events[3 * steps + 1 + nchunk];
        if (k1) RLA_CUDA(cudaEventRecord(ev_prefix, cx.stream));
        const size_t kr = k - k1;                       // k range of the wavefront
        auto tile_gemm = [&](size_t mm, size_t nn, const T *pa, const T *pb, T *pc, cudaStream_t strm) -> int {
            if (k1 == 0) return gemm_dev<T>(mm, k, nn, alpha, pa, lda, pb, ldb, beta, pc, ldc, strm);
            return gemm_dev_acc<T>(mm, kr, nn, alpha, pa + k1, lda, pb + k1 * ldb, ldb, pc, ldc, strm);
        };
        if (k1) RLA_CUDA(cudaStreamWaitEvent(cx.stream2, ev_prefix, 0));
        for (size_t st = 0; st < steps; ++st) {
            const size_t r0 = st < sm ? rb[st] : m, c0 = st < sn ? cb[st] : n;
            const size_t rows = st < sm ? rb[st + 1] - r0 : 0;
            const size_t cols = st < sn ? cb[st + 1] - c0 : 0;
            if (rows) RLA_TRY(up(dA + r0 * lda + k1, lda, ha + r0 * hrsa + k1, hrsa, rows, kr, pin_a));
            if (cols) RLA_TRY(up(dB + k1 * ldb + c0, ldb, hb + k1 * hrsb + c0, hrsb, kr, cols, pin_b));
            cudaEvent_t ev_in = cx.events[3 * st], ev_row = cx.events[3 * st + 1], ev_col = cx.events[3 * st + 2];
            RLA_CUDA(cudaEventRecord(ev_in, cx.copy_in));
            // row strip: rows of panel st against every chunk uploaded so far (including this step's);
            // column strip: panels before st against this step's chunk.  The two strips write disjoint tiles of C and
            // run on two streams, so the partial last wave of one is filled by the other (and by the next step's).
            const size_t ncols_avail = (st + 1 < sn ? cb[st + 1] : n);
            const size_t nrows_prev = (st < sm ? rb[st] : m);
            if (rows) {
                RLA_CUDA(cudaStreamWaitEvent(cx.stream, ev_in, 0));
                RLA_TRY(tile_gemm(rows, ncols_avail, dA + r0 * lda, dB, dC + r0 * ldc, cx.stream));
                RLA_CUDA(cudaEventRecord(ev_row, cx.stream));
            }
            if (cols && nrows_prev) {
                RLA_CUDA(cudaStreamWaitEvent(cx.stream2, ev_in, 0));
                RLA_TRY(tile_gemm(nrows_prev, cols, dA, dB + c0, dC + c0, cx.stream2));
                RLA_CUDA(cudaEventRecord(ev_col, cx.stream2));
            }
            if (rows) {
                RLA_CUDA(cudaStreamWaitEvent(cx.copy_out, ev_row, 0));
                RLA_TRY(down(hc + r0 * hrsc, hrsc, dC + r0 * ldc, ldc, rows, ncols_avail));
            }
            if (cols && nrows_prev) {
                RLA_CUDA(cudaStreamWaitEvent(cx.copy_out, ev_col, 0));
                RLA_TRY(down(hc + c0, hrsc, dC + c0, ldc, nrows_prev, cols));
            }
        }
    } else {
    // row-panel pipeline: panel height chosen so that a panel is >= ~32 MiB of A+C traffic
    size_t panel = m;
    if (big) {
        panel = ((size_t(32) << 20) / bytes_per_row + 127) / 128 * 128;
        if (panel < 128) panel = 128;
        if (panel > m) panel = m;
    }
    const size_t npanels = (m + panel - 1) / panel;
    while (cx.events.size() < 2 * npanels + 1) {
        cudaEvent_t e;
        RLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cx.events.push_back(e);
    }
    RLA_TRY(up(dB, ldb, hb, hrsb, k, n, pin_b));
    for (size_t p = 0; p < npanels; ++p) {
        const size_t r0 = p * panel, rows = (r0 + panel <= m) ? panel : m - r0;
        RLA_TRY(up(dA + r0 * lda, lda, ha + r0 * hrsa, hrsa, rows, k, pin_a));
        if (beta != T(0)) RLA_TRY(up(dC + r0 * ldc, ldc, hc + r0 * hrsc, hrsc, rows, n, pin_c));
        RLA_CUDA(cudaEventRecord(cx.events[2 * p], cx.copy_in));
        RLA_CUDA(cudaStreamWaitEvent(cx.stream, cx.events[2 * p], 0));
        RLA_TRY(gemm_dev<T>(rows, k, n, alpha, dA + r0 * lda, lda, dB, ldb, beta, dC + r0 * ldc, ldc, cx.stream));
        RLA_CUDA(cudaEventRecord(cx.events[2 * p + 1], cx.stream));
        RLA_CUDA(cudaStreamWaitEvent(cx.copy_out, cx.events[2 * p + 1], 0));
        RLA_TRY(down(hc + r0 * hrsc, hrsc, dC + r0 * ldc, ldc, rows, n));
    }
    }
    RLA_CUDA(cudaStreamSynchronize(cx.copy_out));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream2));
    return stg.finish();             // pageable C: the last staged pieces are copied out by the drainer
    };
    const int pst = pipeline();
    if (pst != RLA_OK) {              // nothing may still be reading / writing the operand copies when they are freed
        cudaStreamSynchronize(cx.copy_in);
        cudaStreamSynchronize(cx.stream);
        cudaStreamSynchronize(cx.stream2);
        (void)cudaGetLastError();
    }
    opA.finish(pst == RLA_OK, cx.device);
    opB.finish(pst == RLA_OK, cx.device);
    RLA_TRY(pst);
    }
    if (!c_direct)
        for (size_t i = 0; i < m; ++i)
            for (size_t j = 0; j < n; ++j) c[ptrdiff_t(i) * rsc + ptrdiff_t(j) * csc] = pc[i * n + j];
    return RLA_OK;
}

// Small-call twins of decompose / solve (see gemm_host_small): device block = [info 16 B | perm n*8 | lu n*ld (| b n)].
template <typename T>
int getrf_host_small(size_t n, T *lu, size_t *perm) {
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    const size_t offP = 16, offA = offP + round16(n * sizeof(int64_t)), total = offA + round16(n * ld * sizeof(T));
    RLA_TRY(cx.hPack.ensure(total));
    RLA_TRY(cx.dPack.ensure(total));
    unsigned char *hp = static_cast<unsigned char *>(cx.hPack.p), *dp = static_cast<unsigned char *>(cx.dPack.p);
    T *ha = reinterpret_cast<T *>(hp + offA);
    for (size_t i = 0; i < n; ++i) memcpy(ha + i * ld, lu + i * n, n * sizeof(T));
    RLA_CUDA(cudaMemcpyAsync(dp + offA, hp + offA, n * ld * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(getrf_launch<T>(n, reinterpret_cast<T *>(dp + offA), ld, reinterpret_cast<int64_t *>(dp + offP),
                            reinterpret_cast<int32_t *>(dp), cx.lu_ws, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hp, dp, total, cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*reinterpret_cast<int32_t *>(hp) != 0) return RLA_ERR_SINGULAR;
    memcpy(perm, hp + offP, n * sizeof(int64_t));
    for (size_t i = 0; i < n; ++i) memcpy(lu + i * n, ha + i * ld, n * sizeof(T));
    return RLA_OK;
}

template <typename T>
int getrs_host_small(size_t n, const T *lu, const size_t *perm, T *b) {
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    const size_t offB = 16, offP = offB + round16(n * sizeof(T)), offA = offP + round16(n * sizeof(int64_t));
    const size_t total = offA + round16(n * ld * sizeof(T));
    RLA_TRY(cx.hPack.ensure(total));
    RLA_TRY(cx.dPack.ensure(total));
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(T)));
    RLA_TRY(cx.dSync.ensure(64));
    unsigned char *hp = static_cast<unsigned char *>(cx.hPack.p), *dp = static_cast<unsigned char *>(cx.dPack.p);
    T *ha = reinterpret_cast<T *>(hp + offA);
    memcpy(hp + offB, b, n * sizeof(T));
    memcpy(hp + offP, perm, n * sizeof(int64_t));
    for (size_t i = 0; i < n; ++i) memcpy(ha + i * ld, lu + i * n, n * sizeof(T));
    RLA_CUDA(cudaMemcpyAsync(dp + offB, hp + offB, total - offB, cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(getrs_launch<T>(n, reinterpret_cast<const T *>(dp + offA), ld, reinterpret_cast<const int64_t *>(dp + offP),
                            reinterpret_cast<T *>(dp + offB), static_cast<T *>(cx.dTrsv.p), reinterpret_cast<int32_t *>(dp),
                            static_cast<int32_t *>(cx.dSync.p), cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hp, dp, offP, cudaMemcpyDeviceToHost, cx.stream));          // [info | x]
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*reinterpret_cast<int32_t *>(hp) != 0) return RLA_ERR_SINGULAR;                   // b left untouched
    memcpy(b, hp + offB, n * sizeof(T));
    return RLA_OK;
}

// PartialPivLu::decompose through host memory (lu.rs:163-195 consumes a Vec): upload, factor, and download every block
// row as soon as it can no longer change (getrf_launch's rows_final hook), so the D2H of the factors runs under the
// factorisation of the rest instead of after it.  Pageable memory travels through the staging ring.
template <typename T>
int getrf_host(size_t n, T *lu, size_t *perm, T **keep_dev, int64_t **keep_perm, size_t *keep_ld) {
    if (n == 0) return RLA_OK;
    if (!lu || !perm) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    if (!keep_dev && multi_device_count() > 1 && n >= 8192)      // rla_set_devices(N): 1D block-cyclic over N GPUs (multi.cu)
        return getrf_host_multi<T>(n, lu, perm, cx.stager);
    if (!keep_dev && n * n * sizeof(T) <= SMALL_CALL_BYTES) return getrf_host_small<T>(n, lu, perm);
    const size_t ld = pad_ld(n, sizeof(T));
    T *dA;
    int64_t *dP;
    void *p1 = nullptr, *p2 = nullptr;
    if (keep_dev) {
        RLA_CUDA(cudaMalloc(&p1, n * ld * sizeof(T)));
        const cudaError_t e = cudaMalloc(&p2, n * sizeof(int64_t));
        if (e != cudaSuccess) { cudaFree(p1); (void)cudaGetLastError(); note_cuda_error(e); return RLA_ERR_NOMEM; }
        dA = static_cast<T *>(p1);
        dP = static_cast<int64_t *>(p2);
    } else {
        RLA_TRY(cx.dA.ensure(n * ld * sizeof(T)));
        RLA_TRY(cx.dPerm.ensure(n * sizeof(int64_t)));
        dA = static_cast<T *>(cx.dA.p);
        dP = static_cast<int64_t *>(cx.dPerm.p);
    }
    auto body = [&]() -> int {
        RLA_TRY(cx.dInfo.ensure(64));
        RLA_TRY(cx.hSmall.ensure(64));
        int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p);
        int32_t *hInfo = static_cast<int32_t *>(cx.hSmall.p);
        Stager &stg = cx.stager;
        const bool pinned = !g_host_stage || host_is_pinned(lu);
        size_t nev = 0;
        // The first outer block's factorisation reads nothing but the first 256 columns: they go up first, the rest of the
        // matrix follows on the copy stream UNDER that block's panel kernels (n = 4096: 0.6 of the 2.4 ms upload hidden).
        const size_t w0 = 256;
        const bool split_up = n >= 2048;
        cudaEvent_t e_first = nullptr, e_rest = nullptr;
        if (split_up) {
            RLA_TRY(cx.event(nev++, &e_first));
            RLA_TRY(cx.event(nev++, &e_rest));
            RLA_TRY(stg.upload2d(dA, ld * sizeof(T), lu, n * sizeof(T), w0 * sizeof(T), n, pinned, cx.device, cx.copy_in));
            RLA_CUDA(cudaEventRecord(e_first, cx.copy_in));
            RLA_TRY(stg.upload2d(dA + w0, ld * sizeof(T), lu + w0, n * sizeof(T), (n - w0) * sizeof(T), n, pinned, cx.device, cx.copy_in));
            RLA_CUDA(cudaEventRecord(e_rest, cx.copy_in));
            RLA_CUDA(cudaStreamWaitEvent(cx.stream, e_first, 0));
        } else {
            RLA_TRY(stg.upload2d(dA, ld * sizeof(T), lu, n * sizeof(T), n * sizeof(T), n, pinned, cx.device, cx.stream));
        }
        LuAfterFirstBlock after_first = [&](cudaStream_t st) -> int {
            RLA_CUDA(cudaStreamWaitEvent(st, e_rest, 0));
            return RLA_OK;
        };
        const bool overlap = n >= 1024;            // below that one download after the factorisation is just as fast
        LuRowsFinal rows_final = [&](int row0, int nrows, cudaStream_t st) -> int {
            cudaEvent_t e;
            RLA_TRY(cx.event(nev++, &e));
            RLA_CUDA(cudaEventRecord(e, st));
            RLA_CUDA(cudaStreamWaitEvent(cx.copy_out, e, 0));
            return stg.download2d(lu + size_t(row0) * n, n * sizeof(T), dA + size_t(row0) * ld, ld * sizeof(T), n * sizeof(T),
                                  size_t(nrows), pinned, cx.device, cx.copy_out);
        };
        RLA_TRY(getrf_launch<T>(n, dA, ld, dP, dInfo, cx.lu_ws, cx.stream, overlap ? &rows_final : nullptr, split_up ? &after_first : nullptr));
        RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
        static_assert(sizeof(size_t) == sizeof(int64_t), "LP64 expected");
        RLA_CUDA(cudaMemcpyAsync(perm, dP, n * sizeof(int64_t), cudaMemcpyDeviceToHost, cx.stream));
        RLA_CUDA(cudaStreamSynchronize(cx.stream));
        const bool singular = *hInfo != 0;          // the reference drops the matrix: `lu` content is unspecified then
        if (!overlap && !singular)
            RLA_TRY(stg.download2d(lu, n * sizeof(T), dA, ld * sizeof(T), n * sizeof(T), n, pinned, cx.device, cx.copy_out));
        RLA_CUDA(cudaStreamSynchronize(cx.copy_out));
        RLA_TRY(stg.finish());
        return singular ? RLA_ERR_SINGULAR : RLA_OK;
    };
    const int st = body();
    if (st != RLA_OK && st != RLA_ERR_SINGULAR) (void)cx.stager.finish();
    if (keep_dev) {
        if (st == RLA_OK) {
            *keep_dev = dA;
            *keep_perm = dP;
            *keep_ld = ld;
        } else {
            cudaFree(p1);
            cudaFree(p2);
            (void)cudaGetLastError();
        }
    }
    return st;
}

template <typename T>
int getrs_core(size_t n, const T *dLU, size_t ld, const int64_t *dP, T *b) {
    Context &cx = thread_ctx();
    RLA_TRY(cx.dVec.ensure(n * sizeof(T)));
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(T)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.dSync.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    T *dB = static_cast<T *>(cx.dVec.p), *dTmp = static_cast<T *>(cx.dTrsv.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p);
    int32_t *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    RLA_CUDA(cudaMemcpyAsync(dB, b, n * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(getrs_launch<T>(n, dLU, ld, dP, dB, dTmp, dInfo, static_cast<int32_t *>(cx.dSync.p), cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*hInfo != 0) return RLA_ERR_SINGULAR;
    RLA_CUDA(cudaMemcpyAsync(b, dB, n * sizeof(T), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    return RLA_OK;
}

template <typename T>
int getrs_host(size_t n, const T *lu, const size_t *perm, T *b) {
    if (n == 0) return RLA_OK;
    if (!lu || !perm || !b) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    if (n * n * sizeof(T) <= SMALL_CALL_BYTES) return getrs_host_small<T>(n, lu, perm, b);
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dPerm.ensure(n * sizeof(int64_t)));
    int64_t *dP = static_cast<int64_t *>(cx.dPerm.p);
    MatOperand<T> mo;                              // held factors (rla_operand_hold on lu) stay in HBM across solves
    RLA_TRY(mo.prepare(lu, n, n, n, ld, cx.dA, cx.stream));
    int st = RLA_OK;
    if (cudaMemcpyAsync(dP, perm, n * sizeof(int64_t), cudaMemcpyHostToDevice, cx.stream) != cudaSuccess) st = RLA_ERR_CUDA;
    if (st == RLA_OK) st = getrs_core<T>(n, mo.op.dev, ld, dP, b);
    if (st != RLA_OK && st != RLA_ERR_SINGULAR) { cudaStreamSynchronize(cx.stream); (void)cudaGetLastError(); }
    mo.done(st == RLA_OK || st == RLA_ERR_SINGULAR);
    return st;
}

template <typename T>
int gemv_host(size_t m, size_t n, const T *a, ptrdiff_t rs, const T *x, T *y) {
    if (m == 0) return RLA_OK;
    if (!y || (n > 0 && (!a || !x)) || rs < ptrdiff_t(n)) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n ? n : 1, sizeof(T));
    RLA_TRY(cx.dVec.ensure((n ? n : 1) * sizeof(T)));
    RLA_TRY(cx.dVec2.ensure(m * sizeof(T)));
    T *dX = static_cast<T *>(cx.dVec.p), *dY = static_cast<T *>(cx.dVec2.p);
    MatOperand<T> mo;                              // a held matrix stays in HBM across products
    RLA_TRY(mo.prepare(a, n ? m : 0, n, size_t(rs), ld, cx.dA, cx.stream));
    auto body = [&]() -> int {
        if (n) RLA_CUDA(cudaMemcpyAsync(dX, x, n * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
        RLA_TRY(gemv_launch<T>(m, n, mo.op.dev, ld, dX, dY, cx.stream));
        RLA_CUDA(cudaMemcpyAsync(y, dY, m * sizeof(T), cudaMemcpyDeviceToHost, cx.stream));
        RLA_CUDA(cudaStreamSynchronize(cx.stream));
        return RLA_OK;
    };
    const int st = body();
    if (st != RLA_OK) { cudaStreamSynchronize(cx.stream); (void)cudaGetLastError(); }
    mo.done(st == RLA_OK);
    return st;
}

template <typename T>
int trsv_host(int lower, size_t n, const T *a, ptrdiff_t rs, T *x) {
    if (n == 0) return RLA_OK;
    if (!a || !x || rs < ptrdiff_t(n)) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dVec.ensure(n * sizeof(T)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.dSync.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(T)));
    T *dX = static_cast<T *>(cx.dVec.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p), *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    MatOperand<T> mo;                              // a held triangle stays in HBM across solves
    RLA_TRY(mo.prepare(a, n, n, size_t(rs), ld, cx.dA, cx.stream));
    auto body = [&]() -> int {
        RLA_CUDA(cudaMemcpyAsync(dX, x, n * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
        RLA_TRY(trsv_launch<T>(lower != 0, n, mo.op.dev, ld, dX, static_cast<T *>(cx.dTrsv.p), dInfo, static_cast<int32_t *>(cx.dSync.p), cx.stream));
        RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
        RLA_CUDA(cudaStreamSynchronize(cx.stream));
        if (*hInfo != 0) return RLA_ERR_SINGULAR;
        RLA_CUDA(cudaMemcpyAsync(x, dX, n * sizeof(T), cudaMemcpyDeviceToHost, cx.stream));
        RLA_CUDA(cudaStreamSynchronize(cx.stream));
        return RLA_OK;
    };
    const int st = body();
    if (st != RLA_OK && st != RLA_ERR_SINGULAR) { cudaStreamSynchronize(cx.stream); (void)cudaGetLastError(); }
    mo.done(st == RLA_OK || st == RLA_ERR_SINGULAR);
    return st;
}

// ---- Cholesky --------------------------------------------------------------------------------
inline int potrf_status(int32_t info) { return info == 0 ? RLA_OK : (info > 0 ? RLA_ERR_SINGULAR : RLA_ERR_NOT_POSITIVE); }

template <typename T>
int potrf_host(size_t n, T *a) {
    if (n == 0) return RLA_OK;
    if (!a) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dA.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dChol.ensure(potrf_workspace_elems(n) * sizeof(T)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    T *dA = static_cast<T *>(cx.dA.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p), *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    RLA_TRY(upload_matrix(dA, ld, a, n, n, n, cx.stream));
    RLA_TRY(potrf_launch<T>(n, dA, ld, static_cast<T *>(cx.dChol.p), dInfo, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*hInfo != 0) return potrf_status(*hInfo);
    RLA_TRY(download_matrix(a, n, dA, ld, n, n, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    return cx.stager.finish();
}

template <typename T>
int potrs_host(size_t n, const T *l, T *b) {
    if (n == 0) return RLA_OK;
    if (!l || !b) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dA.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dC.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dVec.ensure(n * sizeof(T)));
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(T)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.dSync.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    T *dA = static_cast<T *>(cx.dA.p), *dX = static_cast<T *>(cx.dVec.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p), *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    RLA_TRY(upload_matrix(dA, ld, l, n, n, n, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(dX, b, n * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(potrs_launch<T>(n, dA, ld, dX, static_cast<T *>(cx.dC.p), static_cast<T *>(cx.dTrsv.p), dInfo, dInfo + 4,
                            static_cast<int32_t *>(cx.dSync.p), cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*hInfo != 0) return RLA_ERR_SINGULAR;
    RLA_CUDA(cudaMemcpyAsync(b, dX, n * sizeof(T), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    return RLA_OK;
}

template <typename T>
int potri_host(size_t n, const T *l, T *inv) {
    if (n == 0) return RLA_OK;
    if (!l || !inv) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dA.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dB.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dC.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dPerm.ensure(n * sizeof(int64_t)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    T *dA = static_cast<T *>(cx.dA.p), *dM = static_cast<T *>(cx.dB.p), *dX = static_cast<T *>(cx.dC.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p), *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    RLA_TRY(upload_matrix(dA, ld, l, n, n, n, cx.stream));
    RLA_TRY(potri_launch<T>(n, dA, ld, dX, ld, dM, static_cast<int64_t *>(cx.dPerm.p), dInfo, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*hInfo != 0) return RLA_ERR_SINGULAR;
    RLA_TRY(download_matrix(inv, n, dX, ld, n, n, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    return cx.stager.finish();
}

template <typename T>
int getri_host(size_t n, const T *lu, const size_t *perm, T *inv) {
    if (n == 0) return RLA_OK;
    if (!lu || !perm || !inv) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    const size_t ld = pad_ld(n, sizeof(T));
    RLA_TRY(cx.dA.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dC.ensure(n * ld * sizeof(T)));
    RLA_TRY(cx.dPerm.ensure(n * sizeof(int64_t)));
    RLA_TRY(cx.dInfo.ensure(64));
    RLA_TRY(cx.hSmall.ensure(64));
    T *dA = static_cast<T *>(cx.dA.p), *dX = static_cast<T *>(cx.dC.p);
    int64_t *dP = static_cast<int64_t *>(cx.dPerm.p);
    int32_t *dInfo = static_cast<int32_t *>(cx.dInfo.p), *hInfo = static_cast<int32_t *>(cx.hSmall.p);
    RLA_TRY(upload_matrix(dA, ld, lu, n, n, n, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(dP, perm, n * sizeof(int64_t), cudaMemcpyHostToDevice, cx.stream));
    RLA_TRY(getri_launch<T>(n, dA, ld, dP, dX, ld, dInfo, cx.stream));
    RLA_CUDA(cudaMemcpyAsync(hInfo, dInfo, sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    if (*hInfo != 0) return RLA_ERR_SINGULAR;
    RLA_TRY(download_matrix(inv, n, dX, ld, n, n, cx.stream));
    RLA_CUDA(cudaStreamSynchronize(cx.stream));
    return cx.stager.finish();
}

}  // namespace

void note_cuda_error(cudaError_t e) { tl_last_cuda = e; }
void note_launch(unsigned n) { tl_launches += n; }
uint64_t launch_count_take() {
    const uint64_t v = tl_launches;
    tl_launches = 0;
    return v;
}

}  // namespace rla

struct rla_lu_handle {
    size_t n, ld;
    double *lu;
    int64_t *perm;
    int device;
};

using namespace rla;

extern "C" {

int rla_dgemm(size_t m, size_t k, size_t n, double alpha, const double *a, ptrdiff_t rsa, ptrdiff_t csa,
              const double *b, ptrdiff_t rsb, ptrdiff_t csb, double beta, double *c, ptrdiff_t rsc, ptrdiff_t csc) {
    return gemm_host<double>(m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc);
}
int rla_sgemm(size_t m, size_t k, size_t n, float alpha, const float *a, ptrdiff_t rsa, ptrdiff_t csa, const float *b,
              ptrdiff_t rsb, ptrdiff_t csb, float beta, float *c, ptrdiff_t rsc, ptrdiff_t csc) {
    return gemm_host<float>(m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc);
}
int rla_dgetrf(size_t n, double *lu, size_t *perm) { return getrf_host<double>(n, lu, perm, nullptr, nullptr, nullptr); }
int rla_sgetrf(size_t n, float *lu, size_t *perm) { return getrf_host<float>(n, lu, perm, nullptr, nullptr, nullptr); }
int rla_dgetrs(size_t n, const double *lu, const size_t *perm, double *b) { return getrs_host<double>(n, lu, perm, b); }
int rla_sgetrs(size_t n, const float *lu, const size_t *perm, float *b) { return getrs_host<float>(n, lu, perm, b); }

int rla_dgetri(size_t n, const double *lu, const size_t *perm, double *inv) { return getri_host<double>(n, lu, perm, inv); }
int rla_sgetri(size_t n, const float *lu, const size_t *perm, float *inv) { return getri_host<float>(n, lu, perm, inv); }
int rla_dgetri_dev(size_t n, const double *lu, size_t ld, const int64_t *d_perm, double *x, size_t ldx, int32_t *d_info,
                   void *stream) {
    RLA_TRY(ensure_ctx());
    return getri_launch<double>(n, lu, ld, d_perm, x, ldx, d_info, pick_stream(stream));
}
int rla_sgetri_dev(size_t n, const float *lu, size_t ld, const int64_t *d_perm, float *x, size_t ldx, int32_t *d_info,
                   void *stream) {
    RLA_TRY(ensure_ctx());
    return getri_launch<float>(n, lu, ld, d_perm, x, ldx, d_info, pick_stream(stream));
}

int rla_dpotrf(size_t n, double *a) { return potrf_host<double>(n, a); }
int rla_spotrf(size_t n, float *a) { return potrf_host<float>(n, a); }
int rla_dpotrs(size_t n, const double *l, double *b) { return potrs_host<double>(n, l, b); }
int rla_spotrs(size_t n, const float *l, float *b) { return potrs_host<float>(n, l, b); }
int rla_dpotri(size_t n, const double *l, double *inv) { return potri_host<double>(n, l, inv); }
int rla_spotri(size_t n, const float *l, float *inv) { return potri_host<float>(n, l, inv); }
size_t rla_potrf_workspace_bytes(size_t n, size_t elem_size) { return potrf_workspace_elems(n) * elem_size; }
int rla_dpotrf_dev(size_t n, double *a, size_t ld, void *ws, int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return potrf_launch<double>(n, a, ld, static_cast<double *>(ws), d_info, pick_stream(stream));
}
int rla_spotrf_dev(size_t n, float *a, size_t ld, void *ws, int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return potrf_launch<float>(n, a, ld, static_cast<float *>(ws), d_info, pick_stream(stream));
}

int rla_dgemv(size_t m, size_t n, const double *a, ptrdiff_t rs, const double *x, double *y) { return gemv_host<double>(m, n, a, rs, x, y); }
int rla_sgemv(size_t m, size_t n, const float *a, ptrdiff_t rs, const float *x, float *y) { return gemv_host<float>(m, n, a, rs, x, y); }
int rla_dgemv_dev(size_t m, size_t n, const double *a, size_t lda, const double *x, double *y, void *stream) {
    RLA_TRY(ensure_ctx());
    return gemv_launch<double>(m, n, a, lda, x, y, pick_stream(stream));
}
int rla_sgemv_dev(size_t m, size_t n, const float *a, size_t lda, const float *x, float *y, void *stream) {
    RLA_TRY(ensure_ctx());
    return gemv_launch<float>(m, n, a, lda, x, y, pick_stream(stream));
}
int rla_dtrsv(int lower, size_t n, const double *a, ptrdiff_t rs, double *x) { return trsv_host<double>(lower, n, a, rs, x); }
int rla_strsv(int lower, size_t n, const float *a, ptrdiff_t rs, float *x) { return trsv_host<float>(lower, n, a, rs, x); }

int rla_dgetrf_keep(size_t n, double *lu, size_t *perm, rla_lu_handle **out) {
    if (!out) return RLA_ERR_INVALID;
    *out = nullptr;
    double *dA = nullptr;
    int64_t *dP = nullptr;
    size_t ld = 0;
    if (n == 0) {
        rla_lu_handle *h = new rla_lu_handle{0, 0, nullptr, nullptr, 0};
        *out = h;
        return RLA_OK;
    }
    int st = getrf_host<double>(n, lu, perm, &dA, &dP, &ld);
    if (st != RLA_OK) return st;
    *out = new rla_lu_handle{n, ld, dA, dP, thread_ctx().device};
    return RLA_OK;
}
int rla_dlu_solve(const rla_lu_handle *h, double *b) {
    if (!h) return RLA_ERR_INVALID;
    if (h->n == 0) return RLA_OK;
    if (!b) return RLA_ERR_INVALID;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); prev = -1; }
    RLA_TRY(ensure_ctx(h->device));                  // the factors live on the device that produced them
    const int st = getrs_core<double>(h->n, h->lu, h->ld, h->perm, b);
    if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
    return st;
}
void rla_lu_free(rla_lu_handle *h) {
    if (!h) return;
    if (h->lu) cudaFree(h->lu);
    if (h->perm) cudaFree(h->perm);
    delete h;
}

int rla_operand_hold(const void *host, size_t bytes) {
    if (!host || bytes == 0) return RLA_ERR_INVALID;
    std::lock_guard<std::mutex> lk(g_held_mu);
    const unsigned char *q = static_cast<const unsigned char *>(host);
    for (HeldRegion &r : g_held_regions)
        if (r.base == q) {
            if (r.bytes != bytes) return RLA_ERR_INVALID;    // the same base held twice must be the same range
            ++r.refs;
            return RLA_OK;
        }
    g_held_regions.push_back(HeldRegion{q, bytes, 1});
    return RLA_OK;
}
int rla_operand_release(const void *host) {
    const unsigned char *q = static_cast<const unsigned char *>(host);
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_held_mu);
        size_t i = 0;
        for (; i < g_held_regions.size(); ++i)
            if (g_held_regions[i].base == q) break;
        if (i == g_held_regions.size()) return RLA_ERR_INVALID;
        if (--g_held_regions[i].refs > 0) return RLA_OK;
        bytes = g_held_regions[i].bytes;
        g_held_regions.erase(g_held_regions.begin() + long(i));
    }
    // device copies of views inside the range go, unless another held range still covers them
    std::vector<HeldView> keep;
    {
        std::lock_guard<std::mutex> lk(g_held_mu);
        for (const HeldView &v : g_held_views) {
            const unsigned char *pv = static_cast<const unsigned char *>(v.p);
            const size_t ext = ((v.rows - 1) * v.rs + v.cols) * v.elem;
            for (const HeldRegion &r : g_held_regions)
                if (pv >= r.base && pv + ext <= r.base + r.bytes) { keep.push_back(v); break; }
        }
    }
    held_free_views(q, bytes, false, &keep);
    return RLA_OK;
}
size_t rla_operand_resident_bytes(void) {
    std::lock_guard<std::mutex> lk(g_held_mu);
    size_t total = 0;
    for (const HeldView &v : g_held_views) total += v.rows * pad_ld(v.cols, v.elem) * v.elem;
    return total;
}

int rla_init(int device) { return ensure_ctx(device); }
int rla_device_count(void) { return probe_device_count(); }
int rla_set_devices(int n_gpus) { return multi_set_devices(n_gpus); }
int rla_get_devices(void) { return multi_device_count(); }
int rla_shutdown(void) {
    // the calling thread's contexts (streams, events, device / pinned buffers, LU workspaces, staging rings) and the
    // multi-device contexts; everything is rebuilt lazily by the next call
    for (auto &c : tl_ctxs.per_dev) c.reset();
    multi_release();
    held_free_views(nullptr, 0, true);               // resident operand copies (the held ranges themselves stay declared)
    return RLA_OK;
}
int rla_dev_alloc(void **p, size_t bytes) {
    if (!p) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    RLA_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    return RLA_OK;
}
int rla_dev_free(void *p) {
    if (p) RLA_CUDA(cudaFree(p));
    return RLA_OK;
}
int rla_host_alloc_pinned(void **p, size_t bytes) {
    if (!p) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    RLA_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    return RLA_OK;
}
int rla_host_free_pinned(void *p) {
    if (p) RLA_CUDA(cudaFreeHost(p));
    return RLA_OK;
}
int rla_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    RLA_TRY(ensure_ctx());
    RLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pick_stream(stream)));
    return RLA_OK;
}
int rla_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    RLA_TRY(ensure_ctx());
    RLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, pick_stream(stream)));
    return RLA_OK;
}
int rla_stream_sync(void *stream) {
    RLA_TRY(ensure_ctx());
    RLA_CUDA(cudaStreamSynchronize(pick_stream(stream)));
    return RLA_OK;
}

int rla_dgemm_dev(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda, const double *b, size_t ldb,
                  double beta, double *c, size_t ldc, void *stream) {
    RLA_TRY(ensure_ctx());
    return dgemm_launch(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, pick_stream(stream));
}
int rla_sgemm_dev(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b, size_t ldb,
                  float beta, float *c, size_t ldc, void *stream) {
    RLA_TRY(ensure_ctx());
    return sgemm_launch(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, pick_stream(stream));
}
int rla_dgetrf_dev(size_t n, double *a, size_t ld, int64_t *d_perm, int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return getrf_launch<double>(n, a, ld, d_perm, d_info, thread_ctx().lu_ws, pick_stream(stream));
}
int rla_sgetrf_dev(size_t n, float *a, size_t ld, int64_t *d_perm, int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return getrf_launch<float>(n, a, ld, d_perm, d_info, thread_ctx().lu_ws, pick_stream(stream));
}
int rla_dgetrs_dev(size_t n, const double *lu, size_t ld, const int64_t *d_perm, double *d_b, int32_t *d_info,
                   void *stream) {
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(double)));
    RLA_TRY(cx.dSync.ensure(64));
    return getrs_launch<double>(n, lu, ld, d_perm, d_b, static_cast<double *>(cx.dTrsv.p), d_info,
                                static_cast<int32_t *>(cx.dSync.p), pick_stream(stream));
}
int rla_sgetrs_dev(size_t n, const float *lu, size_t ld, const int64_t *d_perm, float *d_b, int32_t *d_info,
                   void *stream) {
    RLA_TRY(ensure_ctx());
    Context &cx = thread_ctx();
    RLA_TRY(cx.dTrsv.ensure(3 * n * sizeof(float)));
    RLA_TRY(cx.dSync.ensure(64));
    return getrs_launch<float>(n, lu, ld, d_perm, d_b, static_cast<float *>(cx.dTrsv.p), d_info,
                               static_cast<int32_t *>(cx.dSync.p), pick_stream(stream));
}
size_t rla_lu_plan_bytes(void) { return lu_plan_bytes(); }
size_t rla_lu_max_n(size_t elem_size) {
    if (ensure_ctx() != RLA_OK || (elem_size != 4 && elem_size != 8)) return 0;
    return lu_max_n(elem_size);
}
int rla_debug_lu_trace(unsigned long long *host512) { return lu_trace_fetch(host512); }
int rla_debug_divcheck(int f32, int mode, unsigned long long seed, unsigned long long count, unsigned long long *mismatches) {
    if (!mismatches) return RLA_ERR_INVALID;
    RLA_TRY(ensure_ctx());
    return lu_divcheck(f32, mode, seed, count, mismatches);
}
int rla_dlu_factor_block_dev(size_t n, double *a_loc, size_t ld, size_t row0, size_t lcol0, size_t w, int32_t *d_info,
                             void *d_plan, void *stream) {
    RLA_TRY(ensure_ctx());
    if (n > 0x3fffffffull || w == 0 || w > 256 || row0 + w > n) return RLA_ERR_INVALID;
    return lu_factor_block_dev<double>(int(n), a_loc, ld, int(row0), int(lcol0), int(w), d_info, d_plan, thread_ctx().lu_ws,
                                       pick_stream(stream));
}
int rla_dlu_laswp_dev(double *a_loc, size_t ld, size_t w, const void *d_plan, const int32_t *d_info, size_t c0a,
                      size_t c1a, size_t c0b, size_t c1b, void *stream) {
    RLA_TRY(ensure_ctx());
    return lu_laswp_dev<double>(a_loc, ld, int(w), d_plan, d_info, int(c0a), int(c1a), int(c0b), int(c1b), pick_stream(stream));
}
int rla_dlu_update_dev(size_t n, double *a_loc, size_t ld, size_t row0, size_t w, const double *panel, size_t ldp,
                       size_t c0, size_t c1, const int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return lu_update_dev<double>(int(n), a_loc, ld, int(row0), int(w), panel, ldp, int(c0), int(c1), d_info, pick_stream(stream));
}
int rla_lu_rowid_init_dev(int32_t *rowid, size_t n, void *stream) {
    RLA_TRY(ensure_ctx());
    return lu_rowid_init_dev(rowid, int(n), pick_stream(stream));
}
int rla_lu_rowid_apply_dev(const void *d_plan, int32_t *rowid, const int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return lu_rowid_apply_dev(d_plan, rowid, d_info, pick_stream(stream));
}
int rla_lu_perm_from_rowid_dev(const int32_t *rowid, int64_t *d_perm, size_t n, const int32_t *d_info, void *stream) {
    RLA_TRY(ensure_ctx());
    return lu_perm_from_rowid_dev(rowid, d_perm, int(n), d_info, pick_stream(stream));
}
int rla_fill_uniform_f64_dev(double *dst, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                             double lo, double scale, void *stream) {
    RLA_TRY(ensure_ctx());
    return fill_uniform_launch<double>(dst, rows, cols, ld, seed, offset, lo, scale, pick_stream(stream));
}
int rla_fill_uniform_f32_dev(float *dst, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset, float lo,
                             float scale, void *stream) {
    RLA_TRY(ensure_ctx());
    return fill_uniform_launch<float>(dst, rows, cols, ld, seed, offset, lo, scale, pick_stream(stream));
}

int rla_measure_peak(int kind, double *tflops) {
    RLA_TRY(ensure_ctx());
    return measure_peak(kind, tflops);
}

int rla_set_tuning(const char *key, int value) {
    if (!key) return RLA_ERR_INVALID;
    if (strcmp(key, "host_gemm_kprefix") == 0) {
        if (value < -1 || value > 15) return RLA_ERR_INVALID;
        g_host_gemm_kprefix = value;
        return RLA_OK;
    }
    if (strcmp(key, "host_gemm_kchunk") == 0) {
        if (value < 32 || value > 8192) return RLA_ERR_INVALID;
        g_host_gemm_kchunk = value;
        return RLA_OK;
    }
    if (strcmp(key, "dgemm_streamk") == 0) {
        if (value < 0 || value > 3) return RLA_ERR_INVALID;
        g_dgemm_streamk = value;
        return RLA_OK;
    }
    if (strcmp(key, "dgemm_cfg") == 0) {
        if (value < -1 || value > 7) return RLA_ERR_INVALID;
        g_dgemm_cfg = value;
        return RLA_OK;
    }
    if (strcmp(key, "sgemm_cfg") == 0) {
        if (value < -1 || value > 1) return RLA_ERR_INVALID;
        g_sgemm_cfg = value;
        return RLA_OK;
    }
    if (strcmp(key, "lu_gmax") == 0) {
        if (value < 1) return RLA_ERR_INVALID;
        g_lu_gmax = value;
        return RLA_OK;
    }
    if (strcmp(key, "host_gemm_s") == 0) {
        if (value < 0 || value > 64) return RLA_ERR_INVALID;
        g_host_gemm_s = value;
        return RLA_OK;
    }
    if (strcmp(key, "host_gemm_2d") == 0) {
        g_host_gemm_2d = value ? 1 : 0;
        return RLA_OK;
    }
    if (strcmp(key, "lu_dbg") == 0) {
        g_lu_dbg = value;
        return RLA_OK;
    }
    if (strcmp(key, "host_gemm_grade") == 0) {
        g_host_gemm_grade = value ? 1 : 0;
        return RLA_OK;
    }
    if (strcmp(key, "host_stage") == 0) {
        g_host_stage = value ? 1 : 0;
        return RLA_OK;
    }
    if (strcmp(key, "lu_cluster") == 0) {
        if (value < 0 || value > 5) return RLA_ERR_INVALID;
        g_lu_cluster = value;
        return RLA_OK;
    }
    if (strcmp(key, "lu_k3e_rows") == 0) {
        if (value < 0) return RLA_ERR_INVALID;
        g_lu_k3e_rows = value;
        return RLA_OK;
    }
    if (strcmp(key, "lu_slab_rows") == 0) {
        if (value < 0) return RLA_ERR_INVALID;
        g_lu_slab_rows = value;
        return RLA_OK;
    }
    return RLA_ERR_INVALID;
}

const char *rla_strerror(int status) {
    switch (status) {
        case RLA_OK: return "ok";
        case RLA_ERR_SINGULAR: return "matrix is singular to working precision (ErrorKind::DivByZero)";
        case RLA_ERR_INVALID: return "invalid argument";
        case RLA_ERR_NOT_POSITIVE: return "diagonal entries of matrix are not all positive (ErrorKind::DecompFailure)";
        case RLA_ERR_CUDA: return "CUDA call failed (see rla_last_cuda_error)";
        case RLA_ERR_NOMEM: return "device or pinned-host allocation failed";
        case RLA_ERR_NO_DEVICE: return "no sm_100 (B200) device available; librla_b200 has no CPU fallback";
        default: return "unknown status";
    }
}
int rla_last_cuda_error(void) { return int(tl_last_cuda); }
const char *rla_version(void) { return "rla_b200 0.1.0 (sm_100a)"; }
uint64_t rla_launch_count(void) { return tl_launches; }
void rla_launch_count_reset(void) { tl_launches = 0; }

}  // extern "C"
