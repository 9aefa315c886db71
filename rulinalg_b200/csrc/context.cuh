// context.cuh -- host-side state of librla_b200 shared by api.cu (single-device entry points), multi.cu
// (rla_set_devices: one process driving several GPUs) and host.cu (staging of pageable host memory).
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

namespace rla {

// Grow-only device (or pinned-host) buffer.
struct Buffer {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    int ensure(size_t bytes);
    void release();
};

// ---- host memory staging (host.cu) -----------------------------------------------------------------------
// rulinalg hands over Vec<T> storage (mat_mul.rs:28-31, lu.rs:166): ordinary pageable memory.  A DMA engine cannot
// read it, and cudaMemcpyAsync on it degrades to the driver's single-threaded bounce buffer.  The Stager moves
// such operands through its own ring of pinned slots with a small pool of host threads doing the
// pageable <-> pinned copies, so PCIe stays busy in both directions; pinned (cudaHostAlloc / cudaHostRegister)
// operands bypass it and are copied in place.
bool host_is_pinned(const void *p);

class Stager {
public:
    Stager() = default;
    ~Stager() { release(); }
    Stager(const Stager &) = delete;
    Stager &operator=(const Stager &) = delete;
    // rows x width_bytes from host (row pitch spitch) to device (row pitch dpitch), asynchronous on `st` of device `dev`
    // (which the caller has made current).
    int upload2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, bool pinned,
                 int dev, cudaStream_t st);
    // device -> host; for pageable destinations the data has landed only after finish()
    int download2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, bool pinned,
                   int dev, cudaStream_t st);
    int finish();                        // wait until every pageable download has been copied out
    void release();
    uint64_t staged_bytes = 0;           // bytes that went through the rings (diagnostics)

private:
    struct Slot {
        unsigned char *p = nullptr;
        cudaEvent_t ev[RLA_MAX_DEVICES] = {};
        int dev = 0;
        bool busy = false;
        void *dst = nullptr;             // download: destination rows
        size_t dpitch = 0, width = 0, rows = 0;
    };
    static constexpr int NUP = 6, NDOWN = 18;
    static constexpr size_t SLOT_BYTES = size_t(16) << 20;
    static constexpr size_t DIRECT_BYTES = size_t(2) << 20;   // smaller pageable copies use the driver's own bounce buffer
    Slot up_[NUP], down_[NDOWN];         // slot memory is pinned on first use, one slot at a time
    int next_up_ = 0, next_down_ = 0;
    struct Drain;                        // drainer thread + its queue (host.cu)
    Drain *drain_ = nullptr;
    int ensure_drain();
    static int ensure_slot(Slot &s);
};

// parallel row copy on the staging thread pool: rows x width bytes, arbitrary pitches
void parallel_copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows);
int staging_threads();

// One per (host thread, device) for the single-device entry points, one per device for the multi-device ones.
// Owns its streams, events and grow-only buffers; everything is returned in destroy() (thread exit,
// rla_shutdown, or re-initialisation on another device).
struct Context {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t stream2 = nullptr;      // second compute stream (column strips of the host GEMM pipeline)
    cudaStream_t copy_in = nullptr;      // H2D
    cudaStream_t copy_out = nullptr;     // D2H
    cudaStream_t p2p = nullptr;          // peer pulls (multi-device paths)
    Buffer dA, dB, dC, dPerm, dInfo, dVec, dVec2, dSync, dTrsv, dChol, dPanel[2], dRowid;
    Buffer hSmall;                       // pinned scalars (info, perm)
    Buffer hPack, dPack;                 // small-call fast path: operands packed into ONE pinned block <-> one device block
    LuWorkspace lu_ws;
    Stager stager;                       // pinned ring for pageable host operands (allocated on first use)
    std::vector<cudaEvent_t> events;
    Context() { hSmall.pinned_host = true; hPack.pinned_host = true; }
    ~Context() { destroy(); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    int init(int dev);                   // create the streams on `dev` (tears down first if bound to another device)
    void destroy();
    int event(size_t i, cudaEvent_t *out);   // i-th timing-disabled event of this context (created on demand)
};

Context &thread_ctx();                   // the calling host thread's context
int ensure_ctx(int device = -1);         // bind it (lazily) and make its device current
int probe_device_count();                // visible devices (0 when none); cached
bool device_usable(int d);               // d is an sm_100 device

// ---- multi-device state (multi.cu) -------------------------------------------------------------------------
int multi_device_count();                // value set by rla_set_devices (1 = single-device behaviour)
int multi_set_devices(int n);
void multi_release();                    // rla_shutdown
template <typename T>
int gemm_host_multi(size_t m, size_t k, size_t n, T alpha, const T *ha, size_t hrsa, const T *hb, size_t hrsb, T *hc,
                    size_t hrsc, Stager &stg);
template <typename T>
int getrf_host_multi(size_t n, T *lu, size_t *perm, Stager &stg);

// shared helpers (api.cu)
template <typename T>
int gemm_dev(size_t m, size_t k, size_t n, T alpha, const T *a, size_t lda, const T *b, size_t ldb, T beta, T *c,
             size_t ldc, cudaStream_t st);
inline size_t pad_ld(size_t cols, size_t elem) {
    const size_t q = 16 / elem;          // keep rows 16-byte aligned so the cp.async fast path applies
    return (cols + q - 1) / q * q;
}
void lu_workspace_release(LuWorkspace &ws);   // lu.cu
int measure_peak(int kind, double *tflops);   // peak.cu

}  // namespace rla
