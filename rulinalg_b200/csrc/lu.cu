// lu.cu -- K3/K4/K5 + driver: blocked right-looking LU with partial (row) pivoting, row-major.
//
// Replaces the body of PartialPivLu::decompose (src/matrix/decomposition/lu.rs:163-195 with
// gaussian_elimination :603-616).  Reference semantics kept exactly:
//   * pivot = FIRST row attaining max |a_ik|, i >= k (strict '>' scan, lu.rs:173-178); a NaN never
//     wins a comparison, a NaN on the diagonal stays the pivot;
//   * |pivot| < epsilon (absolute) -> DivByZero (lu.rs:179-183) -> *info = k+1;
//   * whole rows are swapped (lu.rs:185) -- including the already computed L part;
//   * elimination arithmetic inside a panel is  m = a_ik / a_kk ; a_ij = a_ij - m*a_kj  with a
//     separate multiply and subtract (no FMA contraction), each element receiving
//     its updates in ascending k, so for n <= PW (one panel) the factors are bit-identical to the
//     reference.  Beyond one panel the trailing updates run on the GEMM kernels (FMA/DMMA), which
//     round differently (tolerance stated in DESIGN.md / tests).
//
// Structure (two-level): outer block W=256 columns, inner panels PW=64 columns.
//   panel  : lu_panel_kernel          -- tall panels: cooperative launch of G row CTAs (+1 hub CTA), each CTA keeps its rows
//                                        of the panel in shared memory; per column ONE grid-wide exchange of 16-byte
//                                        self-validating messages through L2, every CTA reduces the posted candidates.
//            lu_panel_cluster_kernel  -- panels of <= 4096 rows: one thread-block cluster (<= 16 CTAs), panel in registers,
//                                        candidates exchanged with st.async + mbarrier over DSMEM, implicit pivoting.
//                                        Bit-identical to lu_panel_kernel.
//   laswp  : laswp_apply_kernel       -- applies an outer block's row interchanges as ONE gather (net permutation "plan"
//                                        built incrementally by the panel kernels) instead of 256 dependent swaps;
//                                        rowid_apply_kernel carries the row-origin vector from which `perm` is produced.
//   trsm   : trsm_unit_lower_kernel   -- U12 = L11^-1 A12, four threads per column, column in registers.
//   gemm   : dgemm_launch / sgemm_launch (alpha=-1, beta=1) for the Schur complement.
//   driver : getrf_launch             -- depth-1 look-ahead: block k+1 is factored on a high-priority side stream while
//                                        the rank-256 update of block k runs on the caller's stream.
#include <cfloat>
#include <climits>
#include <math_constants.h>

#include "common.cuh"

namespace rla {
namespace {

constexpr int PW = 64;          // inner panel width
constexpr int OUTER_W = 256;    // outer block width
constexpr int PANEL_THREADS = 512;
constexpr int PLDS = PW + 1;    // padded smem row (odd => column sweeps are conflict-free)

template <typename T> struct Eps;
template <> struct Eps<double> { static __device__ __forceinline__ double v() { return DBL_EPSILON; } };
template <> struct Eps<float> { static __device__ __forceinline__ float v() { return FLT_EPSILON; } };

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// Candidate = (key, idx): key = bit pattern of |value| as an f64 (order-preserving for non-negative values; 0 with
// idx INT_MAX = "no candidate").  Winner = max key, ties -> lowest row index (the reference's strict '>' scan).
// Three redux.sync instead of 15 dependent shuffles: the shuffle version cost ~0.5 us per column.
__device__ __forceinline__ unsigned long long key_of(double absval) {
    return absval < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(absval);
}
// (A variant that decides with ONE redux + ballot + shuffles when a single lane holds the maximal high word was measured:
// slower -- n = 4096 LU 11.08 vs 10.69 ms; CREDUX results land in uniform registers and are cheaper than shuffles.)
__device__ __forceinline__ void warp_argmax(unsigned long long &key, int &idx, unsigned &winner_mask) {
    const unsigned hi = unsigned(key >> 32), lo = unsigned(key);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool top = (hi == mhi) && (lo == mlo);
    const int midx = __reduce_min_sync(0xffffffffu, top ? idx : INT_MAX);
    winner_mask = __ballot_sync(0xffffffffu, top && idx == midx);
    key = ((unsigned long long)mhi << 32) | mlo;
    idx = midx;
}

// -------------------------------------------------------------------------------------------
// Panel factorisation of A[J:n, J:J+jb].  Cooperative grid of G "row" CTAs + 1 "hub" CTA.
// Row CTA b keeps rows [J + b*R, J + (b+1)*R) of the panel in shared memory.
//
// One grid-wide exchange per column, through 16-byte self-validating messages {payload, check}.  No fence sits between
// a message and the data it announces because every piece validates itself: the check word carries a tag that is
// unique per (panel launch, column) AND a hash of the payload word.  The PTX memory model only promises single-copy
// atomicity for naturally aligned accesses of up to 8 bytes -- a 16-byte vector access is formally two 8-byte
// accesses in unspecified order -- so a reader could in principle see a new check word next to an old payload; the
// hash turns that torn view into "not there yet" (the reader polls again) instead of a silently wrong pivot row.
// (On sm_100 an aligned 16-byte global access has never been observed to tear; over DSMEM it does, see below.)
//   every row CTA publishes a candidate packet {|max| key, row index, tag} + the candidate row as 64
//   tagged chunks, then reads ALL G packets with one whole warp (ceil(G/32) per lane), reduces them
//   (every CTA reaches the same verdict), and fetches the winner's tagged row chunks.
// Buffers alternate by column parity; a row CTA can be at most one column ahead of the slowest one
// because it needs everybody's packet to advance.
//
// Row CTA 0 also appends each verdict to a tagged pivot log (one slot per column, never reused within a
// launch).  The hub CTA follows that log at its own pace -- nobody ever waits for it: one warp folds each
// interchange into the outer block's net-permutation plan (consumed by laswp_apply_kernel), ten warps apply
// it to the columns of the outer block that lie outside this panel.
// -------------------------------------------------------------------------------------------
struct __align__(16) Msg {
    unsigned long long lo, hi;
};
__device__ __forceinline__ void msg_store(Msg *p, unsigned long long lo, unsigned long long hi) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void msg_load(const Msg *p, unsigned long long &lo, unsigned long long &hi) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
// 64 -> 32 bit mix of a payload word (two odd multipliers + xor-shift: any single-word change flips ~half the bits)
__device__ __forceinline__ unsigned mix32(unsigned long long v) {
    unsigned h = unsigned(v) * 0x9E3779B1u ^ unsigned(v >> 32) * 0x85EBCA6Bu;
    return h ^ (h >> 15);
}
// value chunk: {bits, tag << 32 | mix32(bits)}
__device__ __forceinline__ void chunk_store(Msg *p, unsigned long long bits, unsigned tag) {
    msg_store(p, bits, ((unsigned long long)tag << 32) | mix32(bits));
}
__device__ __forceinline__ bool chunk_load(const Msg *p, unsigned tag, unsigned long long &bits) {
    unsigned long long hi;
    msg_load(p, bits, hi);
    return hi == (((unsigned long long)tag << 32) | mix32(bits));
}
// candidate packet: {key, tag24 << 40 | mix16(key) << 24 | idx24}; idx24 = 0xffffff means "no candidate" (INT_MAX).
// Tags stay below 2^24 (ensure_workspace re-zeroes the scratch before they would wrap), rows below 2^24 - 1.
__device__ __forceinline__ void packet_store(Msg *p, unsigned long long key, int idx, unsigned tag) {
    const unsigned long long i24 = idx == INT_MAX ? 0xffffffull : (unsigned long long)(unsigned)idx;
    msg_store(p, key, ((unsigned long long)tag << 40) | ((unsigned long long)(mix32(key) & 0xffffu) << 24) | i24);
}
__device__ __forceinline__ bool packet_load(const Msg *p, unsigned tag, unsigned long long &key, int &idx) {
    unsigned long long hi;
    msg_load(p, key, hi);
    const unsigned i24 = unsigned(hi) & 0xffffffu;
    idx = i24 == 0xffffffu ? INT_MAX : int(i24);
    return (hi >> 24) == (((unsigned long long)tag << 16) | (mix32(key) & 0xffffu));
}
__device__ __forceinline__ unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }
__device__ __forceinline__ void from_bits(unsigned long long b, double &v) { v = __longlong_as_double((long long)b); }
__device__ __forceinline__ void from_bits(unsigned long long b, float &v) { v = __uint_as_float(unsigned(b)); }

constexpr int GMAX = 256;
constexpr int LASWP_MAXJB = OUTER_W;
// Net effect of an outer block's interchanges, built incrementally by the hub CTA.
//   "Touched" rows: index i < w -> row J0+i; index w+f -> far row fr[f].  origin[i] = touched index
//   whose OLD contents end up in touched row i.
struct LaswpPlan {
    int nt;
    int rows[2 * LASWP_MAXJB];
    int origin[2 * LASWP_MAXJB];
};
struct PlanState {
    int nf;
    int od[LASWP_MAXJB];             // origin of dense touched rows
    int fr[LASWP_MAXJB];             // far row numbers
    int of[LASWP_MAXJB];             // origin of far touched rows
};
struct PanelScratch {
    Msg *packets;                    // [2][GMAX]       {bits(|max| as f64), tag24 | hash16 | idx24}
    Msg *rowbuf;                     // [2][GMAX][PW]   {bits(value), tag32 | hash32}
    Msg *diagbuf;                    // [2][PW]         {bits(value), tag32 | hash32}
    unsigned long long *piv_log;     // [PW]            tag<<32 | pivot row (0xffffffff = singular), written by row CTA 0
    PlanState *state;
    LaswpPlan *plan;
    int32_t *rowid;
    unsigned long long *trace;       // optional [64 columns][8 stamps] of globaltimer ns (lu_dbg bit 3), else nullptr
};
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot) do { if (sc.trace && tid == 0) sc.trace[c * 8 + (slot)] = gtime(); } while (0)
// ---- thread-block-cluster / DSMEM primitives (used by the cluster variants of the panel kernels) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
// remote 16-byte store that completes 16 tx-bytes on the destination CTA's mbarrier when the data has landed
__device__ __forceinline__ void st_async_v2(unsigned cluster_addr, unsigned long long lo, unsigned long long hi,
                                            unsigned cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr),
                 "l"(lo), "l"(hi), "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ double ld_cluster(const double *p, unsigned rank) {
    unsigned long long v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(mapa(smem_u32(p), rank)) : "memory");
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ float ld_cluster(const float *p, unsigned rank) {
    unsigned v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(mapa(smem_u32(p), rank)) : "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
constexpr int HUB_PLAN_WARP = 5;                  // hub warp 0 follows the pivot log, warp 5 maintains the plan
constexpr int HUB_SWAP_T0 = 192;                  // warps 6..15 apply the interchanges

// CL (cluster mode, panels whose rows fit ONE thread-block cluster of G <= 16 CTAs): the same kernel with the exchange over the
// SM-to-SM network instead of L2 -- the candidate packets go to every CTA with st.async + mbarrier complete_tx, the CTA's
// candidate row (and the diagonal row) are staged in ITS shared memory before the packet leaves, and the winner's row is
// pulled with ld.shared::cluster after the verdict.  Launched as two clusters: CTAs 0..G-1 hold the rows, CTA G is the hub,
// CTAs G+1.. exit.  Everything else (deferred update, arithmetic, pivot rule) is shared with the L2 path => bit-identical.
template <typename T, bool CL>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
lu_panel_kernel(T *__restrict__ A, size_t ld, int n, int J, int jb, int R, int G, int32_t *__restrict__ ipiv,
                int32_t *__restrict__ info, PanelScratch sc, unsigned tag_base, int J0, int w, int dbg) {
    if (*info != 0) return;   // an earlier panel hit a tiny pivot: written by a previous kernel => uniform
    extern __shared__ __align__(16) unsigned char panel_smem[];
    T *s = reinterpret_cast<T *>(panel_smem);
    __shared__ T prow_s[2][PW];                    // pivot rows of the current and the previous column (deferred update)
    __shared__ unsigned long long red_key[2][PANEL_THREADS / 32];
    __shared__ int red_idx[2][PANEL_THREADS / 32];
    __shared__ T sh_abs;
    __shared__ int sh_idx, sh_win, sh_sing;
    __shared__ int piv_sm[PW];
    __shared__ __align__(16) Msg cl_cand[2][16];           // CL: [parity][source rank], written remotely (st.async)
    __shared__ __align__(8) unsigned long long cl_bar[2];  // CL: one mbarrier per parity (1 arrival + 16*G tx-bytes)
    __shared__ __align__(16) T cl_crow[2][PW];             // CL: my candidate row, staged for remote readers
    __shared__ __align__(16) T cl_drow[2][PW];             // CL: the diagonal row (its owner's copy)

    const int b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (CL && b > G) return;                               // the hub's cluster: only its first CTA works

    if (b >= G) {
        // =========================== hub CTA ===========================
        __shared__ PlanState st;
        const bool first = (J == J0), last = (J + jb == J0 + w);
        if (first) {
            for (int i = tid; i < w; i += PANEL_THREADS) st.od[i] = i;
            if (tid == 0) st.nf = 0;
        } else {
            for (int i = tid; i < LASWP_MAXJB; i += PANEL_THREADS) {
                st.od[i] = sc.state->od[i];
                st.fr[i] = sc.state->fr[i];
                st.of[i] = sc.state->of[i];
            }
            if (tid == 0) st.nf = sc.state->nf;
        }
        for (int i = tid; i < PW; i += PANEL_THREADS) piv_sm[i] = INT_MIN;
        __syncthreads();

        if (warp == 0) {
            // ---- follower: the pivot log written by row CTA 0 (one tagged 8-byte entry per column, never
            //      overwritten within a launch, so the hub may lag by any number of columns) -> shared memory ----
            for (int c = 0; c < jb; ++c) {
                const unsigned want = tag_base + unsigned(c) + 1u;
                unsigned long long v;
                do {
                    v = *((volatile unsigned long long *)(sc.piv_log + c));
                } while (unsigned(v >> 32) != want);
                const int p = int(unsigned(v & 0xffffffffull));      // -1 = singular
                if (lane == 0) *((volatile int *)&piv_sm[c]) = p;
                if (p < 0) return;
            }
            return;
        }

        // ---- workers: follow the pivots through shared memory ----
        const int sw_c0a = J0, sw_c1a = J, sw_c0b = J + jb, sw_c1b = J0 + w;
        const int na = sw_c1a - sw_c0a, ncols = na + (sw_c1b - sw_c0b);
        const int t2 = tid - HUB_SWAP_T0;
        const int col = (t2 < na) ? sw_c0a + t2 : sw_c0b + (t2 - na);
        int nf = st.nf;
        if (warp != HUB_PLAN_WARP && (t2 < 0 || t2 >= ncols)) return;
        for (int c = 0; c < jb; ++c) {
            int p;
            while ((p = *((volatile int *)&piv_sm[c])) == INT_MIN) __nanosleep(200);   // never compete with the root warps
            if (p < 0) return;                       // singular
            const int d = J + c;
            if (p == d) continue;
            if (warp == HUB_PLAN_WARP) {
                // plan warp
                const int k = d - J0;                // dense index of the diagonal row
                if (p < J0 + w) {
                    if (lane == 0) { const int t = st.od[k]; st.od[k] = st.od[p - J0]; st.od[p - J0] = t; }
                } else {
                    int f = -1;
                    for (int base = 0; base < nf; base += 32) {
                        const int q = base + lane;
                        const unsigned hit = __ballot_sync(0xffffffffu, q < nf && st.fr[q] == p);
                        if (hit) { f = base + __ffs(hit) - 1; break; }
                    }
                    if (f < 0) {
                        f = nf++;
                        if (lane == 0) { st.fr[f] = p; st.of[f] = w + f; }
                    }
                    __syncwarp();
                    if (lane == 0) { const int t = st.od[k]; st.od[k] = st.of[f]; st.of[f] = t; }
                }
                __syncwarp();
            } else if (!(dbg & 1)) {
                T *rd = A + size_t(d) * ld + col, *rp = A + size_t(p) * ld + col;
                const T vd = *rd, vp = *rp;
                *rd = vp;
                *rp = vd;
            }
        }
        if (warp == HUB_PLAN_WARP) {
            if (!last) {
                for (int i = lane; i < LASWP_MAXJB; i += 32) {
                    sc.state->od[i] = st.od[i];
                    sc.state->fr[i] = st.fr[i];
                    sc.state->of[i] = st.of[i];
                }
                if (lane == 0) sc.state->nf = nf;
            } else {
                // outer block complete: publish the plan (rowid_apply_kernel / laswp_apply_kernel consume it)
                const int nt = w + nf;
                if (lane == 0) sc.plan->nt = nt;
                for (int i = lane; i < nt; i += 32) {
                    sc.plan->rows[i] = (i < w) ? J0 + i : st.fr[i - w];
                    sc.plan->origin[i] = (i < w) ? st.od[i] : st.of[i - w];
                }
            }
        }
        return;
    }

    // =========================== row CTAs ===========================
    const int r0 = J + b * R;
    const int r1 = min(n, r0 + R);
    const int nrows = max(0, r1 - r0);

    if (CL) {
        if (tid == 0) {
            mbar_init(smem_u32(&cl_bar[0]), 1);
            mbar_init(smem_u32(&cl_bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    for (int idx = tid; idx < nrows * jb; idx += PANEL_THREADS) {
        const int r = idx / jb, c = idx - r * jb;
        s[r * PLDS + c] = A[size_t(r0 + r) * ld + J + c];
    }
    __syncthreads();
    if (CL) cluster_sync_all();                            // every CTA's mbarriers are initialised before any packet travels

    // The rank-1 update of column c is DEFERRED: only column c+1 is brought up to date before the next pivot search
    // (it is all the search needs); the update of columns c+2.. runs in the next iteration, under the grid-wide exchange
    // of the candidates -- first the two rows that have to be published (the CTA's candidate row and the diagonal row),
    // then everything else by warps 1..15 while warp 0 polls the packets.  For tall panels the update is shared-memory
    // bound (n = 32768: 393 rows x 63 columns per CTA, ~0.8 us per column) and used to sit between two exchanges; now
    // a column costs max(exchange, update) instead of their sum.  Every element still receives the same single
    // a - m*u per column, in ascending column order => bit-identical to the other panel kernels (asserted).
    for (int c = 0; c < jb; ++c) {
        const int d = J + c;                       // global diagonal row of this column
        const int par = c & 1;
        const unsigned tag = tag_base + unsigned(c) + 1u;
        const unsigned long long tag_hi = (unsigned long long)tag << 32;   // (pivot log: 8-byte entries, atomic as they are)
        const int lo = max(0, d - r0);             // first local row still active (== first row the deferred update touches)
        const bool owns_d = (d >= r0 && d < r1);
        const T *pu = prow_s[(c + 1) & 1];         // pivot row of column c-1 (deferred update)

        if (b == 0) TRACE(0);
        // ---- local argmax over active rows of column c (first max wins) ----
        T best = T(-1);
        int bidx = INT_MAX;
        for (int r = lo + tid; r < nrows; r += PANEL_THREADS) {
            const T v = fabs(s[r * PLDS + c]);
            if (v > best) { best = v; bidx = r0 + r; }     // NaN: comparison false, never wins
        }
        unsigned long long bkey = key_of(double(best));
        if (bkey == 0ull && best < T(0)) bidx = INT_MAX;
        {
            unsigned wm;
            warp_argmax(bkey, bidx, wm);
        }
        if (lane == 0) { red_key[par][warp] = bkey; red_idx[par][warp] = bidx; }
        __syncthreads();
        // every warp reduces the 16 warp candidates itself (one CTA barrier instead of two; the buffers alternate by column)
        bkey = (lane < PANEL_THREADS / 32) ? red_key[par][lane] : 0ull;
        bidx = (lane < PANEL_THREADS / 32) ? red_idx[par][lane] : INT_MAX;
        {
            unsigned wm;
            warp_argmax(bkey, bidx, wm);
        }
        // reference starts the scan with curr_max = a_dd even when it is NaN (lu.rs:170-171):
        // then no later row can win.  Post +inf for that row so it wins the global reduce.
        if (owns_d) {
            const T dv = s[(d - r0) * PLDS + c];
            if (dv != dv) { bkey = key_of(CUDART_INF); bidx = d; }
        }
        if (!CL && tid == 0) {
            // candidate packet first: it is what everybody is waiting for
            packet_store(sc.packets + par * GMAX + b, bkey, bidx, tag);
            if (b == 0 && sc.trace) sc.trace[c * 8 + 1] = gtime();
        }
        // ---- candidate row (and the diagonal row) as self-validating chunks; their deferred update comes first ----
        const int li = bidx;
        const int li_loc = (li != INT_MAX) ? li - r0 : -1;
        {
            if (li != INT_MAX && tid < jb) {
                T v = s[li_loc * PLDS + tid];
                if (c > 0 && tid > c) { v = sub_rn(v, mul_rn(s[li_loc * PLDS + c - 1], pu[tid])); s[li_loc * PLDS + tid] = v; }
                if (CL) {
                    cl_crow[par][tid] = v;
                    if (li == d) cl_drow[par][tid] = v;
                } else {
                    chunk_store(sc.rowbuf + size_t(par * GMAX + b) * PW + tid, bits_of(v), tag);
                    if (li == d) chunk_store(sc.diagbuf + par * PW + tid, bits_of(v), tag);     // the candidate IS the diagonal row
                }
            }
            if (owns_d && li != d && tid >= 64 && tid < 64 + jb) {
                const int cc = tid - 64, dl = d - r0;
                T v = s[dl * PLDS + cc];
                if (c > 0 && cc > c) { v = sub_rn(v, mul_rn(s[dl * PLDS + c - 1], pu[cc])); s[dl * PLDS + cc] = v; }
                if (CL) cl_drow[par][cc] = v;
                else chunk_store(sc.diagbuf + par * PW + cc, bits_of(v), tag);
            }
        }
        if (CL) __syncthreads();                           // the rows are in shared memory BEFORE the packet that announces them leaves
        // ---- the verdict: every row CTA reads the G candidate packets itself with ONE whole warp (ceil(G/32)
        //      packets per lane; lanes past G duplicate packet G-1 so the warp is never partially active); the other
        //      15 warps meanwhile finish the deferred update of column c-1 on all other rows ----
        if (warp == 0) {
            unsigned long long gk = 0ull;
            int gi = INT_MAX, gw = 0;
            if (CL) {
                if (lane == 0) mbar_expect_tx(smem_u32(&cl_bar[par]), 16u * unsigned(G));
                __syncwarp();
                if (lane < G)
                    st_async_v2(mapa(smem_u32(&cl_cand[par][b]), unsigned(lane)), bkey, (unsigned long long)(unsigned)bidx,
                                mapa(smem_u32(&cl_bar[par]), unsigned(lane)));
                if (b == 0 && sc.trace && lane == 0) sc.trace[c * 8 + 1] = gtime();
                mbar_wait(smem_u32(&cl_bar[par]), unsigned(c >> 1) & 1u);
                if (lane < G) {
                    gk = ((volatile Msg *)&cl_cand[par][lane])->lo;
                    gi = int(unsigned(((volatile Msg *)&cl_cand[par][lane])->hi));
                    gw = lane;
                }
            } else
            for (int base = 0; base < G; base += 32) {
                const int q = min(base + lane, G - 1);
                unsigned long long lo_;
                int i1;
                while (!packet_load(sc.packets + par * GMAX + q, tag, lo_, i1)) {}
                if (lo_ > gk || (lo_ == gk && i1 < gi)) { gk = lo_; gi = i1; gw = q; }
            }
            unsigned wm;
            warp_argmax(gk, gi, wm);
            gw = __shfl_sync(0xffffffffu, gw, wm ? __ffs(wm) - 1 : 0);
            if (lane == 0) {
                const int sing = (T(__longlong_as_double((long long)gk)) < Eps<T>::v()) ? 1 : 0;   // lu.rs:179-183
                sh_idx = gi;
                sh_win = gw;
                sh_sing = sing;
                if (b == 0) {                      // CTA 0 keeps the books: pivot log for the hub, ipiv, info
                    *((volatile unsigned long long *)(sc.piv_log + c)) =
                        tag_hi | (sing ? 0xffffffffull : (unsigned long long)(unsigned)gi);
                    if (sing) *info = d + 1; else ipiv[d] = gi;
                    if (sc.trace) sc.trace[c * 8 + 7] = (unsigned long long)(unsigned)gi;
                }
            }
        } else if (c > 0 && c + 1 < jb) {
            // rows [lo, nrows) except the two handled above, columns c+1.. : four rows in flight per warp
            constexpr int NW = PANEL_THREADS / 32 - 1;
            const int dl = owns_d ? d - r0 : -1;
            for (int rb = lo + (warp - 1); rb < nrows; rb += 4 * NW) {
                T m[4];
                bool on[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * NW;
                    on[u] = r < nrows && r != li_loc && r != dl;
                    m[u] = on[u] ? s[r * PLDS + c - 1] : T(0);
                }
                for (int cc = c + 1 + lane; cc < jb; cc += 32) {
                    const T pv = pu[cc];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + u * NW;
                        if (on[u]) s[r * PLDS + cc] = sub_rn(s[r * PLDS + cc], mul_rn(m[u], pv));
                    }
                }
            }
        }
        __syncthreads();
        if (b == 0) TRACE(2);
        const int prow_idx = sh_idx, win = sh_win;
        if (sh_sing) return;                       // uniform across the grid (every CTA reduces the same packets)
        T *prow_c = prow_s[par];
        // the pivot row arrives (warps 0-1) and, if rows d <-> prow_idx swap inside the panel, goes straight into row d
        // while warps 2-3 bring the old diagonal row into row prow_idx: one phase, one barrier
        if (CL) {
            if (tid < jb) {
                const T v = ld_cluster(&cl_crow[par][tid], unsigned(win));
                prow_c[tid] = v;
                if (prow_idx != d && owns_d) s[(d - r0) * PLDS + tid] = v;
            } else if (prow_idx != d && prow_idx >= r0 && prow_idx < r1 && tid >= 64 && tid < 64 + jb) {
                s[(prow_idx - r0) * PLDS + (tid - 64)] = ld_cluster(&cl_drow[par][tid - 64], unsigned((d - J) / R));
            }
        } else if ((tid & ~31) < jb) {            // whole warps poll (lanes past jb re-read chunk jb-1)
            unsigned long long vlo;
            const Msg *src = sc.rowbuf + size_t(par * GMAX + win) * PW + min(tid, jb - 1);
            while (!chunk_load(src, tag, vlo)) {}
            T v;
            from_bits(vlo, v);
            if (tid < jb) {
                prow_c[tid] = v;
                if (prow_idx != d && owns_d) s[(d - r0) * PLDS + tid] = v;
            }
        } else if (prow_idx != d && prow_idx >= r0 && prow_idx < r1 && tid >= 64 && ((tid - 64) & ~31) < jb) {
            unsigned long long vlo;
            const Msg *src = sc.diagbuf + par * PW + min(tid - 64, jb - 1);
            while (!chunk_load(src, tag, vlo)) {}
            T v;
            from_bits(vlo, v);
            if (tid - 64 < jb) s[(prow_idx - r0) * PLDS + (tid - 64)] = v;
        }
        __syncthreads();
        if (b == 0) TRACE(3);
        // ---- multipliers (one IEEE division per row) and column c+1 (mul, sub) -- all the next pivot search needs;
        //      the rest of this column's rank-1 update is deferred into the next iteration ----
        const T piv = prow_c[c];
        const int lo2 = max(0, d + 1 - r0);
        if (c + 1 < jb) {
            const T pn = prow_c[c + 1];
            for (int r = lo2 + tid; r < nrows; r += PANEL_THREADS) {
                const T m = div_rn(s[r * PLDS + c], piv);
                s[r * PLDS + c] = m;
                s[r * PLDS + c + 1] = sub_rn(s[r * PLDS + c + 1], mul_rn(m, pn));
            }
        } else {
            for (int r = lo2 + tid; r < nrows; r += PANEL_THREADS) s[r * PLDS + c] = div_rn(s[r * PLDS + c], piv);
        }
        // no CTA barrier here: the next column's scan reads, per thread, exactly the rows this thread has just written
        // (same r = lo + tid mapping); everything that crosses threads is behind the barrier after the warp candidates
        if (b == 0) TRACE(4);
    }
    __syncthreads();

    for (int idx = tid; idx < nrows * jb; idx += PANEL_THREADS) {
        const int r = idx / jb, c = idx - r * jb;
        A[size_t(r0 + r) * ld + J + c] = s[r * PLDS + c];
    }
    if (CL) cluster_sync_all();                            // remote shared memory stays alive while anyone may still read it
}

// -------------------------------------------------------------------------------------------
// Cluster variant of the panel factorisation, for panels of at most 4096 rows: the whole panel lives in the
// REGISTERS of one thread-block cluster (<= 16 CTAs x 256 rows; thread (row, half) holds 32 consecutive columns of
// its row), and the per-column exchange never leaves the SM-to-SM network.  Same arithmetic, same verdicts, same
// outputs as lu_panel_kernel (bit-identical factors), ~2.5x less latency per column:
//   * every CTA sends its candidate packet into the `cand` slot it owns in EVERY CTA's shared memory with
//     st.async (16 tx-bytes completed on the destination's mbarrier when the data has landed); one try_wait later
//     each CTA holds all candidates locally and reaches the same verdict -- no fence, no L2 round trip;
//   * rows are never moved: a row's position in the reference's (physically swapped) ordering is a register, the
//     picked row retires to position J + c and the panel is written back permuted; the two threads that hold a
//     CTA's candidate row stage it in shared memory BEFORE the packet leaves (CTA barrier in between), everybody
//     fetches the winner's row with ld.shared::cluster after the verdict.  (A first version staged the row while
//     the packets were in flight and let readers poll a tag in each 16-byte {value, tag} entry: ~5 % of 4-panel
//     blocks came out wrong -- a 16-byte DSMEM load is not single-copy atomic against a local 16-byte store.);
//   * the rank-1 update runs on registers (no shared-memory traffic for the panel); the thread that updates
//     column c+1 keeps the new value aside: it is the next column's pivot candidate and multiplier numerator.
//     The column loop is unrolled by 4 so that `c % 4` is static and register indices stay compile-time.
// Buffers and mbarriers alternate by column parity: a CTA can be at most one column ahead of the slowest one.
// The hub's duties run after the column loop (16 warps x 128 registers leave no room for extra warps): one warp of
// CTA rank 0 folds the 64 pivots into the outer block's net-permutation plan while one warp of EVERY CTA folds them
// into the panel's own net permutation, which all CTAs then apply to the outer block's columns outside the panel as
// one gather (each CTA a slice of the columns): ~3 us per panel instead of 64 dependent swaps.  No thread exits
// before the final cluster barrier (remote shared memory must stay alive while anyone may still read it).
// -------------------------------------------------------------------------------------------
constexpr int CL_WORKERS = 512;
constexpr int CL_THREADS = CL_WORKERS;
constexpr int CL_MAX = 16;
constexpr int CL_ROWS = CL_WORKERS / 2;                          // rows per CTA
constexpr int CL_HC = PW / 2;                                    // columns per thread
constexpr int CL_GC = 16;                                        // columns per pass of the closing gather

__device__ __forceinline__ void worker_sync() { __syncthreads(); }

// One warp folds the interchange (row d <-> row p) into a net permutation over "touched" rows: index i < w is row
// base + i, index w + f is far row fr[f]; od[i] / of[f] = touched index whose OLD contents end up there.
__device__ __forceinline__ void fold_pivot(int *od, int *fr, int *of, int &nf, int base, int w, int d, int p, int lane) {
    if (p == d) return;
    const int k = d - base;
    if (p < base + w) {
        if (lane == 0) { const int t = od[k]; od[k] = od[p - base]; od[p - base] = t; }
    } else {
        int f = -1;
        for (int b0 = 0; b0 < nf; b0 += 32) {
            const int q = b0 + lane;
            const unsigned hit = __ballot_sync(0xffffffffu, q < nf && fr[q] == p);
            if (hit) { f = b0 + __ffs(hit) - 1; break; }
        }
        if (f < 0) {
            f = nf++;
            if (lane == 0) { fr[f] = p; of[f] = w + f; }
        }
        __syncwarp();
        if (lane == 0) { const int t = od[k]; od[k] = of[f]; of[f] = t; }
    }
    __syncwarp();
}

// trace stamps without leaving warp 0 diverged (a diverged warp takes the slow BRA.DIV path of redux.sync)
#define CTRACE(slot)                                                   \
    do {                                                               \
        if (sc.trace && rank == 0 && warp == 0) {                      \
            const unsigned long long t_ = gtime();                     \
            if (lane == 0) sc.trace[c * 8 + (slot)] = t_;              \
            __syncwarp();                                              \
        }                                                              \
    } while (0)

template <typename T>
__global__ void __launch_bounds__(CL_THREADS, 1)
lu_panel_cluster_kernel(T *__restrict__ A, size_t ld, int n, int J, int jb, int R, int32_t *__restrict__ ipiv,
                        int32_t *__restrict__ info, PanelScratch sc, int J0, int w, int dbg) {
    if (*info != 0) return;   // written by an earlier kernel => uniform over the cluster
    extern __shared__ __align__(16) unsigned char panel_smem[];   // [CL_ROWS][PLDS] staging for coalesced panel load / store
    T *stg = reinterpret_cast<T *>(panel_smem);
    __shared__ __align__(16) Msg cand[2][CL_MAX];          // [parity][source rank], written remotely (st.async)
    __shared__ __align__(8) unsigned long long bar[2];     // one mbarrier per parity: 1 arrival (mine) + 16*CS tx-bytes
    __shared__ __align__(16) T crow[2][PW];                // my candidate row, staged for remote readers
    __shared__ __align__(16) T prow_s[PW];                 // the pivot row of the current column
    __shared__ __align__(16) T colv[CL_ROWS];              // column c+1 of my rows (from the half that holds it to the other)
    __shared__ unsigned long long red_key[CL_WORKERS / 32];
    __shared__ int red_idx[CL_WORKERS / 32];
    __shared__ int sh_cpos, sh_idx, sh_win, sh_sing, sh_nt;
    __shared__ int piv_sm[PW];
    __shared__ PlanState st;                               // CTA rank 0: the outer block's plan state
    __shared__ int od_l[PW], fr_l[PW], of_l[PW];           // the panel's own net permutation
    __shared__ int rows_l[2 * PW], org_l[2 * PW];
    __shared__ int posv[CL_ROWS];                          // final position of my rows (write-back)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = cluster_ctarank(), CS = cluster_nctarank();
    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 0] = gtime();

    if (tid == 0) {
        mbar_init(smem_u32(&bar[0]), 1);
        mbar_init(smem_u32(&bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 1] = gtime();

    // Thread (lrow, half) keeps columns [32*half, 32*half + 32) of one row in registers for the whole panel.  Rows are
    // never moved: `pos` is the row's current position in the reference's (physically swapped) ordering -- what
    // ties are broken by and what ipiv records; a picked row goes to position J + c when the panel is written back.
    const int lrow = tid & (CL_ROWS - 1), half = tid >> 8;
    const int r0 = J + int(rank) * R;
    const int nrows = max(0, min(n, r0 + R) - r0);
    const bool has_row = lrow < nrows;
    int pos = r0 + lrow;
    bool active = has_row;                                  // not picked yet
    T a[CL_HC];
    if (jb == PW) {
#pragma unroll 8
        for (int idx = tid; idx < nrows * PW; idx += CL_WORKERS)          // coalesced: 64 consecutive columns per row
            stg[(idx >> 6) * PLDS + (idx & 63)] = A[size_t(r0 + (idx >> 6)) * ld + J + (idx & 63)];
    } else {
        for (int idx = tid; idx < nrows * jb; idx += CL_WORKERS) {
            const int r = idx / jb, cc = idx - r * jb;
            stg[r * PLDS + cc] = A[size_t(r0 + r) * ld + J + cc];
        }
    }
    worker_sync();
#pragma unroll
    for (int k = 0; k < CL_HC; ++k) a[k] = (has_row && half * CL_HC + k < jb) ? stg[lrow * PLDS + half * CL_HC + k] : T(0);
    // the hub's bookkeeping runs inside the column loop, in the shadow of the packet exchange (see below)
    int nf_l = 0, nf_g = 0;
    constexpr int LOCAL_FOLD_WARP = CL_WORKERS / 32 - 1, PLAN_FOLD_WARP = CL_WORKERS / 32 - 2;
    if (warp == LOCAL_FOLD_WARP) {
        for (int i = lane; i < jb; i += 32) od_l[i] = i;
    } else if (warp == PLAN_FOLD_WARP && rank == 0) {
        if (J == J0) {
            for (int i = lane; i < w; i += 32) st.od[i] = i;
        } else {
            for (int i = lane; i < LASWP_MAXJB; i += 32) {
                st.od[i] = sc.state->od[i];
                st.fr[i] = sc.state->fr[i];
                st.of[i] = sc.state->of[i];
            }
            nf_g = sc.state->nf;
        }
    }
    if (half == 0) colv[lrow] = a[0];
    // candidates of column 0 (later columns get theirs from the rank-1 update)
    {
        unsigned long long bkey = 0ull;
        int bidx = INT_MAX;
        if (active && half == 0) {
            const T av = fabs(a[0]);
            if (av == av) { bkey = key_of(double(av)); bidx = pos; }                     // a NaN never wins ...
            else if (pos == J) { bkey = key_of(CUDART_INF); bidx = pos; }                  // ... unless it is the diagonal
        }
        unsigned wm;
        warp_argmax(bkey, bidx, wm);
        if (lane == 0) { red_key[warp] = bkey; red_idx[warp] = bidx; }
    }
    worker_sync();

    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 2] = gtime();
    bool singular = false;
    for (int c4 = 0; c4 < jb; c4 += 4) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int c = c4 + s4;
            if (c >= jb || singular) continue;
            const int d = J + c;
            const int par = c & 1;
            const int rel = c - half * CL_HC;               // column c relative to my 32 (negative: all of mine are right of it)
            const T colc = colv[lrow];                      // my row's value in column c
            CTRACE(0);
            // ---- my CTA's candidate: reduce, stage its row, THEN announce it ----
            unsigned long long ckey = 0ull;
            int cidx = INT_MAX;
            if (warp == 0) {
                ckey = (lane < CL_WORKERS / 32) ? red_key[lane] : 0ull;
                cidx = (lane < CL_WORKERS / 32) ? red_idx[lane] : INT_MAX;
                unsigned wm;
                warp_argmax(ckey, cidx, wm);
                if (lane == 0) sh_cpos = cidx;
            }
            worker_sync();                                  // A: sh_cpos
            if (active && pos == sh_cpos && rel < CL_HC) {  // the two threads that hold the candidate row
#pragma unroll
                for (int k = 0; k < CL_HC; ++k) crow[par][half * CL_HC + k] = a[k];
            }
            worker_sync();                                  // A2: the row is in shared memory before the packet that
                                                            //     announces it leaves (remote readers fetch it after the verdict)
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(smem_u32(&bar[par]), 16u * CS);
                __syncwarp();
                if (unsigned(lane) < CS)
                    st_async_v2(mapa(smem_u32(&cand[par][rank]), unsigned(lane)), ckey, (unsigned long long)(unsigned)cidx,
                                mapa(smem_u32(&bar[par]), unsigned(lane)));
                CTRACE(1);
            } else if (c > 0) {
                // while warp 0 waits for the packets: fold the previous pivot into the panel's own net permutation
                // (every CTA) and into the outer block's plan (CTA 0)
                if (warp == LOCAL_FOLD_WARP) fold_pivot(od_l, fr_l, of_l, nf_l, J, jb, d - 1, piv_sm[c - 1], lane);
                else if (warp == PLAN_FOLD_WARP && rank == 0) fold_pivot(st.od, st.fr, st.of, nf_g, J0, w, d - 1, piv_sm[c - 1], lane);
            }
            // ---- the verdict: every CTA reduces the same CS packets ----
            if (warp == 0) {
                CTRACE(5);
                mbar_wait(smem_u32(&bar[par]), unsigned(c >> 1) & 1u);
                CTRACE(6);
                const int q = min(lane, int(CS) - 1);
                unsigned long long gk = ((volatile Msg *)&cand[par][q])->lo;
                int gi = int(unsigned(((volatile Msg *)&cand[par][q])->hi));
                int gw = q;
                if (unsigned(lane) >= CS) { gk = 0ull; gi = INT_MAX; }       // lanes past the cluster size carry no candidate
                unsigned wm;
                warp_argmax(gk, gi, wm);
                gw = __shfl_sync(0xffffffffu, gw, wm ? __ffs(wm) - 1 : 0);
                if (lane == 0) {
                    const int sing = (T(__longlong_as_double((long long)gk)) < Eps<T>::v()) ? 1 : 0;   // lu.rs:179-183
                    sh_idx = gi;
                    sh_win = gw;
                    sh_sing = sing;
                    piv_sm[c] = gi;
                    if (rank == 0) {               // CTA 0 keeps the books
                        if (sing) *info = d + 1; else ipiv[d] = gi;
                        if (sc.trace) sc.trace[c * 8 + 7] = (unsigned long long)(unsigned)gi;
                    }
                }
            }
            worker_sync();                                  // B: verdict
            CTRACE(2);
            if (sh_sing) { singular = true; continue; }     // uniform over the cluster
            const int p = sh_idx;                           // position of the pivot row
            if (tid >= c && tid < jb) prow_s[tid] = ld_cluster(&crow[par][tid], unsigned(sh_win));
            // bookkeeping of the interchange d <-> p: the picked row retires to position d, the row at d moves to p
            if (active) {
                if (pos == p) { active = false; pos = d; }
                else if (pos == d) pos = p;
            }
            worker_sync();                                  // C: pivot row here
            CTRACE(3);
            // ---- multiplier (one IEEE division per row, computed by both halves), rank-1 update on registers with
            //      mul, sub; a[rel] <- m; the new value of column c+1 is kept aside ----
            T nxt = T(0);
            if (active && rel < CL_HC) {
                const T m = div_rn(colc, prow_s[c]);
                const int gidx = rel >> 2;                  // floor(rel / 4); rel mod 4 == s4
                const T *u = prow_s + half * CL_HC;
#pragma unroll
                for (int gi = 0; gi < CL_HC / 4; ++gi) {
                    if (gi >= gidx) {                       // warp-uniform
                        const bool p_gt = gi > gidx, p_eq = gi == gidx;
#pragma unroll
                        for (int sub = 0; sub < 4; ++sub) {
                            const int k = 4 * gi + sub;
                            const bool act = p_gt || (p_eq && sub > s4);
                            if (act) a[k] = sub_rn(a[k], mul_rn(m, u[k]));
                            if (p_eq && sub == s4) a[k] = m;
                            const bool cap = (s4 < 3) ? (p_eq && sub == s4 + 1) : (gi == gidx + 1 && sub == 0);
                            if (cap) nxt = a[k];
                        }
                    }
                }
            }
            // ---- candidates of column c+1, and its values for the half that does not hold it ----
            if (c + 1 < jb) {
                unsigned long long bkey = 0ull;
                int bidx = INT_MAX;
                if (half == ((c + 1) >> 5)) {
                    colv[lrow] = nxt;
                    if (active) {
                        const T av = fabs(nxt);
                        if (av == av) { bkey = key_of(double(av)); bidx = pos; }
                        else if (pos == d + 1) { bkey = key_of(CUDART_INF); bidx = pos; }     // NaN diagonal stays (lu.rs:170-171)
                    }
                }
                unsigned wm;
                warp_argmax(bkey, bidx, wm);
                if (lane == 0) { red_key[warp] = bkey; red_idx[warp] = bidx; }
            }
            worker_sync();                                  // E
            CTRACE(4);
        }
    }

    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 3] = gtime();
    if (!singular) {
        // write the panel back permuted: registers -> shared memory -> coalesced rows at their final positions
        if (has_row) {
#pragma unroll
            for (int k = 0; k < CL_HC; ++k) stg[lrow * PLDS + half * CL_HC + k] = a[k];
            if (half == 0) posv[lrow] = pos;
        }
        if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 4] = gtime();
        // ---- the hub's duties: the last pivot, then publish ----
        if (warp == PLAN_FOLD_WARP && rank == 0) {
            // the outer block's net-permutation plan (consumed by rowid_apply_kernel / laswp_apply_kernel)
            fold_pivot(st.od, st.fr, st.of, nf_g, J0, w, J + jb - 1, piv_sm[jb - 1], lane);
            if (J + jb != J0 + w) {
                for (int i = lane; i < LASWP_MAXJB; i += 32) {
                    sc.state->od[i] = st.od[i];
                    sc.state->fr[i] = st.fr[i];
                    sc.state->of[i] = st.of[i];
                }
                if (lane == 0) sc.state->nf = nf_g;
            } else {
                const int nt = w + nf_g;
                if (lane == 0) sc.plan->nt = nt;
                for (int i = lane; i < nt; i += 32) {
                    sc.plan->rows[i] = (i < w) ? J0 + i : st.fr[i - w];
                    sc.plan->origin[i] = (i < w) ? st.od[i] : st.of[i - w];
                }
            }
        }
        if (warp == LOCAL_FOLD_WARP) {
            // the panel's own net permutation, for the outer block's columns outside the panel
            fold_pivot(od_l, fr_l, of_l, nf_l, J, jb, J + jb - 1, piv_sm[jb - 1], lane);
            const int nt = jb + nf_l;
            for (int i = lane; i < nt; i += 32) {
                rows_l[i] = (i < jb) ? J + i : fr_l[i - jb];
                org_l[i] = (i < jb) ? od_l[i] : of_l[i - jb];
            }
            if (lane == 0) sh_nt = nt;
        }
        worker_sync();
        if (jb == PW) {
#pragma unroll 8
            for (int idx = tid; idx < nrows * PW; idx += CL_WORKERS)
                A[size_t(posv[idx >> 6]) * ld + J + (idx & 63)] = stg[(idx >> 6) * PLDS + (idx & 63)];
        } else {
            for (int idx = tid; idx < nrows * jb; idx += CL_WORKERS) {
                const int r = idx / jb, cc = idx - r * jb;
                A[size_t(posv[r]) * ld + J + cc] = stg[r * PLDS + cc];
            }
        }
        if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 5] = gtime();
        if (!(dbg & 1)) {
            const int na = J - J0, ncols = na + (J0 + w - J - jb);     // columns [J0, J) and [J + jb, J0 + w)
            const int cpc = (ncols + int(CS) - 1) / int(CS);
            const int c_lo = int(rank) * cpc, c_hi = min(ncols, c_lo + cpc);
            const int nt = sh_nt;
            for (int g0 = c_lo; g0 < c_hi; g0 += CL_GC) {             // uniform per CTA
                T v[4];
                bool mv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = tid + CL_WORKERS * u, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                    mv[u] = i < nt && t2 < c_hi && org_l[i] != i;
                    if (mv[u]) {
                        const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                        v[u] = A[size_t(rows_l[org_l[i]]) * ld + col];
                    }
                }
                worker_sync();                                          // every source is read before any destination is written
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = tid + CL_WORKERS * u, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                    if (mv[u]) {
                        const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                        A[size_t(rows_l[i]) * ld + col] = v[u];
                    }
                }
            }
        }
    }
    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 6] = gtime();
    cluster_sync_all();
    if (sc.trace && rank == 0 && tid == 0 && w == PW) sc.trace[1536 + 7] = gtime();
}

// -------------------------------------------------------------------------------------------
// Cluster panel kernel, second generation ("pushed rows").  Same data placement idea as lu_panel_cluster_kernel
// (panel in the registers of one thread-block cluster, implicit pivoting, identical arithmetic and verdicts =>
// bit-identical factors), but the per-column dependency chain is cut from five CTA barriers + a DSMEM round trip
// (~1.5 us per column) to two one-way DSMEM hops and no CTA-wide barrier at all:
//   * a row lives in the two lanes 2j, 2j+1 of ONE warp (lane parity h holds the columns of parity h), so the value
//     in the next pivot column travels between them by shuffle and both know at once whether their row is the
//     warp's candidate;
//   * only column c+1 is brought up to date before the pivot search (one mul + one sub); the candidates are reduced
//     (redux.sync), posted to warp 0 with bar.arrive, and the rest of the rank-1 update runs in the shadow of the
//     exchange;
//   * every warp stages its own candidate row (post-update) in a per-warp slot, speculatively; warp 0 of every CTA
//     reduces the 16 warp candidates, sends the CTA's packet to every CTA (st.async + mbarrier complete_tx, as
//     before) and reaches the verdict; warp 0 of the WINNER CTA then PUSHES the winner's staged row + a 16-byte
//     verdict header into every CTA's pivot-row buffer as 16-byte st.async chunks that complete tx bytes on the
//     destination's row mbarrier.  Nobody pulls, nobody waits on a CTA barrier: all 512 threads of every CTA sleep
//     on the row mbarrier and wake with the pivot row in their shared memory;
//   * the IEEE division per row (the multiplier) left the critical path: the staged row carries RN(1/candidate),
//     computed in the shadow of the exchange, and every row forms colc / pivot from it with five FMAs (div_via_rcp:
//     correctly rounded, bit-identical to the division).
// Buffers (slots, pivot rows, packets, mbarriers) alternate by column parity; a CTA can be at most one column ahead.
// -------------------------------------------------------------------------------------------
// Correctly rounded a / b from a correctly rounded reciprocal y = RN(1/b) that is computed ONCE per pivot (by the lane
// that stages the candidate row, in the shadow of the exchange) instead of a full IEEE division per row on the
// critical path:  q0 = a*y;  q1 = q0 + (a - b*q0)*y  (faithful);  q2 = q1 + (a - b*q1)*y = RN(a/b)  (Markstein's
// theorem: a faithful quotient corrected once with the exactly computed residual and the correctly rounded reciprocal
// is the correctly rounded quotient).  The residuals are exact FMAs as long as nothing under- or overflows, which the
// exponent guard ensures; everything else (zeros, subnormals, huge / tiny magnitudes, Inf, NaN) takes the IEEE division.
// Bit-identity with __ddiv_rn / __fdiv_rn is asserted over 2^32 operand pairs in tests (rla_debug_divcheck).
__device__ __forceinline__ double rcp_rn(double b) { return __drcp_rn(b); }
__device__ __forceinline__ float rcp_rn(float b) { return __frcp_rn(b); }
__device__ __forceinline__ bool div_fast_ok(double a, double b) {
    const unsigned ea = (unsigned(__double2hiint(a)) >> 20) & 0x7ffu, eb = (unsigned(__double2hiint(b)) >> 20) & 0x7ffu;
    return ea - 0x300u < 0x200u && eb - 0x300u < 0x200u;          // both in [2^-255, 2^256)
}
__device__ __forceinline__ bool div_fast_ok(float a, float b) {
    const unsigned ea = (__float_as_uint(a) >> 23) & 0xffu, eb = (__float_as_uint(b) >> 23) & 0xffu;
    return ea - 0x60u < 0x40u && eb - 0x60u < 0x40u;              // both in [2^-31, 2^33)
}
__device__ __forceinline__ double div_via_rcp(double a, double b, double y) {
    if (!div_fast_ok(a, b)) return __ddiv_rn(a, b);
    const double q0 = __dmul_rn(a, y);
    const double q1 = __fma_rn(__fma_rn(-b, q0, a), y, q0);
    return __fma_rn(__fma_rn(-b, q1, a), y, q1);
}
__device__ __forceinline__ float div_via_rcp(float a, float b, float y) {
    if (!div_fast_ok(a, b)) return __fdiv_rn(a, b);
    const float q0 = __fmul_rn(a, y);
    const float q1 = __fmaf_rn(__fmaf_rn(-b, q0, a), y, q0);
    return __fmaf_rn(__fmaf_rn(-b, q1, a), y, q1);
}
// brute-force check of the above against the IEEE division (development / test aid)
template <typename T>
__global__ void divcheck_kernel(unsigned long long seed, unsigned long long per_thread, int mode, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    unsigned long long x = splitmix64(seed + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull);
    for (unsigned long long i = 0; i < per_thread; ++i) {
        x = splitmix64(x);
        const unsigned long long r1 = x;
        x = splitmix64(x);
        const unsigned long long r2 = x;
        T a, b;
        if (sizeof(T) == 8) {
            double da, db;
            if (mode == 0) {            // arbitrary bit patterns (all exponents, NaN, Inf, subnormals)
                da = __longlong_as_double((long long)r1);
                db = __longlong_as_double((long long)r2);
            } else {                    // LU-like: |a| <= |b|, magnitudes spread over 2^-40 .. 2^8, random mantissas and signs
                const int eb = 1023 - 40 + int((r2 >> 52) % 48);
                db = __longlong_as_double((long long)((r2 & 0x800fffffffffffffull) | ((unsigned long long)eb << 52)));
                const int ea = eb - int((r1 >> 52) % 60);
                da = __longlong_as_double((long long)((r1 & 0x800fffffffffffffull) | ((unsigned long long)(ea > 1 ? ea : 1) << 52)));
            }
            a = T(da); b = T(db);
        } else {
            float fa, fb;
            if (mode == 0) {
                fa = __uint_as_float(unsigned(r1));
                fb = __uint_as_float(unsigned(r2));
            } else {
                const int eb = 127 - 20 + int((r2 >> 32) % 28);
                fb = __uint_as_float((unsigned(r2) & 0x807fffffu) | (unsigned(eb) << 23));
                const int ea = eb - int((r1 >> 32) % 30);
                fa = __uint_as_float((unsigned(r1) & 0x807fffffu) | (unsigned(ea > 1 ? ea : 1) << 23));
            }
            a = T(fa); b = T(fb);
        }
        const T want = div_rn(a, b);
        const T got = div_via_rcp(a, b, rcp_rn(b));
        const bool same = (sizeof(T) == 8) ? (__double_as_longlong(double(want)) == __double_as_longlong(double(got)))
                                           : (__float_as_uint(float(want)) == __float_as_uint(float(got)));
        const bool both_nan = (want != want) && (got != got);
        if (!same && !both_nan) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

template <int V>
struct IC { static constexpr int value = V; };           // integral constant for generic-lambda dispatch

template <typename T>
struct RowLay {
    static constexpr int HALF = PW / 2;                       // columns per lane
    static constexpr int PADE = 16 / int(sizeof(T));          // 16 bytes between the two halves: the lanes of a pair hit
    static constexpr int ODD0 = HALF + PADE;                  //   different banks when they read "their" half
    static constexpr int ELEMS = 2 * HALF + PADE;             // staged row: [even columns][pad: RN(1/candidate)][odd columns]
    static constexpr int RCP = HALF;                          // the reciprocal of the candidate value sits in the pad
    static constexpr int TOTAL = ELEMS + 16 / int(sizeof(T)); // + verdict header {pivot position, singular}
    static constexpr unsigned ROW_BYTES = unsigned(ELEMS * sizeof(T));
    static constexpr unsigned TX_BYTES = ROW_BYTES + 16u;
};
constexpr int C2_BAR_CAND = 1, C2_BAR_STAGED = 2;             // named barriers (0 is __syncthreads)

__device__ __forceinline__ void named_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(CL_THREADS) : "memory"); }
__device__ __forceinline__ void named_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(CL_THREADS) : "memory"); }
// trace stamps (lu_dbg bit 3; separate instantiation so the production kernel carries none of it): warp 0 of CTA
// rank 0, slot s of column tc_
#define C2TRACE(tc_, slot)                                              \
    do {                                                                \
        if (TRACE && trace && rank == 0 && warp == 0) {                 \
            const unsigned long long t_ = gtime();                      \
            if (lane == 0) trace[(tc_) * 8 + (slot)] = t_;              \
            __syncwarp();                                               \
        }                                                               \
    } while (0)

// stamps of three other warps of CTA rank 0 (pages 1..3 of the trace buffer): when do THEY see the row, post, stage
#define C2TRACE_W(tc_, slot)                                                                   \
    do {                                                                                       \
        if (TRACE && trace && rank == 0 && (warp == 5 || warp == 10 || warp == 15)) {          \
            const unsigned long long t_ = gtime();                                             \
            if (lane == 0) trace[(warp / 5) * 512 + (tc_) * 8 + (slot)] = t_;                  \
            __syncwarp();                                                                      \
        }                                                                                      \
    } while (0)

// Shared state of the exchange.  One struct so the out-of-line chain functions below take a single pointer: the column
// loop is unrolled 8x (static register indices) and everything inlined into it is replicated 8 times -- with the chain
// inline the kernel was 186 KB of code and ran out of instruction cache (every column step fetched its instructions
// from L2: 2.4 us per column).
template <typename T>
struct C2Shared {
    Msg cand[2][CL_MAX];                                 // [parity][source rank], written remotely (st.async)
    unsigned long long bar_pk[2];                        // packets: 1 arrival (mine) + 16*CS tx bytes
    unsigned long long bar_row[2];                       // pivot row: 1 arrival (mine) + TX_BYTES
    T slot[2][CL_THREADS / 32][RowLay<T>::ELEMS];        // [parity][warp]: the warp's candidate row
    T prow[2][RowLay<T>::TOTAL];                         // [parity]: the pivot row + verdict header (pushed)
    unsigned long long red_key[CL_THREADS / 32];
    int red_idx[CL_THREADS / 32];
    unsigned long long best[2];                          // [parity]: running maximum of the warps' candidate keys (staging filter)
};

// warp 0, first half: reduce the 16 warp candidates, send the CTA's candidate to every CTA.  Returns the warp that
// holds the CTA's candidate.
template <typename T>
__device__ __noinline__ void c2_chain_send(C2Shared<T> *sh, int par, unsigned rank, unsigned CS) {
    constexpr int NW = CL_THREADS / 32;
    const int lane = threadIdx.x & 31;
    named_sync(C2_BAR_CAND);                 // every warp's candidate is in red_key / red_idx
    // reset the staging filter of the NEXT column: its last users passed this barrier two columns ago, its next users
    // wake only after the row that this CTA's packet (sent below) helps to decide has travelled back
    if (lane == 0) sh->best[par ^ 1] = 0ull;
    unsigned long long ckey = (lane < NW) ? sh->red_key[lane] : 0ull;
    int cidx = (lane < NW) ? sh->red_idx[lane] : INT_MAX;
    unsigned wmk;
    warp_argmax(ckey, cidx, wmk);
    const int ww = wmk ? __ffs(wmk) - 1 : 0; // lane index == warp index of the CTA's candidate
    if (lane == 0) mbar_expect_tx(smem_u32(&sh->bar_pk[par]), 16u * CS);
    __syncwarp();
    if (unsigned(lane) < CS)
        st_async_v2(mapa(smem_u32(&sh->cand[par][rank]), unsigned(lane)), ckey,
                    (unsigned long long)(unsigned)cidx | ((unsigned long long)(unsigned)ww << 32),
                    mapa(smem_u32(&sh->bar_pk[par]), unsigned(lane)));
}
// warp 0, second half: verdict; the winner CTA pushes [row | header] into every CTA's pivot-row buffer as 16-byte
// st.async chunks (lane l sends chunk l to every CTA).  (cp.async.bulk was measured here first: no faster.)
template <typename T, bool TRACE>
__device__ __noinline__ void c2_chain_finish(C2Shared<T> *sh, int cn, int par, unsigned rank, unsigned CS, unsigned long long *trace) {
    using RL = RowLay<T>;
    const int lane = threadIdx.x & 31, warp = 0;
    mbar_wait(smem_u32(&sh->bar_pk[par]), unsigned(cn >> 1) & 1u);
    C2TRACE(cn, 3);
    const int q = min(lane, int(CS) - 1);
    unsigned long long gk = ((volatile Msg *)&sh->cand[par][q])->lo;
    const unsigned long long hi = ((volatile Msg *)&sh->cand[par][q])->hi;
    int gi = int(unsigned(hi));
    int gslot = int(unsigned(hi >> 32)), gw = q;
    if (unsigned(lane) >= CS) { gk = 0ull; gi = INT_MAX; }   // lanes past the cluster size carry no candidate
    unsigned wm2;
    warp_argmax(gk, gi, wm2);
    const int src_lane = wm2 ? __ffs(wm2) - 1 : 0;
    gw = __shfl_sync(0xffffffffu, gw, src_lane);
    gslot = __shfl_sync(0xffffffffu, gslot, src_lane);
    const int sing = (T(__longlong_as_double((long long)gk)) < Eps<T>::v()) ? 1 : 0;   // lu.rs:179-183
    if (lane == 0) mbar_expect_tx(smem_u32(&sh->bar_row[par]), RL::TX_BYTES);
    if (unsigned(gw) != rank) {              // not my row: nothing to wait for (the barrier still needs my arrival)
        named_arrive(C2_BAR_STAGED);
        C2TRACE(cn, 4);
        C2TRACE(cn, 5);
        return;
    }
    named_sync(C2_BAR_STAGED);               // every warp of this CTA has staged its candidate row
    C2TRACE(cn, 4);
    {
        constexpr int NCH = int(RL::TX_BYTES / 16);                         // 34 (f64) / 18 (f32)
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&sh->slot[par][gslot][0]);
#pragma unroll
        for (int ch0 = 0; ch0 < NCH; ch0 += 32) {
            const int ch = ch0 + lane;
            if (ch < NCH) {
                unsigned long long v0, v1;
                if (ch < NCH - 1) { v0 = src[2 * ch]; v1 = src[2 * ch + 1]; }
                else { v0 = (unsigned long long)(unsigned)gi; v1 = (unsigned long long)(unsigned)sing; }
                const unsigned dst = smem_u32(&sh->prow[par][0]) + 16u * unsigned(ch), bar = smem_u32(&sh->bar_row[par]);
                for (unsigned r = 0; r < CS; ++r) st_async_v2(mapa(dst, r), v0, v1, mapa(bar, r));
            }
        }
    }
    C2TRACE(cn, 5);
}
__device__ __noinline__ void fold_pivot_ool(int *od, int *fr, int *of, int *nf, int base, int w, int d, int p) {
    int v = *nf;
    fold_pivot(od, fr, of, v, base, w, d, p, int(threadIdx.x & 31));
    *nf = v;
}

template <typename T, bool TRACE>
__global__ void __launch_bounds__(CL_THREADS, 1)
lu_panel_cluster2_kernel(T *__restrict__ A, size_t ld, int n, int J, int jb, int R, int32_t *__restrict__ ipiv,
                         int32_t *__restrict__ info, PanelScratch sc, int J0, int w, int dbg) {
    if (*info != 0) return;   // written by an earlier kernel => uniform over the cluster
    using RL = RowLay<T>;
    extern __shared__ __align__(16) unsigned char panel_smem[];   // [CL_ROWS][PLDS] staging for coalesced panel load / store
    T *stg = reinterpret_cast<T *>(panel_smem);
    __shared__ __align__(16) C2Shared<T> shx;
    C2Shared<T> *sh = &shx;
    __shared__ int sh_nt;
    __shared__ PlanState st;                                      // CTA rank 0: the outer block's plan state
    __shared__ int od_l[PW], fr_l[PW], of_l[PW];                  // the panel's own net permutation
    __shared__ int rows_l[2 * PW], org_l[2 * PW];
    __shared__ int posv[CL_ROWS];                                 // final position of my rows (write-back)
    __shared__ int nf_sh[2];                                      // far-row counts of the two folds
    unsigned long long *const trace = sc.trace;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = cluster_ctarank(), CS = cluster_nctarank();
    if (tid == 0) {
        mbar_init(smem_u32(&sh->bar_pk[0]), 1);
        mbar_init(smem_u32(&sh->bar_pk[1]), 1);
        mbar_init(smem_u32(&sh->bar_row[0]), 1);
        mbar_init(smem_u32(&sh->bar_row[1]), 1);
        sh->best[0] = sh->best[1] = 0ull;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();

    // lane pair (2j, 2j+1) = one row; lane parity h holds columns 2k + h, k = 0..31, in registers for the whole panel.
    // Rows never move: `pos` is the row's position in the reference's (physically swapped) ordering.
    const int lrow = tid >> 1, h = tid & 1;
    const int r0 = J + int(rank) * R;
    const int nrows = max(0, min(n, r0 + R) - r0);
    const bool has_row = lrow < nrows;
    int pos = r0 + lrow;
    bool active = has_row;
    T a[RL::HALF];
    if (jb == PW) {
#pragma unroll 8
        for (int idx = tid; idx < nrows * PW; idx += CL_THREADS)
            stg[(idx >> 6) * PLDS + (idx & 63)] = A[size_t(r0 + (idx >> 6)) * ld + J + (idx & 63)];
    } else {
        for (int idx = tid; idx < nrows * jb; idx += CL_THREADS) {
            const int r = idx / jb, cc = idx - r * jb;
            stg[r * PLDS + cc] = A[size_t(r0 + r) * ld + J + cc];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RL::HALF; ++k) a[k] = (has_row && 2 * k + h < jb) ? stg[lrow * PLDS + 2 * k + h] : T(0);

    constexpr int NW = CL_THREADS / 32, LOCAL_FOLD_WARP = NW - 1, PLAN_FOLD_WARP = NW - 2;
    if (warp == LOCAL_FOLD_WARP) {
        for (int i = lane; i < jb; i += 32) od_l[i] = i;
        if (lane == 0) nf_sh[0] = 0;
    } else if (warp == PLAN_FOLD_WARP && rank == 0) {
        if (J == J0) {
            for (int i = lane; i < w; i += 32) st.od[i] = i;
            if (lane == 0) nf_sh[1] = 0;
        } else {
            for (int i = lane; i < LASWP_MAXJB; i += 32) {
                st.od[i] = sc.state->od[i];
                st.fr[i] = sc.state->fr[i];
                st.of[i] = sc.state->of[i];
            }
            if (lane == 0) nf_sh[1] = sc.state->nf;
        }
    }
    __syncwarp();

    // ---- the pieces of one column step -----------------------------------------------------------------------
    // candidates of column cn from `val` (valid in the lanes with h == (cn & 1)); posts the warp's candidate
    auto post_candidate = [&](int cn, T val, unsigned &wm) {
        unsigned long long bkey = 0ull;
        int bidx = INT_MAX;
        if (active && h == (cn & 1)) {
            const T av = fabs(val);
            if (av == av) { bkey = key_of(double(av)); bidx = pos; }                     // a NaN never wins ...
            else if (pos == J + cn) { bkey = key_of(CUDART_INF); bidx = pos; }             // ... unless it is the diagonal (lu.rs:170-171)
        }
        warp_argmax(bkey, bidx, wm);
        // staging filter: only a warp whose candidate is not already beaten by an earlier poster stages its row (the CTA's
        // eventual winner -- maximal key, lowest position among equals -- always passes: nothing posted can exceed it)
        unsigned long long seen = 0ull;
        if (lane == 0) {
            sh->red_key[warp] = bkey;
            sh->red_idx[warp] = bidx;
            seen = atomicMax(&sh->best[cn & 1], bkey);
        }
        seen = __shfl_sync(0xffffffffu, seen, 0);
        if (bkey < seen) wm = 0u;
    };
    // my (post-update) row into the warp's slot if it is the warp's candidate
    // (the lane that holds the candidate VALUE also leaves RN(1/value) in the pad: the next column's multipliers are
    // formed from it, see div_via_rcp)
    auto stage_row = [&](int par, unsigned wm, int cn, T val) {
        if ((wm >> (lane & ~1)) & 3u) {
            T *dst = &sh->slot[par][warp][h * RL::ODD0];
#pragma unroll
            for (int k = 0; k < RL::HALF; k += int(16 / sizeof(T))) {
                if constexpr (sizeof(T) == 8) *reinterpret_cast<double2 *>(dst + k) = make_double2(a[k], a[k + 1]);
                else *reinterpret_cast<float4 *>(dst + k) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
            }
            if (h == (cn & 1)) sh->slot[par][warp][RL::RCP] = rcp_rn(val);
        }
        __syncwarp();
    };

    // ---- prologue: candidates of column 0 from the raw panel ----
    {
        unsigned wm;
        post_candidate(0, a[0], wm);
        if (warp == 0) c2_chain_send<T>(sh, 0, rank, CS); else named_arrive(C2_BAR_CAND);
        stage_row(0, wm, 0, a[0]);
        if (warp == 0) c2_chain_finish<T, TRACE>(sh, 0, 0, rank, CS, trace); else named_arrive(C2_BAR_STAGED);
    }
    T colc;                                      // my row's value in the current pivot column (both lanes of the pair)
    T a1;                                        // my value in the NEXT pivot column before this column's update (its holder lane)
    {
        const T mine = a[0], other = __shfl_xor_sync(0xffffffffu, mine, 1);
        colc = (h == 0) ? mine : other;
        a1 = a[0];                               // column 1 = element 0 of the odd lane
    }

    bool singular = false;
    // One pivot column.  S8 = c mod 8 is a compile-time constant so that every register index below is static:
    // column c = 2*kc + s with kc = 4*gidx + sub0, s = S8 & 1, sub0 = S8 >> 1 static and gidx = c / 8 warp-uniform.
    auto column_step = [&](auto S8, const int c8) {
        constexpr int s8 = decltype(S8)::value;
        constexpr int s = s8 & 1, sub0 = s8 >> 1;
        const int c = c8 + s8;
        const int d = J + c;
        const int par = c & 1;
        const int gidx = c8 >> 3;
        const int kc = 4 * gidx + sub0;
        // ---- the pivot row of column c (and the verdict) has been pushed into prow[par] ----
        mbar_wait(smem_u32(&sh->bar_row[par]), unsigned(c >> 1) & 1u);
        C2TRACE(c, 0);
        C2TRACE_W(c, 0);
        const T *u = sh->prow[par];
        const ulonglong2 hv = *reinterpret_cast<const ulonglong2 *>(&sh->prow[par][RL::ELEMS]);
        const int p = int(unsigned(hv.x));
        if (hv.y != 0ull) {                                 // |pivot| < eps: uniform over the cluster
            if (rank == 0 && tid == 0) *info = d + 1;
            singular = true;
            return;
        }
        if (rank == 0 && tid == 0) ipiv[d] = p;
        // bookkeeping of the interchange d <-> p: the picked row retires to position d, the row at d moves to p
        if (pos == p) { active = false; pos = d; }
        else if (pos == d) pos = p;
        const T *uh = u + h * RL::ODD0;                    // my half of the pivot row: uh[k] = U[c][2k + h]
        const T piv = u[s * RL::ODD0 + kc];
        const T m = active ? div_via_rcp(colc, piv, u[RL::RCP]) : T(0);   // == colc / piv, IEEE-rounded (both lanes compute it)
        const bool more = c + 1 < jb;
        unsigned wm = 0;
        T nxt = T(0);
        if (more) {
            // ---- column c+1 first (one mul + one sub from the value captured one step earlier), candidates, post ----
            const int k1 = kc + s;                          // column c+1 = 2*k1 + (1 - s)
            if (active && h == 1 - s) nxt = sub_rn(a1, mul_rn(m, uh[k1]));
            const T other = __shfl_xor_sync(0xffffffffu, nxt, 1);
            colc = (h == 1 - s) ? nxt : other;              // both lanes: my row's value in column c+1
            post_candidate(c + 1, nxt, wm);
            C2TRACE(c + 1, 6);
            C2TRACE_W(c + 1, 1);
            if (warp == 0) c2_chain_send<T>(sh, par ^ 1, rank, CS); else named_arrive(C2_BAR_CAND);
            C2TRACE(c + 1, 1);
        }
        // ---- the rank-1 update with mul, sub, in the shadow of the exchange.  Groups of 4 right of the pivot group:
        //      one indexed jump, then straight-line code.  The pivot group (static sub0) looks at the lane parity, stores
        //      the multiplier in the lane that holds column c, and captures column c+2 for the next step. ----
        if (active) {
            auto upd4 = [&](auto G) {
                constexpr int k = 4 * decltype(G)::value;
                if constexpr (sizeof(T) == 8) {
                    const double2 u0 = *reinterpret_cast<const double2 *>(uh + k), u1 = *reinterpret_cast<const double2 *>(uh + k + 2);
                    a[k] = sub_rn(a[k], mul_rn(m, T(u0.x)));
                    a[k + 1] = sub_rn(a[k + 1], mul_rn(m, T(u0.y)));
                    a[k + 2] = sub_rn(a[k + 2], mul_rn(m, T(u1.x)));
                    a[k + 3] = sub_rn(a[k + 3], mul_rn(m, T(u1.y)));
                } else {
                    const float4 uu = *reinterpret_cast<const float4 *>(uh + k);
                    a[k] = sub_rn(a[k], mul_rn(m, T(uu.x)));
                    a[k + 1] = sub_rn(a[k + 1], mul_rn(m, T(uu.y)));
                    a[k + 2] = sub_rn(a[k + 2], mul_rn(m, T(uu.z)));
                    a[k + 3] = sub_rn(a[k + 3], mul_rn(m, T(uu.w)));
                }
            };
            switch (gidx) {
                case 0: upd4(IC<1>{}); [[fallthrough]];
                case 1: upd4(IC<2>{}); [[fallthrough]];
                case 2: upd4(IC<3>{}); [[fallthrough]];
                case 3: upd4(IC<4>{}); [[fallthrough]];
                case 4: upd4(IC<5>{}); [[fallthrough]];
                case 5: upd4(IC<6>{}); [[fallthrough]];
                case 6: upd4(IC<7>{}); [[fallthrough]];
                default: break;
            }
            auto pivgroup = [&](auto G) {
                constexpr int g = decltype(G)::value;
#pragma unroll
                for (int sub = sub0 + 1; sub < 4; ++sub) a[4 * g + sub] = sub_rn(a[4 * g + sub], mul_rn(m, uh[4 * g + sub]));
                constexpr int k = 4 * g + sub0;
                if constexpr (s == 0) a[k] = h ? sub_rn(a[k], mul_rn(m, uh[k])) : m;   // column c is even: the odd lane still updates 2k+1
                else { if (h) a[k] = m; }                                               // column c is odd: it is the odd lane's own
                if constexpr (k + 1 < RL::HALF) a1 = a[k + 1];                           // column c+2 = 2(kc+1) + s, after this update
            };
            switch (gidx) {
                case 0: pivgroup(IC<0>{}); break;
                case 1: pivgroup(IC<1>{}); break;
                case 2: pivgroup(IC<2>{}); break;
                case 3: pivgroup(IC<3>{}); break;
                case 4: pivgroup(IC<4>{}); break;
                case 5: pivgroup(IC<5>{}); break;
                case 6: pivgroup(IC<6>{}); break;
                default: pivgroup(IC<7>{}); break;
            }
        }
        if (more) {
            stage_row(par ^ 1, wm, c + 1, nxt);
            C2TRACE(c + 1, 2);
            C2TRACE_W(c + 1, 2);
            if (warp == 0) c2_chain_finish<T, TRACE>(sh, c + 1, par ^ 1, rank, CS, trace); else named_arrive(C2_BAR_STAGED);
        }
        // ---- the hub's bookkeeping, off the critical path ----
        if (warp == LOCAL_FOLD_WARP) fold_pivot_ool(od_l, fr_l, of_l, &nf_sh[0], J, jb, d, p);
        else if (warp == PLAN_FOLD_WARP && rank == 0) fold_pivot_ool(st.od, st.fr, st.of, &nf_sh[1], J0, w, d, p);
    };
    for (int c8 = 0; c8 < jb && !singular; c8 += 8) {
        column_step(IC<0>{}, c8);
        if (c8 + 1 < jb && !singular) column_step(IC<1>{}, c8);
        if (c8 + 2 < jb && !singular) column_step(IC<2>{}, c8);
        if (c8 + 3 < jb && !singular) column_step(IC<3>{}, c8);
        if (c8 + 4 < jb && !singular) column_step(IC<4>{}, c8);
        if (c8 + 5 < jb && !singular) column_step(IC<5>{}, c8);
        if (c8 + 6 < jb && !singular) column_step(IC<6>{}, c8);
        if (c8 + 7 < jb && !singular) column_step(IC<7>{}, c8);
    }


    if (!singular) {
        // write the panel back permuted: registers -> shared memory -> coalesced rows at their final positions
        if (has_row) {
#pragma unroll
            for (int k = 0; k < RL::HALF; ++k) stg[lrow * PLDS + 2 * k + h] = a[k];
            if (h == 0) posv[lrow] = pos;
        }
        if (warp == PLAN_FOLD_WARP && rank == 0) {
            if (J + jb != J0 + w) {
                for (int i = lane; i < LASWP_MAXJB; i += 32) {
                    sc.state->od[i] = st.od[i];
                    sc.state->fr[i] = st.fr[i];
                    sc.state->of[i] = st.of[i];
                }
                if (lane == 0) sc.state->nf = nf_sh[1];
            } else {
                const int nt = w + nf_sh[1];
                if (lane == 0) sc.plan->nt = nt;
                for (int i = lane; i < nt; i += 32) {
                    sc.plan->rows[i] = (i < w) ? J0 + i : st.fr[i - w];
                    sc.plan->origin[i] = (i < w) ? st.od[i] : st.of[i - w];
                }
            }
        }
        if (warp == LOCAL_FOLD_WARP) {
            const int nt = jb + nf_sh[0];
            for (int i = lane; i < nt; i += 32) {
                rows_l[i] = (i < jb) ? J + i : fr_l[i - jb];
                org_l[i] = (i < jb) ? od_l[i] : of_l[i - jb];
            }
            if (lane == 0) sh_nt = nt;
        }
        __syncthreads();
        if (jb == PW) {
#pragma unroll 8
            for (int idx = tid; idx < nrows * PW; idx += CL_THREADS)
                A[size_t(posv[idx >> 6]) * ld + J + (idx & 63)] = stg[(idx >> 6) * PLDS + (idx & 63)];
        } else {
            for (int idx = tid; idx < nrows * jb; idx += CL_THREADS) {
                const int r = idx / jb, cc = idx - r * jb;
                A[size_t(posv[r]) * ld + J + cc] = stg[r * PLDS + cc];
            }
        }
        if (!(dbg & 1)) {
            const int na = J - J0, ncols = na + (J0 + w - J - jb);     // columns [J0, J) and [J + jb, J0 + w)
            const int cpc = (ncols + int(CS) - 1) / int(CS);
            const int c_lo = int(rank) * cpc, c_hi = min(ncols, c_lo + cpc);
            const int nt = sh_nt;
            for (int g0 = c_lo; g0 < c_hi; g0 += CL_GC) {             // uniform per CTA
                T v[4];
                bool mv[4];
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int e = tid + CL_THREADS * uu, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                    mv[uu] = i < nt && t2 < c_hi && org_l[i] != i;
                    if (mv[uu]) {
                        const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                        v[uu] = A[size_t(rows_l[org_l[i]]) * ld + col];
                    }
                }
                __syncthreads();                                        // every source is read before any destination is written
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int e = tid + CL_THREADS * uu, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                    if (mv[uu]) {
                        const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                        A[size_t(rows_l[i]) * ld + col] = v[uu];
                    }
                }
            }
        }
    }
    cluster_sync_all();       // remote shared memory must stay alive while anyone may still write into it
}

// -------------------------------------------------------------------------------------------
// Column-slab panel kernel (K3d): panels of <= 3840 rows, LEFT-looking across CTAs.
//
// The row-distributed kernels above pay one cross-CTA exchange per COLUMN (the pivot search needs every CTA's candidate):
// ~2 us x 64 columns per panel whatever the panel height.  Here the 64 panel columns are dealt to the CTAs instead:
// CTA k owns columns [k*NC, (k+1)*NC) of ALL rows (ROWS rows x NC columns per thread in registers, ROWS*NC = 32,
// ROWS = 1/2/4/8 for panels of <= 480/960/1920/3840 rows => 2/4/8/16 CTAs), so the pivot search of a column is local to
// one CTA (two warp reductions and ONE CTA barrier per column, the candidate rows staged speculatively per warp) and the
// cross-CTA traffic is one-way: when a column is finished its owner publishes {pivot position, multiplier column}
// through L2 (multipliers: plain stores; then a release store of a 64-bit header epoch << 32 | position), and every CTA
// that owns later columns applies it (a[r][j] -= m[r] * u[j], u = its own columns' values in the pivot row, broadcast
// through shared memory with one barrier) at its own pace -- nobody ever answers.  The dependency chain of a panel is
// 64 local column steps + (number of CTAs) hand-overs instead of 64 exchanges.
//   * arithmetic and pivot rule are those of K3 / K3b (unfused multiply and subtract, updates of an element in
//     ascending column order, first maximum in the swapped ordering wins, NaN never wins unless it is the diagonal,
//     |pivot| < eps -> info): factors, permutation and status are bit-identical (asserted in tests);
//   * implicit pivoting as in K3b: rows never move, every CTA tracks each row's position in the reference's
//     swapped ordering; the panel is written back permuted at the end;
//   * a 16th warp per CTA keeps the books so the 15 worker warps never wait for anything but data: it polls the
//     headers of foreign columns (ld.acquire.gpu) and hands them to the workers through shared memory, publishes the
//     CTA's own columns (the workers count themselves done in shared memory; st.release.gpu by this warp is the only
//     gpu-scope fence in the kernel), and folds every pivot into the panel's net permutation (and, in CTA 0, into the
//     outer block's plan) as it goes;
//   * no cluster, no cooperative launch: CTAs take a ticket on entry and logical CTA k only ever waits for logical
//     CTAs < k, all of which are running by then.
// -------------------------------------------------------------------------------------------
constexpr int SL_WORKERS = 480;                  // 15 worker warps + the bookkeeping warp = 512 threads => 128 registers each
constexpr int SL_WARPS = SL_WORKERS / 32;        // (a 17-warp CTA is capped at 96: one scheduler would hold five warps)
constexpr int SL_THREADS = SL_WORKERS + 32;
constexpr int SL_MAXROWS = 8 * SL_WORKERS;       // 3840
constexpr int SL_GPASS = (2 * PW * CL_GC + SL_WORKERS - 1) / SL_WORKERS;   // closing gather: entries per thread and pass
struct SlabScratch {
    unsigned long long *hdr;     // [PW]  epoch << 32 | pivot position (0xffffffff: |pivot| < eps)
    unsigned *ticket;            // monotonic CTA counter (never reset; the host passes this launch's base)
    void *mbuf;                  // [PW][rstride] multiplier columns, indexed by row slot
    unsigned epoch, ticket_base;
    int rstride;
};
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_cta_s(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s(unsigned *p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_cta_s(unsigned *p, unsigned v) {
    asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void slab_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SL_WORKERS) : "memory"); }

template <typename T, int ROWS>
__global__ void __launch_bounds__(SL_THREADS, 1)
lu_panel_slab_kernel(T *__restrict__ A, size_t ld, int n, int J, int jb, int32_t *__restrict__ ipiv,
                     int32_t *__restrict__ info, PanelScratch sc, SlabScratch ss, int J0, int w, int dbg) {
    constexpr int NC = 32 / ROWS;
    constexpr int VN = 16 / int(sizeof(T));                 // elements per 16-byte shared-memory access
    struct alignas(16) Vec { T v[VN]; };
    __shared__ __align__(16) T ubuf[2][NC];                 // pivot row of a foreign column, my NC columns
    __shared__ __align__(16) T wrow[2][SL_WARPS][NC];       // own columns: every warp's candidate row (speculative)
    __shared__ unsigned long long wkey[2][SL_WARPS];
    __shared__ int widx[2][SL_WARPS];
    __shared__ int piv_sm[PW];                              // pivot position per column (-1: singular)
    __shared__ int own_piv[NC];
    __shared__ unsigned ready, done_cnt;                    // columns known (bookkeeper -> workers); worker warps done (-> bookkeeper)
    __shared__ int sh_k, sh_skip, sh_nt;
    __shared__ PlanState st;                                // hub CTA: the outer block's plan state
    __shared__ int dst_l[2 * PW], src_l[2 * PW];            // the panel's net permutation: row dst_l[e] <- old row src_l[e]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        sh_k = int(atomicAdd(ss.ticket, 1u) - ss.ticket_base);
        sh_skip = (*info != 0);
        ready = 0u;
        done_cnt = 0u;
        sh_nt = 0;
    }
    __syncthreads();
    if (sh_skip) return;
    const int k = sh_k;
    unsigned long long *const tr = (w == PW) ? sc.trace : nullptr;   // optional %globaltimer stamps (lu_dbg bit 3)
    if (tr && tid == 0) tr[1536 + 8 * k + 0] = gtime();
    const int ncta = (jb + NC - 1) / NC;                    // column CTAs; logical CTA ncta is the hub
    if (k >= ncta) {
        // ---------------- the hub CTA: one warp follows the headers and folds every pivot into the outer block's
        // net-permutation plan (consumed by rowid_apply_kernel / laswp_apply_kernel); nobody waits for it ----------------
        if (warp != 0) return;
        int nf_g = 0;
        if (J == J0) {
            for (int i = lane; i < w; i += 32) st.od[i] = i;
        } else {
            for (int i = lane; i < LASWP_MAXJB; i += 32) {
                st.od[i] = sc.state->od[i];
                st.fr[i] = sc.state->fr[i];
                st.of[i] = sc.state->of[i];
            }
            nf_g = sc.state->nf;
        }
        __syncwarp();
        for (int c = 0; c < jb; ++c) {
            unsigned long long h;
            do {
                h = ld_relaxed_gpu(&ss.hdr[c]);
                h = __shfl_sync(0xffffffffu, h, 0);
            } while (unsigned(h >> 32) != ss.epoch);
            const int p = int(unsigned(h));
            if (p < 0) return;                               // singular: the owner has set *info, the plan is not needed
            fold_pivot(st.od, st.fr, st.of, nf_g, J0, w, J + c, p, lane);
        }
        if (J + jb != J0 + w) {
            for (int i = lane; i < LASWP_MAXJB; i += 32) {
                sc.state->od[i] = st.od[i];
                sc.state->fr[i] = st.fr[i];
                sc.state->of[i] = st.of[i];
            }
            if (lane == 0) sc.state->nf = nf_g;
        } else {
            const int nt = w + nf_g;
            if (lane == 0) sc.plan->nt = nt;
            for (int i = lane; i < nt; i += 32) {
                sc.plan->rows[i] = (i < w) ? J0 + i : st.fr[i - w];
                sc.plan->origin[i] = (i < w) ? st.od[i] : st.of[i - w];
            }
        }
        return;
    }
    const int c0 = k * NC;                                  // my first column (panel-relative)
    const int nown = min(NC, jb - c0);
    const int nrows = n - J;
    const size_t RS = size_t(ss.rstride);
    T *const mbuf = static_cast<T *>(ss.mbuf);

    if (tid >= SL_WORKERS) {
        // ---------------- the bookkeeping warp ----------------
        // Columns become KNOWN in order: own columns when all worker warps have counted themselves done (every column
        // finished by then is published behind ONE gpu-scope fence), foreign columns when their header shows this
        // launch's epoch; known columns are handed to the workers at once.
        int nknown = 0;
        while (nknown < jb) {
            const int c = nknown;
            if (c >= c0 && c < c0 + nown) {
                const int ndone = int(ld_acquire_cta_s(&done_cnt)) / SL_WARPS;           // own columns complete
                const int nnew = c0 + ndone - c;
                if (nnew <= 0) continue;
                if (lane == 0 && tr) for (int q = 0; q < nnew; ++q) tr[(c + q) * 8 + 2] = gtime();
                if (lane < nnew) {
                    const int p = own_piv[c - c0 + lane];
                    if (p < 0) *info = J + c + lane + 1; else ipiv[J + c + lane] = p;       // lu.rs:179-183
                    piv_sm[c + lane] = p;
                }
                __syncwarp();
                __threadfence();                             // the workers' multiplier columns (cumulativity) before the headers
                int stop = nnew;
                bool sing = false;
                for (int q = 0; q < nnew; ++q) {
                    const int p = piv_sm[c + q];
                    if (lane == 0) st_relaxed_gpu(&ss.hdr[c + q], ((unsigned long long)ss.epoch << 32) | (unsigned long long)(unsigned)p);
                    if (p < 0) { sing = true; stop = q + 1; break; }
                }
                if (lane == 0 && tr) for (int q = 0; q < stop; ++q) { tr[(c + q) * 8 + 3] = gtime(); tr[(c + q) * 8 + 7] = (unsigned long long)(unsigned)piv_sm[c + q]; }
                nknown = c + stop;
                __syncwarp();
                if (lane == 0) st_release_cta_s(&ready, unsigned(nknown));
                if (sing) break;
            } else {
                unsigned long long h = ld_relaxed_gpu(&ss.hdr[c]);
                h = __shfl_sync(0xffffffffu, h, 0);
                if (unsigned(h >> 32) != ss.epoch) continue;
                __threadfence();                             // acquire: the multiplier column behind the header
                const int p = int(unsigned(h));
                if (tr && lane == 0 && c0 == (c / NC + 1) * NC) tr[c * 8 + 4] = gtime();   // the next owner saw it
                if (lane == 0) piv_sm[c] = p;
                __syncwarp();
                nknown = c + 1;
                if (lane == 0) st_release_cta_s(&ready, unsigned(nknown));
                if (p < 0) break;
            }
        }
        return;
    }

    // ---------------- workers: thread tid keeps row slots tid, tid + 480, ... of my NC columns ----------------
    // The column steps are ISSUE-bound (15 warps x a few hundred instructions per column on 4 schedulers), so the row
    // loops are written branch-free: predicated selects instead of per-row branches, floating-point comparisons for the
    // candidate search, one reciprocal per column + five FMAs per row (div_via_rcp: bit-identical to the IEEE division)
    // instead of a division per row when a thread holds four or more rows.
    int pos[ROWS];
    unsigned act = 0u;                                       // bit r: row r not picked yet
    T a[ROWS][NC];
    // My slab of the panel is nrows segments of NC*sizeof(T) bytes: moved between global memory and registers through
    // shared memory in 16-byte chunks so that every 32-byte sector is requested exactly once (a thread reading its own
    // row with scalar loads requests each sector four times: ~8 us per CTA instead of ~2).
    extern __shared__ __align__(16) unsigned char slab_smem[];
    T *const stg = reinterpret_cast<T *>(slab_smem);         // [row slot][NC + VN]
    constexpr int SLD = NC + VN, LPR = NC / VN;
    const bool vec_io = nown == NC && ((size_t(ld) * sizeof(T)) & 15) == 0 &&
                        (reinterpret_cast<size_t>(A + size_t(J) * ld + J + c0) & 15) == 0;
    if (vec_io) {
        for (int q = tid; q < nrows * LPR; q += SL_WORKERS) {
            const int row = q / LPR, part = q % LPR;
            *reinterpret_cast<Vec *>(stg + size_t(row) * SLD + part * VN) =
                *reinterpret_cast<const Vec *>(A + size_t(J + row) * ld + J + c0 + part * VN);
        }
        slab_worker_sync();
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int s = tid + SL_WORKERS * r;
        const bool has = s < nrows;
        pos[r] = J + s;
        if (has) act |= 1u << r;
        if (vec_io) {
#pragma unroll
            for (int jv = 0; jv < NC; jv += VN) {
                Vec t;
#pragma unroll
                for (int e = 0; e < VN; ++e) t.v[e] = T(0);
                if (has) t = *reinterpret_cast<const Vec *>(stg + size_t(s) * SLD + jv);
#pragma unroll
                for (int e = 0; e < VN; ++e) a[r][jv + e] = t.v[e];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC; ++j) a[r][j] = (has && j < nown) ? A[size_t(J + s) * ld + J + c0 + j] : T(0);
        }
    }
    T m[ROWS];
    bool sing = false;
    unsigned known = 0u;

    // ---- columns of the CTAs before me, one at a time, as they are published ----
    {
        bool have = false;                                   // m[] already holds column c (requested during column c - 1)
        for (int c = 0; c < c0; ++c) {
            if (unsigned(c) + 1u >= known) {
                known = ld_acquire_cta_s(&ready);
                while (known <= unsigned(c)) known = ld_acquire_cta_s(&ready);
            }
            const int p = piv_sm[c];
            if (p < 0) { sing = true; break; }
            if (!have) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) m[r] = __ldcg(mbuf + size_t(c) * RS + tid + SL_WORKERS * r);
            }
            const int d = J + c, par = c & 1;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const bool on = (act >> r) & 1u;
                const bool isp = on && pos[r] == p;          // I hold the pivot row: it retires to position d
                const bool isd = pos[r] == d;                // (only a live row can sit at d)
                if (isp) {
#pragma unroll
                    for (int jv = 0; jv < NC; jv += VN) {
                        Vec t;
#pragma unroll
                        for (int e = 0; e < VN; ++e) t.v[e] = a[r][jv + e];
                        *reinterpret_cast<Vec *>(&ubuf[par][jv]) = t;
                    }
                }
                pos[r] = isp ? d : (isd ? p : pos[r]);
                act &= ~(unsigned(isp) << r);
            }
            slab_worker_sync();
            if (tr && tid == 0 && c0 == (c / NC + 1) * NC) tr[c * 8 + 5] = gtime();        // the next owner's workers start applying it
            have = (c + 1 < c0) && (unsigned(c) + 1u < known);          // uniform per warp
            constexpr int UCH = NC < 8 ? NC : 8;
#pragma unroll
            for (int j0 = 0; j0 < NC; j0 += UCH) {
                T u[UCH];
#pragma unroll
                for (int jv = 0; jv < UCH; jv += VN) {
                    const Vec t = *reinterpret_cast<const Vec *>(&ubuf[par][j0 + jv]);
#pragma unroll
                    for (int e = 0; e < VN; ++e) u[jv + e] = t.v[e];
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    if ((act >> r) & 1u) {
#pragma unroll
                        for (int jj = 0; jj < UCH; ++jj) a[r][j0 + jj] = sub_rn(a[r][j0 + jj], mul_rn(m[r], u[jj]));
                    }
                    // the multiplier of this row in the next column: requested now, consumed one column later
                    if (j0 + UCH == NC && have) m[r] = __ldcg(mbuf + size_t(c + 1) * RS + tid + SL_WORKERS * r);
                }
            }
        }
    }

    // ---- my own columns: pivot search local to this CTA ----
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        if (i < nown && !sing) {
            const int c = c0 + i, d = J + c, par = i & 1;
            if (tr && tid == 0) tr[c * 8 + 0] = gtime();
            // my best row: first maximum of |a| in the swapped ordering; a NaN never wins unless it is the diagonal
            // (lu.rs:170-178); "no candidate" = (-1, INT_MAX)
            T bav = T(-1);
            int bidx = INT_MAX, br = 0;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                T av = fabs(a[r][i]);
                if (!(av == av) && pos[r] == d) av = T(CUDART_INF);
                const bool gt = ((act >> r) & 1u) && (av > bav || (av == bav && pos[r] < bidx));
                bav = gt ? av : bav;
                bidx = gt ? pos[r] : bidx;
                br = gt ? r : br;
            }
            unsigned long long wk = key_of(double(bav));
            int wi = bidx;
            unsigned wm;
            warp_argmax(wk, wi, wm);
            if (lane == 0) { wkey[par][warp] = wk; widx[par][warp] = wi; }
            if (bidx == wi && wi != INT_MAX) {                           // the one lane that holds the warp's candidate row
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    if (r == br) {
#pragma unroll
                        for (int jv = (i / VN) * VN; jv < NC; jv += VN) {
                            Vec t;
#pragma unroll
                            for (int e = 0; e < VN; ++e) t.v[e] = a[r][jv + e];
                            *reinterpret_cast<Vec *>(&wrow[par][warp][jv]) = t;
                        }
                    }
                }
            }
            slab_worker_sync();
            unsigned long long gk = (lane < SL_WARPS) ? wkey[par][lane] : 0ull;
            int gi = (lane < SL_WARPS) ? widx[par][lane] : INT_MAX;
            unsigned gm;
            warp_argmax(gk, gi, gm);
            const int ww = gm ? __ffs(gm) - 1 : 0;
            const bool sing_now = T(__longlong_as_double((long long)gk)) < Eps<T>::v();      // lu.rs:179-183
            if (tr && tid == 0) tr[c * 8 + 1] = gtime();
            if (tid == 0) own_piv[i] = sing_now ? -1 : gi;
            if (sing_now) {
                sing = true;
            } else {
                const int p = gi;
                const T *u = &wrow[par][ww][0];
                const T piv = u[i];
                T y = T(0);
                if (ROWS >= 4) y = rcp_rn(piv);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const bool on = (act >> r) & 1u;
                    const bool isp = on && pos[r] == p;
                    const bool isd = pos[r] == d;
                    pos[r] = isp ? d : (isd ? p : pos[r]);
                    act &= ~(unsigned(isp) << r);
                    if (on && !isp) {                        // (rows past the panel's end hold zeros: they would take div_via_rcp's slow path)
                        const T mm = (ROWS >= 4) ? div_via_rcp(a[r][i], piv, y) : div_rn(a[r][i], piv);
                        a[r][i] = mm;
                        m[r] = mm;
                        __stcg(mbuf + size_t(c) * RS + tid + SL_WORKERS * r, mm);
                    }
                }
                if (NC <= 8) {
                    T ur[NC];
#pragma unroll
                    for (int jv = ((i + 1) / VN) * VN; jv < NC; jv += VN) {
                        const Vec t = *reinterpret_cast<const Vec *>(u + jv);
#pragma unroll
                        for (int e = 0; e < VN; ++e) ur[jv + e] = t.v[e];
                    }
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        if ((act >> r) & 1u) {
#pragma unroll
                            for (int j = i + 1; j < NC; ++j) a[r][j] = sub_rn(a[r][j], mul_rn(m[r], ur[j]));
                        }
                    }
                } else {
#pragma unroll
                    for (int jv = ((i + 1) / VN) * VN; jv < NC; jv += VN) {
                        const Vec t = *reinterpret_cast<const Vec *>(u + jv);
#pragma unroll
                        for (int e = 0; e < VN; ++e) {
                            if (jv + e > i) {
#pragma unroll
                                for (int r = 0; r < ROWS; ++r)
                                    if ((act >> r) & 1u) a[r][jv + e] = sub_rn(a[r][jv + e], mul_rn(m[r], t.v[e]));
                            }
                        }
                    }
                }
            }
            if (tr && tid == 0) tr[c * 8 + 6] = gtime();
            __syncwarp();
            if (lane == 0) red_release_cta_s(&done_cnt, 1u);
        }
    }

    // ---- the pivots of the columns after mine only move my rows' positions ----
    if (!sing) {
        for (int c = c0 + nown; c < jb; ++c) {
            if (unsigned(c) >= known) {
                known = ld_acquire_cta_s(&ready);
                while (known <= unsigned(c)) known = ld_acquire_cta_s(&ready);
            }
            const int p = piv_sm[c];
            if (p < 0) { sing = true; break; }
            const int d = J + c;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const bool isp = ((act >> r) & 1u) && pos[r] == p;
                const bool isd = pos[r] == d;
                pos[r] = isp ? d : (isd ? p : pos[r]);
                act &= ~(unsigned(isp) << r);
            }
        }
    }
    if (tr && tid == 0) tr[1536 + 8 * k + 1] = gtime();
    if (sing) return;                                        // uniform over the workers (and over the CTAs)
    // the panel's net permutation, straight from the rows' positions: row pos[r] <- old row J + s wherever they differ
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int s = tid + SL_WORKERS * r;
        if (s < nrows && pos[r] != J + s) {
            const int e = atomicAdd(&sh_nt, 1);
            dst_l[e] = pos[r];
            src_l[e] = J + s;
        }
    }
    if (tr && tid == 0) tr[1536 + 8 * k + 2] = gtime();
    // write my columns back, every row at its final position
    if (vec_io) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int s = tid + SL_WORKERS * r;
            if (s < nrows) {
#pragma unroll
                for (int jv = 0; jv < NC; jv += VN) {
                    Vec t;
#pragma unroll
                    for (int e = 0; e < VN; ++e) t.v[e] = a[r][jv + e];
                    *reinterpret_cast<Vec *>(stg + size_t(pos[r] - J) * SLD + jv) = t;
                }
            }
        }
        slab_worker_sync();
        for (int q = tid; q < nrows * LPR; q += SL_WORKERS) {
            const int row = q / LPR, part = q % LPR;
            *reinterpret_cast<Vec *>(A + size_t(J + row) * ld + J + c0 + part * VN) =
                *reinterpret_cast<const Vec *>(stg + size_t(row) * SLD + part * VN);
        }
    } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int s = tid + SL_WORKERS * r;
            if (s < nrows) {
#pragma unroll
                for (int j = 0; j < NC; ++j)
                    if (j < nown) A[size_t(pos[r]) * ld + J + c0 + j] = a[r][j];
            }
        }
    }
    if (tr && tid == 0) tr[1536 + 8 * k + 3] = gtime();
    slab_worker_sync();                                      // the list of moved rows is complete
    // the panel's interchanges on the outer block's columns outside the panel: one gather, split over the CTAs
    if (!(dbg & 1)) {
        const int na = J - J0, ncols = na + (J0 + w - J - jb);     // columns [J0, J) and [J + jb, J0 + w)
        const int cpc = (ncols + ncta - 1) / ncta;
        const int c_lo = k * cpc, c_hi = min(ncols, c_lo + cpc);
        const int nt = sh_nt;
        for (int g0 = c_lo; g0 < c_hi; g0 += CL_GC) {             // uniform per CTA
            T v[SL_GPASS];
            bool mv[SL_GPASS];
#pragma unroll
            for (int uu = 0; uu < SL_GPASS; ++uu) {
                const int e = tid + SL_WORKERS * uu, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                mv[uu] = i < nt && t2 < c_hi;
                if (mv[uu]) {
                    const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                    v[uu] = A[size_t(src_l[i]) * ld + col];
                }
            }
            slab_worker_sync();                                     // every source is read before any destination is written
#pragma unroll
            for (int uu = 0; uu < SL_GPASS; ++uu) {
                const int e = tid + SL_WORKERS * uu, i = e / CL_GC, t2 = g0 + (e % CL_GC);
                if (mv[uu]) {
                    const int col = (t2 < na) ? J0 + t2 : J + jb + (t2 - na);
                    A[size_t(dst_l[i]) * ld + col] = v[uu];
                }
            }
        }
    }
    if (tr && tid == 0) tr[1536 + 8 * k + 4] = gtime();
}

// -------------------------------------------------------------------------------------------
// laswp for a whole outer block: interchanges (J+k <-> ipiv[J+k]), k = 0..jb-1 (jb <= 256), applied
// to the columns left and right of the block as ONE gather instead of jb dependent swaps.
//   plan kernel (one warp): simulate the interchanges on indices.  "Touched" rows: index i < jb ->
//     row J+i; index jb+f -> far row fr[f].  origin[i] = touched index whose OLD contents end up in
//     touched row i.  Also permutes the row-origin vector from which `perm` is produced.
//   apply kernel: each CTA owns 32 columns; reads every source row segment into shared memory
//     (8 independent loads in flight per lane), then writes the destinations.
// -------------------------------------------------------------------------------------------
constexpr int LASWP_THREADS = 512;
constexpr int LASWP_CW = 32;

template <typename T>
__global__ void __launch_bounds__(LASWP_THREADS)
laswp_apply_kernel(T *__restrict__ A, size_t ld, const LaswpPlan *__restrict__ plan, const int32_t *__restrict__ info,
                   int c0a, int c1a, int c0b, int c1b) {
    if (*info != 0) return;
    extern __shared__ __align__(16) unsigned char laswp_smem[];
    T *tile = reinterpret_cast<T *>(laswp_smem);                 // [nt][LASWP_CW]
    __shared__ int rows_s[2 * LASWP_MAXJB];
    __shared__ int org_s[2 * LASWP_MAXJB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = plan->nt;
    for (int i = tid; i < nt; i += LASWP_THREADS) {
        rows_s[i] = plan->rows[i];
        org_s[i] = plan->origin[i];
    }
    __syncthreads();
    const int na = c1a - c0a, ncols = na + (c1b - c0b);
    const int q = blockIdx.x * LASWP_CW + lane;
    const bool ok = q < ncols;
    const int col = ok ? ((q < na) ? c0a + q : c0b + (q - na)) : 0;
    constexpr int NW = LASWP_THREADS / 32, UNR = 8;
    // read phase: tile[i] <- old contents of the row that ends up in touched row i
    for (int i0 = warp; i0 < nt; i0 += NW * UNR) {
        T v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * NW;
            v[u] = T(0);
            if (i < nt && ok && org_s[i] != i) v[u] = A[size_t(rows_s[org_s[i]]) * ld + col];
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * NW;
            if (i < nt) tile[i * LASWP_CW + lane] = v[u];
        }
    }
    __syncthreads();
    for (int i = warp; i < nt; i += NW) {
        if (ok && org_s[i] != i) A[size_t(rows_s[i]) * ld + col] = tile[i * LASWP_CW + lane];
    }
}

// -------------------------------------------------------------------------------------------
// trsm: B <- L^-1 B, L = unit lower jb x jb (jb <= 64) at A[j:j+jb, j:j+jb], B = A[j:j+jb, c0:c1).
// Four threads per column (rows interleaved mod 4, 16 values each in registers); x_k is broadcast
// inside the 4-thread group with a shuffle, L is read from shared memory.
// -------------------------------------------------------------------------------------------
constexpr int TRSM_THREADS = 128;
constexpr int TRSM_COLS = TRSM_THREADS / 4;

template <typename T>
__global__ void __launch_bounds__(TRSM_THREADS)
trsm_unit_lower_kernel(const T *__restrict__ L, size_t ldl, int jb, T *__restrict__ B, size_t ldb, int ncols,
                       const int32_t *__restrict__ info) {
    // L: jb x jb unit lower (top-left element), B: jb x ncols (top-left element); L and B may alias one matrix
    if (*info != 0) return;
    __shared__ T Ls[PW * (PW + 1)];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < PW * PW; idx += TRSM_THREADS) {
        const int r = idx / PW, c = idx - r * PW;
        Ls[r * (PW + 1) + c] = (r < jb && c < r) ? L[size_t(r) * ldl + c] : T(0);
    }
    __syncthreads();
    const int lane = tid & 31;
    const int r = lane & 3;                                   // row residue owned by this thread
    const int col = blockIdx.x * TRSM_COLS + (tid >> 5) * 8 + (lane >> 2);
    const bool ok = col < ncols;
    T x[PW / 4];
#pragma unroll
    for (int ii = 0; ii < PW / 4; ++ii) {
        const int i = 4 * ii + r;
        x[ii] = (ok && i < jb) ? B[size_t(i) * ldb + col] : T(0);
    }
    const T *Lr = Ls + r * (PW + 1);
#pragma unroll
    for (int k = 0; k < PW - 1; ++k) {
        const T xk = __shfl_sync(0xffffffffu, x[k >> 2], (lane & ~3) | (k & 3));
#pragma unroll
        for (int ii = (k >> 2); ii < PW / 4; ++ii) {
            if (ii > (k >> 2) || r > (k & 3)) x[ii] -= Lr[(4 * ii) * (PW + 1) + k] * xk;
        }
    }
#pragma unroll
    for (int ii = 0; ii < PW / 4; ++ii) {
        const int i = 4 * ii + r;
        if (ok && i >= 1 && i < jb) B[size_t(i) * ldb + col] = x[ii];
    }
}

// B <- U^-1 B, U = jb x jb upper NON-unit (jb <= 64): same 4-threads-per-column scheme, k descending.
// Diagonal reciprocals are formed once (the singularity test |u_ii| < eps is done by the caller).
template <typename T>
__global__ void __launch_bounds__(TRSM_THREADS)
trsm_upper_kernel(const T *__restrict__ U, size_t ldu, int jb, T *__restrict__ B, size_t ldb, int ncols,
                  const int32_t *__restrict__ info) {
    if (*info != 0) return;
    __shared__ T Us[PW * (PW + 1)];
    __shared__ T rdiag[PW];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < PW * PW; idx += TRSM_THREADS) {
        const int r = idx / PW, c = idx - r * PW;
        T v = (r == c) ? T(1) : T(0);
        if (r < jb && c < jb && c >= r) v = U[size_t(r) * ldu + c];
        Us[r * (PW + 1) + c] = v;
    }
    __syncthreads();
    if (tid < PW) rdiag[tid] = T(1) / Us[tid * (PW + 1) + tid];
    __syncthreads();
    const int lane = tid & 31;
    const int r = lane & 3;
    const int col = blockIdx.x * TRSM_COLS + (tid >> 5) * 8 + (lane >> 2);
    const bool ok = col < ncols;
    T x[PW / 4];
#pragma unroll
    for (int ii = 0; ii < PW / 4; ++ii) {
        const int i = 4 * ii + r;
        x[ii] = (ok && i < jb) ? B[size_t(i) * ldb + col] : T(0);
    }
    const T *Ur = Us + r * (PW + 1);
#pragma unroll
    for (int k = PW - 1; k >= 0; --k) {
        if (r == (k & 3)) x[k >> 2] *= rdiag[k];
        const T xk = __shfl_sync(0xffffffffu, x[k >> 2], (lane & ~3) | (k & 3));
#pragma unroll
        for (int ii = 0; ii <= (k >> 2); ++ii) {
            if (ii < (k >> 2) || r < (k & 3)) x[ii] -= Ur[(4 * ii) * (PW + 1) + k] * xk;
        }
    }
#pragma unroll
    for (int ii = 0; ii < PW / 4; ++ii) {
        const int i = 4 * ii + r;
        if (ok && i < jb) B[size_t(i) * ldb + col] = x[ii];
    }
}

// X <- P as a dense matrix: X[r][c] = (r == perm[c])   (column c of the inverse is solve(e_c), and P e_c = e_perm[c])
template <typename T>
__global__ void perm_matrix_kernel(T *__restrict__ X, size_t ldx, int n, const int64_t *__restrict__ perm) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(n) * n) return;
    const int r = int(idx / n), c = int(idx - size_t(r) * n);
    X[size_t(r) * ldx + c] = (int64_t(r) == perm[c]) ? T(1) : T(0);
}
// back_substitution's test (src/matrix/mod.rs:333-336) for every diagonal entry; reports the LAST failing row
// like the reference's descending loop would hit first
template <typename T>
__global__ void diag_check_kernel(const T *__restrict__ lu, size_t ld, int n, int32_t *__restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && fabs(lu[size_t(i) * ld + i]) < Eps<T>::v()) atomicMax(info, i + 1);
}

// rowid <- plan applied to it (the row-origin vector from which `perm` is produced)
__global__ void __launch_bounds__(256)
rowid_apply_kernel(const LaswpPlan *__restrict__ plan, int32_t *__restrict__ rowid, const int32_t *__restrict__ info) {
    if (*info != 0) return;
    __shared__ int ids[2 * LASWP_MAXJB];
    const int nt = plan->nt;
    for (int i = threadIdx.x; i < nt; i += 256) ids[i] = rowid[plan->rows[i]];
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += 256) {
        const int o = plan->origin[i];
        if (o != i) rowid[plan->rows[i]] = ids[o];
    }
}

__global__ void iota_kernel(int32_t *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
// perm[rowid[i]] = i  (PermutationMatrix::inverse, permutation_matrix.rs:137-148)
__global__ void invert_perm_kernel(const int32_t *__restrict__ rowid, int64_t *__restrict__ perm, int n,
                                   const int32_t *__restrict__ info) {
    if (*info != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[rowid[i]] = i;
}

template <typename T>
int gemm_update(size_t m, size_t k, size_t n, const T *a, size_t lda, const T *b, size_t ldb, T *c, size_t ldc,
                cudaStream_t st);
template <>
int gemm_update<double>(size_t m, size_t k, size_t n, const double *a, size_t lda, const double *b, size_t ldb,
                        double *c, size_t ldc, cudaStream_t st) {
    return dgemm_launch(m, k, n, -1.0, a, lda, b, ldb, 1.0, c, ldc, st);
}
template <>
int gemm_update<float>(size_t m, size_t k, size_t n, const float *a, size_t lda, const float *b, size_t ldb,
                       float *c, size_t ldc, cudaStream_t st) {
    return sgemm_launch(m, k, n, -1.f, a, lda, b, ldb, 1.f, c, ldc, st);
}

unsigned long long *g_lu_trace = nullptr;
int g_lu_gmax_ref();
int g_lu_dbg_ref();
int g_lu_cluster_ref();
int g_lu_slab_rows_ref();
int g_lu_k3e_rows_ref();

// scratch layout (bytes): packets | rowbuf | diagbuf | result | laswp plan | plan state
constexpr size_t SC_PACKETS = 0;
constexpr size_t SC_ROWBUF = SC_PACKETS + 2 * GMAX * sizeof(Msg);
constexpr size_t SC_DIAGBUF = SC_ROWBUF + size_t(2) * GMAX * PW * sizeof(Msg);
constexpr size_t SC_RESULT = SC_DIAGBUF + 2 * PW * sizeof(Msg);
constexpr size_t SC_PLAN = SC_RESULT + ((PW * 8 + 255) / 256) * 256;
constexpr size_t SC_STATE = SC_PLAN + ((sizeof(LaswpPlan) + 255) / 256) * 256;
constexpr size_t SC_TOTAL = SC_STATE + sizeof(PlanState) + 256;

int ensure_workspace(LuWorkspace &ws, int n, cudaStream_t st) {
    if (ws.ipiv_cap < size_t(2) * n) {                 // ipiv[n] + rowid[n]
        if (ws.ipiv) RLA_CUDA(cudaFree(ws.ipiv));
        ws.ipiv = nullptr;
        ws.ipiv_cap = 0;
        RLA_CUDA(cudaMalloc(&ws.ipiv, sizeof(int32_t) * 2 * size_t(n)));
        ws.ipiv_cap = size_t(2) * n;
    }
    if (ws.scratch_cap < SC_TOTAL) {
        if (ws.scratch) RLA_CUDA(cudaFree(ws.scratch));
        ws.scratch = nullptr;
        ws.scratch_cap = 0;
        RLA_CUDA(cudaMalloc(&ws.scratch, SC_TOTAL));
        ws.scratch_cap = SC_TOTAL;
        ws.tag = 0;
        RLA_CUDA(cudaMemsetAsync(ws.scratch, 0, SC_TOTAL, st));
    }
    // tags are unique per (launch, column) for the lifetime of the scratch buffer and stay below 2^24 (the candidate
    // packets carry 24 tag bits); re-zero before they would wrap
    if (ws.tag > 0xffffffu - 2u * unsigned(n) - 16u) {
        RLA_CUDA(cudaMemsetAsync(ws.scratch, 0, SC_TOTAL, st));
        ws.tag = 0;
    }
    return RLA_OK;
}

PanelScratch scratch_view(LuWorkspace &ws) {
    unsigned char *sp = static_cast<unsigned char *>(ws.scratch);
    PanelScratch sc;
    sc.packets = reinterpret_cast<Msg *>(sp + SC_PACKETS);
    sc.rowbuf = reinterpret_cast<Msg *>(sp + SC_ROWBUF);
    sc.diagbuf = reinterpret_cast<Msg *>(sp + SC_DIAGBUF);
    sc.piv_log = reinterpret_cast<unsigned long long *>(sp + SC_RESULT);
    sc.plan = reinterpret_cast<LaswpPlan *>(sp + SC_PLAN);
    sc.state = reinterpret_cast<PlanState *>(sp + SC_STATE);
    sc.rowid = nullptr;
    sc.trace = nullptr;
    return sc;
}

constexpr size_t SLAB_HDR_BYTES = PW * sizeof(unsigned long long);
constexpr size_t SLAB_MBUF_OFF = 1024;
constexpr size_t SLAB_BYTES = SLAB_MBUF_OFF + size_t(PW) * SL_MAXROWS * sizeof(double);

template <typename T, int ROWS>
int launch_slab_rows(const SlabScratch &ss, T *a, size_t ld, int n, int j, int jb, int32_t *ipiv, int32_t *d_info,
                     const PanelScratch &sc, int J0, int w, int ncta, cudaStream_t s) {
    constexpr int NC = 32 / ROWS, VN = 16 / int(sizeof(T));
    constexpr size_t smem = size_t(ROWS) * SL_WORKERS * (NC + VN) * sizeof(T);       // staging for the slab's load / write-back
    static DeviceOnce attr_once;
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_slab_kernel<T, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_once.done(od_);
    }
    lu_panel_slab_kernel<T, ROWS><<<ncta + 1, SL_THREADS, smem, s>>>(a, ld, n, j, jb, ipiv, d_info, sc, ss, J0, w, g_lu_dbg_ref());
    RLA_LAUNCHED();
    return RLA_OK;
}

template <typename T>
int launch_slab_panel(LuWorkspace &ws, T *a, size_t ld, int n, int j, int jb, int32_t *ipiv, int32_t *d_info,
                      const PanelScratch &sc, int J0, int w, cudaStream_t s) {
    if (!ws.slab) {
        RLA_CUDA(cudaMalloc(&ws.slab, SLAB_BYTES));
        RLA_CUDA(cudaMemsetAsync(ws.slab, 0, SLAB_MBUF_OFF, s));
        ws.slab_epoch = 0;
        ws.slab_tickets = 0;
    }
    const int nrem = n - j;
    int rows = 1;
    while (rows * SL_WORKERS < nrem) rows *= 2;
    const int nc = 32 / rows, ncta = (jb + nc - 1) / nc;
    unsigned char *base = static_cast<unsigned char *>(ws.slab);
    SlabScratch ss;
    ss.hdr = reinterpret_cast<unsigned long long *>(base);
    ss.ticket = reinterpret_cast<unsigned *>(base + SLAB_HDR_BYTES);
    ss.mbuf = base + SLAB_MBUF_OFF;
    ss.epoch = ++ws.slab_epoch;                 // never 0: a zeroed header is never valid
    if (ws.slab_epoch == 0xffffffffu) {         // wrap: start over with clean headers
        RLA_CUDA(cudaMemsetAsync(ws.slab, 0, SLAB_HDR_BYTES, s));
        ws.slab_epoch = 0;
        ss.epoch = ++ws.slab_epoch;
    }
    ss.ticket_base = ws.slab_tickets;
    ss.rstride = rows * SL_WORKERS;
    int st = RLA_OK;
    switch (rows) {
    case 1: st = launch_slab_rows<T, 1>(ss, a, ld, n, j, jb, ipiv, d_info, sc, J0, w, ncta, s); break;
    case 2: st = launch_slab_rows<T, 2>(ss, a, ld, n, j, jb, ipiv, d_info, sc, J0, w, ncta, s); break;
    case 4: st = launch_slab_rows<T, 4>(ss, a, ld, n, j, jb, ipiv, d_info, sc, J0, w, ncta, s); break;
    default: st = launch_slab_rows<T, 8>(ss, a, ld, n, j, jb, ipiv, d_info, sc, J0, w, ncta, s); break;
    }
    // every CTA of a launch that went out takes exactly one ticket (the hub CTA included); a refused launch takes none.
    // The host count wraps together with the device counter (unsigned arithmetic).
    if (st == RLA_OK) ws.slab_tickets += unsigned(ncta + 1);
    return st;
}

// Factor the outer block whose diagonal starts at (J0, J0) of `a` (w <= 256 columns, rows J0..n-1).
// `a` may be a shifted base pointer so that "column J0" is wherever the block lives in a local matrix
// (distributed layout); only columns [J0, J0+w) are touched.  The block's net permutation goes to *plan.
template <typename T>
int factor_block(LuWorkspace &ws, T *a, size_t ld, int n, int J0, int w, int32_t *ipiv, int32_t *d_info,
                 LaswpPlan *plan, cudaStream_t s) {
    static DeviceOnce attr_once;
    // largest cluster each device schedules for the cluster panel kernel (16 needs the non-portable opt-in)
    static int cluster_max_dev[RLA_MAX_DEVICES];
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_kernel<T, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int cluster_max = 0;
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(CL_ROWS) * PLDS * sizeof(T))));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster2_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(CL_ROWS) * PLDS * sizeof(T))));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster2_kernel<T, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster2_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(CL_ROWS) * PLDS * sizeof(T))));
        RLA_CUDA(cudaFuncSetAttribute(lu_panel_cluster2_kernel<T, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        for (int cs = CL_MAX; cs >= 2; cs /= 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs);
            cfg.blockDim = dim3(CL_THREADS);
            cfg.dynamicSmemBytes = size_t(CL_ROWS) * PLDS * sizeof(T);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = unsigned(cs);
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, lu_panel_cluster_kernel<T>, &cfg) == cudaSuccess && nclusters > 0) {
                cluster_max = cs;
                break;
            }
            (void)cudaGetLastError();
        }
        cluster_max_dev[od_] = cluster_max;
        attr_once.done(od_);
    }
    const int cluster_max = cluster_max_dev[current_device()];
    const int num_sms = device_num_sms();
    PanelScratch sc = scratch_view(ws);
    sc.plan = plan;
    if (g_lu_dbg_ref() & 8) {
        static unsigned long long *trace = nullptr;
        if (!trace) RLA_CUDA(cudaMalloc(&trace, 4 * 64 * 8 * sizeof(unsigned long long)));
        sc.trace = trace;
        g_lu_trace = trace;
    }
    unsigned long long *const trace_base = sc.trace;
    for (int j = J0; j < J0 + w; j += PW) {
        const int jb = min(PW, J0 + w - j);
        const int nrem = n - j;
        if (trace_base) sc.trace = trace_base + size_t((j - J0) / PW) * 64 * 8;   // one page per inner panel
        // panels of <= 4096 rows, column-slab kernel: columns dealt to the CTAs, pivot search local to one CTA
        const int slab_rows = g_lu_cluster_ref() == 3 ? SL_MAXROWS : (g_lu_cluster_ref() == 1 ? min(g_lu_slab_rows_ref(), SL_MAXROWS) : 0);
        // panels whose rows fit the shared memory of ONE cluster (<= 16 x 393 rows in f64): the grid kernel in cluster mode --
        // rows in shared memory, exchange over DSMEM (lu_cluster 5: wherever it fits; automatic rule: lu_k3e_rows)
        const int k3e_rows = (cluster_max >= 2) ? min(g_lu_cluster_ref() == 5 ? INT_MAX : (g_lu_cluster_ref() == 1 ? g_lu_k3e_rows_ref() : 0),
                                                      cluster_max * int((200 * 1024) / (PLDS * sizeof(T)))) : 0;
        if (nrem <= slab_rows) {
            RLA_TRY(launch_slab_panel<T>(ws, a, ld, n, j, jb, ipiv, d_info, sc, J0, w, s));
        } else if (nrem <= k3e_rows) {
            const int CS = cluster_max;
            int R_ = (nrem + CS - 1) / CS, G_ = CS, n__ = n, J_ = j, jb_ = jb, J0_ = J0, w_ = w, dbg_ = g_lu_dbg_ref();
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * CS);                     // cluster 0: the row CTAs; cluster 1: the hub (its other CTAs exit)
            cfg.blockDim = dim3(PANEL_THREADS);
            cfg.dynamicSmemBytes = size_t(R_) * PLDS * sizeof(T);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = unsigned(CS);
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            RLA_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_kernel<T, true>, a, ld, n__, J_, jb_, R_, G_, ipiv, d_info, sc, ws.tag, J0_, w_, dbg_));
            note_launch();
            ws.tag += unsigned(jb);
        } else
        // panels of <= 16 x 256 rows: one thread-block cluster, panel in registers, exchange over DSMEM
        if (g_lu_cluster_ref() && nrem <= cluster_max * CL_ROWS) {
            int CS = 1;
            while (CS * CL_ROWS < nrem) CS *= 2;
            const int R = (nrem + CS - 1) / CS;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(CS);
            cfg.blockDim = dim3(CL_THREADS);
            cfg.dynamicSmemBytes = size_t(CL_ROWS) * PLDS * sizeof(T);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = unsigned(CS);
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (g_lu_cluster_ref() == 2 && sc.trace)
                RLA_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_cluster2_kernel<T, true>, a, ld, n, j, jb, R, ipiv, d_info, sc, J0, w,
                                            g_lu_dbg_ref()));
            else if (g_lu_cluster_ref() == 2)
                RLA_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_cluster2_kernel<T, false>, a, ld, n, j, jb, R, ipiv, d_info, sc, J0, w,
                                            g_lu_dbg_ref()));
            else
                RLA_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel<T>, a, ld, n, j, jb, R, ipiv, d_info, sc, J0, w,
                                            g_lu_dbg_ref()));
            note_launch();
        } else {
        int G = min(min(num_sms - 1, g_lu_gmax_ref()), max(1, (nrem + 63) / 64));
        while (size_t((nrem + G - 1) / G) * PLDS * sizeof(T) > 200 * 1024 && G < num_sms - 1) ++G;
        int R = (nrem + G - 1) / G;
        size_t smem = size_t(R) * PLDS * sizeof(T);
        if (smem > 200 * 1024) return RLA_ERR_INVALID;   // n beyond ~58k rows per panel: not supported yet
        {
            // +1 hub CTA: reduces the candidates, applies the interchanges to the rest of the outer block and
            // builds the outer block's net permutation plan on the fly
            const int grid = G + 1;
            T *a_ = a;
            size_t ld_ = ld;
            int n__ = n, J_ = j, jb_ = jb, R_ = R, G_ = G, J0_ = J0, w_ = w, dbg_ = g_lu_dbg_ref();
            unsigned tag_base = ws.tag;
            void *args[] = {&a_, &ld_, &n__, &J_, &jb_, &R_, &G_, &ipiv, &d_info, &sc, &tag_base, &J0_, &w_, &dbg_};
            RLA_CUDA(cudaLaunchCooperativeKernel((void *)lu_panel_kernel<T, false>, dim3(grid), dim3(PANEL_THREADS), args, smem, s));
            note_launch();
            ws.tag += unsigned(jb);
        }
        }
        // U12 and the Schur update inside the outer block
        if (j + jb < J0 + w) {
            const int nc = J0 + w - j - jb;
            trsm_unit_lower_kernel<T><<<(nc + TRSM_COLS - 1) / TRSM_COLS, TRSM_THREADS, 0, s>>>(
                a + size_t(j) * ld + j, ld, jb, a + size_t(j) * ld + j + jb, ld, nc, d_info);
            RLA_LAUNCHED();
            if (j + jb < n)
                RLA_TRY(gemm_update<T>(size_t(n - j - jb), size_t(jb), size_t(nc), a + size_t(j + jb) * ld + j, ld,
                                       a + size_t(j) * ld + j + jb, ld, a + size_t(j + jb) * ld + j + jb, ld, s));
        }
    }
    return RLA_OK;
}

// Apply a block's net permutation to columns [c0a,c1a) U [c0b,c1b) of `a`.
template <typename T>
int apply_laswp(T *a, size_t ld, int w, const LaswpPlan *plan, const int32_t *info, int c0a, int c1a, int c0b, int c1b,
                cudaStream_t st) {
    const int ncols = (c1a - c0a) + (c1b - c0b);
    if (ncols <= 0) return RLA_OK;
    const size_t smem = size_t(2) * w * LASWP_CW * sizeof(T);
    static DeviceOnce attr_once;
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(laswp_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LASWP_MAXJB * LASWP_CW * 8));
        attr_once.done(od_);
    }
    const int blocks = (ncols + LASWP_CW - 1) / LASWP_CW;
    laswp_apply_kernel<T><<<blocks, LASWP_THREADS, smem, st>>>(a, ld, plan, info, c0a, c1a, c0b, c1b);
    RLA_LAUNCHED();
    return RLA_OK;
}

// U12 = L11^-1 B by blocks of PW rows (L11: w x w unit lower at L, B: w x ncols), all in place.
template <typename T>
int trsm_block(const T *L, size_t ldl, int w, T *B, size_t ldb, int ncols, const int32_t *info, cudaStream_t st) {
    if (ncols <= 0) return RLA_OK;
    for (int kb = 0; kb < w; kb += PW) {
        const int jb = min(PW, w - kb);
        if (jb > 1) {
            trsm_unit_lower_kernel<T><<<(ncols + TRSM_COLS - 1) / TRSM_COLS, TRSM_THREADS, 0, st>>>(
                L + size_t(kb) * ldl + kb, ldl, jb, B + size_t(kb) * ldb, ldb, ncols, info);
            RLA_LAUNCHED();
        }
        if (kb + jb < w)
            RLA_TRY(gemm_update<T>(size_t(w - kb - jb), size_t(jb), size_t(ncols), L + size_t(kb + jb) * ldl + kb, ldl,
                                   B + size_t(kb) * ldb, ldb, B + size_t(kb + jb) * ldb, ldb, st));
    }
    return RLA_OK;
}

}  // namespace

int g_lu_gmax = 32;           // rla_set_tuning("lu_gmax", v): cap on the grid panel kernel's row CTAs (raised automatically until the
                              // rows fit in shared memory).  The panel is latency-bound, its CTAs only take SMs from the overlapped
                              // Schur update: tools/lu_gmax_sweep.py, n = 16384: 120.9 / 128.0 / 133.1 ms at 32 / 112 / 147
int g_lu_dbg = 0;             // rla_set_tuning("lu_dbg", bits): experiments (bit0: hub skips row swaps, bit2: no look-ahead)
int g_lu_cluster = 1;          // rla_set_tuning("lu_cluster", v): panel kernel selection.  0 = always the grid-wide kernel (K3);
                               // 1 (default) = automatic: column-slab kernel (K3d) for panels of <= lu_slab_rows rows, cluster pull
                               // kernel (K3b) up to 4096 rows, K3 above; 2 = pushed-row cluster kernel (K3c, experimental: bit-identical,
                               // measured slower); 3 = K3d wherever it fits (<= 3840 rows), K3b / K3 above; 4 = K3b / K3 only;
                               // 5 = K3 in cluster mode (rows in shared memory, DSMEM exchange) wherever one cluster holds the rows
int g_lu_k3e_rows = 0;         // rla_set_tuning("lu_k3e_rows", v): tallest panel the automatic rule gives to the grid kernel in cluster mode
                               // (rows in shared memory, exchange over DSMEM); panels of <= lu_slab_rows rows still go to K3d
int g_lu_slab_rows = 1920;     // rla_set_tuning("lu_slab_rows", v): tallest panel the automatic rule gives to K3d (measured crossover
                               // against K3b: profiles/r02_lu_slab_sweep.jsonl)
namespace { int g_lu_gmax_ref() { return g_lu_gmax; } int g_lu_dbg_ref() { return g_lu_dbg; } int g_lu_cluster_ref() { return g_lu_cluster; } int g_lu_slab_rows_ref() { return g_lu_slab_rows; } int g_lu_k3e_rows_ref() { return g_lu_k3e_rows; } }

int device_num_sms() {
    static std::atomic<int> sms[RLA_MAX_DEVICES];
    const int dev = current_device();
    int v = sms[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { (void)cudaGetLastError(); v = 148; }
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

void lu_workspace_release(LuWorkspace &ws) {
    if (ws.ipiv) cudaFree(ws.ipiv);
    if (ws.scratch) cudaFree(ws.scratch);
    if (ws.slab) cudaFree(ws.slab);
    if (ws.side) cudaStreamDestroy(ws.side);
    if (ws.ev_head) cudaEventDestroy(ws.ev_head);
    if (ws.ev_fact) cudaEventDestroy(ws.ev_fact);
    (void)cudaGetLastError();
    ws = LuWorkspace();
}

size_t lu_plan_bytes() { return sizeof(LaswpPlan); }

// Largest n the panel kernels take: a 64-column panel of n rows must fit the shared memory of (SMs - 1) row CTAs at
// 200 KB each (57 771 rows in f64, 115 689 in f32 on a 148-SM B200; an n = 57 771 f64 matrix is 26.7 GB).
size_t lu_max_n(size_t elem_size) {
    const size_t rows_per_cta = (size_t(200) * 1024) / (size_t(PLDS) * elem_size);
    return rows_per_cta * size_t(device_num_sms() - 1);
}

// development / test aid: div_via_rcp against the IEEE division over `count` pseudo-random operand pairs
int lu_divcheck(int f32, int mode, unsigned long long seed, unsigned long long count, unsigned long long *mismatches) {
    unsigned long long *d = nullptr;
    RLA_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    RLA_CUDA(cudaMemset(d, 0, sizeof(unsigned long long)));
    const unsigned blocks = 148 * 8, threads = 256;
    const unsigned long long per = (count + blocks * threads - 1) / (blocks * threads);
    if (f32) divcheck_kernel<float><<<blocks, threads>>>(seed, per, mode, d);
    else divcheck_kernel<double><<<blocks, threads>>>(seed, per, mode, d);
    RLA_LAUNCHED();
    RLA_CUDA(cudaMemcpy(mismatches, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return RLA_OK;
}
int lu_trace_fetch(unsigned long long *host512) {
    if (!g_lu_trace) return RLA_ERR_INVALID;
    RLA_CUDA(cudaMemcpy(host512, g_lu_trace, 4 * 64 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return RLA_OK;
}

template <typename T>
int getrf_launch(size_t n_, T *a, size_t ld, int64_t *d_perm, int32_t *d_info, LuWorkspace &ws, cudaStream_t st,
                 const LuRowsFinal *rows_final, const LuAfterFirstBlock *after_first) {
    if (n_ > lu_max_n(sizeof(T))) return RLA_ERR_INVALID;      // checked up front: nothing has been enqueued yet
    const int n = int(n_);
    RLA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    if (n == 0) return RLA_OK;
    RLA_TRY(ensure_workspace(ws, n, st));
    int32_t *ipiv = ws.ipiv;
    int32_t *rowid = ws.ipiv + n;
    LaswpPlan *plan = scratch_view(ws).plan;
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowid, n);
    RLA_LAUNCHED();

    // A[J0+w:n, c0:c1) -= L21 * U12[:, c0:c1)   (k = w)
    auto schur = [&](int J0, int w, int c0, int c1, cudaStream_t s) -> int {
        if (c1 <= c0 || J0 + w >= n) return RLA_OK;
        return gemm_update<T>(size_t(n - J0 - w), size_t(w), size_t(c1 - c0), a + size_t(J0 + w) * ld + J0, ld,
                              a + size_t(J0) * ld + c0, ld, a + size_t(J0 + w) * ld + c0, ld, s);
    };

    // Look-ahead (depth 1): while the rank-256 Schur update of block k runs on the caller's stream over the
    // columns right of block k+1, block k+1 -- whose columns were updated first -- is factored on a
    // high-priority side stream.  The latency-bound panel kernels then overlap the tensor-pipe-bound update.
    const bool lookahead = (g_lu_dbg & 4) == 0 && n > 2 * OUTER_W;
    if (lookahead && ws.side == nullptr) {
        int lo = 0, hi = 0;
        RLA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        RLA_CUDA(cudaStreamCreateWithPriority(&ws.side, cudaStreamNonBlocking, hi));
        RLA_CUDA(cudaEventCreateWithFlags(&ws.ev_head, cudaEventDisableTiming));
        RLA_CUDA(cudaEventCreateWithFlags(&ws.ev_fact, cudaEventDisableTiming));
    }
    RLA_TRY(factor_block<T>(ws, a, ld, n, 0, min(OUTER_W, n), ipiv, d_info, plan, st));
    if (after_first) RLA_TRY((*after_first)(st));
    for (int J0 = 0; J0 < n; J0 += OUTER_W) {
        const int w = min(OUTER_W, n - J0);
        // interchanges of the whole outer block applied left and right of it, and to the row-origin vector
        rowid_apply_kernel<<<1, 256, 0, st>>>(plan, rowid, d_info);
        RLA_LAUNCHED();
        RLA_TRY(apply_laswp<T>(a, ld, w, plan, d_info, 0, J0, J0 + w, n, st));
        if (J0 + w >= n) {
            if (rows_final) RLA_TRY((*rows_final)(J0, w, st));
            break;
        }
        RLA_TRY(trsm_block<T>(a + size_t(J0) * ld + J0, ld, w, a + size_t(J0) * ld + J0 + w, ld, n - J0 - w, d_info, st));
        // rows [J0, J0+w) are final from here on: later interchanges only touch rows below them
        if (rows_final) RLA_TRY((*rows_final)(J0, w, st));
        const int next = J0 + w;
        const int wn = min(OUTER_W, n - next);
        if (lookahead && next + wn < n) {
            RLA_TRY(schur(J0, w, next, next + wn, st));                 // head: the next block's columns first
            RLA_CUDA(cudaEventRecord(ws.ev_head, st));
            RLA_CUDA(cudaStreamWaitEvent(ws.side, ws.ev_head, 0));
            RLA_TRY(factor_block<T>(ws, a, ld, n, next, wn, ipiv, d_info, plan, ws.side));   // || with the tail update
            RLA_CUDA(cudaEventRecord(ws.ev_fact, ws.side));
            RLA_TRY(schur(J0, w, next + wn, n, st));                    // tail
            RLA_CUDA(cudaStreamWaitEvent(st, ws.ev_fact, 0));
        } else {
            RLA_TRY(schur(J0, w, next, n, st));
            RLA_TRY(factor_block<T>(ws, a, ld, n, next, wn, ipiv, d_info, plan, st));
        }
    }
    invert_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowid, d_perm, n, d_info);
    RLA_LAUNCHED();
    return RLA_OK;
}

template int getrf_launch<double>(size_t, double *, size_t, int64_t *, int32_t *, LuWorkspace &, cudaStream_t, const LuRowsFinal *, const LuAfterFirstBlock *);
template int getrf_launch<float>(size_t, float *, size_t, int64_t *, int32_t *, LuWorkspace &, cudaStream_t, const LuRowsFinal *, const LuAfterFirstBlock *);

// PartialPivLu::inverse (lu.rs:251-285) as a blocked multi-RHS solve: X = U^-1 L^-1 P with all n unit vectors at once
// (SURVEY 8f rank 1).  The reference performs n separate solves; this is 2n^3 flops on the GEMM kernels instead.
template <typename T>
int getri_launch(size_t n_, const T *lu, size_t ld, const int64_t *d_perm, T *x, size_t ldx, int32_t *d_info,
                 cudaStream_t st) {
    if (n_ > 0x3fffffffull) return RLA_ERR_INVALID;
    const int n = int(n_);
    RLA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    if (n == 0) return RLA_OK;
    if (n <= 64) return getri_small_launch<T>(n, lu, ld, d_perm, x, ldx, d_info, st);
    diag_check_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(lu, ld, n, d_info);
    RLA_LAUNCHED();
    {
        const size_t total = size_t(n) * n;
        perm_matrix_kernel<T><<<unsigned((total + 255) / 256), 256, 0, st>>>(x, ldx, n, d_perm);
        RLA_LAUNCHED();
    }
    // L Y = P : forward, outer blocks of 256
    for (int J0 = 0; J0 < n; J0 += OUTER_W) {
        const int w = min(OUTER_W, n - J0);
        RLA_TRY(trsm_block<T>(lu + size_t(J0) * ld + J0, ld, w, x + size_t(J0) * ldx, ldx, n, d_info, st));
        if (J0 + w < n)
            RLA_TRY(gemm_update<T>(size_t(n - J0 - w), size_t(w), size_t(n), lu + size_t(J0 + w) * ld + J0, ld,
                                   x + size_t(J0) * ldx, ldx, x + size_t(J0 + w) * ldx, ldx, st));
    }
    // U X = Y : backward
    const int nblk = (n + OUTER_W - 1) / OUTER_W;
    for (int bI = nblk - 1; bI >= 0; --bI) {
        const int J0 = bI * OUTER_W, w = min(OUTER_W, n - J0);
        const int nsub = (w + PW - 1) / PW;
        for (int sb = nsub - 1; sb >= 0; --sb) {
            const int kb = sb * PW, jb = min(PW, w - kb);
            trsm_upper_kernel<T><<<(n + TRSM_COLS - 1) / TRSM_COLS, TRSM_THREADS, 0, st>>>(
                lu + size_t(J0 + kb) * ld + J0 + kb, ld, jb, x + size_t(J0 + kb) * ldx, ldx, n, d_info);
            RLA_LAUNCHED();
            if (kb > 0)   // rows of this outer block above the sub-block
                RLA_TRY(gemm_update<T>(size_t(kb), size_t(jb), size_t(n), lu + size_t(J0) * ld + J0 + kb, ld,
                                       x + size_t(J0 + kb) * ldx, ldx, x + size_t(J0) * ldx, ldx, st));
        }
        if (J0 > 0)
            RLA_TRY(gemm_update<T>(size_t(J0), size_t(w), size_t(n), lu + J0, ld, x + size_t(J0) * ldx, ldx, x, ldx, st));
    }
    return RLA_OK;
}
template int getri_launch<double>(size_t, const double *, size_t, const int64_t *, double *, size_t, int32_t *, cudaStream_t);
template int getri_launch<float>(size_t, const float *, size_t, const int64_t *, float *, size_t, int32_t *, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// Building blocks of the 1D block-cyclic multi-GPU LU (rulinalg_b200/sharded_lu.py drives them).
// Local matrix: n rows x ncols_loc columns (row stride ld); global column block J (256 wide) of rank
// J mod g sits at local columns [(J/g)*256, ...).  All row indices are global.
// ---------------------------------------------------------------------------------------------
// Owner: factor the block whose global diagonal starts at row0 and whose columns sit at local column lcol0.
template <typename T>
int lu_factor_block_dev(int n, T *a_loc, size_t ld, int row0, int lcol0, int w, int32_t *d_info, void *d_plan,
                        LuWorkspace &ws, cudaStream_t st) {
    RLA_TRY(ensure_workspace(ws, n, st));
    T *shifted = a_loc + lcol0 - row0;      // "column row0" of `shifted` is local column lcol0
    return factor_block<T>(ws, shifted, ld, n, row0, w, ws.ipiv, d_info, static_cast<LaswpPlan *>(d_plan), st);
}
// Everyone: apply the block's interchanges to local columns [c0a,c1a) U [c0b,c1b).
template <typename T>
int lu_laswp_dev(T *a_loc, size_t ld, int w, const void *d_plan, const int32_t *d_info, int c0a, int c1a, int c0b,
                 int c1b, cudaStream_t st) {
    return apply_laswp<T>(a_loc, ld, w, static_cast<const LaswpPlan *>(d_plan), d_info, c0a, c1a, c0b, c1b, st);
}
// Everyone: with the broadcast panel P ((n-row0) x w, row stride ldp; L11 on top of L21), update local columns
// [c0,c1): U12 = L11^-1 A12, then A22 -= L21 U12.
template <typename T>
int lu_update_dev(int n, T *a_loc, size_t ld, int row0, int w, const T *panel, size_t ldp, int c0, int c1,
                  const int32_t *d_info, cudaStream_t st) {
    if (c1 <= c0) return RLA_OK;
    RLA_TRY(trsm_block<T>(panel, ldp, w, a_loc + size_t(row0) * ld + c0, ld, c1 - c0, d_info, st));
    if (row0 + w < n)
        RLA_TRY(gemm_update<T>(size_t(n - row0 - w), size_t(w), size_t(c1 - c0), panel + size_t(w) * ldp, ldp,
                               a_loc + size_t(row0) * ld + c0, ld, a_loc + size_t(row0 + w) * ld + c0, ld, st));
    return RLA_OK;
}
int lu_rowid_init_dev(int32_t *rowid, int n, cudaStream_t st) {
    if (n <= 0) return RLA_OK;
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowid, n);
    RLA_LAUNCHED();
    return RLA_OK;
}
int lu_rowid_apply_dev(const void *d_plan, int32_t *rowid, const int32_t *d_info, cudaStream_t st) {
    rowid_apply_kernel<<<1, 256, 0, st>>>(static_cast<const LaswpPlan *>(d_plan), rowid, d_info);
    RLA_LAUNCHED();
    return RLA_OK;
}
int lu_perm_from_rowid_dev(const int32_t *rowid, int64_t *perm, int n, const int32_t *d_info, cudaStream_t st) {
    if (n <= 0) return RLA_OK;
    invert_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowid, perm, n, d_info);
    RLA_LAUNCHED();
    return RLA_OK;
}
template int lu_factor_block_dev<double>(int, double *, size_t, int, int, int, int32_t *, void *, LuWorkspace &, cudaStream_t);
template int lu_laswp_dev<double>(double *, size_t, int, const void *, const int32_t *, int, int, int, int, cudaStream_t);
template int lu_update_dev<double>(int, double *, size_t, int, int, const double *, size_t, int, int, const int32_t *, cudaStream_t);
template int lu_factor_block_dev<float>(int, float *, size_t, int, int, int, int32_t *, void *, LuWorkspace &, cudaStream_t);
template int lu_laswp_dev<float>(float *, size_t, int, const void *, const int32_t *, int, int, int, int, cudaStream_t);
template int lu_update_dev<float>(int, float *, size_t, int, int, const float *, size_t, int, int, const int32_t *, cudaStream_t);

}  // namespace rla
