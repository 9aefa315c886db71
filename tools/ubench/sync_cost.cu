// micro-benchmark of the synchronisation primitives the LU panel kernels are built from (cycles per operation, one CTA /
// a 2-CTA cluster, 512 or 256 threads).  JSON lines.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }

template <int MODE>
__global__ void k(long long *out, int iters) {
    __shared__ unsigned long long red[32];
    __shared__ __align__(16) unsigned long long slot[4];
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank = 0, csz = 1;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (csz > 1) { asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory"); }
    unsigned v = tid * 2654435761u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { __syncthreads(); }
        else if (MODE == 1) {        // warp argmax (3 redux + ballot), lane 0 -> smem, barrier, warp 0 reduces again
            unsigned m1 = __reduce_max_sync(0xffffffffu, v);
            unsigned m2 = __reduce_max_sync(0xffffffffu, v == m1 ? v ^ 0x55u : 0u);
            int m3 = __reduce_min_sync(0xffffffffu, (v == m1) ? lane : 99);
            unsigned b = __ballot_sync(0xffffffffu, lane == m3);
            if (lane == 0) red[warp] = (unsigned long long)m1 << 32 | m2 | b;
            __syncthreads();
            v = unsigned(red[(warp + 1) & 15]) + it;
        } else if (MODE == 2) {      // producer/consumer named barrier: 15 warps arrive, warp 0 syncs; then everyone waits on __syncthreads
            if (warp == 0) asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
            else asm volatile("bar.arrive 1, %0;" ::"r"(blockDim.x) : "memory");
            __syncthreads();
        } else if (MODE == 3) {      // mbarrier round trip: lane 0 of warp 0 st.async 16 B to (peer or self) CTA; all threads of the destination wait
            const int par = it & 1;
            const unsigned peer = csz > 1 ? rank ^ 1u : rank;
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(smem_u32(&bar[par])) : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(mapa(smem_u32(&slot[2 * par]), peer)),
                             "l"((unsigned long long)it), "l"(1ull), "r"(mapa(smem_u32(&bar[par]), peer)) : "memory");
            }
            unsigned ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar[par])), "r"(unsigned(it >> 1) & 1u) : "memory");
            } while (!ok);
            v += unsigned(slot[2 * par]);
        } else if (MODE == 4) {      // dependent shuffle chain x4
            v = __shfl_xor_sync(0xffffffffu, v, 1) + 1; v = __shfl_xor_sync(0xffffffffu, v, 2) + 1;
            v = __shfl_xor_sync(0xffffffffu, v, 4) + 1; v = __shfl_xor_sync(0xffffffffu, v, 8) + 1;
        } else if (MODE == 5) {      // f64 division
            double d = __ddiv_rn(double(v | 1u), double((v >> 3) | 1u));
            v = unsigned(__double2hiint(d));
        }
    }
    long long t1 = clock64();
    if (tid == 0 && rank == 0) out[0] = t1 - t0;
    if (v == 0x12345u) out[1] = v;
    if (csz > 1) { asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory"); }
}
template <int MODE>
void run(const char *name, int threads, int cluster) {
    long long *d; cudaMalloc(&d, 16);
    const int iters = 2000;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int r = 0; r < 2; ++r) cudaLaunchKernelEx(&cfg, k<MODE>, d, iters);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\":\"%s\",\"threads\":%d,\"cluster\":%d,\"cycles_per_iter\":%.1f,\"err\":\"%s\"}\n", name, threads, cluster, double(h) / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    for (int th : {128, 256, 512}) {
        run<0>("syncthreads", th, 1);
        run<1>("warp_argmax+sts+syncthreads", th, 1);
        run<2>("named arrive/sync + syncthreads", th, 1);
        run<3>("st.async self + mbarrier wait (all threads)", th, 1);
        run<3>("st.async peer + mbarrier wait (all threads)", th, 2);
        run<4>("4 dependent shuffles", th, 1);
        run<5>("f64 division", th, 1);
    }
    return 0;
}
