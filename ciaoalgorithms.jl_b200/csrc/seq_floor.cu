// seq_floor.cu — measures the latency floor of the sequential kernels' cluster exchange (seq_impl.cuh step 2) on the
// device it runs on: the same st.async → remote mbarrier (tx-count) → try_wait wake-up path, with nothing else in the loop.
//
//   mode 0  exchange only: every warp sends its 16-byte partial to all C CTAs, waits for the C·W partials of the round
//   mode 1  exchange + what every step must do with it: load the partials, tensor-core warp sum (bit-identical scalars
//           everywhere), and make the next round's message depend on that sum (a true chain, like a_i·w_{k+1})
//
// All clusters that fit on the GPU run the loop at the same time and report their own round time and the SMs they sit on:
// cross-SM DSMEM latency is a per-SM-pair constant (B300_MICROARCH.md "CGA; DSMEM": 186–240 cycles, pair-deterministic), a
// round ends with the slowest pair of the cluster, so the floor depends on where the cluster was placed.  bench.py reports
// the figure next to the measured step instead of a literal; DESIGN.md §4.2 uses the per-cluster spread.
#include "seq_impl.cuh"

struct FloorArgs {
    int iters, npart_pad;
    float *ns_per_round;   // [n_clusters]
    float *cyc_per_round;  // [n_clusters]
    int *smid;             // [n_clusters][C]
};

template <int MODE>
__global__ void __launch_bounds__(288, 1) exchange_floor_kernel(const FloorArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Tc = blockDim.x - 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = Tc >> 5;
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const int cluster_id = blockIdx.x / C;
    double *part = reinterpret_cast<double *>(smem_raw);  // [2][npart_pad][2]
    uint64_t *part_bar = reinterpret_cast<uint64_t *>(part + 2 * (size_t)p.npart_pad * 2);
    const uint32_t part_bytes = C * W * 16;
    for (int i = tid; i < 2 * p.npart_pad * 2; i += blockDim.x) part[i] = 0.0;
    if (tid == 0) {
        mbar_init(&part_bar[0], 1);
        mbar_init(&part_bar[1], 1);
        fence_mbar_init();
        uint32_t s;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        p.smid[cluster_id * C + rank] = (int)s;
    }
    __syncthreads();
    cluster_sync_all();
    const int K = p.iters;
    if (warp == W) {
        if (lane == 0) {   // re-arms the exchange barrier two rounds ahead, like the producer lane of seq_kernel
            const uint32_t pbs = smem_u32(part_bar);
            if (K > 0) mbar_arrive_expect_tx_s(pbs, part_bytes);
            if (K > 1) mbar_arrive_expect_tx_s(pbs + 8, part_bytes);
            for (int k = 0; k < K; ++k) {
                const uint32_t pb = pbs + ((uint32_t)k & 1u) * 8;
                mbar_wait_s(pb, ((uint32_t)k >> 1) & 1u);
                if (k + 2 < K) mbar_arrive_expect_tx_s(pb, part_bytes);
            }
        }
    } else {
        uint32_t send_dst[2], send_bar[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t peer = lane < C ? lane : 0;
            send_dst[q] = mapa_u32(smem_u32(part + ((size_t)q * p.npart_pad + rank * W + warp) * 2), peer);
            send_bar[q] = mapa_u32(smem_u32(&part_bar[q]), peer);
        }
        double v = 1.0 + 1e-3 * (double)(rank * W + warp);
        const long long c0 = clock64();
        uint64_t t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int k = 0; k < K; ++k) {
            const int par = k & 1;
            if (lane < C) st_async_v2f64(par ? send_dst[1] : send_dst[0], v, 0.0, par ? send_bar[1] : send_bar[0]);
            mbar_wait(&part_bar[par], (uint32_t)((k >> 1) & 1));
            if (MODE == 1) {
                const double2 w = reinterpret_cast<const double2 *>(part + (size_t)par * p.npart_pad * 2)[lane];
                const double u = warp_sum_mma(w.x, lane);
                v = fma(u, 1e-6, 1.0);   // the next message depends on this round's sum
            }
        }
        const long long c1 = clock64();
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (rank == 0 && tid == 0 && K > 0) {
            p.ns_per_round[cluster_id] = (float)((double)(t1 - t0) / K) + (v > 1e30 ? 1.0f : 0.0f);
            p.cyc_per_round[cluster_id] = (float)((double)(c1 - c0) / K);
        }
    }
    cluster_sync_all();
}

// ns_per_round / cyc_per_round: one entry per cluster (host arrays of max_clusters entries); smid: max_clusters·cluster entries
extern "C" int ciao_measure_exchange(ciao_ctx *c, int cluster, int warps, int iters, int mode, int max_clusters, int *n_clusters,
                                     float *ns_per_round, float *cyc_per_round, int *smid) {
    if (!c || !n_clusters || !ns_per_round || !cyc_per_round || !smid) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_measure_exchange: null argument");
    if ((cluster != 2 && cluster != 4 && cluster != 8 && cluster != 16) || warps < 1 || warps > 8 || cluster * warps > SEQ_MAX_PART ||
        iters < 1 || (mode != 0 && mode != 1) || max_clusters < 1)
        CIAO_FAIL(CIAO_ERR_INVALID, "ciao_measure_exchange: cluster in {2,4,8,16}, 1..8 warps, iters ≥ 1, mode 0/1");
    CUDA_TRY(cudaSetDevice(c->device));
    const int nc = std::min(max_clusters, c->num_sms / cluster);
    const int npart_pad = (cluster * warps + 31) / 32 * 32;
    const size_t smem = 2 * (size_t)npart_pad * 2 * 8 + 2 * 8 + 128;
    float *dev_f = nullptr;
    int *dev_i = nullptr;
    CUDA_TRY(cudaMalloc(&dev_f, 2 * (size_t)nc * sizeof(float)));
    CUDA_TRY(cudaMalloc(&dev_i, (size_t)nc * cluster * sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(dev_f, 0, 2 * (size_t)nc * sizeof(float), c->stream));
    CUDA_TRY(cudaMemsetAsync(dev_i, 0xff, (size_t)nc * cluster * sizeof(int), c->stream));
    FloorArgs a{iters, npart_pad, dev_f, dev_f + nc, dev_i};
    auto kern = mode == 0 ? exchange_floor_kernel<0> : exchange_floor_kernel<1>;
    if (cluster > 8) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc * cluster);
    cfg.blockDim = dim3(warps * 32 + 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(ns_per_round, dev_f, (size_t)nc * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cyc_per_round, dev_f + nc, (size_t)nc * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(smid, dev_i, (size_t)nc * cluster * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dev_f);
    cudaFree(dev_i);
    if (e != cudaSuccess) {
        ciao_set_error("ciao_measure_exchange: %s", cudaGetErrorString(e));
        return CIAO_ERR_CUDA;
    }
    c->timing.launches += 1;
    *n_clusters = nc;
    return CIAO_OK;
}
