// batch.cu — minibatch steps of Finito and LFinito as streaming passes.
//
// Inside a minibatch every component gradient is evaluated at the same z (Finito_basic.jl:110-117,
// Finito_LFinito.jl:93-99), so the rows of a batch are independent and only Σ over the batch enters av.
// For the reference's static batches (contiguous rows r(j−1)+1..rj, Finito_basic.jl:52-57) a batch of
// ≥ BATCH_MIN_ROWS rows is therefore one HBM-bound pass over a row window — the same TMA-ring / column-
// owner structure as row_pass_kernel (pass.cu), plus the table read-modify-write for Finito:
//
//   BATCH_FINITO   t_i = z − (γ_i/N)∇f_i(z);  Σ_i (t_i − s_i)·(γ̂/γ_i);  s_i ← t_i        24·d bytes per row
//   BATCH_LFINITO  Σ_i (γ̂/N)(∇f_i(z_full) − ∇f_i(z)),  Σ_i γ̂/γ_i                           8·d bytes per row
//
// followed by a fixed-order reduction of the CTA partials and a d-sized finish kernel (av update, prox).
// Summation order inside a batch differs from the reference's sequential loop (rounding-level, covered by
// the parity tolerance); results are bitwise reproducible run to run.
#include <type_traits>

#include "common.cuh"

// the local rows [*lo_loc, *lo_loc + *n_loc) of the global window [lo, lo + n) (minibatches on row shards)
int local_window(ciao_ctx *c, int64_t lo, int64_t n, int64_t *lo_loc, int64_t *n_loc) {
    if (c->il_block) {
        const int64_t sb = c->il_block * c->il_world;
        if (lo % sb != 0)
            CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "interleaved row shards: a minibatch must start at a multiple of block_rows·world = %lld rows (starts at %lld)",
                      (long long)sb, (long long)lo);
        *lo_loc = lo / c->il_world;
        *n_loc = il_count(n, c->il_block, c->il_rank, c->il_world);
        return CIAO_OK;
    }
    const int64_t a = std::max(lo, c->row0), b = std::min(lo + n, c->row0 + c->n_rows);
    *lo_loc = std::max<int64_t>(0, a - c->row0);
    *n_loc = std::max<int64_t>(0, b - a);
    return CIAO_OK;
}

enum { BATCH_FINITO = 0, BATCH_LFINITO = 1,
       BATCH_LFINITO_CZ = 2 };   // LFinito with c_i(z_full) of every row cached by the pass at z_full (batch_sm_kernel only)
constexpr int BATCH_MIN_ROWS = 256;
enum { BATCH_WINDOWS_DISJOINT = 0, BATCH_WINDOWS_ALIGNED = 1, BATCH_WINDOWS_ANY = 2 };

struct BatchArgs {
    const double *rec;   // first record of the batch window
    int64_t n_rows, ld, d_pad;
    const double *z, *zf;
    double *table;       // first table row of the window (Finito)
    double *ws, *fws;
    double cN;           // γ̂/N
    int stages;
};

template <int CPT, int MODE, int LOSS>
__global__ void __launch_bounds__(256, 2) batch_pass_kernel(const BatchArgs p) {
    constexpr int RPG = 16 / CPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const size_t stage_doubles = (size_t)RPG * p.ld;
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *red = ring + (size_t)S * stage_doubles;  // [2][RPG][32][2]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * RPG * 32 * 2);

    const int64_t n_groups = (p.n_rows + RPG - 1) / RPG;
    const int64_t my_count = (blockIdx.x < n_groups) ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint64_t policy = 0;
    for (int i = tid; i < 2 * RPG * 32 * 2; i += T) red[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        policy = l2_policy_evict_first();
    }
    __syncthreads();
    auto issue = [&](int64_t it) {
        const int64_t r0 = (blockIdx.x + it * (int64_t)gridDim.x) * RPG;
        const int rows = (int)min((int64_t)RPG, p.n_rows - r0);
        const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(double));
        const int slot = (int)(it % S);
        mbar_arrive_expect_tx(&full[slot], bytes);
        tma_load_1d_stream(ring + (size_t)slot * stage_doubles, p.rec + r0 * p.ld, bytes, &full[slot], policy);
    };
    if (tid == 0)
        for (int64_t it = 0; it < my_count && it < S; ++it) issue(it);

    double zr[CPT], zfr[CPT], acc[CPT];
    int col[CPT / 2];
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k) {
        col[k] = 2 * (tid + T * k);
        const bool v = col[k] < p.d_pad;
        if (!v) col[k] = -1;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            zr[2 * k + e] = v ? p.z[col[k] + e] : 0.0;
            zfr[2 * k + e] = (v && MODE == BATCH_LFINITO) ? p.zf[col[k] + e] : 0.0;
            acc[2 * k + e] = 0.0;
        }
    }
    double fsum = 0.0;  // thread 0: Σ γ̂/γ_i (LFinito)

    for (int64_t it = 0; it < my_count; ++it) {
        const int slot = (int)(it % S);
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const int par = (int)(it & 1);
        const int64_t r0 = (blockIdx.x + it * (int64_t)gridDim.x) * RPG;
        const int rows = (int)min((int64_t)RPG, p.n_rows - r0);
        // old table rows: plain coalesced loads, issued before the wait so that they overlap it
        double2 so[RPG][CPT / 2];
        if (MODE == BATCH_FINITO) {
#pragma unroll
            for (int r = 0; r < RPG; ++r)
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k)
                    so[r][k] = (r < rows && col[k] >= 0)
                                   ? __ldcs(reinterpret_cast<const double2 *>(p.table + (r0 + r) * p.d_pad + col[k]))
                                   : make_double2(0.0, 0.0);
        }
        mbar_wait(&full[slot], parity);
        const double *sp = ring + (size_t)slot * stage_doubles;
        double a[RPG][CPT], p0[RPG], p1[RPG], tb[RPG], tl[RPG], tgn[RPG], thg[RPG];
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            p0[r] = p1[r] = 0.0;
            const bool rv = r < rows;
            const double *rp = sp + (size_t)r * p.ld;
#pragma unroll
            for (int k = 0; k < CPT / 2; ++k) {
                double2 v = make_double2(0.0, 0.0);
                if (rv && col[k] >= 0) v = *reinterpret_cast<const double2 *>(rp + col[k]);
                a[r][2 * k] = v.x;
                a[r][2 * k + 1] = v.y;
                p0[r] = fma(v.x, zr[2 * k], p0[r]);
                p0[r] = fma(v.y, zr[2 * k + 1], p0[r]);
                if (MODE == BATCH_LFINITO) {
                    p1[r] = fma(v.x, zfr[2 * k], p1[r]);
                    p1[r] = fma(v.y, zfr[2 * k + 1], p1[r]);
                }
            }
            tb[r] = rv ? rp[p.d_pad + TAIL_B] : 0.0;
            tl[r] = rv ? rp[p.d_pad + TAIL_LAM] : 0.0;
            tgn[r] = rv ? rp[p.d_pad + TAIL_GAM_N] : 0.0;
            thg[r] = rv ? rp[p.d_pad + TAIL_HAT_GAM] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            p0[r] = warp_sum_mma(p0[r], lane);
            if (MODE == BATCH_LFINITO) p1[r] = warp_sum_mma(p1[r], lane);
            if (lane == 0) {
                red[((par * RPG + r) * 32 + warp) * 2] = p0[r];
                red[((par * RPG + r) * 32 + warp) * 2 + 1] = p1[r];
            }
        }
        __syncthreads();
        if (tid == 0 && it + S < my_count) issue(it + S);
        // warp partials (entries ≥ W stay zero) summed by the same tensor-core reduction in every warp; then the (row, dot)
        // coefficients, one evaluation per warp (loss_coef_lanes)
        constexpr int NP = (MODE == BATCH_LFINITO ? 2 : 1) * RPG;
        double uq[NP], bq[NP], lq[NP], cq[NP];
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            const double2 pr = *reinterpret_cast<const double2 *>(red + ((par * RPG + r) * 32 + lane) * 2);
            uq[r] = warp_sum_mma(pr.x, lane); bq[r] = tb[r]; lq[r] = tl[r];
            if (MODE == BATCH_LFINITO) { uq[RPG + r] = warp_sum_mma(pr.y, lane); bq[RPG + r] = tb[r]; lq[RPG + r] = tl[r]; }
        }
        loss_coef_lanes<LOSS, NP>(uq, bq, lq, lane, cq);
#pragma unroll
        for (int r = 0; r < RPG; ++r) {
            if (r >= rows) continue;
            const double cz = cq[r];
            if (MODE == BATCH_FINITO) {  // Finito_basic.jl:112-116
                const double cneg = -tgn[r], rr2 = thg[r];
                double *trow = p.table + (r0 + r) * p.d_pad;
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k) {
                    double t0 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k], cz, tl[r]), cneg), zr[2 * k]);
                    double t1 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k + 1], cz, tl[r]), cneg), zr[2 * k + 1]);
                    acc[2 * k] += __dmul_rn(__dsub_rn(t0, so[r][k].x), rr2);
                    acc[2 * k + 1] += __dmul_rn(__dsub_rn(t1, so[r][k].y), rr2);
                    if (col[k] >= 0) __stcs(reinterpret_cast<double2 *>(trow + col[k]), make_double2(t0, t1));
                }
            } else {  // Finito_LFinito.jl:94-98
                // (γ̂/N)(∇f_i(z_full) − ∇f_i(z)) = a_i · [(γ̂/N)·λ_i·(c_i(z_full) − c_i(z))]: one scalar per row and ONE fma per
                // element instead of ten fp64 operations (the pass was fp64-issue bound at 3.3 TB/s); the batch sum has its own
                // summation order anyway, and forming the coefficient difference first loses less to cancellation near z = z_full
                const double czf = cq[(MODE == BATCH_LFINITO ? RPG : 0) + r];
                const double wrow = p.cN * ((LOSS == CIAO_LOSS_LS ? tl[r] : 1.0) * (czf - cz));
#pragma unroll
                for (int e = 0; e < CPT; ++e) acc[e] = fma(a[r][e], wrow, acc[e]);
                if (tid == 0) fsum += thg[r];
            }
        }
    }
    double *wrow = p.ws + (size_t)blockIdx.x * p.d_pad;
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k)
        if (col[k] >= 0) *reinterpret_cast<double2 *>(wrow + col[k]) = make_double2(acc[2 * k], acc[2 * k + 1]);
    if (tid == 0) p.fws[blockIdx.x] = fsum;
}

// Fixed-order reduction of the CTA partials fused with the batch's closing update (one launch instead of two):
// Finito: av += Σ; z = prox_g(av, γ̂) (Finito_basic.jl:118).   LFinito: av += Σ + (z − z_full)·Σ γ̂/γ_i (:98)
__global__ void __launch_bounds__(32 * REDUCE_SLICES) batch_reduce_finish_kernel(const double *ws, const double *fws, int G,
                                                                                  double *av, double *z, const double *zf,
                                                                                  int64_t d_pad, int mode, double hat_gamma,
                                                                                  RegParams reg) {
    __shared__ double sm[REDUCE_SLICES][33];
    __shared__ double fsh;
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t j = blockIdx.x * 32ll + x;
    double s = 0.0;
    if (j < d_pad)
        for (int b = y; b < G; b += REDUCE_SLICES) s += ws[(size_t)b * d_pad + j];
    sm[y][x] = s;
    if (mode == BATCH_LFINITO && x == 0 && y == 0) {
        double f = 0.0;
        for (int b = 0; b < G; ++b) f += fws[b];
        fsh = f;
    }
    __syncthreads();
    if (y != 0 || j >= d_pad) return;
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < REDUCE_SLICES; ++q) t += sm[q][x];
    double a = __dadd_rn(av[j], t);
    if (mode == BATCH_LFINITO) {
        a = __dadd_rn(a, __dmul_rn(fsh, __dsub_rn(z[j], zf[j])));
        av[j] = a;
    } else {
        av[j] = a;
        const double lo = reg.lo_v ? reg.lo_v[j] : reg.lo_s, hi = reg.hi_v ? reg.hi_v[j] : reg.hi_s;
        const double gl = hat_gamma * reg.lambda;
        z[j] = prox_rt(reg.kind, a, gl, lo, hi);
    }
}

template <int CPT, int MODE>
static int launch_batch_loss(ciao_ctx *c, const BatchArgs &a, int grid, int T, size_t smem) {
    if (c->loss_kind == CIAO_LOSS_LS) {
        auto kern = batch_pass_kernel<CPT, MODE, CIAO_LOSS_LS>;
        static size_t configured[CIAO_MAX_DEVICES] = {};
        if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[c->device % CIAO_MAX_DEVICES] = smem;
        }
        kern<<<grid, T, smem, c->stream>>>(a);
    } else {
        auto kern = batch_pass_kernel<CPT, MODE, CIAO_LOSS_LOGISTIC>;
        static size_t configured[CIAO_MAX_DEVICES] = {};
        if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[c->device % CIAO_MAX_DEVICES] = smem;
        }
        kern<<<grid, T, smem, c->stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return CIAO_OK;
}

// ---------------------------------------------------------------------------
// All batches of a call in ONE persistent kernel (cooperative launch, 2 CTAs per SM).  The per-batch version above costs two
// to three dependent launches per batch (≈ 10 µs each of start-up and ramp) around ≈ 15 µs of HBM time; here a CTA walks the
// sequence of its row groups across all batches, its TMA ring keeps prefetching the rows of the NEXT batch while the grid
// closes the current one (rows do not depend on z), and a batch boundary is
//     partial d-vectors → workspace | grid barrier | distributed fixed-order reduction + av/z update | grid barrier | reload z.
// The grid barrier is a monotone global counter (release: __threadfence + atomicAdd by one thread per CTA; acquire: a
// ld.acquire.gpu spin) — co-residency is guaranteed by cudaLaunchCooperativeKernel.  Same arithmetic, same fixed summation
// order per batch as the per-batch path (CTA partials in CTA order), so results are bitwise reproducible run to run.
struct BatchPArgs {
    const double *rec;    // all row records
    int64_t ld, d_pad;
    double *table;        // all table rows (Finito) or nullptr
    const int64_t *b_lo;  // [n_batches] first row of batch j (0-based)
    const int64_t *b_n;   // [n_batches] rows of batch j
    int64_t n_batches;
    double *z, *av;       // state vectors (global): read at every batch start, written by the finish
    const double *zf;
    double *ws, *fws;     // [grid][d_pad], [grid]
    unsigned int *bar;    // grid barrier counter, zero at launch
    double cN, hat_gamma;
    RegParams reg;
    int stages;
    int red_cols;         // columns per CTA in the distributed reduction (4, 8, 16 or 32)
    int l2_prefetch;      // pull the CTA's rows (and table rows) of the NEXT batch into L2 while the grid closes the current one
};

#ifdef CIAO_SEQ_PROFILE
static __device__ long long g_batch_prof[8];   // CTA 0, thread 0: cycles in [rows, partial write, barrier 1, reduction, barrier 2, z reload]
#define BPROF_T(v) const long long v = clock64()
#define BPROF_ADD(i, a, b) if (bid == 0 && tid == 0) bprof[i] += (b) - (a)
#else
#define BPROF_T(v)
#define BPROF_ADD(i, a, b)
#endif
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

template <int CPT, int MODE, int LOSS>
__global__ void __launch_bounds__(256) batch_persistent_kernel(const BatchPArgs p) {
    constexpr int RPG = 16 / CPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const int64_t G = gridDim.x, bid = blockIdx.x;
    const size_t stage_doubles = (size_t)RPG * p.ld;
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *red = ring + (size_t)S * stage_doubles;          // [2][RPG][32][2]
    double *rsm = red + 2 * RPG * 32 * 2;                     // [9][33]: the distributed reduction
    uint64_t *full = reinterpret_cast<uint64_t *>(rsm + 9 * 33 + 1);

    uint64_t policy = 0;
    for (int i = tid; i < 2 * RPG * 32 * 2; i += T) red[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        policy = l2_policy_evict_first();
    }
    __syncthreads();

    // the sequence of (batch, group) items of this CTA: groups bid, bid + G, … of batch 0, then of batch 1, …
    auto first_item = [&](int64_t &b, int64_t &g) {
        b = 0; g = bid;
        while (b < p.n_batches && g >= (p.b_n[b] + RPG - 1) / RPG) { ++b; g = bid; }
    };
    auto next_item = [&](int64_t &b, int64_t &g) {
        g += G;
        while (b < p.n_batches && g >= (p.b_n[b] + RPG - 1) / RPG) { ++b; g = bid; }
    };
    int64_t pb, pg, issued = 0;  // producer cursor (thread 0)
    auto issue = [&]() {
        const int64_t r0 = p.b_lo[pb] + pg * RPG;
        const int rows = (int)min((int64_t)RPG, p.b_n[pb] - pg * RPG);
        const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(double));
        const int slot = (int)(issued % S);
        mbar_arrive_expect_tx(&full[slot], bytes);
        tma_load_1d_stream(ring + (size_t)slot * stage_doubles, p.rec + r0 * p.ld, bytes, &full[slot], policy);
        ++issued;
        next_item(pb, pg);
    };
    if (tid == 0) {
        first_item(pb, pg);
        for (int s = 0; s < S && pb < p.n_batches; ++s) issue();
    }

    int col[CPT / 2];
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k) {
        col[k] = 2 * (tid + T * k);
        if (col[k] >= p.d_pad) col[k] = -1;
    }
    int64_t cb, cg, it = 0;  // consumer cursor (all threads), items consumed so far
    first_item(cb, cg);
    unsigned int bar_target = 0;

#ifdef CIAO_SEQ_PROFILE
    long long bprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int64_t b = 0; b < p.n_batches; ++b) {
        BPROF_T(t_0);
        double zr[CPT], zfr[CPT], acc[CPT];
#pragma unroll
        for (int k = 0; k < CPT / 2; ++k)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool v = col[k] >= 0;
                zr[2 * k + e] = v ? __ldcg(p.z + col[k] + e) : 0.0;          // written by other CTAs in the last finish
                zfr[2 * k + e] = (v && MODE == BATCH_LFINITO) ? p.zf[col[k] + e] : 0.0;
                acc[2 * k + e] = 0.0;
            }
        double fsum = 0.0;
        const int64_t lo_b = p.b_lo[b], n_b = p.b_n[b];
        BPROF_T(t_1);
        BPROF_ADD(5, t_0, t_1);
        while (cb == b) {
            const int slot = (int)(it % S);
            const uint32_t parity = (uint32_t)((it / S) & 1);
            const int par = (int)(it & 1);
            const int64_t r0 = lo_b + cg * RPG;
            const int rows = (int)min((int64_t)RPG, n_b - cg * RPG);
            double2 so[RPG][CPT / 2];
            if (MODE == BATCH_FINITO) {
#pragma unroll
                for (int r = 0; r < RPG; ++r)
#pragma unroll
                    for (int k = 0; k < CPT / 2; ++k)
                        so[r][k] = (r < rows && col[k] >= 0)
                                       ? __ldcg(reinterpret_cast<const double2 *>(p.table + (r0 + r) * p.d_pad + col[k]))
                                       : make_double2(0.0, 0.0);
            }
            mbar_wait(&full[slot], parity);
            const double *sp = ring + (size_t)slot * stage_doubles;
            double a[RPG][CPT], p0[RPG], p1[RPG], tb[RPG], tl[RPG], tgn[RPG], thg[RPG];
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                p0[r] = p1[r] = 0.0;
                const bool rv = r < rows;
                const double *rp = sp + (size_t)r * p.ld;
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k) {
                    double2 v = make_double2(0.0, 0.0);
                    if (rv && col[k] >= 0) v = *reinterpret_cast<const double2 *>(rp + col[k]);
                    a[r][2 * k] = v.x;
                    a[r][2 * k + 1] = v.y;
                    p0[r] = fma(v.x, zr[2 * k], p0[r]);
                    p0[r] = fma(v.y, zr[2 * k + 1], p0[r]);
                    if (MODE == BATCH_LFINITO) {
                        p1[r] = fma(v.x, zfr[2 * k], p1[r]);
                        p1[r] = fma(v.y, zfr[2 * k + 1], p1[r]);
                    }
                }
                tb[r] = rv ? rp[p.d_pad + TAIL_B] : 0.0;
                tl[r] = rv ? rp[p.d_pad + TAIL_LAM] : 0.0;
                tgn[r] = rv ? rp[p.d_pad + TAIL_GAM_N] : 0.0;
                thg[r] = rv ? rp[p.d_pad + TAIL_HAT_GAM] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                p0[r] = warp_sum_mma(p0[r], lane);
                if (MODE == BATCH_LFINITO) p1[r] = warp_sum_mma(p1[r], lane);
                if (lane == 0) {
                    red[((par * RPG + r) * 32 + warp) * 2] = p0[r];
                    red[((par * RPG + r) * 32 + warp) * 2 + 1] = p1[r];
                }
            }
            __syncthreads();
            if (tid == 0 && pb < p.n_batches) issue();  // the slot just read is free: prefetch runs ahead across batch boundaries
            // warp partials (entries ≥ W stay zero) summed by the same tensor-core reduction in every warp; then the (row, dot)
            // coefficients, one evaluation per warp (loss_coef_lanes)
            constexpr int NP = (MODE == BATCH_LFINITO ? 2 : 1) * RPG;
            double uq[NP], bq[NP], lq[NP], cq[NP];
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                const double2 pr = *reinterpret_cast<const double2 *>(red + ((par * RPG + r) * 32 + lane) * 2);
                uq[r] = warp_sum_mma(pr.x, lane); bq[r] = tb[r]; lq[r] = tl[r];
                if (MODE == BATCH_LFINITO) { uq[RPG + r] = warp_sum_mma(pr.y, lane); bq[RPG + r] = tb[r]; lq[RPG + r] = tl[r]; }
            }
            loss_coef_lanes<LOSS, NP>(uq, bq, lq, lane, cq);
#pragma unroll
            for (int r = 0; r < RPG; ++r) {
                if (r >= rows) continue;
                const double cz = cq[r];
                if (MODE == BATCH_FINITO) {  // Finito_basic.jl:112-116
                    const double cneg = -tgn[r], rr2 = thg[r];
                    double *trow = p.table + (r0 + r) * p.d_pad;
#pragma unroll
                    for (int k = 0; k < CPT / 2; ++k) {
                        double t0 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k], cz, tl[r]), cneg), zr[2 * k]);
                        double t1 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k + 1], cz, tl[r]), cneg), zr[2 * k + 1]);
                        acc[2 * k] += __dmul_rn(__dsub_rn(t0, so[r][k].x), rr2);
                        acc[2 * k + 1] += __dmul_rn(__dsub_rn(t1, so[r][k].y), rr2);
                        if (col[k] >= 0) __stcg(reinterpret_cast<double2 *>(trow + col[k]), make_double2(t0, t1));
                    }
                } else {  // Finito_LFinito.jl:94-98
                    const double czf = cq[(MODE == BATCH_LFINITO ? RPG : 0) + r];   // see batch_pass_kernel: one fma per element
                    const double wrow = p.cN * ((LOSS == CIAO_LOSS_LS ? tl[r] : 1.0) * (czf - cz));
#pragma unroll
                    for (int e = 0; e < CPT; ++e) acc[e] = fma(a[r][e], wrow, acc[e]);
                    if (tid == 0) fsum += thg[r];
                }
            }
            next_item(cb, cg);
            ++it;
        }
        // ---- close the batch ----
        BPROF_T(t_2);
        BPROF_ADD(0, t_1, t_2);
        double *wrow = p.ws + (size_t)bid * p.d_pad;
#pragma unroll
        for (int k = 0; k < CPT / 2; ++k)
            if (col[k] >= 0) __stcg(reinterpret_cast<double2 *>(wrow + col[k]), make_double2(acc[2 * k], acc[2 * k + 1]));
        if (tid == 0) __stcg(p.fws + bid, fsum);
        // The batch boundary (two grid barriers + the reduction, ≈ 8 µs) is dead time for HBM: the ring holds only S row groups.
        // L2 (126 MB) holds a whole batch, so the rows this CTA will stream next are requested now and the boundary overlaps
        // with their HBM traffic; the next rows phase then runs out of L2.
        if (p.l2_prefetch && warp == 0 && b + 1 < p.n_batches) {
            const int64_t lo_n = p.b_lo[b + 1], n_n = p.b_n[b + 1];
            const int64_t ng = (n_n + RPG - 1) / RPG;
            for (int64_t g = bid + (int64_t)lane * G; g < ng; g += 32 * G) {
                const int64_t r0 = lo_n + g * RPG;
                const uint32_t nr = (uint32_t)min((int64_t)RPG, n_n - g * RPG);
                tma_prefetch_l2(p.rec + r0 * p.ld, nr * (uint32_t)(p.ld * sizeof(double)));
                if (MODE == BATCH_FINITO) tma_prefetch_l2(p.table + r0 * p.d_pad, nr * (uint32_t)(p.d_pad * sizeof(double)));
            }
        }
        bar_target += (unsigned int)G;
        BPROF_T(t_3);
        BPROF_ADD(1, t_2, t_3);
        grid_barrier(p.bar, bar_target);
        BPROF_T(t_4);
        BPROF_ADD(2, t_3, t_4);
        // distributed fixed-order reduction over ALL CTAs: CTA c owns the RC columns RC·c … (RC = 4…32, a whole number of
        // sectors); thread (col, slice) sums the CTA partials slice, slice + NS, … with four independent chains, the slices
        // are combined by shuffles inside a warp and through shared memory across the 8 warps — always in the same order
        {
            const int RC = p.red_cols, NS = T / RC, Wn = T >> 5;
            const int cl = tid % RC, sl = tid / RC;
            const int64_t j = bid * RC + cl;
            const bool jv = j < p.d_pad;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (jv) {
                int64_t q = sl;
                for (; q + 3 * NS < G; q += 4 * NS) {
                    s0 += __ldcg(p.ws + (size_t)q * p.d_pad + j);
                    s1 += __ldcg(p.ws + (size_t)(q + NS) * p.d_pad + j);
                    s2 += __ldcg(p.ws + (size_t)(q + 2 * NS) * p.d_pad + j);
                    s3 += __ldcg(p.ws + (size_t)(q + 3 * NS) * p.d_pad + j);
                }
                for (; q < G; q += NS) s0 += __ldcg(p.ws + (size_t)q * p.d_pad + j);
            }
            double sacc = (s0 + s1) + (s2 + s3);
            for (int o = RC; o < 32; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane < RC) rsm[warp * 33 + lane] = sacc;
            double f = 0.0;
            if (MODE == BATCH_LFINITO) {  // Σ γ̂/γ_i of the batch: every CTA needs it; thread t takes fws[t], fws[t+256], …
                for (int64_t q = tid; q < G; q += T) f += __ldcg(p.fws + q);
                f = warp_sum(f);
                if (lane == 0) rsm[8 * 33 + warp] = f;
            }
            __syncthreads();
            if (tid < RC && jv) {
                double t = 0.0;
                for (int w = 0; w < Wn; ++w) t += rsm[w * 33 + tid];
                double anew = __dadd_rn(__ldcg(p.av + j), t);
                const double lo = p.reg.lo_v ? p.reg.lo_v[j] : p.reg.lo_s, hi = p.reg.hi_v ? p.reg.hi_v[j] : p.reg.hi_s;
                const double gl = p.hat_gamma * p.reg.lambda;
                if (MODE == BATCH_LFINITO) {
                    double fs = 0.0;
                    for (int w = 0; w < Wn; ++w) fs += rsm[8 * 33 + w];
                    anew = __dadd_rn(anew, __dmul_rn(fs, __dsub_rn(__ldcg(p.z + j), p.zf[j])));              // :98
                    __stcg(p.av + j, anew);
                    if (b + 1 < p.n_batches) __stcg(p.z + j, prox_rt(p.reg.kind, anew, gl, lo, hi));          // :92 of the next batch
                } else {
                    __stcg(p.av + j, anew);
                    __stcg(p.z + j, prox_rt(p.reg.kind, anew, gl, lo, hi));                                   // Finito_basic.jl:118
                }
            }
            __syncthreads();  // rsm is reused by the next batch
        }
        bar_target += (unsigned int)G;
        BPROF_T(t_5);
        BPROF_ADD(3, t_4, t_5);
        grid_barrier(p.bar, bar_target);
        BPROF_T(t_6);
        BPROF_ADD(4, t_5, t_6);
    }
#ifdef CIAO_SEQ_PROFILE
    if (bid == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) g_batch_prof[i] = bprof[i];
#endif
}
#ifdef CIAO_SEQ_PROFILE
extern "C" int ciao_debug_batch_prof(ciao_ctx *c, long long *out8) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out8, g_batch_prof, 8 * sizeof(long long)));
    return CIAO_OK;
}
#endif

// ---------------------------------------------------------------------------
// The persistent kernel with ONE CTA PER SM and the batch boundary as a flagged-word exchange (default; the grid-barrier
// version above stays selectable with CIAO_BATCH_EXCHANGE=barrier).
//
// What was wrong with the barrier version (profiles/ncu_lfinito_batch_barrier_r2.csv, warp samples, LFinito, batch 4096, 592 CTAs):
// 32 % of the warp time in the row loop, 20 % waiting in grid barrier 1, 14 % in the reduction, 25 % in grid barrier 2, 9 %
// reloading z.  The boundary is a chain of dependent L2 hops (≈ 0.6 µs each when 300–600 CTAs poll), two fences, and 2·G atomics
// on one address.  Three changes:
//
// * One exchange participant per SM.  The CTA has NSG sub-groups of TS threads (what used to be 2–4 CTAs per SM); a sub-group
//   walks its own items with its own TMA ring and named barrier, exactly like a CTA of the old kernel with the virtual id
//   sg·G + bid.  At the end of a batch the sub-groups' partial vectors are added through shared memory (fixed order), so the
//   grid exchanges G = #SMs partials instead of 2–4 times as many, and z is fetched once per SM.
// * Data carries its own readiness.  A double travels as two 64-bit words {32 data bits | 32-bit epoch} (the layout of NCCL's LL
//   protocol), written with one relaxed 16-byte store and read with one relaxed 16-byte load; a 64-bit word cannot tear, so a
//   word whose epoch matches holds this batch's data.  No fence, no atomic, no barrier: the boundary is two hops,
//       CTA partials → owner of RC columns (fixed-order sum of the G partials, av/z update) → every CTA (new z).
//   One buffer per hop suffices: a CTA writes its partial of batch b+1 only after it has received z of batch b for the same
//   columns, which the owner produced after reading every partial of batch b.  The epoch counter lives in the context and
//   keeps counting across launches; the buffers start zeroed.
// * Σ γ̂/γ_i of a batch (LFinito) does not depend on the iterate: batch_fsum_kernel forms it for all batches before the launch.
//
// Tried and dropped (profiles/batch_exchange_r2.md): the same exchange between 2–4 CTAs per SM (every owner polls 592 partials,
// every CTA polls z: the polls serialise in a few L2 slices, the slowest owner finished 8–11 µs after the last partial), 16
// global arrival counters as a gate (8 µs), and a two-level version with groups of 8 CTAs (six hops: 6 µs).
// Summation order: sub-groups, then CTAs (then ranks) in a fixed order per column — bitwise reproducible run to run.
//
// Table rows (Finito) are read and rewritten by the same thread of the same SM whenever a window repeats inside a call as the same
// window (program order: no fence); when repeated windows are not aligned — a row may then move to another SM — `fence` puts a
// __threadfence() before every store and after every successful poll of the exchange (the chain partial → owner → z → reader
// is per column).  (Staging the table rows in the TMA ring with cp.async, so that the prefetch crosses the boundary, was built
// and measured: 19.3 µs per 4096-row batch against 18.9 without, 270 against 242 µs at 65 536 rows — dropped; so was an L2
// prefetch of the item's table rows issued together with its TMA copy: 18.8 / 262 µs.)
struct BatchLArgs {
    const double *rec;
    int64_t ld, d_pad;
    double *table;
    const int64_t *b_lo, *b_n;
    int64_t n_batches;
    double *z, *av;
    const double *zf;
    unsigned long long *llz;    // [d_pad][2]        new z, written by the column owners
    unsigned long long *llws;   // [grid][d_pad][2]  CTA partials
    const double *bfs;          // LFinito: Σ γ̂/γ_i of every batch (batch_fsum_kernel)
    const double *ss;           // LFinito with cached coefficients: {b_i, λ_i, 0, c_i(z_full)} of this context's rows
    uint32_t epoch0;            // flags of this launch: epoch0 + b + 1
    int fence;
    double cN, hat_gamma;
    RegParams reg;
    int stages;
    int red_cols;
    int ts, nsg;                // threads per sub-group, sub-groups per CTA
    // row shards (one process per GPU): the owners of a column on the `world` ranks exchange their rank sums through the peer
    // arenas (flagged words, system scope) and add them in rank order, so every rank forms the same av and z
    int world, rank;
    unsigned long long *xw[CIAO_MAX_PEERS];   // every rank's flagged-word area as addressed from this GPU (own entry: local)
    uint32_t xepoch0;
    unsigned long long xtimeout_ns;           // a peer that does not show up for this long (ciao_ctx::p2p_timeout_ns) ends the kernel with a trap
};

__device__ __forceinline__ void ll_store(unsigned long long *p, double v, uint32_t flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), f = (unsigned long long)flag << 32;
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | f), "l"((b >> 32) | f) : "memory");
}
__device__ __forceinline__ bool ll_load(const unsigned long long *p, uint32_t flag, double &v) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return (uint32_t)(lo >> 32) == flag && (uint32_t)(hi >> 32) == flag;
}
// a partner that never writes would hang the GPU: after ≈ 2^22 polls (seconds) the kernel traps instead
#define LL_SPIN_LIMIT (1 << 22)
// between two unsuccessful polls: nothing (the load's own round trip paces the loop; __nanosleep(32) cost ≈ 1.5 µs per batch)
#ifdef CIAO_LL_NANOSLEEP
#define LL_BACKOFF() __nanosleep(CIAO_LL_NANOSLEEP)
#else
#define LL_BACKOFF()
#endif
__device__ __forceinline__ void ll_store_sys(unsigned long long *p, double v, uint32_t flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), f = (unsigned long long)flag << 32;
    asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | f), "l"((b >> 32) | f) : "memory");
}
__device__ __forceinline__ bool ll_load_sys(const unsigned long long *p, uint32_t flag, double &v) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return (uint32_t)(lo >> 32) == flag && (uint32_t)(hi >> 32) == flag;
}
__device__ __forceinline__ void bar_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// profile build: g_batch_prof = cycles of CTA 0, thread 0 in [rows, combine + partial write, owner: poll + sum + update, z poll];
// g_batch_trace[bid] = {SM id, global timer at rows start, rows end, owner part done, new z received} of the middle batch
#ifdef CIAO_SEQ_PROFILE
static __device__ unsigned long long g_batch_trace[1200 * 8];
static __device__ unsigned int g_batch_dur[1200 * 4];   // rows phase (ns) of four consecutive batches from the middle one on
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define BTRACE(i) if (tid == 0 && b == p.n_batches / 2 && bid < 1200) g_batch_trace[bid * 8 + (i)] = gtimer()
#define BDUR_START() const unsigned long long dur_t0 = gtimer()
#define BDUR_END() if (tid == 0 && bid < 1200 && b >= p.n_batches / 2 && b < p.n_batches / 2 + 4) g_batch_dur[bid * 4 + (b - p.n_batches / 2)] = (unsigned int)(gtimer() - dur_t0)
#else
#define BTRACE(i)
#define BDUR_START()
#define BDUR_END()
#endif

// Σ γ̂/γ_i over the rows of every batch window (Finito_LFinito.jl:98), one CTA per batch, fixed order
__global__ void __launch_bounds__(256) batch_fsum_kernel(const double *rec, int64_t ld, int64_t d_pad, const int64_t *b_lo,
                                                         const int64_t *b_n, double *out) {
    __shared__ double sm[8];
    const int64_t lo = b_lo[blockIdx.x], n = b_n[blockIdx.x];
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) s += rec[(lo + r) * ld + d_pad + TAIL_HAT_GAM];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        out[blockIdx.x] = t;
    }
}

// (The warp sums stay on the fp64 tensor core.  A DMMA.8x8x4 occupies the sub-partition's fp64 pipe for 16 cycles against 2 for a
// DFMA — scripts/micro/fp64_tput.cu — so the 16 DMMAs of an LFinito item are 57 % of its fp64 pipe time; replacing them by a
// transpose-reduce with 6 double shuffles for the 4 sums of an item, and the second-stage sums by plain adds, halved the pipe
// load and changed nothing: 12.0 vs 11.9 µs per 4096-row batch, 115 vs 112 µs per 65 536-row batch.  The row loop is bound by
// the dependent chain of an item, not by the pipe.)
// threads per CTA: LFinito at 4 or 8 columns per thread keeps ≤ 128 registers (4 sub-groups of 128 or 2 of 256 threads)
template <int CPT, int MODE>
struct BatchSmShape {
    static constexpr int MAXT = (MODE != BATCH_FINITO && (CPT == 4 || CPT == 8)) ? 512 : 256;
};

template <int CPT, int MODE, int LOSS>
__global__ void __launch_bounds__(BatchSmShape<CPT, MODE>::MAXT, 1) batch_sm_kernel(const BatchLArgs p) {
    constexpr int RPG = 16 / CPT;
    // LF: an LFinito sweep.  CZ: c_i(z_full) of the rows comes from the scalars the pass at z_full left (ciao_ctx::ss, one 32-byte
    // sector per row, staged with the rows), so an item needs ONE dot product per row, half the warp sums, and no z_full in
    // shared memory — the item loop is bound by its dependent chain and by the fp64 pipe, not by HBM (§4.3).
    constexpr bool LF = MODE != BATCH_FINITO, CZ = MODE == BATCH_LFINITO_CZ, TWO_DOTS = LF && !CZ;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Tall = blockDim.x;
    const int TS = p.ts, NSG = p.nsg;
    const int sg = tid / TS, lt = tid - sg * TS, lwarp = lt >> 5;   // sub-group, thread and warp inside it
    const int S = p.stages;
    const int64_t G = gridDim.x, bid = blockIdx.x;
    const int64_t VG = G * NSG, vb = (int64_t)sg * G + bid;         // virtual CTA of the item distribution
    const size_t rec_doubles = (size_t)RPG * p.ld;
    const size_t stage_doubles = rec_doubles + (CZ ? RPG * 4 : 0);
    // shared memory: per sub-group {ring[S], red[2][RPG][32][2]}; then comb[NSG][d_pad] (sub-group partials, NSG > 1), zsm[d_pad] (z of
    // the current batch), zfs[d_pad] (z_full, LFinito), rsm[17][33], own[5][32], full[NSG][4].  z and z_full are read from
    // shared memory in every item: 32 registers less, which is what lets LFinito run 512 threads without spilling.
    const size_t sg_doubles = (size_t)S * stage_doubles + 2 * RPG * 32 * 2;
    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)sg * sg_doubles;
    double *red = ring + (size_t)S * stage_doubles;
    double *comb = reinterpret_cast<double *>(smem_raw) + (size_t)NSG * sg_doubles;
    double *zsm = comb + (size_t)(NSG > 1 ? NSG : 0) * p.d_pad;
    double *zfs = zsm + p.d_pad;
    double *rsm = zfs + (TWO_DOTS ? p.d_pad : 0);
    double *own = rsm + 17 * 33 + 1;
    uint64_t *full = reinterpret_cast<uint64_t *>(own + 5 * 32) + sg * 4;

    uint64_t policy = 0;
    for (int i = lt; i < 2 * RPG * 32 * 2; i += TS) red[i] = 0.0;
    for (int i = tid; i < 17 * 33 + 1; i += Tall) rsm[i] = 0.0;
    for (int64_t c = tid; c < p.d_pad; c += Tall) {
        zsm[c] = p.z[c];
        if (TWO_DOTS) zfs[c] = p.zf[c];
    }
    if (lt == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        policy = l2_policy_evict_first();
    }
    __syncthreads();

    int col[CPT / 2];
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k) {
        col[k] = 2 * (lt + TS * k);
        if (col[k] >= p.d_pad) col[k] = -1;
    }
    const bool all_cols = p.d_pad == (int64_t)CPT * TS;   // every thread owns CPT valid columns
    // the sequence of (batch, group) items of this sub-group: groups vb, vb + VG, … of batch 0, then of batch 1, …
    // Producer cursor (thread 0 of the sub-group drives the TMA): batch pb with its window (p_lo, p_n) and group count p_ng in
    // registers, group pg, ring slot p_slot — no division and no reload of the window on the per-item path.
    int64_t pb = -1, pg = 0, p_lo = 0, p_n = 0, p_ng = 0;
    int p_slot = 0;
    auto next_batch = [&]() {   // first batch after pb in which this sub-group has an item
        for (++pb; pb < p.n_batches; ++pb) {
            p_n = p.b_n[pb];
            p_ng = (p_n + RPG - 1) / RPG;
            if (vb < p_ng) { p_lo = p.b_lo[pb]; pg = vb; return; }
        }
    };
    auto issue = [&]() {
        if (lt == 0 && pb < p.n_batches) {
            const int64_t r0 = p_lo + pg * RPG;
            const int rows = (int)min((int64_t)RPG, p_n - pg * RPG);
            const uint32_t bytes = (uint32_t)(rows * p.ld * sizeof(double));
            mbar_arrive_expect_tx(&full[p_slot], bytes + (CZ ? (uint32_t)rows * 32u : 0u));
            tma_load_1d_stream(ring + (size_t)p_slot * stage_doubles, p.rec + r0 * p.ld, bytes, &full[p_slot], policy);
            if (CZ) tma_load_1d_stream(ring + (size_t)p_slot * stage_doubles + rec_doubles, p.ss + r0 * 4, (uint32_t)rows * 32u, &full[p_slot], policy);
            p_slot = p_slot + 1 == S ? 0 : p_slot + 1;
            pg += VG;
            if (pg >= p_ng) next_batch();
        }
    };
    if (lt == 0) next_batch();
    for (int s = 0; s < S; ++s) issue();
    // Consumer: ring slot, its mbarrier phase, parity of the red[] buffer.  LFinito keeps them incrementally; Finito derives
    // slot and phase from the item count with a division by the (run-time) ring depth — measured, three runs each on two GPUs
    // (profiles/batch_exchange_r2.md §4): Finito 18.9 / 64.8 / 242 µs per batch of 4096 / 16 384 / 65 536 rows with the
    // division against 22.2 / 80 / 306 µs without; LFinito 11.1 / 30.8 / 111.6 against 10.8 / 29.4 / 106.1 (one-dot sweep: 9.8 / 25.8 /
    // 91.8 against 9.4 / 24.1 / 86.2).  The SASS of the two
    // Finito builds has the same loads, barriers and stores in the same order; like the sequential kernels (profiles/
    // seq_kernel_history_r2.md §6) the item loop is sensitive to ptxas' schedule.  Re-measure after touching this loop.
    int64_t it = 0;
    int c_slot = 0;
    uint32_t c_phase = 0;
    int par = 0;

    // owner role: CTA c owns the RC columns RC·c … of av and z (a whole number of 32-byte sectors), for the whole call
    const int RC = p.red_cols, NS = Tall / RC, Wn = Tall >> 5;
    const int cl = tid % RC, sl = tid / RC;
    const bool owner = bid * RC < p.d_pad;
    const int64_t j_red = bid * RC + cl;       // column this thread sums (slice sl of the partials)
    const int64_t j_own = bid * RC + tid;      // column thread tid < RC finishes
    const bool fin = owner && tid < RC && j_own < p.d_pad;
    if (fin) {   // only thread tid touches own[·][tid]: no synchronisation needed
        own[tid] = p.av[j_own];
        own[32 + tid] = p.z[j_own];
        own[64 + tid] = LF ? p.zf[j_own] : 0.0;
        own[96 + tid] = p.reg.lo_v ? p.reg.lo_v[j_own] : p.reg.lo_s;
        own[128 + tid] = p.reg.hi_v ? p.reg.hi_v[j_own] : p.reg.hi_s;
    }
    const double gl = p.hat_gamma * p.reg.lambda;

#ifdef CIAO_SEQ_PROFILE
    long long bprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int64_t b = 0; b < p.n_batches; ++b) {
        BPROF_T(t_1);
        BTRACE(1);
        BDUR_START();
        const uint32_t ep = p.epoch0 + (uint32_t)b + 1u;
        double acc[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) acc[e] = 0.0;
        const int64_t lo_b = p.b_lo[b], n_b = p.b_n[b], ng_b = (n_b + RPG - 1) / RPG;
        for (int64_t cg = vb; cg < ng_b; cg += VG) {
            const int slot = MODE == BATCH_FINITO ? (int)(it % S) : c_slot;
            const uint32_t parity = MODE == BATCH_FINITO ? (uint32_t)((it / S) & 1) : c_phase;
            const int64_t r0 = lo_b + cg * RPG;
            const int rows = (int)min((int64_t)RPG, n_b - cg * RPG);
            const double *sp = ring + (size_t)slot * stage_doubles;
            // FULL: all RPG rows present and every thread owns CPT valid columns — no predicates in the item (d = 1024, 4096, …)
            auto item = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                double2 so[RPG][CPT / 2];
                if (MODE == BATCH_FINITO) {   // old table rows: plain loads, issued before the wait so that they overlap it
#pragma unroll
                    for (int r = 0; r < RPG; ++r)
#pragma unroll
                        for (int k = 0; k < CPT / 2; ++k)
                            so[r][k] = (FULL || (r < rows && col[k] >= 0))
                                           ? __ldcg(reinterpret_cast<const double2 *>(p.table + (r0 + r) * p.d_pad + col[k]))
                                           : make_double2(0.0, 0.0);
                }
                double zr[CPT], zfr[CPT];
#pragma unroll
                for (int k = 0; k < CPT / 2; ++k) {
                    const bool v = FULL || col[k] >= 0;
                    const double2 zv = v ? *reinterpret_cast<const double2 *>(zsm + col[k]) : make_double2(0.0, 0.0);
                    zr[2 * k] = zv.x; zr[2 * k + 1] = zv.y;
                    if (TWO_DOTS) {
                        const double2 fv = v ? *reinterpret_cast<const double2 *>(zfs + col[k]) : make_double2(0.0, 0.0);
                        zfr[2 * k] = fv.x; zfr[2 * k + 1] = fv.y;
                    }
                }
                mbar_wait(&full[slot], parity);
                double a[RPG][CPT], p0[RPG], p1[RPG], tb[RPG], tl[RPG], tgn[RPG], thg[RPG], czfv[RPG];
#pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    p0[r] = p1[r] = 0.0;
                    const bool rv = FULL || r < rows;
                    const double *rp = sp + (size_t)r * p.ld;
#pragma unroll
                    for (int k = 0; k < CPT / 2; ++k) {
                        double2 v = make_double2(0.0, 0.0);
                        if (FULL || (rv && col[k] >= 0)) v = *reinterpret_cast<const double2 *>(rp + col[k]);
                        a[r][2 * k] = v.x;
                        a[r][2 * k + 1] = v.y;
                        p0[r] = fma(v.x, zr[2 * k], p0[r]);
                        p0[r] = fma(v.y, zr[2 * k + 1], p0[r]);
                        if (TWO_DOTS) {
                            p1[r] = fma(v.x, zfr[2 * k], p1[r]);
                            p1[r] = fma(v.y, zfr[2 * k + 1], p1[r]);
                        }
                    }
                    tb[r] = rv ? rp[p.d_pad + TAIL_B] : 0.0;
                    tl[r] = rv ? rp[p.d_pad + TAIL_LAM] : 0.0;
                    if (CZ) czfv[r] = rv ? sp[rec_doubles + (size_t)r * 4 + 3] : 0.0;   // {b_i, λ_i, 0, c_i(z_full)}: read before the slot is handed back
                    if (MODE == BATCH_FINITO) {
                        tgn[r] = rv ? rp[p.d_pad + TAIL_GAM_N] : 0.0;
                        thg[r] = rv ? rp[p.d_pad + TAIL_HAT_GAM] : 0.0;
                    }
                }
#pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    p0[r] = warp_sum_mma(p0[r], lane);
                    if (TWO_DOTS) p1[r] = warp_sum_mma(p1[r], lane);
                    if (lane == 0) {
                        red[((par * RPG + r) * 32 + lwarp) * 2] = p0[r];
                        red[((par * RPG + r) * 32 + lwarp) * 2 + 1] = p1[r];
                    }
                }
                bar_named(1 + sg, TS);
                issue();  // the slot just read is free: the prefetch runs ahead across batch boundaries
                constexpr int NP = (TWO_DOTS ? 2 : 1) * RPG;
                double uq[NP], bq[NP], lq[NP], cq[NP];
#pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    const double2 pr = *reinterpret_cast<const double2 *>(red + ((par * RPG + r) * 32 + lane) * 2);
                    uq[r] = warp_sum_mma(pr.x, lane); bq[r] = tb[r]; lq[r] = tl[r];
                    if (TWO_DOTS) { uq[RPG + r] = warp_sum_mma(pr.y, lane); bq[RPG + r] = tb[r]; lq[RPG + r] = tl[r]; }
                }
                loss_coef_lanes<LOSS, NP>(uq, bq, lq, lane, cq);
#pragma unroll
                for (int r = 0; r < RPG; ++r) {
                    if (!FULL && r >= rows) continue;
                    const double cz = cq[r];
                    if (MODE == BATCH_FINITO) {  // Finito_basic.jl:112-116
                        const double cneg = -tgn[r], rr2 = thg[r];
                        double *trow = p.table + (r0 + r) * p.d_pad;
#pragma unroll
                        for (int k = 0; k < CPT / 2; ++k) {
                            double t0 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k], cz, tl[r]), cneg), zr[2 * k]);
                            double t1 = __dadd_rn(__dmul_rn(grad_elem<LOSS>(a[r][2 * k + 1], cz, tl[r]), cneg), zr[2 * k + 1]);
                            acc[2 * k] += __dmul_rn(__dsub_rn(t0, so[r][k].x), rr2);
                            acc[2 * k + 1] += __dmul_rn(__dsub_rn(t1, so[r][k].y), rr2);
                            if (FULL || col[k] >= 0) __stcg(reinterpret_cast<double2 *>(trow + col[k]), make_double2(t0, t1));
                        }
                    } else {  // Finito_LFinito.jl:94-98, one fma per element (see batch_pass_kernel)
                        const double czf = CZ ? czfv[r] : cq[(TWO_DOTS ? RPG : 0) + r];
                        const double wrow = p.cN * ((LOSS == CIAO_LOSS_LS ? tl[r] : 1.0) * (czf - cz));
#pragma unroll
                        for (int e = 0; e < CPT; ++e) acc[e] = fma(a[r][e], wrow, acc[e]);
                    }
                }
            };
            if (all_cols && rows == RPG) item(std::true_type{});
            else item(std::false_type{});
            par ^= 1;
            if (++c_slot == S) { c_slot = 0; c_phase ^= 1u; }
            ++it;
        }
        // ---- close the batch ----
        BPROF_T(t_2);
        BPROF_ADD(0, t_1, t_2);
        BTRACE(2);
        BDUR_END();
        const double bfs_b = (LF && fin) ? p.bfs[b] : 0.0;   // needed at the very end of the boundary: loaded now
        // (1) + (2) the sub-groups' partials meet in shared memory; every thread of the CTA then adds the NSG values of its
        // columns in a fixed order and sends them to the column owners as flagged words (hop 1)
        if (p.fence) __threadfence();
        unsigned long long *wrow = p.llws + (size_t)bid * p.d_pad * 2;
        if (NSG > 1) {
            double *dst = comb + (size_t)sg * p.d_pad;
#pragma unroll
            for (int k = 0; k < CPT / 2; ++k)
                if (col[k] >= 0) *reinterpret_cast<double2 *>(dst + col[k]) = make_double2(acc[2 * k], acc[2 * k + 1]);
            __syncthreads();
            BTRACE(5);
            if (p.fence) __threadfence();   // the storing thread's own fence, after the barrier that ordered the sub-groups' table stores before it
            for (int64_t c = 2 * (int64_t)tid; c < p.d_pad; c += 2 * (int64_t)Tall) {
                double2 v0 = *reinterpret_cast<const double2 *>(comb + c);
                const double2 v1 = *reinterpret_cast<const double2 *>(comb + p.d_pad + c);
                if (NSG == 4) {
                    const double2 v2 = *reinterpret_cast<const double2 *>(comb + 2 * p.d_pad + c);
                    const double2 v3 = *reinterpret_cast<const double2 *>(comb + 3 * p.d_pad + c);
                    v0.x = (v0.x + v1.x) + (v2.x + v3.x);
                    v0.y = (v0.y + v1.y) + (v2.y + v3.y);
                } else {
                    v0.x += v1.x;
                    v0.y += v1.y;
                }
                ll_store(wrow + (size_t)c * 2, v0.x, ep);
                ll_store(wrow + (size_t)c * 2 + 2, v0.y, ep);
            }
        } else {
#pragma unroll
            for (int k = 0; k < CPT / 2; ++k)
                if (col[k] >= 0) {
                    ll_store(wrow + (size_t)col[k] * 2, acc[2 * k], ep);
                    ll_store(wrow + (size_t)col[k] * 2 + 2, acc[2 * k + 1], ep);
                }
        }
        BPROF_T(t_3);
        BPROF_ADD(1, t_2, t_3);
        BTRACE(6);
        if (owner) {
            // thread (column cl, slice sl) sums the CTA partials sl, sl + NS, … of its column in a fixed order: polled eight at a
            // time (independent loads, one L2 round trip when the partials are there), four summation chains
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (j_red < p.d_pad) {
                const unsigned long long *src = p.llws + (size_t)j_red * 2;
                for (int64_t q0 = sl; q0 < G; q0 += 8 * NS) {
                    double v[8];
                    for (int spins = 0;; ++spins) {
                        bool ok = true;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int64_t q = q0 + (int64_t)u * NS;
                            v[u] = 0.0;
                            if (q < G) ok &= ll_load(src + (size_t)q * p.d_pad * 2, ep, v[u]);
                        }
                        if (ok) break;
                        LL_BACKOFF();
                        if (spins > LL_SPIN_LIMIT) __trap();
                    }
                    s0 += v[0]; s1 += v[1]; s2 += v[2]; s3 += v[3];
                    s0 += v[4]; s1 += v[5]; s2 += v[6]; s3 += v[7];
                }
                if (p.fence) __threadfence();
            }
            BTRACE(7);
            double sacc = (s0 + s1) + (s2 + s3);
            for (int o = RC; o < 32; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane < RC) rsm[warp * 33 + lane] = sacc;
            __syncthreads();
            if (fin) {
                double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;   // warps w ≡ 0..3 (mod 4); rows ≥ Wn of rsm are zero
#pragma unroll
                for (int w = 0; w < 16; w += 4) {
                    t0 += rsm[w * 33 + tid]; t1 += rsm[(w + 1) * 33 + tid];
                    t2 += rsm[(w + 2) * 33 + tid]; t3 += rsm[(w + 3) * 33 + tid];
                }
                double tsum = (t0 + t1) + (t2 + t3);
                if (p.world > 1) {
                    // row shards: this rank's sum of the column goes to the same owner thread on every rank (one flagged word per
                    // peer over NVLink, one buffer per batch parity: a rank is at most one batch ahead of its slowest peer), and
                    // the sums of all ranks are added in rank order — the same bits on every rank
                    const uint32_t epx = p.xepoch0 + (uint32_t)b + 1u;
                    const size_t slot = ((size_t)(b & 1) * CIAO_MAX_PEERS + p.rank) * P2P_CAP + j_own;
                    for (int q = 0; q < p.world; ++q) ll_store_sys(p.xw[q] + slot * 2, tsum, epx);
                    const unsigned long long *mine = p.xw[p.rank] + ((size_t)(b & 1) * CIAO_MAX_PEERS * P2P_CAP + j_own) * 2;
                    double v[CIAO_MAX_PEERS];
                    unsigned long long t_start = 0;   // the ranks' hosts launch at different times: the limit is wall time, not polls
                    for (int spins = 0;; ++spins) {
                        bool ok = true;
#pragma unroll
                        for (int q = 0; q < CIAO_MAX_PEERS; ++q) {
                            v[q] = 0.0;
                            if (q < p.world) ok &= ll_load_sys(mine + (size_t)q * P2P_CAP * 2, epx, v[q]);
                        }
                        if (ok) break;
                        if ((spins & 1023) == 1023) {
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (!t_start) t_start = now;
                            else if (now - t_start > p.xtimeout_ns) __trap();
                        }
                    }
                    tsum = 0.0;
#pragma unroll
                    for (int q = 0; q < CIAO_MAX_PEERS; ++q) tsum += v[q];   // entries ≥ world are zero
                }
                double anew = __dadd_rn(own[tid], tsum);
                bool new_z = true;
                if (LF) {
                    anew = __dadd_rn(anew, __dmul_rn(bfs_b, __dsub_rn(own[32 + tid], own[64 + tid])));   // Finito_LFinito.jl:98
                    new_z = b + 1 < p.n_batches;                                                  // :92 of the next batch
                }
                own[tid] = anew;
                p.av[j_own] = anew;
                if (new_z) {
                    const double zn = prox_rt(p.reg.kind, anew, gl, own[96 + tid], own[128 + tid]);   // Finito_basic.jl:118
                    own[32 + tid] = zn;
                    p.z[j_own] = zn;
                    if (p.fence) __threadfence();
                    ll_store(p.llz + (size_t)j_own * 2, zn, ep);                                  // hop 2
                }
            }
        }
        BPROF_T(t_4);
        BPROF_ADD(2, t_3, t_4);
        BTRACE(3);
        if (b + 1 < p.n_batches) {
            // hop 2, receiving side: the CTA's threads fetch the new z once per SM into shared memory
            for (int64_t c0 = tid; c0 < p.d_pad; c0 += 4 * (int64_t)Tall) {   // up to four words of the thread in flight together
                double v[4];
                for (int spins = 0;; ++spins) {
                    bool ok = true;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t c = c0 + (int64_t)u * Tall;
                        if (c < p.d_pad) ok &= ll_load(p.llz + (size_t)c * 2, ep, v[u]);
                    }
                    if (ok) break;
                    LL_BACKOFF();
                    if (spins > LL_SPIN_LIMIT) __trap();
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t c = c0 + (int64_t)u * Tall;
                    if (c < p.d_pad) zsm[c] = v[u];
                }
            }
            if (p.fence) __threadfence();
            __syncthreads();   // every sub-group has left its row loop long ago (the comb tree): zsm is free to be rewritten
        }
        BPROF_T(t_5);
        BPROF_ADD(3, t_4, t_5);
        BTRACE(4);
    }
#ifdef CIAO_SEQ_PROFILE
    if (tid == 0 && bid < 1200) {
        unsigned int sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_batch_trace[bid * 8] = sm;
    }
    if (bid == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) g_batch_prof[i] = bprof[i];
#endif
}

#ifdef CIAO_SEQ_PROFILE
extern "C" int ciao_debug_batch_trace(ciao_ctx *c, unsigned long long *out9600) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out9600, g_batch_trace, 9600 * sizeof(unsigned long long)));
    void *sym = nullptr;
    CUDA_TRY(cudaGetSymbolAddress(&sym, g_batch_trace));
    CUDA_TRY(cudaMemset(sym, 0, 9600 * sizeof(unsigned long long)));
    return CIAO_OK;
}
extern "C" int ciao_debug_batch_dur(ciao_ctx *c, unsigned int *out4800) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out4800, g_batch_dur, 4800 * sizeof(unsigned int)));
    return CIAO_OK;
}
#endif

// One minibatch over the contiguous rows [row_lo, row_lo + n) (0-based, GLOBAL row numbers): pass over the rows this context
// holds + tail kernel (fixed-order reduction of the CTA partials, exchange over the ranks when the rows are sharded, closing
// update), on the context stream.  Row-sharded problems (SURVEY.md §8e last item, §8f rank 3): every rank streams the part of
// the batch it owns — rows and table rows live together — and the batch's Σ is all-reduced in the tail kernel
// (Finito_basic.jl:110-118 / Finito_LFinito.jl:91-100 with the batch split by row owner); `last` = last batch of the call.
int run_batch_step(ciao_ctx *c, int mode, int64_t row_lo, int64_t n, bool last = false) {
    NvtxRange nvtx("ciao:minibatch:step");
    const int64_t d_pad = c->d_pad;
    int cpt = 2;
    while (cpt < 16 && (d_pad + cpt - 1) / cpt > 256) cpt *= 2;
    const int64_t Tn = ((d_pad + cpt - 1) / cpt + 31) / 32 * 32;
    if (Tn > 256) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "minibatch pass: d = %lld exceeds 4096", (long long)c->d);
    const int T = (int)Tn, rpg = 16 / cpt;
    const size_t stage_bytes = (size_t)rpg * c->ld * sizeof(double);
    const size_t fixed = 2 * rpg * 32 * 2 * sizeof(double) + 16 * sizeof(uint64_t) + 256;
    int S = 3;  // 2 CTAs per SM: a batch is short, so parallelism across CTAs matters more than ring depth
    while (S > 1 && (size_t)S * stage_bytes + fixed > (size_t)112 * 1024) --S;
    const size_t smem = (size_t)S * stage_bytes + fixed;
    // the part of the batch this context holds
    int64_t lo_loc = 0, n_loc = 0;
    CIAO_TRY(local_window(c, row_lo, n, &lo_loc, &n_loc));
    const int64_t n_groups = (n_loc + rpg - 1) / rpg;
    const int grid = (int)std::min<int64_t>(n_groups, 2 * c->num_sms);
    const size_t need = ((size_t)std::max(grid, 1) * d_pad + std::max(grid, 1) + 16) * sizeof(double);
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    BatchArgs a;
    a.rec = c->rec + lo_loc * c->ld; a.n_rows = n_loc; a.ld = c->ld; a.d_pad = d_pad;
    a.z = ctx_vec(c, CIAO_VEC_Z); a.zf = ctx_vec(c, CIAO_VEC_Z_FULL);
    a.table = c->table ? c->table + lo_loc * d_pad : nullptr;
    a.ws = c->ws; a.fws = c->ws + (size_t)std::max(grid, 1) * d_pad;
    a.cN = c->hat_gamma / (double)c->N_total; a.stages = S;
    int rc = CIAO_OK;
    if (grid > 0) {
        if (mode == BATCH_FINITO) {
            switch (cpt) {
                case 2: rc = launch_batch_loss<2, BATCH_FINITO>(c, a, grid, T, smem); break;
                case 4: rc = launch_batch_loss<4, BATCH_FINITO>(c, a, grid, T, smem); break;
                case 8: rc = launch_batch_loss<8, BATCH_FINITO>(c, a, grid, T, smem); break;
                default: rc = launch_batch_loss<16, BATCH_FINITO>(c, a, grid, T, smem); break;
            }
        } else {
            switch (cpt) {
                case 2: rc = launch_batch_loss<2, BATCH_LFINITO>(c, a, grid, T, smem); break;
                case 4: rc = launch_batch_loss<4, BATCH_LFINITO>(c, a, grid, T, smem); break;
                case 8: rc = launch_batch_loss<8, BATCH_LFINITO>(c, a, grid, T, smem); break;
                default: rc = launch_batch_loss<16, BATCH_LFINITO>(c, a, grid, T, smem); break;
            }
        }
        CIAO_TRY(rc);
        c->timing.launches += 1;
    }
    const bool sharded = c->world > 1 && c->n_rows != c->N_total;
    if (sharded && !c->p2p_ready)
        CIAO_FAIL(CIAO_ERR_STATE, "minibatch steps on row shards need the peer exchange (ciao_comm_p2p_handle / ciao_comm_p2p_attach)");
    TailArgs t;
    memset(&t, 0, sizeof(t));
    t.ws = a.ws; t.fws = a.fws; t.G = grid; t.d_pad = d_pad; t.len = (int)d_pad + 1;
    t.fmax_mode = 0; t.with_vec = 1; t.chunk0 = 0;
    t.sum_out = c->partial;
    t.fin_mode = mode == BATCH_FINITO ? FIN_FINITO : FIN_LFINITO;
    t.last_batch = last ? 1 : 0;
    t.av = ctx_vec(c, CIAO_VEC_AV); t.z = ctx_vec(c, CIAO_VEC_Z); t.zf = ctx_vec(c, CIAO_VEC_Z_FULL);
    t.hat_gamma = c->hat_gamma; t.reg = c->reg;
    fill_exchange(c, t, sharded);
    pass_tail_kernel<<<(int)((d_pad + 1 + 31) / 32), dim3(32, REDUCE_SLICES), 0, c->stream>>>(t);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

// All batches in one cooperative launch.  rows/lens: device arrays of n_batches batch windows.  Returns CIAO_ERR_UNSUPPORTED
// when a cooperative grid of 2 CTAs/SM is not available (the caller then falls back to one pass per batch).
template <int CPT, int MODE, int LOSS>
static int launch_batch_persistent(ciao_ctx *c, BatchPArgs &a, int T, size_t smem, int a_max_ctas) {
    auto kern = batch_persistent_kernel<CPT, MODE, LOSS>;
    static size_t configured[CIAO_MAX_DEVICES] = {};
    if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device % CIAO_MAX_DEVICES] = smem;
    }
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "persistent minibatch kernel does not fit on an SM");
    const int grid = std::min(occ, a_max_ctas) * c->num_sms;
    const size_t need = ((size_t)grid * a.d_pad + grid + 16) * sizeof(double);
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    a.ws = c->ws; a.fws = c->ws + (size_t)grid * a.d_pad;
    int rc_cols = 4;
    while (rc_cols < 32 && (int64_t)rc_cols * grid < a.d_pad) rc_cols *= 2;
    if (rc_cols > T) rc_cols = T;
    if ((int64_t)rc_cols * grid < a.d_pad) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "persistent minibatch kernel: d too large for the distributed reduction");
    a.red_cols = rc_cols;
    void *args[] = {(void *)&a};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(T), args, smem, c->stream));
    return CIAO_OK;
}

template <int CPT, int MODE>
static int launch_batch_persistent_loss(ciao_ctx *c, BatchPArgs &a, int T, size_t smem, int max_ctas) {
    return c->loss_kind == CIAO_LOSS_LS ? launch_batch_persistent<CPT, MODE, CIAO_LOSS_LS>(c, a, T, smem, max_ctas)
                                        : launch_batch_persistent<CPT, MODE, CIAO_LOSS_LOGISTIC>(c, a, T, smem, max_ctas);
}

// the one-CTA-per-SM version: sub-groups instead of CTAs per SM; the exchange buffers live in the context (zeroed when (re)allocated)
template <int CPT, int MODE, int LOSS>
static int launch_batch_sm(ciao_ctx *c, BatchLArgs &a, int TS, int nsg, size_t smem) {
    auto kern = batch_sm_kernel<CPT, MODE, LOSS>;
    static size_t configured[CIAO_MAX_DEVICES] = {};
    if (smem > configured[c->device % CIAO_MAX_DEVICES]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device % CIAO_MAX_DEVICES] = smem;
    }
    const int Tall = TS * nsg;
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Tall, smem));
    if (occ < 1) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "persistent minibatch kernel does not fit on an SM");
    const int grid = c->num_sms;
    int rc_cols = 4;
    while (rc_cols < 32 && (int64_t)rc_cols * grid < a.d_pad) rc_cols *= 2;
    if (rc_cols > Tall) rc_cols = Tall;
    if ((int64_t)rc_cols * grid < a.d_pad) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "persistent minibatch kernel: d too large for the distributed reduction");
    a.red_cols = rc_cols;
    a.ts = TS; a.nsg = nsg;
    const size_t words = 2 * ((size_t)a.d_pad + (size_t)grid * a.d_pad);
    const bool fresh = !c->ll_buf || c->ll_grid != grid || c->ll_d != a.d_pad;
    if (fresh && words * 8 > c->ll_bytes) {
        if (c->ll_buf) cudaFree(c->ll_buf);
        c->ll_buf = nullptr;
        c->ll_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ll_buf, words * 8));
        c->ll_bytes = words * 8;
    }
    if (fresh || (uint64_t)c->ll_epoch + (uint64_t)a.n_batches + 2 > 0xfffffff0ull) {   // new layout, or the 32-bit epoch would wrap
        CUDA_TRY(cudaMemsetAsync(c->ll_buf, 0, c->ll_bytes, c->stream));
        c->ll_grid = grid; c->ll_d = a.d_pad; c->ll_epoch = 0;
    }
    a.llz = c->ll_buf;
    a.llws = a.llz + 2 * (size_t)a.d_pad;
    if (MODE != BATCH_FINITO) {   // Σ γ̂/γ_i per batch, in the (otherwise unused) CTA-partial workspace
        const size_t need = (size_t)a.n_batches * sizeof(double);
        if (need > c->ws_bytes) {
            if (c->ws) cudaFree(c->ws);
            c->ws = nullptr;
            c->ws_bytes = 0;
            CUDA_TRY(cudaMalloc(&c->ws, need));
            c->ws_bytes = need;
        }
        batch_fsum_kernel<<<(unsigned)a.n_batches, 256, 0, c->stream>>>(a.rec, a.ld, a.d_pad, a.b_lo, a.b_n, c->ws);
        CUDA_TRY(cudaGetLastError());
        c->timing.launches += 1;
        if (a.world > 1)   // the batch's rows are spread over the ranks: rank-ordered sums of the per-rank parts
            for (int64_t o = 0; o < a.n_batches; o += P2P_CAP)
                CIAO_TRY(run_p2p_allreduce(c, c->ws + o, std::min<int64_t>(P2P_CAP, a.n_batches - o), 0));
        a.bfs = c->ws;
    }
    a.epoch0 = c->ll_epoch;
    c->ll_epoch += (uint32_t)a.n_batches;
    if (a.world > 1) {   // collective: all ranks launch the same sequence of batches
        for (int q = 0; q < a.world; ++q)
            a.xw[q] = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(c->p2p_peer[q]) + P2P_LL_OFFSET);
        a.xepoch0 = c->p2p_ll_epoch;
        a.xtimeout_ns = c->p2p_timeout_ns;
        c->p2p_ll_epoch += (uint32_t)a.n_batches;
    }
    void *args[] = {(void *)&a};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(Tall), args, smem, c->stream));
    return CIAO_OK;
}
// shape of the launch: sub-groups, ring depth; then the instantiation
template <int CPT, int MODE>
static int launch_batch_sm_shape(ciao_ctx *c, BatchLArgs &a, int TS, int max_sub, int S_want) {
    constexpr int RPG = 16 / CPT;
    int nsg = 1;
    while (nsg * 2 * TS <= BatchSmShape<CPT, MODE>::MAXT && nsg * 2 <= std::min(max_sub, 4)) nsg *= 2;   // 1, 2 or 4 (the combine step)
    const size_t stage_rows = (size_t)RPG * a.ld * sizeof(double) + (MODE == BATCH_LFINITO_CZ ? (size_t)RPG * 32 : 0);
    const size_t shared_part = ((size_t)((nsg > 1 ? nsg : 0) + 1 + (MODE == BATCH_LFINITO ? 1 : 0)) * a.d_pad + 17 * 33 + 1 + 5 * 32) * sizeof(double) + (size_t)nsg * 4 * sizeof(uint64_t) + 256;
    const size_t limit = 227 * 1024;   // the opt-in maximum of dynamic shared memory per CTA
    auto smem_for = [&](int S) { return (size_t)nsg * ((size_t)S * stage_rows + 2 * RPG * 32 * 2 * sizeof(double)) + shared_part; };
    int S = S_want;
    while (S > 1 && smem_for(S) > limit) --S;
    a.stages = S;
    const size_t smem = smem_for(S);
    return c->loss_kind == CIAO_LOSS_LS ? launch_batch_sm<CPT, MODE, CIAO_LOSS_LS>(c, a, TS, nsg, smem)
                                        : launch_batch_sm<CPT, MODE, CIAO_LOSS_LOGISTIC>(c, a, TS, nsg, smem);
}

// b_lo_dev / b_n_dev: device arrays (n_batches) of batch windows; z (and z_full for LFinito) as the first batch needs them
// windows: BATCH_WINDOWS_DISJOINT (no row twice in the call), _ALIGNED (windows repeat, always as the same window: a row stays with
// its SM and thread) or _ANY (a row may move between SMs: the exchange is bracketed by fences)
// Row shards: b_lo / b_n are the parts of the windows this context holds (local row numbers, possibly empty), batch_rows the
// length of the whole batch, and the call is collective (every rank launches the same batches; needs the peer exchange).
int run_batch_sequence(ciao_ctx *c, int mode, const int64_t *b_lo_dev, const int64_t *b_n_dev, int64_t n_batches, int64_t batch_rows,
                       int windows, bool sharded = false) {
    NvtxRange nvtx("ciao:minibatch:persistent");
    if (n_batches <= 0) return CIAO_OK;
    const int64_t d_pad = c->d_pad;
    // scripts/batch_probe.py at C2: 128 threads (8 columns each) beat 256 for batches of 4096 rows (170 vs 159 Finito epochs/s,
    // 219 vs 191 LFinito sweeps/s) and lose for batches of 512 (52 vs 58)
    bool barrier_version = false;
    if (const char *xv = getenv("CIAO_BATCH_EXCHANGE")) barrier_version = !strcmp(xv, "barrier");
    if (sharded) {
        if (!c->p2p_ready || c->world < 2) return CIAO_ERR_UNSUPPORTED;   // the caller falls back to one pass + tail kernel per batch
        barrier_version = false;
    }
    int T_target = (batch_rows >= 2048 && d_pad <= 2048) ? 128 : 256;
    // batch_sm_kernel (profiles/batch_sweep_sm_r2.log): Finito is as fast or faster with one sub-group of 256 threads and four rows per
    // item (18.4 vs 19.0 µs per 4096-row batch, 241.8 vs 242.6 µs at 65 536); LFinito wants four sub-groups of 128 (9.4 vs 10.2, 86.5 vs 100.8)
    if (!barrier_version && mode == BATCH_FINITO) T_target = 256;
    if (const char *tv = getenv("CIAO_BATCH_T")) T_target = std::max(32, std::min(256, atoi(tv)));
    int cpt = 2;
    while (cpt < 16 && (d_pad + cpt - 1) / cpt > T_target) cpt *= 2;
    const int64_t Tn = ((d_pad + cpt - 1) / cpt + 31) / 32 * 32;
    if (Tn > 256) CIAO_FAIL(CIAO_ERR_UNSUPPORTED, "minibatch pass: d = %lld exceeds 4096", (long long)c->d);
    const int T = (int)std::max<int64_t>(Tn, 32), rpg = 16 / cpt;
    const size_t stage_bytes = (size_t)rpg * c->ld * sizeof(double);
    const size_t fixed = (2 * rpg * 32 * 2 + 9 * 33 + 1 + 5 * 32) * sizeof(double) + 16 * sizeof(uint64_t) + 256;
    // a batch with no more row groups than SMs cannot use a second CTA per SM; it only makes the grid barriers and the
    // reduction over the CTA partials longer (measured at C2, batch 512: 58 sweeps/s with 148 CTAs, 48 with 296)
    // LFinito's rows phase is latency-bound per CTA (two dots, two coefficients per row; 127 registers): four CTAs per SM lift a
    // 65 536-row batch from 408 to 542 sweeps/s at C2 (profiles/batch_shape_sweep_r2.log); Finito's is HBM-bound already at two
    int max_ctas = (batch_rows + rpg - 1) / rpg <= (int64_t)c->num_sms ? 1 : (mode == BATCH_LFINITO ? 4 : 2);
    if (const char *cv = getenv("CIAO_BATCH_CTAS")) max_ctas = std::max(1, std::min(8, atoi(cv)));
    int S = 3;
    if (const char *sv = getenv("CIAO_BATCH_STAGES")) S = std::max(1, std::min(3, atoi(sv)));
    const int S_want = S;
    while (S > 1 && (size_t)S * stage_bytes + fixed > (size_t)(220 * 1024) / max_ctas) --S;
    const size_t smem = (size_t)S * stage_bytes + fixed;
    if (!barrier_version) {
        // LFinito's row loop is bound by fp64 issue, not by the ring: two stages measured faster than three (12.9 vs 13.5 µs per batch)
        const int S_ll = (mode == BATCH_LFINITO && !getenv("CIAO_BATCH_STAGES")) ? 2 : S_want;
        BatchLArgs a;
        memset(&a, 0, sizeof(a));
        a.rec = c->rec; a.ld = c->ld; a.d_pad = d_pad; a.table = c->table; a.b_lo = b_lo_dev; a.b_n = b_n_dev; a.n_batches = n_batches;
        a.z = ctx_vec(c, CIAO_VEC_Z); a.av = ctx_vec(c, CIAO_VEC_AV); a.zf = ctx_vec(c, CIAO_VEC_Z_FULL);
        a.cN = c->hat_gamma / (double)c->N_total; a.hat_gamma = c->hat_gamma; a.reg = c->reg;
        a.fence = (mode == BATCH_FINITO && windows == BATCH_WINDOWS_ANY) ? 1 : 0;
        a.world = sharded ? c->world : 1;
        a.rank = c->rank;
        int rc;
        if (mode == BATCH_FINITO) {
            switch (cpt) {
                case 2: rc = launch_batch_sm_shape<2, BATCH_FINITO>(c, a, T, max_ctas, S_want); break;
                case 4: rc = launch_batch_sm_shape<4, BATCH_FINITO>(c, a, T, max_ctas, S_want); break;
                case 8: rc = launch_batch_sm_shape<8, BATCH_FINITO>(c, a, T, max_ctas, S_want); break;
                default: rc = launch_batch_sm_shape<16, BATCH_FINITO>(c, a, T, max_ctas, S_want); break;
            }
        } else if (c->cache_cz && c->cz_local_valid && c->ss && !getenv("CIAO_BATCH_TWO_DOTS")) {
            // the pass at z_full that opened this sweep (ciao_lfinito_outer) left c_i(z_full) of the local rows: one dot per row
            a.ss = c->ss + 4 * c->ss_row0;
            switch (cpt) {
                case 2: rc = launch_batch_sm_shape<2, BATCH_LFINITO_CZ>(c, a, T, max_ctas, S_ll); break;
                case 4: rc = launch_batch_sm_shape<4, BATCH_LFINITO_CZ>(c, a, T, max_ctas, S_ll); break;
                case 8: rc = launch_batch_sm_shape<8, BATCH_LFINITO_CZ>(c, a, T, max_ctas, S_ll); break;
                default: rc = launch_batch_sm_shape<16, BATCH_LFINITO_CZ>(c, a, T, max_ctas, S_ll); break;
            }
        } else {
            switch (cpt) {
                case 2: rc = launch_batch_sm_shape<2, BATCH_LFINITO>(c, a, T, max_ctas, S_ll); break;
                case 4: rc = launch_batch_sm_shape<4, BATCH_LFINITO>(c, a, T, max_ctas, S_ll); break;
                case 8: rc = launch_batch_sm_shape<8, BATCH_LFINITO>(c, a, T, max_ctas, S_ll); break;
                default: rc = launch_batch_sm_shape<16, BATCH_LFINITO>(c, a, T, max_ctas, S_ll); break;
            }
        }
        CIAO_TRY(rc);
        c->timing.launches += 1;
        return CIAO_OK;
    }
    if (!c->grid_bar) CUDA_TRY(cudaMalloc(&c->grid_bar, 64));
    CUDA_TRY(cudaMemsetAsync(c->grid_bar, 0, 64, c->stream));
    BatchPArgs a;
    a.rec = c->rec; a.ld = c->ld; a.d_pad = d_pad; a.table = c->table; a.b_lo = b_lo_dev; a.b_n = b_n_dev; a.n_batches = n_batches;
    a.z = ctx_vec(c, CIAO_VEC_Z); a.av = ctx_vec(c, CIAO_VEC_AV); a.zf = ctx_vec(c, CIAO_VEC_Z_FULL);
    a.bar = c->grid_bar; a.cN = c->hat_gamma / (double)c->N_total; a.hat_gamma = c->hat_gamma; a.reg = c->reg; a.stages = S;
    a.ws = nullptr; a.fws = nullptr;
    // measured at C2, batch 4096: 25.6 µs per batch with the prefetch against 22.7 without — the rows phase is bound by the
    // per-CTA latency of a row group, not by HBM, and the extra requests delay the barrier; off unless CIAO_BATCH_L2PF=1
    a.l2_prefetch = 0;
    if (const char *pf = getenv("CIAO_BATCH_L2PF")) a.l2_prefetch = atoi(pf) != 0;
    int rc;
    if (mode == BATCH_FINITO) {
        switch (cpt) {
            case 2: rc = launch_batch_persistent_loss<2, BATCH_FINITO>(c, a, T, smem, max_ctas); break;
            case 4: rc = launch_batch_persistent_loss<4, BATCH_FINITO>(c, a, T, smem, max_ctas); break;
            case 8: rc = launch_batch_persistent_loss<8, BATCH_FINITO>(c, a, T, smem, max_ctas); break;
            default: rc = launch_batch_persistent_loss<16, BATCH_FINITO>(c, a, T, smem, max_ctas); break;
        }
    } else {
        switch (cpt) {
            case 2: rc = launch_batch_persistent_loss<2, BATCH_LFINITO>(c, a, T, smem, max_ctas); break;
            case 4: rc = launch_batch_persistent_loss<4, BATCH_LFINITO>(c, a, T, smem, max_ctas); break;
            case 8: rc = launch_batch_persistent_loss<8, BATCH_LFINITO>(c, a, T, smem, max_ctas); break;
            default: rc = launch_batch_persistent_loss<16, BATCH_LFINITO>(c, a, T, smem, max_ctas); break;
        }
    }
    CIAO_TRY(rc);
    c->timing.launches += 1;
    return CIAO_OK;
}
