// proshi.cu — K7: ProShI block updates for the sharing problem
//   (1/N) Σ f_i(x_i) + g(Σ x_i),   f_i = ½x'diag(q_i)x + c_i'x + (η/2)dist²(x,[lo,hi])   (test_sharing.jl:15-22)
//
//   proshi_init_kernel      s_i = x0 − (γ_i/N)∇f_i(x0);  Σ s_i           ProShI_basic.jl:76-83
//   proshi_steps_kernel     block steps + dual update                     ProShI_basic.jl:111-123
//   proshi_solution_kernel  s_i += γ_i z  IN PLACE                        ProShI_basic.jl:127-132
//
// Design.  With a diagonal Q_i the block update has no coupling
// between coordinates: every column j of (s, av, z) evolves independently given the
// index sequence.  proshi_steps_kernel therefore gives each thread two columns, keeps
// z_j and av_j in registers for the whole call and walks the host-generated index
// sequence with a P-deep register prefetch of (γ_i, q_i, c_i, s_i) — no barrier, no
// reduction, no kernel launch per step; the n columns are spread over many SMs so
// the gathers of different columns overlap.
#include <algorithm>

#include "common.cuh"

struct ProshiArgs {
    const double *qd, *ql;  // [N][n_pad]
    double *table;          // [N][n_pad]
    const double *gam;      // [N]
    const double *gam_n;    // [N]  γ_i/N, precomputed (ProShI_basic.jl:116)
    const int64_t *idx;     // prepared
    int64_t K, N, n_pad;
    double *v_z, *v_av;
    double box_lo, box_hi, eta, Nd, hat_gamma;
    RegParams reg;
};

__device__ __forceinline__ double proshi_grad(double q, double c, double s, double lo, double hi, double eta) {
    // Sum(Quadratic, SqrDistL2): (q·s + c) + η·(s − Π_box s)
    const double g1 = __dadd_rn(__dmul_rn(q, s), c);
    const double pr = s < lo ? lo : (s > hi ? hi : s);
    const double g2 = __dmul_rn(eta, __dsub_rn(s, pr));
    return __dadd_rn(g1, g2);
}


// Batch-1 steps: a true dependency chain per column (z_j → s_ij → av_j → z_j), so the kernel is latency bound and
// everything that is not on that chain has to stay off it.  Each thread stages the 16-byte slices of (q_i, c_i, s_i)
// and the scalars (γ_i, γ_i/N) of the block it will need PROSHI_D steps later into ITS OWN shared-memory cells with
// cp.async, one commit group per step; `cp.async.wait_group D−1` then guarantees exactly the oldest group — a counted
// in-order pipeline.  (The first version prefetched into registers with ld.global: ptxas tracks all those loads with one
// scoreboard, so the first use of step k's registers also waited for the load just issued for step k+D — every step paid
// a full DRAM latency: 0.74 µs/block.)  No barrier, no mbarrier, no reduction, no kernel launch per step; a warp reads the
// index sequence 32 entries at a time (one coalesced load per 32 steps, prefetched a block ahead) and broadcasts by shuffle.
constexpr int PROSHI_D = 16;  // steps of prefetch; a staged table slice is stale if the block recurs within D steps → HAZARD flag
static_assert(PROSHI_D < 32 && (PROSHI_D & (PROSHI_D - 1)) == 0, "ring depth: power of two below the index block of 32");
static_assert(PROSHI_D <= CIAO_HAZARD_WINDOW - 1, "prep_indices_kernel must flag repeats within the prefetch window");
constexpr int PROSHI_SLOT_BYTES = 4 * 32 * 16;  // per warp and slot: q | c | s | (γ, γ/N), 16 bytes per lane each

__device__ __forceinline__ void cp_async_cg16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_ca8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(64) proshi_steps_kernel(const ProshiArgs p) {
    constexpr int D = PROSHI_D;
    extern __shared__ __align__(16) unsigned char proshi_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t col_raw = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
    const bool active = col_raw < p.n_pad;       // inactive lanes of the last warp shadow column 0 and never store
    const int64_t col = active ? col_raw : 0;
    double z0 = p.v_z[col], z1 = p.v_z[col + 1], av0 = p.v_av[col], av1 = p.v_av[col + 1];
    const double lo0 = p.reg.lo_v ? p.reg.lo_v[col] : p.reg.lo_s, lo1 = p.reg.lo_v ? p.reg.lo_v[col + 1] : p.reg.lo_s;
    const double hi0 = p.reg.hi_v ? p.reg.hi_v[col] : p.reg.hi_s, hi1 = p.reg.hi_v ? p.reg.hi_v[col + 1] : p.reg.hi_s;
    const double gl = p.hat_gamma * p.reg.lambda;
    const double rhat = __ddiv_rn(1.0, p.hat_gamma);
    const int64_t K = p.K;
    const uint32_t cell0 = smem_u32(proshi_smem) + (uint32_t)warp * (D * PROSHI_SLOT_BYTES) + (uint32_t)lane * 16;
    const unsigned char *cellp = proshi_smem + (size_t)warp * (D * PROSHI_SLOT_BYTES) + (size_t)lane * 16;

    auto issue = [&](int slot, int64_t pidx) {
        const int64_t i = pidx & CIAO_IDX_MASK;
        const uint32_t cb = cell0 + (uint32_t)slot * PROSHI_SLOT_BYTES;
        cp_async_cg16(cb, p.qd + i * p.n_pad + col);
        cp_async_cg16(cb + 512, p.ql + i * p.n_pad + col);
        cp_async_cg16(cb + 1024, p.table + i * p.n_pad + col);
        cp_async_ca8(cb + 1536, p.gam + i);
        cp_async_ca8(cb + 1544, p.gam_n + i);
    };
    struct Blk {
        double2 q, c, s;
        double gi, gn;
        int64_t ik;
    };
    auto take = [&](int slot, int64_t pidx, Blk &b) {
        const unsigned char *cb = cellp + (size_t)slot * PROSHI_SLOT_BYTES;
        b.q = *reinterpret_cast<const double2 *>(cb);
        b.c = *reinterpret_cast<const double2 *>(cb + 512);
        b.s = *reinterpret_cast<const double2 *>(cb + 1024);
        const double2 g = *reinterpret_cast<const double2 *>(cb + 1536);
        b.gi = g.x;
        b.gn = g.y;
        b.ik = pidx;
    };

    // index blocks of 32 steps: iv_i covers the step being staged (k + D), iv_prev the block before it, iv_n the next one
    int64_t iv_i = (lane < K) ? __ldg(p.idx + lane) : 0;
    int64_t iv_n = (32 + lane < K) ? __ldg(p.idx + 32 + lane) : 0;
    int64_t iv_prev = iv_i;
#pragma unroll
    for (int s = 0; s < D; ++s) {
        const int64_t pidx = __shfl_sync(0xffffffffu, iv_i, s);
        if (s < K) issue(s, pidx);
        cp_async_commit();
    }
    Blk cur;
    cp_async_wait<D - 1>();
    take(0, __shfl_sync(0xffffffffu, iv_i, 0), cur);

    for (int64_t k = 0; k < K; ++k) {
        double2 *srow = reinterpret_cast<double2 *>(p.table + (cur.ik & CIAO_IDX_MASK) * p.n_pad + col);
        double2 s = cur.s;
        if (cur.ik & CIAO_FLAG_HAZARD) s = __ldcg(srow);  // rewritten after its copy was issued: re-read behind our own store
        const double gi = cur.gi;
        const double cneg = -cur.gn;
        // ProShI_basic.jl:113-119
        av0 = __dsub_rn(av0, s.x);
        av1 = __dsub_rn(av1, s.y);
        const double x0 = __dadd_rn(s.x, __dmul_rn(gi, z0)), x1 = __dadd_rn(s.y, __dmul_rn(gi, z1));
        double t0 = __dmul_rn(proshi_grad(cur.q.x, cur.c.x, x0, p.box_lo, p.box_hi, p.eta), cneg);
        double t1 = __dmul_rn(proshi_grad(cur.q.y, cur.c.y, x1, p.box_lo, p.box_hi, p.eta), cneg);
        t0 = __dadd_rn(t0, x0);
        t1 = __dadd_rn(t1, x1);
        av0 = __dadd_rn(av0, t0);
        av1 = __dadd_rn(av1, t1);
        if (active) __stcg(srow, make_double2(t0, t1));
        if (cur.ik & CIAO_FLAG_PROX) {  // :121-123
            z0 = div_by(__dsub_rn(prox_rt(p.reg.kind, av0, gl, lo0, hi0), av0), p.hat_gamma, rhat);
            z1 = div_by(__dsub_rn(prox_rt(p.reg.kind, av1, gl, lo1, hi1), av1), p.hat_gamma, rhat);
        }
        // stage step k + D into the slot just consumed (issued after this step's store: only repeats inside the window are stale)
        const int64_t sk = k + D;
        if ((sk & 31) == 0) {
            iv_prev = iv_i;
            iv_i = iv_n;
            iv_n = (sk + 32 + lane < K) ? __ldg(p.idx + sk + 32 + lane) : 0;
        }
        const int64_t pidx_s = __shfl_sync(0xffffffffu, iv_i, (int)(sk & 31));
        if (sk < K) issue((int)(k & (D - 1)), pidx_s);
        cp_async_commit();
        // registers of step k + 1
        const int64_t k1 = k + 1;
        const int l1 = (int)(k1 & 31);
        // iv_i is the index block of step k + D = k1 + D − 1; step k1 lies in the same block iff l1 + D − 1 < 32
        const int64_t pidx_1 = __shfl_sync(0xffffffffu, (l1 + D - 1 < 32) ? iv_i : iv_prev, l1);
        cp_async_wait<D - 1>();
        if (k1 < K) take((int)(k1 & (D - 1)), pidx_1, cur);
    }
    cp_async_wait<0>();
    if (active) {
        p.v_z[col] = z0; p.v_z[col + 1] = z1;
        p.v_av[col] = av0; p.v_av[col + 1] = av1;
    }
}

// ---------------------------------------------------------------------------
// Minibatch path (batch ≥ PROSHI_BATCH_MIN blocks): inside a batch every block uses the same z
// (ProShI_basic.jl:111-120), so the blocks are independent and av only needs Σ_i (t_i − s_i).  A CTA owns
// 8 columns (one 64-byte chunk of every row) for the WHOLE call, its threads work through the blocks of
// a batch in parallel, a fixed-order CTA reduction closes the batch and the dual update of the CTA's own
// columns follows — columns never interact, so there is no grid-wide synchronisation and the kernel
// streams the three block arrays at HBM rate.  Bitwise reproducible (fixed thread ↔ block mapping).
constexpr int PROSHI_BATCH_MIN = 64;
constexpr int PROSHI_BT = 256;  // threads per CTA

__global__ void __launch_bounds__(PROSHI_BT) proshi_batch_kernel(const ProshiArgs p, const int64_t *ptr, int64_t n_batches) {
    __shared__ double red[2][PROSHI_BT / 32][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t col = 8 * (int64_t)blockIdx.x;
    bool cv[4];
    double z[8], av[8], lo[8], hi[8];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        cv[h] = col + 2 * h < p.n_pad;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int q = 2 * h + e;
            const int64_t g = col + q;
            z[q] = cv[h] ? p.v_z[g] : 0.0;
            av[q] = cv[h] ? p.v_av[g] : 0.0;
            lo[q] = (cv[h] && p.reg.lo_v) ? p.reg.lo_v[g] : p.reg.lo_s;
            hi[q] = (cv[h] && p.reg.hi_v) ? p.reg.hi_v[g] : p.reg.hi_s;
        }
    }
    const double gl = p.hat_gamma * p.reg.lambda;
    const double rhat = __ddiv_rn(1.0, p.hat_gamma);

    for (int64_t b = 0; b < n_batches; ++b) {
        const int64_t lo_t = ptr[b], hi_t = ptr[b + 1];
        double ds[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) ds[q] = 0.0;
        // two blocks per iteration: all twelve 16-byte loads of both blocks are in flight before the first use
        for (int64_t t = lo_t + tid; t < hi_t; t += 2 * PROSHI_BT) {
            const bool second = t + PROSHI_BT < hi_t;
            int64_t off[2];
            double gi[2], cneg[2];
            double2 q2[2][4], c2[2][4], s2[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !second) continue;
                const int64_t i = __ldg(p.idx + t + u * PROSHI_BT) & CIAO_IDX_MASK;
                gi[u] = __ldg(p.gam + i);
                cneg[u] = -__ldg(p.gam_n + i);
                off[u] = i * p.n_pad + col;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (!cv[h]) continue;
                    q2[u][h] = __ldcs(reinterpret_cast<const double2 *>(p.qd + off[u]) + h);
                    c2[u][h] = __ldcs(reinterpret_cast<const double2 *>(p.ql + off[u]) + h);
                    s2[u][h] = __ldcg(reinterpret_cast<const double2 *>(p.table + off[u]) + h);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !second) continue;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (!cv[h]) continue;
                    // ProShI_basic.jl:114-119 for two columns
                    const double x0 = __dadd_rn(s2[u][h].x, __dmul_rn(gi[u], z[2 * h]));
                    const double x1 = __dadd_rn(s2[u][h].y, __dmul_rn(gi[u], z[2 * h + 1]));
                    const double t0 = __dadd_rn(__dmul_rn(proshi_grad(q2[u][h].x, c2[u][h].x, x0, p.box_lo, p.box_hi, p.eta), cneg[u]), x0);
                    const double t1 = __dadd_rn(__dmul_rn(proshi_grad(q2[u][h].y, c2[u][h].y, x1, p.box_lo, p.box_hi, p.eta), cneg[u]), x1);
                    ds[2 * h] += __dsub_rn(t0, s2[u][h].x);
                    ds[2 * h + 1] += __dsub_rn(t1, s2[u][h].y);
                    __stcg(reinterpret_cast<double2 *>(p.table + off[u]) + h, make_double2(t0, t1));
                }
            }
        }
        // close the batch: fixed-order CTA reduction of Σ(t − s), then av and the dual variable z (:121-123)
        const int par = (int)(b & 1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            ds[q] = warp_sum(ds[q]);
            if (lane == 0) red[par][warp][q] = ds[q];
        }
        __syncthreads();  // also orders this batch's table writes before the next batch's reads (same CTA, same columns)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < PROSHI_BT / 32; ++w) s += red[par][w][q];
            av[q] = __dadd_rn(av[q], s);
            z[q] = div_by(__dsub_rn(prox_rt(p.reg.kind, av[q], gl, lo[q], hi[q]), av[q]), p.hat_gamma, rhat);
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (!cv[h]) continue;
            p.v_z[col + 2 * h] = z[2 * h];
            p.v_z[col + 2 * h + 1] = z[2 * h + 1];
            p.v_av[col + 2 * h] = av[2 * h];
            p.v_av[col + 2 * h + 1] = av[2 * h + 1];
        }
    }
}

// grid (row groups, column chunks of 512); ws[blockIdx.x][n_pad] = partial Σ s_i
__global__ void __launch_bounds__(256) proshi_init_kernel(const double *qd, const double *ql, const double *gam,
                                                          double *gam_n, const double *x0, double *table, double *ws, int64_t N,
                                                          int64_t n_pad, double box_lo, double box_hi, double eta,
                                                          double Nd) {
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    const double xa = x0[col], xb = x0[col + 1];
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        const double2 q = __ldcs(reinterpret_cast<const double2 *>(qd + i * n_pad + col));
        const double2 c = __ldcs(reinterpret_cast<const double2 *>(ql + i * n_pad + col));
        const double cg = __ddiv_rn(__ldg(gam + i), Nd);
        if (col == 0) gam_n[i] = cg;
        const double s0 = __dsub_rn(xa, __dmul_rn(cg, proshi_grad(q.x, c.x, xa, box_lo, box_hi, eta)));
        const double s1 = __dsub_rn(xb, __dmul_rn(cg, proshi_grad(q.y, c.y, xb, box_lo, box_hi, eta)));
        __stcs(reinterpret_cast<double2 *>(table + i * n_pad + col), make_double2(s0, s1));
        a0 += s0;
        a1 += s1;
    }
    ws[(size_t)blockIdx.x * n_pad + col] = a0;
    ws[(size_t)blockIdx.x * n_pad + col + 1] = a1;
}

// ws[blockIdx.x][n_pad] = partial Σ_i table_i   (sum(x_proshi), test_sharing.jl:42)
__global__ void __launch_bounds__(256) table_colsum_kernel(const double *table, double *ws, int64_t N, int64_t n_pad) {
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        const double2 s = __ldcs(reinterpret_cast<const double2 *>(table + i * n_pad + col));
        a0 += s.x;
        a1 += s.y;
    }
    ws[(size_t)blockIdx.x * n_pad + col] = a0;
    ws[(size_t)blockIdx.x * n_pad + col + 1] = a1;
}

__global__ void __launch_bounds__(256) proshi_solution_kernel(double *table, const double *gam, const double *z, int64_t N,
                                                              int64_t n_pad) {
    const int64_t col = 2 * (blockIdx.y * (int64_t)blockDim.x + threadIdx.x);
    if (col >= n_pad) return;
    const double z0 = z[col], z1 = z[col + 1];
    for (int64_t i = blockIdx.x; i < N; i += gridDim.x) {
        double2 *sp = reinterpret_cast<double2 *>(table + i * n_pad + col);
        double2 s = *sp;
        const double gi = __ldg(gam + i);
        s.x = __dadd_rn(s.x, __dmul_rn(gi, z0));
        s.y = __dadd_rn(s.y, __dmul_rn(gi, z1));
        *sp = s;
    }
}

// z = (prox_g(av, γ̂) − av)/γ̂     ProShI_basic.jl:84-86
__global__ void proshi_dual_kernel(const double *av, double *z, int64_t n_pad, double hat_gamma, RegParams reg) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const double lo = reg.lo_v ? reg.lo_v[j] : reg.lo_s, hi = reg.hi_v ? reg.hi_v[j] : reg.hi_s;
    const double a = av[j];
    z[j] = __ddiv_rn(__dsub_rn(prox_rt(reg.kind, a, hat_gamma * reg.lambda, lo, hi), a), hat_gamma);
}

static int ws_reserve(ciao_ctx *c, size_t need) {
    if (need > c->ws_bytes) {
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->ws, need));
        c->ws_bytes = need;
    }
    return CIAO_OK;
}

static void grid2d(ciao_ctx *c, int64_t N, int64_t n_pad, dim3 *grid, int *G) {
    const int chunks = (int)((n_pad / 2 + 255) / 256);
    int rows = std::max(1, c->num_sms * 8 / chunks);
    if (rows > N) rows = (int)N;
    *grid = dim3(rows, chunks);
    *G = rows;
}

// leaves Σ (column sums) in c->partial[0..n_pad)
static int reduce_partials(ciao_ctx *c, int G) {
    const int nb = (int)((c->d_pad + 255) / 256);
    launch_reduce_ws(c, c->ws, c->ws, G, c->partial, c->partial + c->d_pad, 0, 1);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int run_proshi_init(ciao_ctx *c, const double *x0_dev) {
    dim3 grid; int G;
    grid2d(c, c->N_total, c->d_pad, &grid, &G);
    CIAO_TRY(ws_reserve(c, ((size_t)G * c->d_pad + 16) * sizeof(double)));
    CUDA_TRY(cudaEventRecord(c->ev_pa, c->stream));
    proshi_init_kernel<<<grid, 256, 0, c->stream>>>(c->qd, c->ql, c->gamma_dev, c->gamma_dev + c->N_total, x0_dev, c->table, c->ws, c->N_total,
                                                    c->d_pad, c->box_lo, c->box_hi, c->eta, (double)c->N_total);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_pb, c->stream));
    c->timing.launches += 1;
    c->timing.last_pass_bytes = c->N_total * c->d_pad * 24;
    c->pass_timed = true;
    return reduce_partials(c, G);
}

int run_proshi_dual(ciao_ctx *c) {
    const int nb = (int)((c->d_pad + 255) / 256);
    proshi_dual_kernel<<<nb, 256, 0, c->stream>>>(ctx_vec(c, CIAO_VEC_AV), ctx_vec(c, CIAO_VEC_Z), c->d_pad, c->hat_gamma, c->reg);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int run_proshi_steps(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, const int64_t *ptr_dev, int64_t n_batches) {
    if (K <= 0) return CIAO_OK;
    ProshiArgs a;
    a.qd = c->qd; a.ql = c->ql; a.table = c->table; a.gam = c->gamma_dev; a.gam_n = c->gamma_dev + c->N_total; a.idx = idx_prepared;
    a.K = K; a.N = c->N_total; a.n_pad = c->d_pad;
    a.v_z = ctx_vec(c, CIAO_VEC_Z); a.v_av = ctx_vec(c, CIAO_VEC_AV);
    a.box_lo = c->box_lo; a.box_hi = c->box_hi; a.eta = c->eta; a.Nd = (double)c->N_total; a.hat_gamma = c->hat_gamma;
    a.reg = c->reg;
    CUDA_TRY(cudaEventRecord(c->ev_sa, c->stream));
    if (n_batches > 0 && K / n_batches >= PROSHI_BATCH_MIN) {
        // minibatches: blocks of a batch in parallel, one CTA per 8 columns
        proshi_batch_kernel<<<(int)((c->d_pad + 7) / 8), PROSHI_BT, 0, c->stream>>>(a, ptr_dev, n_batches);
    } else {
        const int T = (c->seq_threads > 32) ? 64 : 32;   // one warp per CTA spreads the columns over the most SMs
        const int grid = (int)((c->d_pad / 2 + T - 1) / T);
        const size_t smem = (size_t)(T / 32) * PROSHI_D * PROSHI_SLOT_BYTES;
        static bool configured = false;
        if (!configured) {
            CUDA_TRY(cudaFuncSetAttribute(proshi_steps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PROSHI_D * PROSHI_SLOT_BYTES));
            configured = true;
        }
        proshi_steps_kernel<<<grid, T, smem, c->stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_sb, c->stream));
    c->timing.launches += 1;
    c->timing.last_seq_steps = K;
    c->seq_timed = true;
    return CIAO_OK;
}

int run_proshi_solution(ciao_ctx *c) {
    dim3 grid; int G;
    grid2d(c, c->N_total, c->d_pad, &grid, &G);
    CUDA_TRY(cudaEventRecord(c->ev_pa, c->stream));
    proshi_solution_kernel<<<grid, 256, 0, c->stream>>>(c->table, c->gamma_dev, ctx_vec(c, CIAO_VEC_Z), c->N_total, c->d_pad);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev_pb, c->stream));
    c->timing.launches += 1;
    c->timing.last_pass_bytes = c->N_total * c->d_pad * 16;
    c->pass_timed = true;
    return CIAO_OK;
}

int run_table_colsum(ciao_ctx *c, int64_t n_rows) {
    dim3 grid; int G;
    grid2d(c, n_rows, c->d_pad, &grid, &G);
    CIAO_TRY(ws_reserve(c, ((size_t)G * c->d_pad + 16) * sizeof(double)));
    table_colsum_kernel<<<grid, 256, 0, c->stream>>>(c->table, c->ws, n_rows, c->d_pad);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return reduce_partials(c, G);
}
