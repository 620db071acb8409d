// gen.cu — K10: counter-based synthetic problems written straight into HBM
// (include/ciao_gen.h; bit-identical to ciao_gen_host and to the oracle's generator).
#include "common.cuh"
#include "../../include/ciao_gen.h"

// row records [n_rows][ld]: a_i | tail = b_i or y_i | scale | 0 …
__global__ void gen_records_kernel(double *rec, int64_t n_rows, int64_t row0, int64_t d, int64_t d_pad, int64_t ld,
                                   int kind, uint64_t seed, double scale, int64_t il_block, int il_rank, int il_world) {
    const int64_t total = n_rows * ld;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / ld, j = e - r * ld, i = il_block ? il_global(r, il_block, il_rank, il_world) : row0 + r;
        double v = 0.0;
        if (j < d) v = ciao_syn_entry(kind, d, seed, i, j);
        else if (j == d_pad) v = ciao_syn_rhs(kind, d, seed, i);
        else if (j == d_pad + 1) v = scale;
        rec[e] = v;
    }
}

// sharing blocks: diag(Q_i) [N][d_pad], linear term ≡ 1 (test_sharing.jl:21)
__global__ void gen_blocks_kernel(double *qd, double *ql, int64_t N, int64_t n, int64_t n_pad, uint64_t seed) {
    const int64_t total = N * n_pad;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / n_pad, j = e - i * n_pad;
        qd[e] = j < n ? ciao_syn_entry(CIAO_SYN_SHARING, n, seed, i, j) : 0.0;
        ql[e] = j < n ? 1.0 : 0.0;
    }
}

int launch_gen_records(ciao_ctx *c, int kind, uint64_t seed, double scale) {
    gen_records_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->rec, c->n_rows, c->row0, c->d, c->d_pad, c->ld, kind, seed, scale,
                                                              c->il_block, c->il_rank, c->il_world);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int launch_gen_blocks(ciao_ctx *c, uint64_t seed) {
    gen_blocks_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->qd, c->ql, c->N_total, c->d, c->d_pad, seed);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

extern "C" int ciao_gen_host(int kind, int64_t d, uint64_t seed, int64_t row0, int64_t n_rows, double *A, double *rhs) {
    if (kind < 0 || kind > 2 || d <= 0 || n_rows < 0 || !A) CIAO_FAIL(CIAO_ERR_INVALID, "ciao_gen_host: bad arguments");
    for (int64_t r = 0; r < n_rows; ++r) {
        for (int64_t j = 0; j < d; ++j) A[r * d + j] = ciao_syn_entry(kind, d, seed, row0 + r, j);
        if (rhs && kind != CIAO_SYN_SHARING) rhs[r] = ciao_syn_rhs(kind, d, seed, row0 + r);
    }
    return CIAO_OK;
}
