// indices.cu — host-generated index sequences are validated
// and prepared on the device: 1-based → 0-based, hazard flags for the table prefetch
// (seq.cu), batch-boundary flags for the prox, and LFinito's batch-order expansion.
#include <algorithm>

#include "common.cuh"

// A table row staged for step k was read from HBM after the exchange of step k − SEQ_D (seq_impl.cuh, SEQ_D = 8), i.e.
// possibly before the writes of steps k − 8 … k − 1; ProShI stages PROSHI_D = 16 steps ahead (proshi.cu).
constexpr int HAZARD_WINDOW = CIAO_HAZARD_WINDOW;  // flags j = 1 … 19 steps back

// out[k] = (raw[k] − 1) | HAZARD if the same row occurs in steps (k − HAZARD_WINDOW, k)
__global__ void prep_indices_kernel(const int64_t *raw, int64_t n, int64_t N, int64_t *out, int *err) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = raw[k];
        if (v < 1 || v > N) {
            atomicExch(err, 1);
            out[k] = 0;
            continue;
        }
        int64_t o = v - 1;
        for (int j = 1; j < HAZARD_WINDOW && j <= k; ++j)
            if (raw[k - j] == v) {
                o |= CIAO_FLAG_HAZARD | ((int64_t)j << CIAO_HAZ_DIST_SHIFT);  // nearest previous occurrence
                break;
            }
        out[k] = o;
    }
}

// batch j = [ptr[j], ptr[j+1]): flag its last step (Finito_basic.jl:118, ProShI_basic.jl:121)
__global__ void mark_batch_ends_kernel(const int64_t *ptr, int64_t n_batches, int64_t n_idx, int64_t *out, int *err) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_batches; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = ptr[j], hi = ptr[j + 1];
        if (lo < 0 || hi < lo || hi > n_idx) {
            atomicExch(err, 1);
            continue;
        }
        if (hi > lo) out[hi - 1] |= CIAO_FLAG_PROX;
    }
}

// LFinito (Finito_LFinito.jl:91-100): position jj of `order` (1-based batch numbers) covers rows
// r(j−1) .. min(rj, N) − 1; all batches hold r rows except batch nb (N − r(nb−1) rows) found at
// position short_pos.  out[step] = row | PROX on the first row of every batch.
__global__ void expand_batches_kernel(const int64_t *order, int64_t n_batches, int64_t r, int64_t N, int64_t nb,
                                      int64_t short_pos, int64_t *out, int *err) {
    const int64_t last_len = N - r * (nb - 1);
    const int64_t total = n_batches * r;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jj = e / r, t = e - jj * r;
        const int64_t j = order[jj];
        if (j < 1 || j > nb) {
            atomicExch(err, 1);
            continue;
        }
        const int64_t len = (j == nb) ? last_len : r;
        if (t >= len) continue;
        int64_t off = jj * r;
        if (short_pos >= 0 && jj > short_pos) off -= (r - last_len);
        out[off + t] = (r * (j - 1) + t) | (t == 0 ? CIAO_FLAG_PROX : 0);
    }
}

int launch_prep_indices(ciao_ctx *c, const int64_t *raw_dev, int64_t n, int64_t N, int64_t *out) {
    if (n <= 0) return CIAO_OK;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, c->num_sms * 8);
    prep_indices_kernel<<<grid, 256, 0, c->stream>>>(raw_dev, n, N, out, c->err_dev);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int launch_mark_batch_ends(ciao_ctx *c, const int64_t *ptr_dev, int64_t n_batches, int64_t n_idx, int64_t *out) {
    if (n_batches <= 0) return CIAO_OK;
    const int grid = (int)std::min<int64_t>((n_batches + 255) / 256, c->num_sms * 8);
    mark_batch_ends_kernel<<<grid, 256, 0, c->stream>>>(ptr_dev, n_batches, n_idx, out, c->err_dev);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}

int launch_expand_batches(ciao_ctx *c, const int64_t *order_dev, int64_t n_batches, int64_t r, int64_t N, int64_t nb,
                          int64_t short_pos, int64_t *out) {
    const int64_t total = n_batches * r;
    if (total <= 0) return CIAO_OK;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, c->num_sms * 8);
    expand_batches_kernel<<<grid, 256, 0, c->stream>>>(order_dev, n_batches, r, N, nb, short_pos, out, c->err_dev);
    CUDA_TRY(cudaGetLastError());
    c->timing.launches += 1;
    return CIAO_OK;
}
