// seq_lfinito.cu — instantiates the persistent cluster kernel (seq_impl.cuh) for ALG_LFINITO; one translation
// unit per algorithm so that the template instances compile in parallel.
#include "seq_impl.cuh"

int run_seq_lfinito(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d) {
    return run_seq_alg<ALG_LFINITO>(c, idx_prepared, K, m_d);
}

#ifdef CIAO_SEQ_PROFILE
// debug builds only (scripts/prof_seq.py): per-phase cycles of the last inner kernel of this translation unit
extern "C" int ciao_debug_seq_prof_lfinito(ciao_ctx *c, long long *out128) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out128, g_seq_prof, 16 * 8 * sizeof(long long)));
    return CIAO_OK;
}
#endif
