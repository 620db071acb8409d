// seq_finito.cu — instantiates the persistent cluster kernel (seq_impl.cuh) for ALG_FINITO; one translation
// unit per algorithm so that the template instances compile in parallel.
#include "seq_impl.cuh"

int run_seq_finito(ciao_ctx *c, const int64_t *idx_prepared, int64_t K, double m_d) {
    return run_seq_alg<ALG_FINITO>(c, idx_prepared, K, m_d);
}
