// st_gemm: argument validation and backend selection.
#include "common.cuh"

int st_gemm_simt(const st_gemm_args* a, cudaStream_t stream);
int st_gemm_tc(const st_gemm_args* a, cudaStream_t stream);          // gemm_tc.cu
int st_gemm_tc_supported(const st_gemm_args* a, const char** why);   // gemm_tc.cu

// bf16 problems that the AUTO backend had to route to the fp32-FMA kernel (a 50x cliff on the fast path): counted so
// that benchmarks and the full-size tests can assert there were none, and the reason of the last one is kept.
static int g_simt_fallbacks = 0;
static const char* g_simt_fallback_why = "";

extern "C" __attribute__((visibility("default"))) int st_gemm_simt_fallbacks(int reset) {
  const int n = g_simt_fallbacks;
  if (reset) g_simt_fallbacks = 0;
  return n;
}
extern "C" __attribute__((visibility("default"))) const char* st_gemm_simt_fallback_reason(void) { return g_simt_fallback_why; }

extern "C" __attribute__((visibility("default"))) int st_gemm(const st_gemm_args* a, void* stream) {
  ST_CHECK_ARG(a != nullptr, "st_gemm: null args");
  ST_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0, "st_gemm: empty problem M=%d N=%d K=%d batch=%d", a->M,
               a->N, a->K, a->batch);
  ST_CHECK_ARG(a->A && a->B && a->C, "st_gemm: null operand");
  ST_CHECK_ARG(a->a_mode == ST_OP_STRIDED || a->a_mode == ST_OP_GATHER, "st_gemm: bad a_mode");
  ST_CHECK_ARG(a->b_mode >= ST_OP_STRIDED && a->b_mode <= ST_OP_DGRADW, "st_gemm: bad b_mode");
  ST_CHECK_ARG(!(a->a_mode == ST_OP_GATHER && a->b_mode == ST_OP_GATHER), "st_gemm: only one gathered operand");
  ST_CHECK_ARG(a->b_mode != ST_OP_DGRADW || a->a_mode == ST_OP_GATHER, "st_gemm: DGRADW needs a gathered A");
  if (a->a_mode == ST_OP_GATHER || a->b_mode == ST_OP_GATHER) {
    const int Ct = a->C1 + a->C2, taps = a->kh * a->kw;
    ST_CHECK_ARG(a->n_img > 0 && a->H > 0 && a->W > 0 && a->C1 > 0 && a->C2 >= 0 && taps > 0 && (a->kh & 1) && (a->kw & 1),
                 "st_gemm: bad gather geometry");
    ST_CHECK_ARG(a->batch == 1, "st_gemm: gathered operands are not batched");
    if (a->a_mode == ST_OP_GATHER) {
      ST_CHECK_ARG(a->K == taps * Ct, "st_gemm: K=%d != taps*C=%d", a->K, taps * Ct);
      ST_CHECK_ARG(a->M == a->n_img * a->H * a->W, "st_gemm: M != n_img*H*W");
      ST_CHECK_ARG(a->C2 == 0 || a->A2, "st_gemm: missing second gather source");
    } else {
      ST_CHECK_ARG(a->N == taps * Ct, "st_gemm: N=%d != taps*C=%d", a->N, taps * Ct);
      ST_CHECK_ARG(a->K == a->n_img * a->H * a->W, "st_gemm: K != n_img*H*W");
      ST_CHECK_ARG(a->C2 == 0 || a->B2, "st_gemm: missing second gather source");
    }
  }
  if (a->accumulate) {
    ST_CHECK_ARG(a->out_dtype == ST_F32, "st_gemm: accumulate needs an fp32 output");
    ST_CHECK_ARG(!a->bias && !a->rowbias && !a->residual, "st_gemm: accumulate excludes bias/residual");
  } else {
    ST_CHECK_ARG(a->split_k <= 1, "st_gemm: split_k needs accumulate");
  }
  ST_CHECK_ARG(!a->rowbias || a->rows_per_rb > 0, "st_gemm: rows_per_rb");
  if (a->gn_rows_out) *a->gn_rows_out = 0;      // set by the tcgen05 backend when it emits GroupNorm partial sums

  int backend = a->backend;
  if (backend == ST_BACKEND_AUTO) {
    const char* why = nullptr;
    backend = (a->in_dtype == ST_BF16 && st_tc_available() && st_gemm_tc_supported(a, &why)) ? ST_BACKEND_TCGEN05
                                                                                            : ST_BACKEND_SIMT;
    if (backend == ST_BACKEND_SIMT && a->in_dtype == ST_BF16) {
      if (g_simt_fallbacks < 0x7fffffff) ++g_simt_fallbacks;
      g_simt_fallback_why = why ? why : "no sm_100 device";
    }
  }
  if (backend == ST_BACKEND_TCGEN05) {
    const char* why = "unsupported";
    if (!st_gemm_tc_supported(a, &why)) {
      st_set_error("st_gemm: tcgen05 backend cannot run this problem: %s", why);
      return ST_ERR_UNSUPPORTED;
    }
    return st_gemm_tc(a, (cudaStream_t)stream);
  }
  return st_gemm_simt(a, (cudaStream_t)stream);
}
