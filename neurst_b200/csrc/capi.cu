// extern "C" surface of libb200st (see include/b200st.h).
#include "../../include/b200st.h"
#include "gemm.cuh"

namespace b200st {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
int64_t g_kernel_launches = 0;
}  // namespace b200st

using namespace b200st;

extern "C" {

const char* b200st_last_error(void) { return g_last_error.c_str(); }
int b200st_version(void) { return 100; }
int64_t b200st_launch_count(void) { return g_kernel_launches + tc_launch_count(); }

static GemmOperand to_operand(const b200st_operand& o) {
  GemmOperand r;
  r.ptr = o.ptr; r.dtype = o.dtype; r.mn_major = o.mn_major; r.ld = o.ld; r.sb1 = o.sb1; r.sb2 = o.sb2;
  return r;
}

static int convert_gemm(const b200st_gemm_args* a, GemmArgs& g);

int b200st_gemm(const b200st_gemm_args* a, void* stream) {
  if (!a) B200ST_FAIL("null args");
  GemmArgs g;
  B200ST_TRY(convert_gemm(a, g));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->force_simt) { if (g.splitk != 1) g.splitk = 1; return gemm_simt_f32(g, s); }
  return gemm(g, s);
}

int b200st_gemm_bench(const b200st_gemm_args* a, int32_t iters, float* ms_per_iter, void* stream) {
  if (!a || !ms_per_iter || iters <= 0) B200ST_FAIL("bad args");
  GemmArgs g;
  B200ST_TRY(convert_gemm(a, g));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  B200ST_CUDA(cudaEventCreate(&e0));
  B200ST_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) B200ST_TRY(gemm(g, s));
  B200ST_CUDA(cudaEventRecord(e0, s));
  for (int i = 0; i < iters; ++i) B200ST_TRY(gemm(g, s));
  B200ST_CUDA(cudaEventRecord(e1, s));
  B200ST_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B200ST_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_iter = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

static int convert_gemm(const b200st_gemm_args* a, GemmArgs& g) {
  g = gemm_defaults();
  g.M = a->M; g.N = a->N; g.K = a->K; g.nb1 = a->nb1 > 0 ? a->nb1 : 1; g.nb2 = a->nb2 > 0 ? a->nb2 : 1;
  g.A = to_operand(a->A); g.B = to_operand(a->B);
  g.C = a->C; g.c_dtype = a->c_dtype; g.ldc = a->ldc; g.c_sb1 = a->c_sb1; g.c_sb2 = a->c_sb2;
  g.epi.alpha = a->alpha;
  g.epi.bias = a->bias;
  g.epi.relu = a->relu;
  g.epi.mask_src = a->mask_src; g.epi.mask_dtype = a->mask_dtype;
  g.epi.mask_ld = a->mask_ld; g.epi.mask_sb1 = a->mask_sb1; g.epi.mask_sb2 = a->mask_sb2;
  if (a->dropout_p > 0.f) {
    B200ST_CHECK(a->dropout_p < 1.f, "dropout_p must be < 1");
    g.epi.drop = DropoutSpec{a->dropout_p, 1.f / (1.f - a->dropout_p), a->dropout_seed, a->dropout_stream};
  }
  g.epi.residual = a->residual; g.epi.res_ld = a->res_ld; g.epi.res_sb1 = a->res_sb1; g.epi.res_sb2 = a->res_sb2;
  g.epi.accumulate = a->accumulate;
  g.splitk = a->splitk;
  return 0;
}

int b200st_debug_tc(uint32_t mn_lbo, uint32_t mn_sbo, uint32_t k_lbo, uint32_t k_sbo, int32_t force_bn,
                    int32_t force_stages, int32_t max_ctas) {
  TcDebug& d = tc_debug();
  d.mn_lbo_bytes = mn_lbo; d.mn_sbo_bytes = mn_sbo; d.k_lbo_bytes = k_lbo; d.k_sbo_bytes = k_sbo;
  d.force_bn = force_bn; d.force_stages = force_stages; d.max_ctas = max_ctas;
  return 0;
}

}  // extern "C"
