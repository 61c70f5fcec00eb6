// C-ABI plumbing: error reporting, device check, GEMM dispatch, dtype cast.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

int crog_gemm_simt(const CrogGemm* g, cudaStream_t stream);
int crog_gemm_tc(const CrogGemm* g, cudaStream_t stream);

static thread_local char g_err[512] = "";

void crog_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* crog_last_error(void) { return g_err; }

bool crog_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    // measured on B200 (profiles/README.md, r1 log): with the forward replayed as a CUDA graph the programmatic edges
    // cost 3 % (15.13 -> 15.65 ms/step), so the attribute is opt-in
    const char* e = getenv("CROG_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}
extern "C" int crog_abi_version(void) { return 6; }

extern "C" int crog_check_device(void) {
  int dev = 0;
  CROG_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  CROG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CROG_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  CROG_REQUIRE(major == 10 && minor == 0, CROG_E_UNSUPPORTED_ARCH,
               "crog_b200 is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)", dev, major, minor);
  return CROG_OK;
}

extern "C" int crog_gemm(const CrogGemm* g, void* stream) {
  CROG_REQUIRE(g != nullptr && g->a && g->w && g->out, CROG_E_BADSHAPE, "gemm: null operand");
  CROG_REQUIRE(g->taps == 1 || g->taps == 9, CROG_E_BADSHAPE, "gemm: taps must be 1 or 9 (got %d)", g->taps);
  CROG_REQUIRE(g->taps == 1 || (g->in_padded && g->H > 0), CROG_E_BADSHAPE, "gemm: 3x3 taps need the padded layout");
  CROG_REQUIRE(g->N > 0 && g->M >= 0 && g->cin > 0, CROG_E_BADSHAPE, "gemm: bad extents");
  CROG_REQUIRE(g->out_ld % 8 == 0 && (!g->residual || g->res_ld % 8 == 0), CROG_E_BADALIGN, "gemm: out/res ld must be multiples of 8");
  CROG_REQUIRE(!g->gate || (g->scale2 && g->bias2), CROG_E_BADSHAPE, "gemm: gate needs scale2/bias2");
  CROG_REQUIRE(!g->addmat || g->addmat_rows > 0, CROG_E_BADSHAPE, "gemm: addmat needs addmat_rows");
  if (g->H > 0) {
    const int per = g->in_padded ? (g->H + 2) * (g->W + 2) : g->H * g->W;
    CROG_REQUIRE(g->sample_rows == per, CROG_E_BADSHAPE, "gemm: sample_rows %d != %d for %dx%d", g->sample_rows, per, g->H, g->W);
  }
  if (g->M == 0) return CROG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (g->dtype == CROG_F32) {
    CROG_REQUIRE(g->impl != CROG_IMPL_TCGEN05, CROG_E_BADSHAPE, "gemm: the tcgen05 path takes bf16 operands");
    return crog_gemm_simt(g, s);
  }
  CROG_REQUIRE(g->dtype == CROG_BF16, CROG_E_BADSHAPE, "gemm: unknown dtype %d", g->dtype);
  if (g->impl == CROG_IMPL_SIMT) return crog_gemm_simt(g, s);
  return crog_gemm_tc(g, s);
}

namespace {
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  pdl_launch();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<TO>(to_f(in[i]));
}
}  // namespace

extern "C" int crog_cast(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t n, void* stream) {
  if (n == 0) return CROG_OK;
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (in_dtype == CROG_F32 && out_dtype == CROG_BF16) crog_launch(cast_kernel<float, bf16>, dim3((unsigned)g), dim3(256), 0, s, (const float*)in, (bf16*)out, (long long)n);
  else if (in_dtype == CROG_BF16 && out_dtype == CROG_F32) crog_launch(cast_kernel<bf16, float>, dim3((unsigned)g), dim3(256), 0, s, (const bf16*)in, (float*)out, (long long)n);
  else if (in_dtype == CROG_F32) crog_launch(cast_kernel<float, float>, dim3((unsigned)g), dim3(256), 0, s, (const float*)in, (float*)out, (long long)n);
  else crog_launch(cast_kernel<bf16, bf16>, dim3((unsigned)g), dim3(256), 0, s, (const bf16*)in, (bf16*)out, (long long)n);
  CROG_LAUNCH_OK("cast");
  return CROG_OK;
}
