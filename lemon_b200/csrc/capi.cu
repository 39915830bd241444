// Context management and error plumbing of the C ABI (include/lemon_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "lemon_common.cuh"

int lemon_set_error(lemon_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

extern "C" int lemon_version(void) { return 100; }  // 0.1.0

extern "C" int lemon_ctx_create(int device, lemon_ctx** out) {
  if (!out) return LEMON_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LEMON_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LEMON_ERR_CUDA;
  if (prop.major != 10) return LEMON_ERR_UNSUPPORTED;   // sm_100a cubin only: no fallback path exists
  if (cudaSetDevice(device) != cudaSuccess) return LEMON_ERR_CUDA;
  lemon_ctx* c = new (std::nothrow) lemon_ctx();
  if (!c) return LEMON_ERR_INVALID;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->launches = 0;
  c->err[0] = 0;
  c->tc_scratch = nullptr;
  c->tc_scratch_bytes = 16 * 256 * sizeof(int32_t);
  c->tc_launch_seq = 0;
  if (cudaMalloc(&c->tc_scratch, c->tc_scratch_bytes) != cudaSuccess) { delete c; return LEMON_ERR_CUDA; }
  c->encode_tiled = nullptr;
  c->tune_kres = c->tune_debug = c->tune_cert = c->tune_boot = c->tune_bn = c->tune_pace = -1;
#ifdef LEMON_TC_EXPERIMENT
  {
    auto env_int = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : -1; };
    c->tune_kres = env_int("LEMON_TC_KRES");       c->tune_debug = env_int("LEMON_TC_DEBUG");
    c->tune_cert = env_int("LEMON_TC_CERT");       c->tune_boot = env_int("LEMON_TC_BOOT");
    c->tune_bn = env_int("LEMON_TC_BN");           c->tune_pace = env_int("LEMON_TC_PACE");
  }
#endif
  *out = c;
  return LEMON_OK;
}

extern "C" int lemon_ctx_destroy(lemon_ctx* ctx) {
  if (!ctx) return LEMON_OK;
  cudaSetDevice(ctx->device);
  if (ctx->tc_scratch) cudaFree(ctx->tc_scratch);
  delete ctx;
  return LEMON_OK;
}

extern "C" const char* lemon_last_error(lemon_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" int64_t lemon_launch_count(lemon_ctx* ctx) { return ctx ? ctx->launches : 0; }
