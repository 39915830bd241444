/* Minimal C host of liblemon_b200.so: the drop-in boundary is a plain C ABI (include/lemon_b200.h) -- no torch, no
 * Python.  Build:  gcc -std=c99 -Iinclude examples/c_abi_host.c -Llemon_b200 -llemon_b200 -Wl,-rpath,$PWD/lemon_b200 -o c_abi_host
 * Without a B200 the library refuses to create a context (there is no CPU fallback); with one, the program normalises
 * two rows through K0 and prints their norms (device memory is the caller's, here cudaMalloc'ed through the runtime the
 * library links). */
#include <stdio.h>
#include <string.h>
#include "lemon_b200.h"

int main(void) {
  printf("lemon_version %d\n", lemon_version());
  lemon_ctx* ctx = NULL;
  int rc = lemon_ctx_create(0, &ctx);
  if (rc != LEMON_OK) {
    printf("lemon_ctx_create failed with status %d (%s): no CPU fallback exists\n", rc,
           rc == LEMON_ERR_UNSUPPORTED ? "device is not sm_100" : rc == LEMON_ERR_CUDA ? "no usable CUDA device" : "error");
    return 3;
  }
  printf("context created on device 0; launches so far: %lld; last error: '%s'\n", (long long)lemon_launch_count(ctx),
         lemon_last_error(ctx));
  /* argument validation happens before any launch and never throws */
  rc = lemon_normalize_cast(ctx, NULL, NULL, NULL, NULL, NULL, 2, 8, 64, 0, 1, NULL);
  printf("normalize_cast(NULL input) -> %d, '%s'\n", rc, lemon_last_error(ctx));
  lemon_ctx_destroy(ctx);
  return rc == LEMON_ERR_INVALID ? 0 : 4;
}
