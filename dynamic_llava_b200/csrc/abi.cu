// C-ABI bookkeeping: version, last-error string.  No global mutable state besides the thread-local error slot.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dynllava_b200.h"

static thread_local char g_err[512] = "";

extern "C" int dl_set_error(int code, const char* where) {
  if (code > 0)
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, code, cudaGetErrorString((cudaError_t)code));
  else
    snprintf(g_err, sizeof(g_err), "%s: invalid argument / unsupported configuration (%d)", where, code);
  return code;
}
extern "C" const char* dl_last_error_string(void) { return g_err; }
extern "C" int dl_version(void) { return DL_B200_ABI_VERSION; }
#ifdef DL_HALF
extern "C" int dl_element_type(void) { return 1; }
#else
extern "C" int dl_element_type(void) { return 0; }
#endif
extern "C" int dl_device_is_sm100(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}
