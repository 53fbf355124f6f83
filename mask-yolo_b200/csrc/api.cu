#include <stdlib.h>
// api.cu -- library-level entry points (version, error string, device check).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace myolo {

int pdl_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MYOLO_PDL");       // default on; MYOLO_PDL=0 launches every kernel fully serialised
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode;
}
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace myolo

extern "C" int myolo_version(void) { return 100; }
extern "C" const char* myolo_last_error(void) { return myolo::g_err; }

extern "C" int myolo_device_check(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    myolo::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return MYOLO_ERR_CUDA;
  }
  if (prop.major != 10) {
    myolo::set_error("device %d is sm_%d%d; libmyolo_sm100 only runs on sm_100-class (B200) GPUs", dev, prop.major, prop.minor);
    return MYOLO_ERR_DEVICE;
  }
  return MYOLO_OK;
}
