// Error reporting and device checks behind the C ABI (include/fdbm_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include <stdlib.h>

namespace fdbm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  cudaGetLastError();        // reset the (non-sticky) error so that the caller's next CUDA call is not poisoned
  set_error("CUDA error '%s' (%d) in %s at %s:%d", cudaGetErrorString(e), static_cast<int>(e), what, file, line);
  return FDBM_ECUDA;
}

static int g_arch_ok[64];     // 0 unknown, 1 ok, -1 not sm_100
static int g_sms[64];

int require_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return FDBM_EINVAL; }
  if (g_arch_ok[dev] == 0) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    g_sms[dev] = prop.multiProcessorCount;
    g_arch_ok[dev] = (prop.major == 10 && prop.minor == 0) ? 1 : -1;
  }
  if (g_arch_ok[dev] < 0) {
    set_error("libfdbm_b200 is built for sm_100a (B200) only; device %d is not sm_100 and there is no fallback", dev);
    return FDBM_EARCH;
  }
  return FDBM_OK;
}

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_sms[dev] == 0) require_sm100();
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

bool pdl_enabled() {
  static const bool on = !(getenv("FDBM_PDL") && atoi(getenv("FDBM_PDL")) == 0);
  return on;
}
// Largest batch whose convolution / GroupNorm-table launches ask for programmatic stream serialisation.  Inference: 8 (measured:
// one utterance 18.78 -> 18.28 ms, 256 utterances 1.2 % slower).  A training plan raises it to its batch while it runs its
// forward / backward (16 crops: 285.0 / 282.9 -> 287.8 / 285.6 crops/s): its launches are short whatever the level.
static thread_local int g_pdl_batch_limit = 8;
int pdl_batch_limit() { return g_pdl_batch_limit; }
PdlBatchScope::PdlBatchScope(int limit) : prev(g_pdl_batch_limit) { g_pdl_batch_limit = limit; }
PdlBatchScope::~PdlBatchScope() { g_pdl_batch_limit = prev; }

}  // namespace fdbm

extern "C" const char* fdbm_last_error(void) { return fdbm::g_error; }
extern "C" int fdbm_version(void) { return 100; }
extern "C" int fdbm_check_device(void) { return fdbm::require_sm100(); }
extern "C" int fdbm_operand_is_bf16(void) { return fdbm::kOperandIsBf16; }
