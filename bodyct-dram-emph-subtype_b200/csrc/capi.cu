// Library-level entry points of libdram_b200.so: version, per-thread error text, device info.
#include <string.h>

#include "umma_common.cuh"

namespace dram {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static thread_local unsigned int *g_sat_counter = nullptr;
unsigned int *current_sat_counter() { return g_sat_counter; }

int sm_count() {
  // Cached per device; the hot path never switches devices inside one process (one rank = one GPU).
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      cudaGetLastError();
      sms = 148;
    }
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

}  // namespace dram

extern "C" int dram_version(void) { return DRAM_ABI_VERSION; }

extern "C" int dram_last_error(char *buf, size_t len) {
  if (!buf || len == 0) return DRAM_E_ARG;
  strncpy(buf, dram::g_error, len - 1);
  buf[len - 1] = '\0';
  return DRAM_OK;
}

extern "C" int dram_sm_count(void) { return dram::sm_count(); }

extern "C" int dram_set_saturation_counter(void *counter_u32) {
  DRAM_REQUIRE((reinterpret_cast<uintptr_t>(counter_u32) & 3) == 0, "dram_set_saturation_counter: counter must be 4-byte aligned");
  dram::g_sat_counter = reinterpret_cast<unsigned int *>(counter_u32);
  return DRAM_OK;
}
