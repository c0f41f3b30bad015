/* counter_tool.c -- a minimal KokkosP tool (test infrastructure): logs every begin/end callback the Kokkos front end fires.
 * Interface: the C callbacks Kokkos looks up with dlsym in a library named by KOKKOS_TOOLS_LIBS
 * (reference: core/src/impl/Kokkos_Profiling.cpp, callback names kokkosp_begin_parallel_for / _reduce / _scan, kokkosp_end_*).
 * Used by tests/test_gpu_adapter.py to show that kernels dispatched to Kokkos::B200 are visible to tools exactly like the
 * reference's own backends (the hooks live in the reference's front end, Kokkos_Parallel.hpp:138-148). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static FILE* out;
static uint64_t next_id;

void kokkosp_init_library(const int load_seq, const uint64_t interface_version, const uint32_t ndev, void* devinfo) {
  (void)load_seq; (void)ndev; (void)devinfo;
  const char* p = getenv("KB200_TOOL_LOG");
  out = fopen(p ? p : "/tmp/kb200_tool.log", "w");
  if (out) fprintf(out, "init interface=%llu\n", (unsigned long long)interface_version);
}
void kokkosp_finalize_library(void) {
  if (out) { fprintf(out, "finalize\n"); fclose(out); out = NULL; }
}
#define HOOK(KIND)                                                                                   \
  void kokkosp_begin_parallel_##KIND(const char* name, const uint32_t dev, uint64_t* kid) {           \
    *kid = next_id++;                                                                                \
    if (out) fprintf(out, "begin " #KIND " %s dev=%u id=%llu\n", name, dev, (unsigned long long)*kid); \
  }                                                                                                  \
  void kokkosp_end_parallel_##KIND(const uint64_t kid) {                                             \
    if (out) fprintf(out, "end " #KIND " id=%llu\n", (unsigned long long)kid);                        \
  }
HOOK(for)
HOOK(reduce)
HOOK(scan)
