// Library-level entry points: version, launch counter, capability probe.
#include "common.cuh"
#include "../../include/camradepth_b200.h"

#include <stdlib.h>

unsigned long long g_crd_launches = 0;

bool crd_pdl_enabled() {
  static int on = -1;
  // measured neutral on this workload (B=32 step 48.1 vs 47.7 ms, B=1 forward 4.91 vs 4.89 ms with / without):
  // the small kernels are bound by their own dependent memory round trips, not by the launch boundary.  Opt-in.
  if (on < 0) { const char* e = getenv("CAMRADEPTH_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}

extern "C" int crd_version(void) { return 1; }
extern "C" unsigned long long crd_launch_count(void) { return g_crd_launches; }
extern "C" int crd_has_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
