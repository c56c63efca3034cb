// los_fast.cu -- FAST-mode kernels (k3-sorted, register-tiled).  Placeholder until the tiled
// kernel lands: reports failure loudly rather than falling back to anything else.
#include <string>
#include "../../include/song_los.h"
#include "los_common.h"
namespace songlos {
struct FastState { int unused; };
FastState *fast_state_create() { return new FastState(); }
void fast_state_destroy(FastState *f) { delete f; }
int fast_compute(FastState *, const ExactParams &, int, cudaStream_t, song_los_timing *, std::string &err)
{
  err = "SONG_LOS_MODE_FAST is not available in this build";
  return 1;
}
}  // namespace songlos
