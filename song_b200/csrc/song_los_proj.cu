// song_los_proj.cu -- device assembly of the projection functions J_Llm(x) (bessel2.c:1041-1588).
#include <cuda_runtime.h>
#include "../../include/song_los.h"
extern "C" int song_los_build_projection(song_los_ctx *, int, const int *, const double *const *, const int *,
                                         const int *, int, const double *, const double *, const int *,
                                         const int *, const int *, const int *, double, double)
{
  return SONG_LOS_FAILURE;
}
extern "C" int song_los_get_projection(song_los_ctx *, int *, int *, double *const *) { return SONG_LOS_FAILURE; }
