// TEMPORARY during bring-up: entry points that are declared but not written yet report so.
#include "sd_internal.h"
extern "C" int sd_filter(sd_ctx *ctx, const sd_batch *, const sd_params *, const uint64_t *, const uint32_t *, const int64_t *,
                         const int64_t *, const uint32_t *, uint32_t, uint8_t *, int64_t *) { return sd_fail(ctx, SD_ERR_STATE, "sd_filter: not built yet"); }
extern "C" int sd_synth_sizes(sd_ctx *ctx, const sd_synth_params *, const uint32_t *, const uint32_t *, const uint8_t *, const double *,
                              uint32_t, uint64_t *, uint64_t *) { return sd_fail(ctx, SD_ERR_STATE, "sd_synth_sizes: not built yet"); }
extern "C" int sd_synth_fill(sd_ctx *ctx, const sd_synth_params *, const uint32_t *, const uint32_t *, const uint8_t *, const double *,
                             uint32_t, const uint64_t *, sd_batch *, int32_t *) { return sd_fail(ctx, SD_ERR_STATE, "sd_synth_fill: not built yet"); }
extern "C" int sd_synth_genome(sd_ctx *ctx, uint64_t, uint32_t *, uint32_t *, uint32_t *, uint64_t, const uint32_t *, const uint32_t *,
                               uint32_t, float) { return sd_fail(ctx, SD_ERR_STATE, "sd_synth_genome: not built yet"); }
