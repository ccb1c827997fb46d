// TEMPORARY during bring-up.
#include "sd_internal.h"
extern "C" int sd_filter(sd_ctx *ctx, const sd_batch *, const sd_params *, const uint64_t *, const uint32_t *, const int64_t *,
                         const int64_t *, const uint32_t *, uint32_t, uint8_t *, int64_t *) { return sd_fail(ctx, SD_ERR_STATE, "sd_filter: not built yet"); }
