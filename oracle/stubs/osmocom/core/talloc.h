/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/talloc.h>
 * (see bits.h in this directory for why).  calloc/free are enough for the
 * per-block primitive + msgb the lower MAC allocates and frees. */
#pragma once
#include <stdlib.h>

#define talloc_zero(ctx, type)  ((type *)calloc(1, sizeof(type)))
#define talloc_free(ptr)        free(ptr)
#define talloc_zero_array(ctx, type, n)        ((type *)calloc((n), sizeof(type)))
#define talloc_realloc(ctx, ptr, type, n)      ((type *)realloc((ptr), (size_t)(n) * sizeof(type)))
