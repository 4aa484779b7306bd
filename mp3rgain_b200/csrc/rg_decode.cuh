// rg_decode.cuh -- device MP3 Layer III decode (stage A). Placeholder until the kernels land.
#pragma once
#include "rg_internal.h"
struct DecodeScratch {
    void release() {}
};
static int decode_init_tables(DecodeScratch&) { return 0; }
