// placeholder, filled below
#ifndef ALETSCH_B200_CSRC_K_FRAGMENTS_H
#define ALETSCH_B200_CSRC_K_FRAGMENTS_H
#include "runtime.h"
struct fragments_state { bool built = false; void release(agpu_ctx *) { built = false; } };
#endif
