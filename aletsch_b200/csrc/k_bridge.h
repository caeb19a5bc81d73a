#ifndef ALETSCH_B200_CSRC_K_BRIDGE_H
#define ALETSCH_B200_CSRC_K_BRIDGE_H
#include "runtime.h"
struct bridge_state { bool built = false; int64_t n_chain_val = 0; void release(agpu_ctx *) { built = false; } };
#endif
