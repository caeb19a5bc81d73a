#ifndef ALETSCH_B200_CSRC_K_BRIDGE_H
#define ALETSCH_B200_CSRC_K_BRIDGE_H
#include "runtime.h"
struct bridge_state { bool built = false; void release(agpu_ctx *) { built = false; } };
#endif
