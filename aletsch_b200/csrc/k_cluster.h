#ifndef ALETSCH_B200_CSRC_K_CLUSTER_H
#define ALETSCH_B200_CSRC_K_CLUSTER_H
#include "runtime.h"
struct cluster_state { bool built = false; void release(agpu_ctx *) { built = false; } };
#endif
