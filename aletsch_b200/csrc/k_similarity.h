#ifndef ALETSCH_B200_CSRC_K_SIMILARITY_H
#define ALETSCH_B200_CSRC_K_SIMILARITY_H
#include "runtime.h"
#endif
