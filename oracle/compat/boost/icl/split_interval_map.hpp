// TEST INFRASTRUCTURE ONLY (oracle/): see interval_map.hpp in this directory.
#ifndef ALETSCH_B200_ORACLE_COMPAT_ICL_SPLIT_INTERVAL_MAP_HPP
#define ALETSCH_B200_ORACLE_COMPAT_ICL_SPLIT_INTERVAL_MAP_HPP
#include "boost/icl/interval_map.hpp"
#endif
