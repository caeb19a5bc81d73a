// TEST INFRASTRUCTURE ONLY (oracle/): see thread_pool.hpp in this directory.
#ifndef ALETSCH_B200_ORACLE_COMPAT_ASIO_POST_HPP
#define ALETSCH_B200_ORACLE_COMPAT_ASIO_POST_HPP
#include "boost/asio/thread_pool.hpp"
namespace boost { namespace asio {
template<class F> inline void post(thread_pool &, F f) { f(); }
} }
#endif
