// TEST INFRASTRUCTURE ONLY (oracle/): declaration-level stand-in so that reference headers
// which merely name boost::asio::thread_pool (meta/bundle_group.h:23) compile.  The oracle
// never runs the reference's thread pool.
#ifndef ALETSCH_B200_ORACLE_COMPAT_ASIO_THREAD_POOL_HPP
#define ALETSCH_B200_ORACLE_COMPAT_ASIO_THREAD_POOL_HPP
namespace boost { namespace asio {
class thread_pool { public: explicit thread_pool(int) {} void join() {} };
} }
#endif
