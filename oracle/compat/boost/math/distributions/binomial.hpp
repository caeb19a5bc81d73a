// TEST INFRASTRUCTURE ONLY (oracle/): declaration-level stand-in.  The only user,
// region::calculate_significance (rnacore/region.cc:200-264), is never called on the live
// path (its call at rnacore/region.cc:28 is commented out), so cdf() aborts if reached.
#ifndef ALETSCH_B200_ORACLE_COMPAT_MATH_BINOMIAL_HPP
#define ALETSCH_B200_ORACLE_COMPAT_MATH_BINOMIAL_HPP
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
namespace boost { namespace math {
template<class R = double> struct binomial_distribution { binomial_distribution(int n_, R p_) : n(n_), p(p_) {} int n; R p; };
template<class D, class X> struct complemented2 { const D &d; X x; };
template<class D, class X> inline complemented2<D, X> complement(const D &d, X x) { complemented2<D, X> c = {d, x}; return c; }
template<class D, class X> inline double cdf(const complemented2<D, X> &) { abort(); return 0; }
} }
#endif
