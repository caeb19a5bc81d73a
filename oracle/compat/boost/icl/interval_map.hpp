// TEST INFRASTRUCTURE ONLY (oracle/): a small stand-in for the subset of Boost.ICL
// that the reference's hot-path translation units use.  Boost is not installed in
// the build container; the reference TUs are compiled UNCHANGED from
// /root/reference against this header (see oracle/Makefile).
//
// Surface covered (call sites, reference file:line):
//   rnacore/interval_map.h:22-43      typedefs: right_open_interval, interval<>::type,
//                                     interval_map / split_interval_map with
//                                     partial_absorber, inplace_plus, inter_section
//   rnacore/bundle_base.cc:125-167    map += make_pair(ROI, int)
//   rnacore/bundle_base.cc:177        map.find(ROI)
//   meta/bundle.cc:102                map += map
//   rnacore/interval_map.cc:11-21     find(point), -= pair
//   rnacore/interval_map.cc:34,58     upper_bound(ROI), lower_bound(ROI)
//   rnacore/region.cc:116,125         size() (ICL: cardinality, used as emptiness test)
//   meta/bundle_group.cc:181          rbegin()
//
// Semantics reproduced (Boost.ICL documentation, "Interval Maps", "Addability",
// "partial_absorber"):
//   * storage is an ordered map keyed by right-open intervals under the
//     "exclusive less" order (a < b iff upper(a) <= lower(b)); overlapping keys
//     are equivalent, so find/lower_bound/upper_bound behave like ICL's.
//   * add(I, v): v is combined (+= / set-union) into every stored segment that
//     overlaps I (segments are split at the borders of I first); the parts of I
//     that fall into gaps become new segments with value v.
//   * partial_absorber: adding an identity value is a no-op; a segment whose
//     value becomes the identity is erased.
//   * split_interval_map: borders introduced by any insertion are kept for ever.
//   * interval_map (joining): touching segments with equal values are merged.
//   * subtract(I, v): only stored segments are affected (no new segments).
#ifndef ALETSCH_B200_ORACLE_COMPAT_ICL_INTERVAL_MAP_HPP
#define ALETSCH_B200_ORACLE_COMPAT_ICL_INTERVAL_MAP_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <set>
#include <utility>

namespace boost { namespace icl {

template<class T>
class right_open_interval
{
public:
	typedef T domain_type;
	right_open_interval() : _lwb(T()), _upb(T()) {}
	explicit right_open_interval(const T &p) : _lwb(p), _upb(p + 1) {}
	right_open_interval(const T &l, const T &u) : _lwb(l), _upb(u) {}
	T lower() const { return _lwb; }
	T upper() const { return _upb; }
	bool operator==(const right_open_interval &o) const { return _lwb == o._lwb && _upb == o._upb; }
private:
	T _lwb;
	T _upb;
};

// icl::interval<T>::type; the reference only ever constructs it as (lower, upper)
// and reads lower()/upper(), i.e. right-open use (rnacore/interval_map.cc:39,58).
template<class T> struct interval { typedef right_open_interval<T> type; };

template<class T> inline T lower(const right_open_interval<T> &x) { return x.lower(); }
template<class T> inline T upper(const right_open_interval<T> &x) { return x.upper(); }
template<class T> inline bool is_empty(const right_open_interval<T> &x) { return !(x.lower() < x.upper()); }

struct partial_absorber {};
struct partial_enricher {};
template<class T> struct inplace_plus {};
template<class T> struct inter_section {};

namespace detail {

template<class I> struct exclusive_less
{
	bool operator()(const I &a, const I &b) const { return !(b.lower() < a.upper()); }
};

template<class C> struct codomain_ops
{
	static void add(C &a, const C &b) { a += b; }
	static void sub(C &a, const C &b) { a -= b; }
	static bool identity(const C &a) { return a == C(); }
};

template<class K, class L, class A> struct codomain_ops< std::set<K, L, A> >
{
	typedef std::set<K, L, A> S;
	static void add(S &a, const S &b) { a.insert(b.begin(), b.end()); }
	static void sub(S &a, const S &b) { for(typename S::const_iterator i = b.begin(); i != b.end(); ++i) a.erase(*i); }
	static bool identity(const S &a) { return a.empty(); }
};

template<class Domain, class Codomain, class Interval, bool Joining>
class basic_interval_map
{
public:
	typedef Interval interval_type;
	typedef Codomain codomain_type;
	typedef std::map<Interval, Codomain, exclusive_less<Interval> > impl_type;
	typedef typename impl_type::iterator iterator;
	typedef typename impl_type::const_iterator const_iterator;
	typedef typename impl_type::const_reverse_iterator const_reverse_iterator;
	typedef std::pair<Interval, Codomain> segment_type;
	typedef std::pair<const Interval, Codomain> value_type;
	typedef codomain_ops<Codomain> ops;

	const_iterator begin() const { return _m.begin(); }
	const_iterator end() const { return _m.end(); }
	const_reverse_iterator rbegin() const { return _m.rbegin(); }
	const_reverse_iterator rend() const { return _m.rend(); }

	const_iterator find(const Interval &k) const { return _m.find(k); }
	const_iterator find(const Domain &p) const { return _m.find(Interval(p)); }
	const_iterator lower_bound(const Interval &k) const { return _m.lower_bound(k); }
	const_iterator upper_bound(const Interval &k) const { return _m.upper_bound(k); }

	// ICL: size() == cardinality (number of domain elements), iterative_size() == #segments
	std::size_t size() const
	{
		std::size_t s = 0;
		for(const_iterator it = _m.begin(); it != _m.end(); ++it) s += (std::size_t)(it->first.upper() - it->first.lower());
		return s;
	}
	std::size_t iterative_size() const { return _m.size(); }
	bool empty() const { return _m.empty(); }
	void clear() { _m.clear(); }
	void swap(basic_interval_map &o) { _m.swap(o._m); }

	basic_interval_map& add(const segment_type &s)
	{
		const Interval &iv = s.first;
		if(is_empty(iv)) return *this;
		if(ops::identity(s.second)) return *this;
		const Domain a = iv.lower();
		const Domain b = iv.upper();
		split_at(a);
		split_at(b);

		Domain cur = a;
		iterator it = _m.lower_bound(Interval(a, a + 1));
		while(cur < b)
		{
			if(it == _m.end() || !(it->first.lower() < b))
			{
				_m.insert(it, value_type(Interval(cur, b), s.second));
				cur = b;
				break;
			}
			if(cur < it->first.lower())
			{
				_m.insert(it, value_type(Interval(cur, it->first.lower()), s.second));
				cur = it->first.lower();
				continue;
			}
			assert(it->first.lower() == cur);
			assert(!(b < it->first.upper()));
			cur = it->first.upper();
			ops::add(it->second, s.second);
			if(ops::identity(it->second)) { iterator d = it; ++it; _m.erase(d); }
			else ++it;
		}
		if(Joining) join_range(a, b);
		return *this;
	}

	basic_interval_map& subtract(const segment_type &s)
	{
		const Interval &iv = s.first;
		if(is_empty(iv)) return *this;
		if(ops::identity(s.second)) return *this;
		const Domain a = iv.lower();
		const Domain b = iv.upper();
		split_at(a);
		split_at(b);
		iterator it = _m.lower_bound(Interval(a, a + 1));
		while(it != _m.end() && it->first.lower() < b)
		{
			ops::sub(it->second, s.second);
			if(ops::identity(it->second)) { iterator d = it; ++it; _m.erase(d); }
			else ++it;
		}
		if(Joining) join_range(a, b);
		return *this;
	}

	basic_interval_map& operator+=(const segment_type &s) { return add(s); }
	basic_interval_map& operator-=(const segment_type &s) { return subtract(s); }
	basic_interval_map& operator+=(const basic_interval_map &o)
	{
		if(&o == this)
		{
			impl_type c(o._m);
			for(const_iterator it = c.begin(); it != c.end(); ++it) add(segment_type(it->first, it->second));
			return *this;
		}
		for(const_iterator it = o._m.begin(); it != o._m.end(); ++it) add(segment_type(it->first, it->second));
		return *this;
	}

private:
	// if p lies strictly inside a stored segment, cut that segment at p
	void split_at(const Domain &p)
	{
		iterator it = _m.find(Interval(p, p + 1));
		if(it == _m.end()) return;
		const Domain l = it->first.lower();
		const Domain u = it->first.upper();
		if(!(l < p)) return;
		Codomain v = it->second;
		iterator nx = it; ++nx;
		_m.erase(it);
		iterator li = _m.insert(nx, value_type(Interval(l, p), v));
		(void)li;
		_m.insert(nx, value_type(Interval(p, u), v));
	}

	// merge touching equal-valued neighbours among the segments that meet [a, b]
	void join_range(const Domain &a, const Domain &b)
	{
		iterator it = _m.lower_bound(Interval(a, a + 1));
		if(it != _m.begin()) --it;
		if(it != _m.begin() && !(it->first.upper() < a)) --it;
		while(it != _m.end())
		{
			iterator nx = it; ++nx;
			if(nx == _m.end()) break;
			if(b < it->first.lower()) break;
			if(it->first.upper() == nx->first.lower() && it->second == nx->second)
			{
				const Domain l = it->first.lower();
				const Domain u = nx->first.upper();
				Codomain v = it->second;
				iterator after = nx; ++after;
				_m.erase(it);
				_m.erase(nx);
				it = _m.insert(after, value_type(Interval(l, u), v));
			}
			else it = nx;
		}
	}

	impl_type _m;
};

} // namespace detail

template<class Domain, class Codomain,
	class Traits = partial_absorber,
	template<class> class Compare = std::less,
	template<class> class Combine = inplace_plus,
	template<class> class Section = inter_section,
	class Interval = right_open_interval<Domain> >
class interval_map : public detail::basic_interval_map<Domain, Codomain, Interval, true>
{
};

template<class Domain, class Codomain,
	class Traits = partial_absorber,
	template<class> class Compare = std::less,
	template<class> class Combine = inplace_plus,
	template<class> class Section = inter_section,
	class Interval = right_open_interval<Domain> >
class split_interval_map : public detail::basic_interval_map<Domain, Codomain, Interval, false>
{
};

} } // namespace boost::icl

#endif
