// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for boost::disjoint_sets<int*, int*>
// as used by the reference (rnacore/disjoint_set.h:12-35, meta/bundle_group.cc:296-342).
// Union by rank with path compression (full compression in find_set), the published
// Boost algorithm (boost/pending/detail/disjoint_sets.hpp: link_sets / find_representative_with_full_compression).
// The reference's results depend only on set membership and on the representative's
// size slot, which the caller re-writes after each link (meta/bundle_group.cc:311-315).
#ifndef ALETSCH_B200_ORACLE_COMPAT_DISJOINT_SETS_HPP
#define ALETSCH_B200_ORACLE_COMPAT_DISJOINT_SETS_HPP
#include <cstddef>
#include <cstdint>
namespace boost {
template<class RankPA, class ParentPA>
class disjoint_sets
{
public:
	disjoint_sets(RankPA r, ParentPA p) : rank(r), parent(p) {}
	template<class E> void make_set(E x) { parent[x] = x; rank[x] = 0; }
	template<class E> E find_set(E x)
	{
		E r = x;
		while(parent[r] != r) r = parent[r];
		while(parent[x] != r) { E n = parent[x]; parent[x] = r; x = n; }
		return r;
	}
	template<class E> void link(E x, E y)
	{
		x = find_set(x);
		y = find_set(y);
		if(x == y) return;
		if(rank[x] > rank[y]) parent[y] = x;
		else
		{
			parent[x] = y;
			if(rank[x] == rank[y]) ++rank[y];
		}
	}
	template<class E> void union_set(E x, E y) { link(find_set(x), find_set(y)); }
private:
	RankPA rank;
	ParentPA parent;
};
}
#endif
