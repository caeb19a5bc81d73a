// TEST INFRASTRUCTURE ONLY (oracle/): named flat result arrays shared by the checkers.
#ifndef ALETSCH_B200_ORACLE_ORC_BAG_H
#define ALETSCH_B200_ORACLE_ORC_BAG_H
#include <stdint.h>
#include <string>
#include <vector>
#include <deque>

struct orc_bag
{
	struct item
	{
		std::string name;
		int kind;                    // 0 int32, 1 float64
		std::vector<int32_t> i;
		std::vector<double> d;
	};
	std::deque<item> items;   // deque: references handed out stay valid across later insertions

	std::vector<int32_t> &ints(const std::string &name)
	{
		for(size_t k = 0; k < items.size(); k++) if(items[k].name == name) { items[k].kind = 0; return items[k].i; }
		items.push_back(item());
		items.back().name = name;
		items.back().kind = 0;
		return items.back().i;
	}

	std::vector<double> &reals(const std::string &name)
	{
		for(size_t k = 0; k < items.size(); k++) if(items[k].name == name) { items[k].kind = 1; return items[k].d; }
		items.push_back(item());
		items.back().name = name;
		items.back().kind = 1;
		return items.back().d;
	}
};
#endif
