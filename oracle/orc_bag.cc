// TEST INFRASTRUCTURE ONLY (oracle/): C accessors of the result bag (orc_api.h).
#include "orc_api.h"
#include "orc_bag.h"

extern "C" {
void *orc_bag_new(void) { return new orc_bag; }
void orc_bag_free(void *bag) { delete (orc_bag*)bag; }
void orc_bag_clear(void *bag) { ((orc_bag*)bag)->items.clear(); }
int orc_bag_count(void *bag) { return (int)((orc_bag*)bag)->items.size(); }
const char *orc_bag_name(void *bag, int i) { return ((orc_bag*)bag)->items[i].name.c_str(); }
int orc_bag_kind(void *bag, int i) { return ((orc_bag*)bag)->items[i].kind; }
int64_t orc_bag_len(void *bag, int i)
{
	orc_bag::item &it = ((orc_bag*)bag)->items[i];
	return it.kind == 0 ? (int64_t)it.i.size() : (int64_t)it.d.size();
}
const void *orc_bag_data(void *bag, int i)
{
	orc_bag::item &it = ((orc_bag*)bag)->items[i];
	return it.kind == 0 ? (const void*)it.i.data() : (const void*)it.d.data();
}
}
