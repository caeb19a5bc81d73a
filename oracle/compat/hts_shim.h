// TEST INFRASTRUCTURE ONLY (oracle/): in-memory "BAM files" behind the htslib stand-in.
#ifndef ALETSCH_B200_ORACLE_COMPAT_HTS_SHIM_H
#define ALETSCH_B200_ORACLE_COMPAT_HTS_SHIM_H
#include <stdint.h>
#include <string>
#include <vector>
#include "htslib/sam.h"

struct hts_shim_record
{
	bam1_core_t core;
	std::vector<uint8_t> data;     // qname(+NULs) | cigar | seq | qual | aux
};

struct hts_shim_file
{
	std::vector<std::string> target_name;
	std::vector<uint32_t> target_len;
	std::vector<hts_shim_record> records;   // in file (coordinate-sorted) order
};

// registers (or replaces) an in-memory file under `name`; sam_open(name) then reads it
void hts_shim_register(const std::string &name, const hts_shim_file &f);
void hts_shim_clear();

// build one record: qname, cigar ops, optional aux tags XS/ts ('A'), NH/HI/NM ('i')
hts_shim_record hts_shim_make_record(int32_t tid, int32_t pos, uint8_t mapq, uint16_t flag,
	int32_t mtid, int32_t mpos, int32_t isize, const std::string &qname,
	const uint32_t *cigar, uint32_t n_cigar, char xs, char ts, int nh, int hi, int nm);

// fill a bam1_t that points into rec.data (no copy; rec must outlive b)
void hts_shim_view(const hts_shim_record &rec, bam1_t *b);
#endif
