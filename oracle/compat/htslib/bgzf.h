/* TEST INFRASTRUCTURE ONLY (oracle/): see sam.h in this directory. */
#ifndef ALETSCH_B200_ORACLE_COMPAT_HTSLIB_BGZF_H
#define ALETSCH_B200_ORACLE_COMPAT_HTSLIB_BGZF_H
#include "htslib/sam.h"
int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence);
int64_t bgzf_tell(BGZF *fp);
BGZF *bgzf_open(const char *path, const char *mode);
int bgzf_close(BGZF *fp);
#endif
