/* oracle/ref_shim/sndfile.h -- stand-in for libsndfile (absent from this image): ofdm-processor.cpp only touches it in its
 * sample-dump branch, which the oracle never switches on (test infrastructure only) */
#ifndef ORC_SHIM_SNDFILE
#define ORC_SHIM_SNDFILE
typedef struct SNDFILE_tag SNDFILE;
typedef long sf_count_t;
static inline sf_count_t sf_writef_short (SNDFILE *, const short *, sf_count_t n) { return n; }
#endif
