/*
 * orc_kernels.h -- internal seam between the Tier-C loops (dab_oracle.c) and the Tier-A/B primitives.
 * The primitives are provided either by orc_port.c (plain-C restatement -> liboracle.so) or by
 * ref_shim/ref_provider.cpp (the reference's own classes, compiled unmodified -> _ref/libdabref.so).
 * Test infrastructure only (see dab_oracle.h).
 */
#ifndef ORC_KERNELS_H
#define ORC_KERNELS_H
#include <stdint.h>
#include "dab_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

const char *k_kind (void);
int   k_perm_table (const orc_params *p, int16_t *out);
float k_phi (int mode, int k);
int   k_pcode (int n, int8_t *out);
int   k_viterbi (int frameBits, const int16_t *soft, uint8_t *out);

/* protection decoder objects (one per sub-channel, like the reference's eep/uep_deconvolve instances) */
typedef struct k_prot k_prot;
k_prot *k_prot_new (int bitRate, int uepFlag, int protLevel);
void    k_prot_free (k_prot *);
int     k_prot_deconvolve (k_prot *, const int16_t *v, int32_t size, uint8_t *out);

typedef struct k_ofdm k_ofdm;
k_ofdm *k_ofdm_new (const orc_params *p, int threshold, int freqSyncMethod);
void    k_ofdm_free (k_ofdm *);
int32_t k_find_index (k_ofdm *, const float *v);
int     k_block0 (k_ofdm *, const float *v, int flag);
void    k_token (k_ofdm *, const float *inv, int16_t *ibits);
void    k_get_phase_reference (k_ofdm *, float *out);
void    k_get_ref_table (k_ofdm *, float *out);
int     k_fft (float *v, int n, int inverse);


/* DAB+ super-frame layer: firecode_checker::check on 11 bytes, reedSolomon (8, 0435, 0, 1, 10)::dec / enc with cutlen 135 */
int  k_firecode_check (const uint8_t *x11);
int  k_rs_dec (const uint8_t *r120, uint8_t *d110);
void k_rs_enc (const uint8_t *d110, uint8_t *r120);

#ifdef __cplusplus
}
#endif
#endif
