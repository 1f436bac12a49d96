// dabgpu_engine.cu -- stream engine (placeholder until the OFDM group lands)
#include "dabgpu_internal.h"
int  dab_engine_init (dabgpu *h) { (void) h; return DABGPU_OK; }
void dab_engine_free (dabgpu *h) { (void) h; }
