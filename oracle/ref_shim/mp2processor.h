/* oracle/ref_shim/mp2processor.h -- stand-in for the reference's MP2 audio decoder (out of scope): records the
 * energy-dispersed bits dabConcurrent::run delivers (dab-concurrent.cpp:190) in the audioSink stand-in. */
#ifndef ORC_SHIM_MP2PROCESSOR
#define ORC_SHIM_MP2PROCESSOR
#include "dab-processor.h"
#include "audiosink.h"
class RadioInterface;
class mp2Processor : public dabProcessor {
public:
	mp2Processor (RadioInterface *, audioSink *as, FILE *, int16_t) : sink (as) {}
	void addtoFrame (uint8_t *v, int16_t n) { if (sink) sink -> take (v, (int) (uint16_t) n); }
	audioSink *sink;
};
#endif
