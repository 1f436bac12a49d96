/* oracle/ref_shim/mp4processor.h -- stand-in for the reference's DAB+ audio decoder front (the super-frame layer is tested
 * separately): records what addtoFrame receives, like the mp2Processor stand-in. */
#ifndef ORC_SHIM_MP4PROCESSOR
#define ORC_SHIM_MP4PROCESSOR
#include "dab-processor.h"
#include "audiosink.h"
class RadioInterface;
class mp4Processor : public dabProcessor {
public:
	mp4Processor (RadioInterface *, audioSink *as, int16_t) : sink (as) {}
	void addtoFrame (uint8_t *v, int16_t n) { if (sink) sink -> take (v, (int) (uint16_t) n); }
	audioSink *sink;
};
#endif
