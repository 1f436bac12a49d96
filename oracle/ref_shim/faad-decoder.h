/* oracle/ref_shim/faad-decoder.h -- stand-in for the reference's AAC decoder wrapper (libfaad is absent and audio decoding is
 * out of scope): records every access unit mp4Processor::processSuperframe hands over after its CRC check
 * (mp4processor.cpp:272-277) in the audioSink stand-in (test infrastructure only). */
#ifndef ORC_SHIM_FAAD_DECODER
#define ORC_SHIM_FAAD_DECODER
#include <stdint.h>
#include "audiosink.h"
class faadDecoder {
public:
	faadDecoder (audioSink *as) : sink (as) {}
	int16_t MP42PCM (uint8_t, uint8_t, int16_t, uint8_t, uint8_t buffer [], int16_t bufferLength) {
		if (sink) sink -> take_au (buffer, bufferLength);
		return 1;                                            /* "samples decoded": anything but 0 (0 counts as a frame error) */
	}
	audioSink *sink;
};
#endif
