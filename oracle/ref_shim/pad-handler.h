/* oracle/ref_shim/pad-handler.h -- stand-in for the reference's PAD (dynamic label / MOT) handler, out of scope: mp4Processor only
 * constructs one and passes it the access units that carry PAD (test infrastructure only). */
#ifndef ORC_SHIM_PAD_HANDLER
#define ORC_SHIM_PAD_HANDLER
#include <stdint.h>
class RadioInterface;
class padHandler {
public:
	padHandler (RadioInterface *) {}
	void processPAD (uint8_t *) {}
};
#endif
