/* oracle/ref_shim/msc-datagroup.h -- stand-in for the reference's packet-data backend (out of scope): msc-handler.cpp only
 * needs the constructor to exist; the tests never select a data channel. */
#ifndef ORC_SHIM_MSC_DATAGROUP
#define ORC_SHIM_MSC_DATAGROUP
#include "dab-virtual.h"
class RadioInterface;
class mscDatagroup : public dabVirtual {
public:
	mscDatagroup (RadioInterface *, uint8_t, int16_t, int16_t, int16_t, int16_t, int16_t, uint8_t, int16_t, bool) {}
};
#endif
