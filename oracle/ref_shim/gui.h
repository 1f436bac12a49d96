/* oracle/ref_shim/gui.h -- stand-in for the reference's gui.h: the OFDM classes only pass a
 * RadioInterface* to connect(); nothing is ever called on it (test infrastructure only). */
#ifndef ORC_SHIM_GUI_H
#define ORC_SHIM_GUI_H
#include <QObject>
class RadioInterface : public QObject {};
#endif
