/* oracle/ref_shim/audiosink.h -- stand-in for the reference's audio output (PortAudio): the backend only calls
 * restart () on it.  Here it is also where the decoded blocks end up: the stand-in mp2Processor / mp4Processor append what
 * dabConcurrent::run hands to addtoFrame (test infrastructure only). */
#ifndef ORC_SHIM_AUDIOSINK
#define ORC_SHIM_AUDIOSINK
#include <stdint.h>
#include <mutex>
#include <vector>
class audioSink {
public:
	void restart () {}
	void stop () {}
	void take (const uint8_t *v, int n) { std::lock_guard<std::mutex> l (m); bits. insert (bits. end (), v, v + n); blocks ++; }
	int nblocks () { std::lock_guard<std::mutex> l (m); return blocks; }
	void take_au (const uint8_t *v, int n) { std::lock_guard<std::mutex> l (m); aus. emplace_back (v, v + n); }
	std::vector<std::vector<uint8_t> > aus;                 /* access units the DAB+ front handed to the AAC decoder stand-in */
	std::mutex m;
	std::vector<uint8_t> bits;
	int blocks = 0;
};
#endif
