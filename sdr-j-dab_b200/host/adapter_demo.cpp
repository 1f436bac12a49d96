// adapter_demo.cpp -- exercises the drop-in adapter classes (dab_adapters.h) exactly the way the reference's
// own code uses the classes they replace.  Built and run by tests/test_adapters_gpu.py on the GPU box;
// prints one hash line per class so the test can compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dab_adapters.h"

static uint32_t lcg (uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
static unsigned long long fnv (const uint8_t *p, size_t n) { unsigned long long h = 1469598103934665603ull; for (size_t i = 0; i < n; i ++) { h ^= p [i]; h *= 1099511628211ull; } return h; }

int main (int argc, char **argv) {
	try {
		uint32_t seed = 12345;
		{	// viterbi (768), as ficHandler does (fic-handler.cpp:84, 293)
			viterbi v (768);
			std::vector<int16_t> in (4 * 774); std::vector<uint8_t> out (768);
			for (auto &x : in) x = (int16_t) (lcg (seed) % 255) - 127;
			v. deconvolve (in. data (), out. data ());
			printf ("viterbi768 %016llx\n", fnv (out. data (), out. size ()));
		}
		{	// eep_deconvolve (128, 3-A) and uep_deconvolve (128, 3), as dabConcurrent does (dab-concurrent.cpp:78-83, 177-180)
			eep_deconvolve e (128, 0103); uep_deconvolve u (128, 3);
			std::vector<int16_t> in (96 * 64); std::vector<uint8_t> out (3072);
			for (auto &x : in) x = (int16_t) (lcg (seed) % 255) - 127;
			e. deconvolve (in. data (), 96 * 64, out. data ());
			printf ("eep128_3A %016llx\n", fnv (out. data (), out. size ()));
			u. deconvolve (in. data (), 96 * 64, out. data ());
			printf ("uep128_3 %016llx\n", fnv (out. data (), out. size ()));
		}
		{	// dabBackend = dabConcurrent: 20 CIF fragments in, 4 frames out
			int frames = 0; unsigned long long acc = 0;
			dabBackend b (1, 12 * 64, 16, 1, 0103, 0, [&] (uint8_t *v, int16_t n) { frames ++; acc ^= fnv (v, n) + frames; });
			std::vector<int16_t> frag (12 * 64);
			for (int t = 0; t < 20; t ++) { for (auto &x : frag) x = (int16_t) (lcg (seed) % 255) - 127; b. process (frag. data (), 12 * 64); }
			printf ("backend %d %016llx\n", frames, acc);
		}
		{	// figSubchannels = fib_processor's FIG 0/1 path on 40 pseudo-random FIC groups (all CRC flags set)
			figSubchannels f (1);
			std::vector<uint8_t> bits (40 * 768), crc (40 * 3, 1);
			for (auto &x : bits) x = (uint8_t) (lcg (seed) & 1);
			f. process_groups (bits. data (), crc. data (), 40);
			printf ("fig01 %016llx\n", fnv (reinterpret_cast<const uint8_t *> (f. ficList), sizeof (f. ficList)));
		}
		if (argc > 1) {	// mp4SuperframeFront = mp4Processor::addtoFrame, one CIF per call, bit rate 32: CIF blocks from a file
			FILE *fp = fopen (argv [1], "rb");
			if (!fp) throw std::runtime_error ("cannot open the CIF file");
			int count = 0; unsigned long long acc = 0;
			mp4SuperframeFront m (1, 32, [&] (const uint8_t *sf, int32_t n, const dabgpu_superframe_info &fi) {
				count ++; acc ^= fnv (sf, n) + (unsigned long long) fi. first_cif * 1315423911ull + fi. corrected;
			});
			std::vector<uint8_t> cif (24 * 32);
			while (fread (cif. data (), 1, cif. size (), fp) == cif. size ()) m. addtoFrame (cif. data (), 24 * 32);
			fclose (fp);
			printf ("dabplus %d %016llx\n", count, acc);
		}
		if (argc > 2) {	// ofdmProcessor (throughput form): a recording of complex floats behind a virtualInput, FIBs and frames through the handlers' sinks
			struct fileInput : public virtualInput {
				FILE *fp;
				explicit fileInput (const char *path) : fp (fopen (path, "rb")) { if (!fp) throw std::runtime_error ("cannot open the IQ file"); }
				~fileInput (void) { fclose (fp); }
				int32_t getSamples (DSPCOMPLEX *v, int32_t size) { return (int32_t) fread (v, sizeof (DSPCOMPLEX), (size_t) size, fp); }
				int32_t Samples (void) { return 1 << 20; }
			} rig (argv [2]);
			DabParams p = { 1, 76, 1536, 2656, 196608, 2552, 2048, 504, 1000 };          // gui.cpp:1361-1371
			int fibs = 0, frames = 0; unsigned long long facc = 0, macc = 0;
			ficHandler fic (nullptr, 1, [&] (uint8_t *fib, uint16_t ficno) { fibs ++; facc = facc * 31 + fnv (fib, 256) + ficno; });
			mscHandler msc (nullptr, &p, nullptr, 1, [&] (uint8_t *v, int16_t n) { frames ++; macc = macc * 31 + fnv (v, (size_t) n); });
			audiodata ad = {}; ad. startAddr = 0; ad. length = 96; ad. bitRate = 128; ad. uepFlag = 1; ad. protLevel = 0103;
			msc. set_audioChannel (&ad);
			ofdmProcessor op (&rig, &p, nullptr, &msc, &fic, 3, nullptr, 1, 5);
			op. run ();                                                                     // returns at the end of the file
			printf ("stream %lld %d %016llx %d %016llx %d\n", (long long) op. frames_decoded (), fibs, facc, frames, macc, (int) fic. get_ficRatio ());
		}
		if (argc > 3) {	// ficHandler::process_ficBlock / mscHandler::process_mscBlock, symbol by symbol as ofdmProcessor::run calls them
			// (ofdm-processor.cpp:421-441); soft bits [frames][75][3072] int16 from a file; once with dabConcurrent behind
			// mscHandler (concurrencyOn = 1, 16-CIF warm-up), once with dabSerial (15-CIF warm-up: one frame more at the start)
			FILE *fp = fopen (argv [3], "rb");
			if (!fp) throw std::runtime_error ("cannot open the soft-bit file");
			std::vector<int16_t> soft;
			{ int16_t buf [4096]; size_t n; while ((n = fread (buf, sizeof (int16_t), 4096, fp)) > 0) soft. insert (soft. end (), buf, buf + n); }
			fclose (fp);
			DabParams p = { 1, 76, 1536, 2656, 196608, 2552, 2048, 504, 1000 };
			const size_t nfr = soft. size () / (75 * 3072);
			for (int conc = 1; conc >= 0; conc --) {
				int fibs = 0, frames = 0; unsigned long long facc = 0, macc = 0, macc_skip1 = 0;
				ficHandler fic (nullptr, 2 * p. K, [&] (uint8_t *fib, uint16_t ficno) { fibs ++; facc = facc * 31 + fnv (fib, 256) + ficno; });
				mscHandler msc (nullptr, &p, nullptr, (uint8_t) conc, [&] (uint8_t *v, int16_t n) {
					frames ++; macc = macc * 31 + fnv (v, (size_t) n);
					if (frames > 1) macc_skip1 = macc_skip1 * 31 + fnv (v, (size_t) n);
				});
				audiodata ad = {}; ad. startAddr = 0; ad. length = 96; ad. bitRate = 128; ad. uepFlag = 1; ad. protLevel = 0103;
				msc. set_audioChannel (&ad);
				for (size_t f = 0; f < nfr; f ++)
					for (int blk = 1; blk < p. L; blk ++) {
						int16_t *ibits = &soft [(f * 75 + (size_t) (blk - 1)) * 3072];
						if (blk < 4) fic. process_ficBlock (ibits, (int16_t) blk); else msc. process_mscBlock (ibits, (int16_t) blk);
					}
				printf ("%s %d %016llx %d %016llx %016llx %d\n", conc ? "handlers_concurrent" : "handlers_serial", fibs, facc, frames, macc, macc_skip1, (int) fic. get_ficRatio ());
			}
		}
		if (argc > 4) {	// dabgpu_group_*: ONE recording (u8 IQ file) decoded by a group of handles -- two GPUs when the box has them, two
			// handles on one GPU otherwise -- against a single handle: parallel scheme, forced chain
			FILE *fp = fopen (argv [4], "rb");
			if (!fp) throw std::runtime_error ("cannot open the u8 IQ file");
			std::vector<uint8_t> iq;
			{ uint8_t buf [65536]; size_t n; while ((n = fread (buf, 1, sizeof (buf), fp)) > 0) iq. insert (iq. end (), buf, buf + n); }
			fclose (fp);
			const size_t nsamp = iq. size () / 2;
			const int cap = (int) (nsamp / 196608) + 4;
			dabgpu_subch sc [2] = { { 0, 96, 128, 1, 0103 }, { 96, 96, 128, 0, 3 } };
			struct Out { std::vector<uint8_t> fic, crc, m0, m1; std::vector<dabgpu_frame_info> info; uint8_t *mp [2]; int32_t nblk [2]; dabgpu_result r; };
			auto mk = [&] (Out &o) {
				o. fic. assign ((size_t) cap * 4 * 768, 0); o. crc. assign ((size_t) cap * 12, 0); o. m0. assign ((size_t) cap * 4 * 3072, 0); o. m1. assign ((size_t) cap * 4 * 3072, 0);
				o. info. resize (cap); o. mp [0] = o. m0. data (); o. mp [1] = o. m1. data ();
				o. r = dabgpu_result {}; o. r. max_frames = cap; o. r. info = o. info. data (); o. r. fic_bits = o. fic. data (); o. r. fic_crc = o. crc. data ();
				o. r. msc_bits = o. mp; o. r. msc_nblocks = o. nblk;
			};
			auto digest = [&] (Out &o) {
				unsigned long long h = fnv (o. fic. data (), (size_t) o. r. nframes * 4 * 768);
				h = h * 31 + fnv (o. m0. data (), (size_t) o. nblk [0] * 3072); h = h * 31 + fnv (o. m1. data (), (size_t) o. nblk [1] * 3072);
				for (int i = 0; i < o. r. nframes; i ++) h = h * 31 + (unsigned long long) o. info [i]. pos * 7 + (unsigned) o. info [i]. fine;
				return h;
			};
			dabgpu_config cfg = {}; cfg. device = 0; cfg. dabMode = 1; cfg. threshold = 3; cfg. freqSyncMethod = 1;
			dabgpu_t *one; if (dabgpu_create (&cfg, &one)) throw std::runtime_error (dabgpu_last_error (nullptr));
			dabgpu_host::check (one, dabgpu_set_subchannels (one, sc, 2));
			Out a; mk (a);
			dabgpu_host::check (one, dabgpu_decode (one, iq. data (), nsamp, &a. r));
			const int ndev = dabgpu_device_count ();
			int32_t devs [2] = { 0, ndev > 1 ? 1 : 0 };
			dabgpu_group_t *grp;
			if (dabgpu_group_create (&cfg, devs, 2, &grp)) throw std::runtime_error (dabgpu_last_error (nullptr));
			if (dabgpu_group_set_subchannels (grp, sc, 2)) throw std::runtime_error (dabgpu_group_last_error (grp));
			for (int scheme = 1; scheme >= 0; scheme --) {
				Out b; mk (b);
				int32_t used = -1;
				if (dabgpu_group_decode (grp, iq. data (), nsamp, &b. r, 40, scheme, &used)) throw std::runtime_error (dabgpu_group_last_error (grp));
				printf ("group_scheme%d %d %d %d %016llx %016llx %d\n", scheme, used, a. r. nframes, b. r. nframes, digest (a), digest (b), devs [1]);
			}
			dabgpu_group_destroy (grp); dabgpu_destroy (one);
		}
	} catch (const std::exception &e) { fprintf (stderr, "adapter_demo: %s\n", e. what ()); return 1; }
	return 0;
}
