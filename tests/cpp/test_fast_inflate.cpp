// FastInflate::inflate_raw against zlib: streams deflated by zlib at every level / strategy from several kinds of data (BGZF-sized and
// smaller), and damaged streams (truncated, bit flips), which must never crash nor write outside the output -- they may only fail or
// produce bytes a CRC would reject.  CPU only.
#include "../../dropest_b200/host/FastInflate.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using Estimation::BamProcessing::FastInflate::inflate_raw;

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &src, int level, int strategy, int mem_level)
{
	z_stream zs{};
	if (deflateInit2(&zs, level, Z_DEFLATED, -15, mem_level, strategy) != Z_OK) { std::puts("deflateInit2 failed"); std::exit(2); }
	std::vector<uint8_t> out(deflateBound(&zs, uLong(src.size())) + 64);
	zs.next_in = const_cast<Bytef *>(src.data()); zs.avail_in = uInt(src.size());
	zs.next_out = out.data(); zs.avail_out = uInt(out.size());
	if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { std::puts("deflate failed"); std::exit(2); }
	out.resize(zs.total_out);
	deflateEnd(&zs);
	return out;
}

int main()
{
	std::mt19937_64 rng(7);
	size_t n_cases = 0, n_damaged = 0, n_damaged_ok = 0;
	auto make = [&](int kind, size_t n) {
		std::vector<uint8_t> v(n);
		switch (kind)
		{
		case 0: for (auto &b : v) b = uint8_t(rng()); break;                                   // incompressible: stored blocks
		case 1: for (auto &b : v) b = "ACGT"[rng() & 3]; break;                                // four symbols, short codes
		case 2: for (size_t i = 0; i < n; ++i) v[i] = uint8_t("read_name:0123456789ABCDEF"[i % 26] + (rng() % 97 == 0)); break; // long matches
		case 3: for (auto &b : v) b = uint8_t(rng() % 3 ? 'I' : 33 + rng() % 60); break;        // skewed: long and short codes
		case 4: for (size_t i = 0; i < n; ++i) v[i] = i && rng() % 5 ? v[i - 1] : uint8_t(rng()); break; // runs: distance 1
		case 5:
		{   // many distinct symbols with a geometric distribution: code lengths up to 15, subtables
			for (auto &b : v) { unsigned s = 0; while (s < 255 && (rng() & 1)) ++s; b = uint8_t(s * 37 + (rng() % 7 == 0 ? rng() : 0)); }
			break;
		}
		default:
		{   // BAM-like records
			size_t i = 0;
			while (i < n)
			{
				std::string rec = "A00123:45:HXXXX:1:" + std::to_string(rng() % 9999) + ":" + std::to_string(rng() % 99999) + std::string(1, '\0');
				for (int k = 0; k < 46; ++k) rec += char(rng());
				rec += std::string(91, 'F');
				rec += "CBZ"; for (int k = 0; k < 16; ++k) rec += "ACGT"[rng() & 3]; rec += std::string(1, '\0');
				rec += "GXZENSG000001" + std::to_string(rng() % 20000) + std::string(1, '\0');
				for (char c : rec) { if (i < n) v[i++] = uint8_t(c); }
			}
		}
		}
		return v;
	};
	const size_t sizes[] = {0, 1, 2, 7, 8, 9, 63, 255, 256, 257, 258, 259, 1000, 4096, 30000, 65280, 65536};
	for (int kind = 0; kind <= 6; ++kind)
		for (size_t n : sizes)
			for (int level : {0, 1, 3, 6, 9})
				for (int strategy : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED})
				{
					const auto src = make(kind, n);
					auto comp = deflate_raw(src, level, strategy, 1 + int(rng() % 9));
					const size_t clen = comp.size();
					comp.resize(clen + 16, 0xA5); // what follows a BGZF block's payload: 16 readable bytes
					std::vector<uint8_t> out(n + 32, 0xEE);
					const bool ok = inflate_raw(comp.data(), clen, out.data() + 16, n);
					++n_cases;
					bool same = ok && std::equal(src.begin(), src.end(), out.begin() + 16);
					for (int k = 0; k < 16; ++k) same = same && out[k] == 0xEE && out[16 + n + k] == 0xEE;
					if (!same) { std::printf("MISMATCH kind %d n %zu level %d strategy %d ok %d\n", kind, n, level, strategy, int(ok)); return 1; }
					// a wrong expected size must be refused
					if (n > 0 && inflate_raw(comp.data(), clen, out.data() + 16, n - 1)) { std::puts("accepted a short output"); return 1; }
					std::vector<uint8_t> big(n + 40);
					if (inflate_raw(comp.data(), clen, big.data(), n + 1)) { std::puts("accepted a long output"); return 1; }
					// damaged streams
					if (n >= 63 && level && (n_cases % 3 == 0))
						for (int rep = 0; rep < 6; ++rep)
						{
							std::vector<uint8_t> bad(comp.begin(), comp.begin() + clen);
							size_t blen = clen;
							if (rep % 2) blen = size_t(rng() % clen);
							else bad[size_t(rng() % clen)] ^= uint8_t(1u << (rng() % 8));
							bad.resize(blen);
							bad.resize(blen + 16, 0x5A);
							std::vector<uint8_t> o2(n + 32, 0xEE);
							const bool ok2 = inflate_raw(bad.data(), blen, o2.data() + 16, n);
							++n_damaged; n_damaged_ok += ok2;
							for (int k = 0; k < 16; ++k) if (o2[k] != 0xEE || o2[16 + n + k] != 0xEE) { std::puts("wrote outside the output"); return 1; }
						}
				}
	// fuzz against zlib's verdict: valid streams with 0-3 damages (bit flips, byte overwrites, truncation), random bytes, wrong output
	// sizes.  Whatever the decoder accepts, zlib accepts too, with the same bytes (the reverse need not hold); nothing is written outside.
	size_t n_fuzz = 0, n_accepted = 0;
	for (int iter = 0; iter < 40000; ++iter)
	{
		const size_t len = 1 + rng() % 3000;
		std::vector<uint8_t> src(len);
		const int kind = int(rng() % 4);
		for (auto &b : src) b = kind == 0 ? uint8_t(rng()) : kind == 1 ? uint8_t("ACGT"[rng() & 3]) : kind == 2 ? uint8_t('a' + rng() % 3) : uint8_t(rng() % 7 ? 'I' : rng());
		auto comp = deflate_raw(src, int(rng() % 10), int(rng() % 5), 1 + int(rng() % 9));
		int damages = int(rng() % 4);
		if (rng() % 50 == 0) { for (auto &b : comp) b = uint8_t(rng()); damages = 1; }
		for (int d = 0; d < damages; ++d)
			switch (rng() % 3)
			{
			case 0: comp[rng() % comp.size()] ^= uint8_t(1u << (rng() % 8)); break;
			case 1: comp[rng() % comp.size()] = uint8_t(rng()); break;
			default: if (comp.size() > 1) comp.resize(1 + rng() % (comp.size() - 1)); break;
			}
		const size_t clen = comp.size();
		std::vector<uint8_t> padded(comp);
		padded.resize(clen + 16, uint8_t(rng()));
		const size_t out_len = rng() % 8 ? len : rng() % (2 * len + 1);
		std::vector<uint8_t> out(out_len + 64, 0xEE), ref(out_len + 64, 0xEE);
		const bool ok = inflate_raw(padded.data(), clen, out.data() + 32, out_len);
		for (int k = 0; k < 32; ++k) if (out[k] != 0xEE || out[32 + out_len + k] != 0xEE) { std::puts("fuzz: wrote outside the output"); return 1; }
		++n_fuzz;
		if (!ok) continue;
		++n_accepted;
		z_stream is{};
		inflateInit2(&is, -15);
		is.next_in = comp.data(); is.avail_in = uInt(clen); is.next_out = ref.data() + 32; is.avail_out = uInt(out_len);
		const int rc = inflate(&is, Z_FINISH);
		const bool zok = rc == Z_STREAM_END && is.avail_out == 0;
		inflateEnd(&is);
		if (!zok || out != ref) { std::printf("fuzz: accepted what zlib refuses or decodes differently (iteration %d)\n", iter); return 1; }
	}
	std::printf("ok\t%zu streams equal to zlib's input\t%zu damaged streams survived (%zu decoded to something)\t%zu fuzz cases, %zu accepted, all as zlib decodes them\n",
	            n_cases, n_damaged, n_damaged_ok, n_fuzz, n_accepted);
	return 0;
}
