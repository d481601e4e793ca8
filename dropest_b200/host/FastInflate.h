// FastInflate.h -- raw DEFLATE (RFC 1951) decoder for BGZF blocks: whole input and whole output in memory, output size known.
// The BAM ingest (BamIngest.cpp) spends most of its host time inflating; zlib's streaming inflate pays for a state
// machine that can stop after any byte.  This one cannot stop: 64-bit bit buffer refilled without a branch, one table lookup per symbol
// (11-bit litlen / 8-bit offset primary tables with subtables for longer codes), up to two literals per refill, word-wise match copies.
// Contract: returns true iff the stream is well formed, ends with its final block and produced EXACTLY out_len bytes.  On false the
// output is unspecified and the caller falls back to zlib (which then decides whether the block is corrupt).  It accepts nothing zlib
// refuses (same rules for the code-length sets; fuzzed against zlib with damaged streams) and produces the same bytes for what both accept.  It may read up to 15 bytes beyond in + in_len (their values do not matter:
// a BGZF block has its 8-byte footer there, then the next block or the reader's spare bytes) and never writes outside [out, out + out_len).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace Estimation
{
namespace BamProcessing
{
namespace FastInflate
{
	constexpr unsigned LITLEN_BITS = 11, OFFSET_BITS = 8, PRECODE_BITS = 7;
	constexpr unsigned LITLEN_CAP = 2400, OFFSET_CAP = 512, PRECODE_CAP = 128; // primary table + the largest possible set of subtables
	// table entry: value << 16 | flags | extra_bits << 8 | bits to consume
	constexpr uint32_t LITERAL = 0x8000u, EXCEPTIONAL = 0x4000u, SUBTABLE = 0x2000u, END_OF_BLOCK = 0x1000u, INVALID = EXCEPTIONAL;

	inline uint64_t load64(const uint8_t *p) { uint64_t v; std::memcpy(&v, p, 8); return v; } // little-endian hosts (x86-64, aarch64)
	inline void store64(uint8_t *p, uint64_t v) { std::memcpy(p, &v, 8); }
	inline unsigned reverse_bits(unsigned code, unsigned len)
	{
		unsigned r = 0;
		for (unsigned i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
		return r;
	}

	// Canonical Huffman decode table for code lengths lens[0, n): false when zlib would refuse the lengths (inftrees.c: over-subscribed, or
	// incomplete unless the set is empty or -- literal/length and distance codes only -- one single code of length 1) or the subtables do not
	// fit.  The entries an accepted incomplete set leaves open are INVALID (decoding one fails the block).
	inline bool build_table(const uint8_t *lens, unsigned n, unsigned table_bits, uint32_t *table, unsigned cap, const uint32_t *symbol_entry,
	                        bool is_precode = false)
	{
		unsigned count[16] = {0};
		for (unsigned s = 0; s < n; ++s) ++count[lens[s]];
		count[0] = 0;
		int left = 1;
		unsigned max_len = 0;
		for (unsigned len = 1; len <= 15; ++len)
		{
			left = left * 2 - int(count[len]);
			if (left < 0) return false;
			if (count[len]) max_len = len;
		}
		if (left > 0 && max_len != 0 && (is_precode || max_len != 1)) return false;
		unsigned next_code[16];
		for (unsigned len = 1, code = 0; len <= 15; ++len)
		{
			code = (code + count[len - 1]) << 1;
			next_code[len] = code;
		}
		const unsigned primary = 1u << table_bits;
		for (unsigned i = 0; i < primary; ++i) table[i] = INVALID;
		uint16_t reversed[288];
		bool any_long = false;
		for (unsigned s = 0; s < n; ++s)
		{
			const unsigned len = lens[s];
			if (!len) continue;
			reversed[s] = uint16_t(reverse_bits(next_code[len]++, len));
			any_long |= len > table_bits;
		}
		if (any_long)
		{   // codes that share their first table_bits bits share a subtable, as wide as the longest of them needs
			uint8_t sub_bits[1u << LITLEN_BITS];
			std::memset(sub_bits, 0, primary);
			for (unsigned s = 0; s < n; ++s)
				if (lens[s] > table_bits)
				{
					uint8_t &b = sub_bits[reversed[s] & (primary - 1)];
					if (lens[s] - table_bits > b) b = uint8_t(lens[s] - table_bits);
				}
			unsigned next = primary;
			for (unsigned i = 0; i < primary; ++i)
				if (sub_bits[i])
				{
					const unsigned size = 1u << sub_bits[i];
					if (next + size > cap) return false;
					table[i] = (next << 16) | EXCEPTIONAL | SUBTABLE | (uint32_t(sub_bits[i]) << 8) | table_bits;
					for (unsigned k = 0; k < size; ++k) table[next + k] = INVALID;
					next += size;
				}
		}
		for (unsigned s = 0; s < n; ++s)
		{
			const unsigned len = lens[s];
			if (!len) continue;
			const unsigned r = reversed[s];
			if (len <= table_bits)
			{
				const uint32_t e = symbol_entry[s] | len;
				for (unsigned i = r; i < primary; i += 1u << len) table[i] = e;
			}
			else
			{
				const uint32_t ptr = table[r & (primary - 1)];
				const unsigned start = ptr >> 16, size = 1u << ((ptr >> 8) & 15u), rest = len - table_bits;
				const uint32_t e = symbol_entry[s] | rest;
				for (unsigned i = r >> table_bits; i < size; i += 1u << rest) table[start + i] = e;
			}
		}
		return true;
	}

	struct SymbolEntries
	{
		uint32_t litlen[288], offset[32], precode[19];
		SymbolEntries()
		{
			static const uint16_t length_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
			static const uint8_t length_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
			static const uint16_t offset_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
			static const uint8_t offset_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
			for (unsigned s = 0; s < 256; ++s) litlen[s] = (s << 16) | LITERAL;
			litlen[256] = EXCEPTIONAL | END_OF_BLOCK;
			for (unsigned s = 257; s < 286; ++s) litlen[s] = (uint32_t(length_base[s - 257]) << 16) | (uint32_t(length_extra[s - 257]) << 8);
			litlen[286] = litlen[287] = INVALID;
			for (unsigned s = 0; s < 30; ++s) offset[s] = (uint32_t(offset_base[s]) << 16) | (uint32_t(offset_extra[s]) << 8);
			offset[30] = offset[31] = INVALID;
			for (unsigned s = 0; s < 19; ++s) precode[s] = s << 16;
		}
	};

	struct FixedTables // BTYPE = 01
	{
		uint32_t litlen[LITLEN_CAP], offset[OFFSET_CAP];
		bool ok;
		explicit FixedTables(const SymbolEntries &se)
		{
			uint8_t lens[288];
			for (unsigned s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
			ok = build_table(lens, 288, LITLEN_BITS, litlen, LITLEN_CAP, se.litlen);
			for (unsigned s = 0; s < 32; ++s) lens[s] = 5;
			ok = build_table(lens, 32, OFFSET_BITS, offset, OFFSET_CAP, se.offset) && ok;
		}
	};

	inline bool inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
	{
		static const SymbolEntries entries;
		static const FixedTables fixed(entries);
		const uint8_t *in_next = in, *const in_end = in + in_len;
		uint8_t *out_next = out, *const out_end = out + out_len;
		uint64_t bitbuf = 0;
		unsigned bitcnt = 0;
		uint32_t dyn_litlen[LITLEN_CAP], dyn_offset[OFFSET_CAP];
		// after REFILL at least 56 bits are valid.  in_next runs up to 7 bytes ahead of the position actually consumed (whole bytes wait in
		// the bit buffer), so a well-formed stream keeps in_next <= in_end + 7 and the 8 bytes read end below in_end + 15
#define DGE_REFILL()                                                                                                                        \
	do {                                                                                                                                    \
		if (in_next > in_end + 7) return false;                                                                                             \
		bitbuf |= load64(in_next) << bitcnt;                                                                                                \
		in_next += (63u - bitcnt) >> 3;                                                                                                     \
		bitcnt |= 56u;                                                                                                                      \
	} while (0)
#define DGE_TAKE(n) (bitbuf >>= (n), bitcnt -= (n))
		bool final_block = false;
		while (!final_block)
		{
			DGE_REFILL();
			final_block = bitbuf & 1u;
			const unsigned type = unsigned(bitbuf >> 1) & 3u;
			DGE_TAKE(3);
			const uint32_t *litlen, *offset;
			if (type == 0)
			{   // stored: byte-aligned LEN, ~LEN, bytes
				DGE_TAKE(bitcnt & 7u);
				const uint8_t *p = in_next - (bitcnt >> 3);
				bitbuf = 0; bitcnt = 0;
				if (p > in_end || in_end - p < 4) return false;
				const unsigned len = unsigned(p[0]) | (unsigned(p[1]) << 8), nlen = unsigned(p[2]) | (unsigned(p[3]) << 8);
				if ((len ^ nlen) != 0xFFFFu) return false;
				p += 4;
				if (size_t(in_end - p) < len || size_t(out_end - out_next) < len) return false;
				std::memcpy(out_next, p, len);
				out_next += len;
				in_next = p + len;
				continue;
			}
			else if (type == 1)
			{
				if (!fixed.ok) return false;
				litlen = fixed.litlen; offset = fixed.offset;
			}
			else if (type == 2)
			{
				const unsigned hlit = (unsigned(bitbuf) & 31u) + 257u, hdist = (unsigned(bitbuf >> 5) & 31u) + 1u, hclen = (unsigned(bitbuf >> 10) & 15u) + 4u;
				DGE_TAKE(14);
				if (hlit > 286 || hdist > 30) return false;
				static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
				uint8_t pre_lens[19] = {0};
				DGE_REFILL(); // 19 * 3 = 57 bits at most: 42 fit for the first 14 lengths, the rest after another refill
				for (unsigned i = 0; i < hclen; ++i)
				{
					if (i == 14) DGE_REFILL();
					pre_lens[order[i]] = uint8_t(bitbuf & 7u);
					DGE_TAKE(3);
				}
				uint32_t precode[PRECODE_CAP];
				if (!build_table(pre_lens, 19, PRECODE_BITS, precode, PRECODE_CAP, entries.precode, true)) return false;
				uint8_t lens[288 + 32 + 138];
				unsigned i = 0;
				const unsigned total = hlit + hdist;
				while (i < total)
				{
					DGE_REFILL();
					const uint32_t e = precode[bitbuf & ((1u << PRECODE_BITS) - 1)];
					if (e & EXCEPTIONAL) return false;
					DGE_TAKE(e & 0xFFu);
					const unsigned sym = e >> 16;
					if (sym < 16) { lens[i++] = uint8_t(sym); continue; }
					unsigned rep;
					uint8_t value = 0;
					if (sym == 16)
					{
						if (i == 0) return false;
						value = lens[i - 1];
						rep = 3 + (unsigned(bitbuf) & 3u); DGE_TAKE(2);
					}
					else if (sym == 17) { rep = 3 + (unsigned(bitbuf) & 7u); DGE_TAKE(3); }
					else { rep = 11 + (unsigned(bitbuf) & 127u); DGE_TAKE(7); }
					if (i + rep > total) return false;
					std::memset(lens + i, value, rep);
					i += rep;
				}
				if (lens[256] == 0) return false; // no end-of-block code
				if (!build_table(lens, hlit, LITLEN_BITS, dyn_litlen, LITLEN_CAP, entries.litlen)) return false;
				if (!build_table(lens + hlit, hdist, OFFSET_BITS, dyn_offset, OFFSET_CAP, entries.offset)) return false;
				litlen = dyn_litlen; offset = dyn_offset;
			}
			else return false;

			// ---- the symbols of one block
			while (true)
			{
				DGE_REFILL();
				uint32_t e = litlen[bitbuf & ((1u << LITLEN_BITS) - 1)];
				if (e & LITERAL)
				{   // the common case first: up to four literals per refill (a literal of the primary table takes at most 11 of the 56 bits)
					if (size_t(out_end - out_next) < 4) goto careful_literal;
					DGE_TAKE(e & 0xFFu);
					*out_next++ = uint8_t(e >> 16);
					e = litlen[bitbuf & ((1u << LITLEN_BITS) - 1)];
					if (e & LITERAL)
					{
						DGE_TAKE(e & 0xFFu);
						*out_next++ = uint8_t(e >> 16);
						e = litlen[bitbuf & ((1u << LITLEN_BITS) - 1)];
						if (e & LITERAL)
						{
							DGE_TAKE(e & 0xFFu);
							*out_next++ = uint8_t(e >> 16);
							e = litlen[bitbuf & ((1u << LITLEN_BITS) - 1)];
							if (e & LITERAL)
							{
								DGE_TAKE(e & 0xFFu);
								*out_next++ = uint8_t(e >> 16);
								continue;
							}
						}
					}
					DGE_REFILL(); // the low bits that selected e are unchanged
					if (false)
					{
					careful_literal: // the last three bytes of the output
						if (out_next == out_end) return false;
						DGE_TAKE(e & 0xFFu);
						*out_next++ = uint8_t(e >> 16);
						continue;
					}
				}
				if (e & EXCEPTIONAL)
				{
					if (e & SUBTABLE)
					{
						DGE_TAKE(LITLEN_BITS);
						e = litlen[(e >> 16) + (unsigned(bitbuf) & ((1u << ((e >> 8) & 15u)) - 1u))];
						if (e & LITERAL)
						{
							if (out_next == out_end) return false;
							DGE_TAKE(e & 0xFFu);
							*out_next++ = uint8_t(e >> 16);
							continue;
						}
					}
					if (e & EXCEPTIONAL)
					{
						if (!(e & END_OF_BLOCK)) return false;
						DGE_TAKE(e & 0xFFu);
						break;
					}
				}
				// a match: length (<= 15 + 5 bits) + offset (<= 15 + 13 bits)
				DGE_TAKE(e & 0xFFu);
				const unsigned length_extra = (e >> 8) & 15u;
				const size_t length = (e >> 16) + (unsigned(bitbuf) & ((1u << length_extra) - 1u));
				DGE_TAKE(length_extra);
				// >= 56 bits at the start of a match, at most 15 + 5 gone: the offset's 15 + 13 are there
				uint32_t d = offset[bitbuf & ((1u << OFFSET_BITS) - 1)];
				if (d & EXCEPTIONAL)
				{
					if (!(d & SUBTABLE)) return false;
					DGE_TAKE(OFFSET_BITS);
					d = offset[(d >> 16) + (unsigned(bitbuf) & ((1u << ((d >> 8) & 15u)) - 1u))];
					if (d & EXCEPTIONAL) return false;
				}
				DGE_TAKE(d & 0xFFu);
				const unsigned offset_extra = (d >> 8) & 15u;
				const size_t distance = (d >> 16) + (unsigned(bitbuf) & ((1u << offset_extra) - 1u));
				DGE_TAKE(offset_extra);
				if (distance > size_t(out_next - out) || length > size_t(out_end - out_next)) return false;
				const uint8_t *src = out_next - distance;
				uint8_t *dst = out_next;
				out_next += length;
				if (size_t(out_end - out_next) >= 16 && (distance >= 8 || distance == 1))
				{   // whole words; may write up to 15 bytes past the match, which are inside the output and rewritten by what follows
					if (distance == 1)
					{
						const uint64_t v = 0x0101010101010101ull * src[0];
						store64(dst, v); store64(dst + 8, v);
						for (dst += 16; dst < out_next; dst += 8) store64(dst, v);
					}
					else
					{   // most matches are short: two words without a loop, in order (the second may read what the first wrote)
						store64(dst, load64(src));
						store64(dst + 8, load64(src + 8));
						for (dst += 16, src += 16; dst < out_next; dst += 8, src += 8) store64(dst, load64(src));
					}
				}
				else
					for (size_t i = 0; i < length; ++i) dst[i] = src[i];
			}
		}
#undef DGE_REFILL
#undef DGE_TAKE
		// everything produced, and no bit taken from beyond the input
		return out_next == out_end && in_next - (bitcnt >> 3) <= in_end;
	}
}
}
}
