// oracle/shim -- TEST INFRASTRUCTURE ONLY.  Just enough of boost::iostreams for Tools/GeneAnnotation/RefGenesContainer.cpp of the reference:
// a filtering_istream onto which an optional gzip_decompressor and one std::istream source are pushed.  zlib does the inflating.
#pragma once
#include <istream>
#include <stdexcept>
#include <streambuf>
#include <zlib.h>

namespace boost { namespace iostreams {
	struct gzip_decompressor {};

	class filtering_istream : public std::istream
	{
		class buf_t : public std::streambuf
		{
		public:
			std::istream *src = nullptr;
			bool gz = false, z_init = false, z_end = false;
			z_stream zs;
			char in[1 << 16], out[1 << 16];
			~buf_t() { if (z_init) inflateEnd(&zs); }
			int_type underflow() override
			{
				if (!src) return traits_type::eof();
				if (!gz)
				{
					src->read(out, sizeof(out));
					const std::streamsize n = src->gcount();
					if (n <= 0) return traits_type::eof();
					setg(out, out, out + n);
					return traits_type::to_int_type(out[0]);
				}
				if (!z_init)
				{
					zs = z_stream();
					if (inflateInit2(&zs, 15 + 32) != Z_OK) throw std::runtime_error("zlib init failed");
					z_init = true;
				}
				while (!z_end)
				{
					if (zs.avail_in == 0)
					{
						src->read(in, sizeof(in));
						zs.next_in = reinterpret_cast<Bytef *>(in);
						zs.avail_in = uInt(src->gcount());
						if (zs.avail_in == 0) break;
					}
					zs.next_out = reinterpret_cast<Bytef *>(out);
					zs.avail_out = sizeof(out);
					const int rc = inflate(&zs, Z_NO_FLUSH);
					if (rc == Z_STREAM_END) { if (zs.avail_in == 0) z_end = true; else inflateReset(&zs); }
					else if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("gzip stream is corrupt");
					const size_t n = sizeof(out) - zs.avail_out;
					if (n) { setg(out, out, out + n); return traits_type::to_int_type(out[0]); }
					if (rc == Z_BUF_ERROR && zs.avail_in == 0) continue;
				}
				return traits_type::eof();
			}
		};
		buf_t _buf;

	public:
		filtering_istream() : std::istream(nullptr) { rdbuf(&_buf); }
		void push(const gzip_decompressor &) { _buf.gz = true; }
		void push(std::istream &source) { _buf.src = &source; }
	};
}}
