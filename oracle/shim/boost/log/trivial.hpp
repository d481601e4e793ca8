// oracle/shim: BOOST_LOG_TRIVIAL -> null sink (Boost.Log is absent from this image).
#pragma once
#include <ostream>
namespace boost { namespace log { namespace trivial {
	enum severity_level { trace, debug, info, warning, error, fatal };
	struct null_stream
	{
		template <class T> null_stream &operator<<(const T &) { return *this; }
		null_stream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; } // std::flush / std::endl
	};
}}}
#define BOOST_LOG_TRIVIAL(lvl) ::boost::log::trivial::null_stream()
