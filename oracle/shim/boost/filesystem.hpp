// oracle/shim: the one boost::filesystem::path use on the compiled path (Tools/UtilFunctions.cpp:146-152).
#pragma once
#include <string>
namespace boost { namespace filesystem {
	class path
	{
		std::string _p;
	public:
		struct codecvt_t {};
		path(const std::string &p = "") : _p(p) {}
		static codecvt_t codecvt() { return codecvt_t(); }
		path parent_path() const
		{
			auto pos = _p.find_last_of('/');
			return pos == std::string::npos ? path("") : path(_p.substr(0, pos));
		}
		path &append(const std::string &s, codecvt_t)
		{
			if (!_p.empty() && _p.back() != '/') _p += '/';
			_p += s;
			return *this;
		}
		std::string string() const { return _p; }
	};
}}
