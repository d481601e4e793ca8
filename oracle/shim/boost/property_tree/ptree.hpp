// oracle/shim -- TEST INFRASTRUCTURE ONLY.  Just enough of boost::property_tree for Estimation/BamProcessing/BamTags.cpp of the reference:
// a flat key -> text store with get<T>(path, default).
#pragma once
#include <map>
#include <sstream>
#include <string>

namespace boost { namespace property_tree {
	class ptree
	{
		std::map<std::string, std::string> _values;

	public:
		void put(const std::string &path, const std::string &value) { _values[path] = value; }
		template <class T> T get(const std::string &path, const T &default_value) const
		{
			auto it = _values.find(path);
			if (it == _values.end()) return default_value;
			std::istringstream in(it->second);
			T v;
			in >> v;
			return v;
		}
		template <class T> T get(const std::string &path, const char *default_value) const { return get<T>(path, T(default_value)); }
	};
}}
