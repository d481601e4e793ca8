// oracle/shim: boost::split / boost::is_any_of (Tools/UtilFunctions.cpp:157, ConstLengthBarcodesParser.cpp:8-9).
#pragma once
#include <string>
#include <vector>
namespace boost {
	struct any_of_pred { std::string chars; bool operator()(char c) const { return chars.find(c) != std::string::npos; } };
	inline any_of_pred is_any_of(const std::string &s) { return any_of_pred{s}; }
	template <class Pred>
	void split(std::vector<std::string> &out, const std::string &in, Pred pred)
	{
		out.clear();
		std::string cur;
		for (char c : in)
		{
			if (pred(c)) { out.push_back(cur); cur.clear(); }
			else cur += c;
		}
		out.push_back(cur);
	}
}
