// oracle/shim -- TEST INFRASTRUCTURE ONLY.  BamTools is absent.  The part of BamTools::BamWriter that the reference's
// Estimation/BamProcessing/BamProcessorAbstract.cpp touches (Open, SaveAlignment, Close).  Alignments are written as TEXT, one line each:
//   name <TAB> TAG:TYPE:VALUE ...      (the tags in the order of the tag block)
#pragma once
#include "BamReader.h"

namespace BamTools
{
	class BamWriter
	{
		std::ofstream _out;

	public:
		bool Open(const std::string &filename, const SamHeader &, const RefVector &)
		{
			_out.open(filename);
			return bool(_out);
		}
		void Close() { if (_out.is_open()) _out.close(); }
		bool SaveAlignment(const BamAlignment &al)
		{
			_out << al.Name;
			for (auto const &t : al.TagOrder)
			{
				auto const &v = al.Tags.at(t);
				_out << '\t' << t << ':' << v.first << ':' << v.second;
			}
			_out << '\n';
			return true;
		}
	};
}
