// oracle/shim -- TEST INFRASTRUCTURE ONLY.  BamTools is absent.  The part of BamTools::BamReader that the reference's
// Estimation/BamProcessing/BamController.cpp touches (Open, GetNextAlignment, GetReferenceData, GetHeader, Close), backed by a TEXT file so
// that the reference's own loop can run unmodified on the alignments a test also writes as a real BAM:
//   @SQ <TAB> name <TAB> length                                         one line per reference sequence, in order
//   name <TAB> ref id <TAB> position <TAB> CIGAR <TAB> flag <TAB> TAG:TYPE:VALUE ...      one line per alignment
#pragma once
#include "BamAlignment.h"
#include <fstream>
#include <sstream>

namespace BamTools
{
	struct RefData
	{
		std::string RefName;
		int32_t RefLength;
		RefData(const std::string &name = "", int32_t length = 0) : RefName(name), RefLength(length) {}
	};
	typedef std::vector<RefData> RefVector;
	struct SamHeader { std::string Text; };

	class BamReader
	{
		std::ifstream _in;
		RefVector _refs;
		SamHeader _header;
		std::string _pending; // first alignment line, read while looking for the end of the header

	public:
		bool Open(const std::string &filename)
		{
			_in.open(filename);
			if (!_in) return false;
			_refs.clear();
			_pending.clear();
			std::string line;
			while (std::getline(_in, line))
			{
				if (line.compare(0, 4, "@SQ\t") != 0) { _pending = line; break; }
				const size_t t = line.find('\t', 4);
				_refs.emplace_back(line.substr(4, t - 4), std::stoi(line.substr(t + 1)));
			}
			return true;
		}
		void Close() { _in.close(); }
		const RefVector &GetReferenceData() const { return _refs; }
		SamHeader GetHeader() const { return _header; }
		bool GetNextAlignment(BamAlignment &al)
		{
			std::string line;
			if (!_pending.empty()) { line.swap(_pending); }
			else
			{
				do { if (!std::getline(_in, line)) return false; } while (line.empty());
			}
			std::vector<std::string> f;
			std::istringstream ls(line);
			std::string tok;
			while (std::getline(ls, tok, '\t')) f.push_back(tok);
			al = BamAlignment();
			al.Name = f.at(0);
			al.RefID = std::stoi(f.at(1));
			al.Position = std::stoi(f.at(2));
			const std::string &cg = f.at(3);
			size_t i = 0;
			while (i < cg.size() && cg != "*")
			{
				size_t j = i;
				while (j < cg.size() && isdigit(cg[j])) ++j;
				al.CigarData.emplace_back(cg[j], uint32_t(std::stoul(cg.substr(i, j - i))));
				i = j + 1;
			}
			al.AlignmentFlag = uint32_t(std::stoul(f.at(4)));
			for (size_t k = 5; k < f.size(); ++k)
				if (f[k].size() >= 5) al.SetTag(f[k].substr(0, 2), f[k][3], f[k].substr(5));
			return true;
		}
	};
}
