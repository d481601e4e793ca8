// oracle/shim -- TEST INFRASTRUCTURE ONLY.  BamTools is absent; this is the part of BamTools::BamAlignment that the reference's
// Estimation/BamProcessing/ReadParamsParser.cpp and FilledBamParamsParser.cpp touch: Name, Position, GetEndPosition() over the CIGAR
// (M, D, N, =, X consume the reference; default arguments = half-open end), string / character tags.  Tags are held decoded.
// For the writers (BamProcessorAbstract::save_alignment) also EditTag, restated from the published BamTools 2.5 API (api/BamAlignment.h:
// EditTag = RemoveTag when the tag exists + AddTag; AddTag refuses tag names that are not two characters and appends to the tag block):
// `TagOrder` keeps the order of the tag block.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace BamTools
{
	struct CigarOp
	{
		char Type;
		uint32_t Length;
		CigarOp(char type = '\0', uint32_t length = 0) : Type(type), Length(length) {}
	};

	struct BamAlignment
	{
		std::string Name, TagData;
		int32_t RefID = -1, Position = -1, Length = 0;
		uint32_t AlignmentFlag = 0;
		std::vector<CigarOp> CigarData;
		std::map<std::string, std::pair<char, std::string>> Tags; // tag -> (type, text)
		std::vector<std::string> TagOrder;                        // order of the tag block (maintained by SetTag / EditTag only)

		void SetTag(const std::string &tag, char type, const std::string &value) // test drivers: append a tag of the input file
		{
			if (Tags.find(tag) == Tags.end()) TagOrder.push_back(tag);
			Tags[tag] = std::make_pair(type, value);
		}
		bool HasTag(const std::string &tag) const { return Tags.find(tag) != Tags.end(); }
		void RemoveTag(const std::string &tag)
		{
			Tags.erase(tag);
			for (size_t k = 0; k < TagOrder.size(); ++k)
				if (TagOrder[k] == tag) { TagOrder.erase(TagOrder.begin() + long(k)); break; }
		}
		bool EditTag(const std::string &tag, const std::string &type, const std::string &value)
		{
			if (HasTag(tag)) RemoveTag(tag);
			if (tag.size() != 2 || type.size() != 1) return false; // AddTag: IsValidSize(tag, type)
			Tags[tag] = std::make_pair(type[0], value);
			TagOrder.push_back(tag);
			return true;
		}

		bool IsMapped() const { return !(AlignmentFlag & 0x4); }
		bool IsPrimaryAlignment() const { return !(AlignmentFlag & 0x100); }
		int GetEndPosition(bool usePadded = false, bool closedInterval = false) const
		{
			int end = Position;
			for (auto const &op : CigarData)
			{
				switch (op.Type)
				{
				case 'M': case 'D': case 'N': case '=': case 'X': end += int(op.Length); break;
				case 'P': if (usePadded) end += int(op.Length); break;
				default: break;
				}
			}
			if (closedInterval) --end;
			return end;
		}
		bool GetTagType(const std::string &tag, char &type) const
		{
			auto it = Tags.find(tag);
			if (it == Tags.end()) return false;
			type = it->second.first;
			return true;
		}
		bool GetTag(const std::string &tag, std::string &destination) const // string tags only, like BamTools' std::string overload
		{
			auto it = Tags.find(tag);
			if (it == Tags.end() || (it->second.first != 'Z' && it->second.first != 'H' && it->second.first != 'A')) return false;
			destination = it->second.second;
			return true;
		}
	};
}
