#pragma once
#include "BamAux.h"
#include <map>
#include <cassert>
namespace BamTools {
class BamAlignment {
public:
    std::string Name;
    int32_t Length = 0;
    std::string QueryBases;
    std::string AlignedBases;     // left empty: every reference use is guarded by .length()
    std::string Qualities;
    std::string TagData;
    int32_t RefID = -1;
    int32_t Position = -1;
    uint16_t Bin = 0;
    uint16_t MapQuality = 0;
    uint32_t AlignmentFlag = 0;
    std::vector<CigarOp> CigarData;
    int32_t MateRefID = -1;
    int32_t MatePosition = -1;
    int32_t InsertSize = 0;
    std::string Filename;
    std::map<std::string, int32_t> shim_int_tags;   // e.g. AS

    bool IsDuplicate() const { return AlignmentFlag & 0x400; }
    bool IsFailedQC() const { return AlignmentFlag & 0x200; }
    bool IsFirstMate() const { return AlignmentFlag & 0x40; }
    bool IsSecondMate() const { return AlignmentFlag & 0x80; }
    bool IsMapped() const { return !(AlignmentFlag & 0x4); }
    bool IsMateMapped() const { return !(AlignmentFlag & 0x8); }
    bool IsMateReverseStrand() const { return AlignmentFlag & 0x20; }
    bool IsPaired() const { return AlignmentFlag & 0x1; }
    bool IsPrimaryAlignment() const { return !(AlignmentFlag & 0x100); }
    bool IsProperPair() const { return AlignmentFlag & 0x2; }
    bool IsReverseStrand() const { return AlignmentFlag & 0x10; }
    bool BuildCharData() { return true; }
    // BamTools: Position + sum of reference-consuming ops (M,=,X,D,N; padded adds P); closedInterval subtracts 1.
    int GetEndPosition(bool usePadded = false, bool closedInterval = false) const {
        int e = Position;
        for (const CigarOp& op : CigarData) {
            switch (op.Type) {
            case 'M': case '=': case 'X': case 'D': case 'N': e += (int)op.Length; break;
            case 'P': if (usePadded) e += (int)op.Length; break;
            default: break;
            }
        }
        if (closedInterval) e -= 1;
        return e;
    }
    bool HasTag(const std::string& t) const { return shim_int_tags.count(t) > 0; }
    bool GetTagType(const std::string& t, char& type) const { if (!HasTag(t)) return false; type = Constants::BAM_TAG_TYPE_INT32; return true; }
    template <class T> bool GetTag(const std::string& t, T& dst) const {
        auto it = shim_int_tags.find(t); if (it == shim_int_tags.end()) return false; dst = (T)it->second; return true;
    }
    bool GetTag(const std::string&, std::string&) const { return false; }
    std::string ErrorString() const { return ""; }
};
}
