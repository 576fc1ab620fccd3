#pragma once
#include "BamAlignment.h"
namespace BamTools {
// Reader over an in-memory record vector; Open() never touches the file system beyond what the
// reference itself asserts (it checks Utilities::fileExists on the BAM path before calling Open).
class BamReader {
public:
    static RefVector& shim_refs() { static RefVector r; return r; }
    static std::vector<BamAlignment>& shim_records() { static std::vector<BamAlignment> r; return r; }
    size_t cursor = 0;
    bool region_on = false; BamRegion region;    // SetRegion: only records on LeftRefID that overlap [LeftPosition, RightPosition) are returned, in vector (= file) order
    bool Open(const std::string&) { cursor = 0; region_on = false; return true; }
    bool Close() { return true; }
    bool IsOpen() const { return true; }
    bool LocateIndex() { return true; }
    bool HasIndex() const { return true; }
    bool Rewind() { cursor = 0; region_on = false; return true; }
    bool SetRegion(const BamRegion& r) { cursor = 0; region = r; region_on = true; return true; }
    bool Jump(int, int = 0) { cursor = 0; return true; }
    bool GetNextAlignment(BamAlignment& a) {
        const std::vector<BamAlignment>& R = shim_records();
        while (cursor < R.size()) {
            const BamAlignment& c = R[cursor++];
            if (region_on && !(c.RefID == region.LeftRefID && c.Position < region.RightPosition && c.GetEndPosition(false, false) > region.LeftPosition)) continue;
            a = c; return true;
        }
        return false;
    }
    bool GetNextAlignmentCore(BamAlignment& a) { return GetNextAlignment(a); }
    const RefVector& GetReferenceData() const { return shim_refs(); }
    int GetReferenceCount() const { return (int)shim_refs().size(); }
    int GetReferenceID(const std::string& n) const { for (size_t i = 0; i < shim_refs().size(); i++) if (shim_refs()[i].RefName == n) return (int)i; return -1; }
    std::string GetErrorString() const { return ""; }
    std::string GetFilename() const { return ""; }
};
}
