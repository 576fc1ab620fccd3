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
    bool Open(const std::string&) { cursor = 0; return true; }
    bool Close() { return true; }
    bool IsOpen() const { return true; }
    bool LocateIndex() { return true; }
    bool HasIndex() const { return true; }
    bool Rewind() { cursor = 0; return true; }
    bool SetRegion(const BamRegion&) { cursor = 0; return true; }
    bool Jump(int, int = 0) { cursor = 0; return true; }
    bool GetNextAlignment(BamAlignment& a) { if (cursor >= shim_records().size()) return false; a = shim_records()[cursor++]; return true; }
    bool GetNextAlignmentCore(BamAlignment& a) { return GetNextAlignment(a); }
    const RefVector& GetReferenceData() const { return shim_refs(); }
    int GetReferenceCount() const { return (int)shim_refs().size(); }
    int GetReferenceID(const std::string& n) const { for (size_t i = 0; i < shim_refs().size(); i++) if (shim_refs()[i].RefName == n) return (int)i; return -1; }
    std::string GetErrorString() const { return ""; }
    std::string GetFilename() const { return ""; }
};
}
