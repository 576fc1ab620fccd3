// In-memory stand-in for the BamTools (>= 2.5) value types the reference's hot path touches.
// Test infrastructure only: lets the UNMODIFIED reference translation units compile and be
// driven from flat synthetic seed batches (oracle/ref_driver.cpp). No BAM file is ever read.
#pragma once
#include <string>
#include <vector>
#include <cstdint>
namespace BamTools {
namespace Constants {
const char BAM_TAG_TYPE_ASCII = 'A';
const char BAM_TAG_TYPE_INT8 = 'c';
const char BAM_TAG_TYPE_UINT8 = 'C';
const char BAM_TAG_TYPE_INT16 = 's';
const char BAM_TAG_TYPE_UINT16 = 'S';
const char BAM_TAG_TYPE_INT32 = 'i';
const char BAM_TAG_TYPE_UINT32 = 'I';
const char BAM_TAG_TYPE_FLOAT = 'f';
const char BAM_TAG_TYPE_STRING = 'Z';
const char BAM_TAG_TYPE_HEX = 'H';
const char BAM_TAG_TYPE_ARRAY = 'B';
}
struct CigarOp {
    char Type; uint32_t Length;
    CigarOp(char t = '\0', uint32_t l = 0) : Type(t), Length(l) {}
};
struct RefData {
    std::string RefName; int32_t RefLength;
    RefData(const std::string& n = "", int32_t l = 0) : RefName(n), RefLength(l) {}
};
typedef std::vector<RefData> RefVector;
struct BamRegion {
    int LeftRefID, LeftPosition, RightRefID, RightPosition;
    BamRegion(int lr = -1, int lp = -1, int rr = -1, int rp = -1) : LeftRefID(lr), LeftPosition(lp), RightRefID(rr), RightPosition(rp) {}
};
}
