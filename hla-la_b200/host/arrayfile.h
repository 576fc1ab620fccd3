// Minimal named-array container used to hand flat batches between the generator, the C-ABI
// library, the oracle drivers and the Python test harness (numpy.frombuffer on the payloads).
// Layout: u32 magic 'HLAB', u32 n_arrays, then per array:
//   char name[24]; u32 dtype; u32 pad; u64 count; payload (count * elemsize, padded to 8 bytes).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

namespace hlala {

enum ArrayDType : uint32_t { DT_U8 = 1, DT_U16 = 2, DT_I32 = 3, DT_U32 = 4, DT_I64 = 5, DT_F64 = 6 };
inline size_t dtype_size(uint32_t dt) {
    switch (dt) { case DT_U8: return 1; case DT_U16: return 2; case DT_I32: case DT_U32: return 4; case DT_I64: case DT_F64: return 8; }
    throw std::runtime_error("arrayfile: bad dtype");
}

struct ArrayEntry { uint32_t dtype = 0; uint64_t count = 0; std::vector<uint8_t> data; };

class ArrayFile {
public:
    std::vector<std::string> order;
    std::map<std::string, ArrayEntry> arrays;

    template <class T> void put(const std::string& name, uint32_t dtype, const std::vector<T>& v) {
        if (sizeof(T) != dtype_size(dtype)) throw std::runtime_error("arrayfile: element size mismatch for " + name);
        ArrayEntry e; e.dtype = dtype; e.count = v.size(); e.data.resize(v.size() * sizeof(T));
        if (!v.empty()) memcpy(e.data.data(), v.data(), e.data.size());
        if (!arrays.count(name)) order.push_back(name);
        arrays[name] = std::move(e);
    }
    template <class T> const T* get(const std::string& name, uint64_t* count = nullptr) const {
        auto it = arrays.find(name);
        if (it == arrays.end()) throw std::runtime_error("arrayfile: missing array " + name);
        if (sizeof(T) != dtype_size(it->second.dtype)) throw std::runtime_error("arrayfile: dtype mismatch for " + name);
        if (count) *count = it->second.count;
        return reinterpret_cast<const T*>(it->second.data.data());
    }
    bool has(const std::string& name) const { return arrays.count(name) > 0; }
    uint64_t count(const std::string& name) const { return arrays.at(name).count; }

    void write(const std::string& path) const {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("arrayfile: cannot write " + path);
        uint32_t magic = 0x42414C48u, n = (uint32_t)order.size();
        fwrite(&magic, 4, 1, f); fwrite(&n, 4, 1, f);
        for (const std::string& name : order) {
            const ArrayEntry& e = arrays.at(name);
            char nm[24]; memset(nm, 0, sizeof nm); strncpy(nm, name.c_str(), 23);
            uint32_t pad = 0;
            fwrite(nm, 1, 24, f); fwrite(&e.dtype, 4, 1, f); fwrite(&pad, 4, 1, f); fwrite(&e.count, 8, 1, f);
            if (!e.data.empty()) fwrite(e.data.data(), 1, e.data.size(), f);
            size_t rem = (8 - e.data.size() % 8) % 8; uint64_t z = 0;
            if (rem) fwrite(&z, 1, rem, f);
        }
        fclose(f);
    }
    void read(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("arrayfile: cannot read " + path);
        uint32_t magic = 0, n = 0;
        if (fread(&magic, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || magic != 0x42414C48u) { fclose(f); throw std::runtime_error("arrayfile: bad header in " + path); }
        for (uint32_t i = 0; i < n; i++) {
            char nm[25]; nm[24] = 0; uint32_t dt, pad; uint64_t cnt;
            if (fread(nm, 1, 24, f) != 24 || fread(&dt, 4, 1, f) != 1 || fread(&pad, 4, 1, f) != 1 || fread(&cnt, 8, 1, f) != 1) { fclose(f); throw std::runtime_error("arrayfile: truncated " + path); }
            ArrayEntry e; e.dtype = dt; e.count = cnt; e.data.resize(cnt * dtype_size(dt));
            if (!e.data.empty() && fread(e.data.data(), 1, e.data.size(), f) != e.data.size()) { fclose(f); throw std::runtime_error("arrayfile: truncated payload " + path); }
            size_t rem = (8 - e.data.size() % 8) % 8; char skip[8];
            if (rem && fread(skip, 1, rem, f) != rem) { fclose(f); throw std::runtime_error("arrayfile: truncated pad " + path); }
            order.push_back(nm); arrays[nm] = std::move(e);
        }
        fclose(f);
    }
};

} // namespace hlala
