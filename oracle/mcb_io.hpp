// oracle/mcb_io.hpp — tiny named-array container ("MCB1") shared by the oracle harness (C++) and the
// python side (mcut_b200/mcbio.py).  TEST INFRASTRUCTURE ONLY.
//
//   file   := "MCB1" u32:count  array*
//   array  := u32:name_len name  u32:dtype  u32:ndim  u64:dims[ndim]  raw little-endian data
//   dtype  := 0 u8 | 1 u32 | 2 i32 | 3 u64 | 4 f32 | 5 f64 | 6 i64
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mcb {

enum dtype_t : uint32_t { U8 = 0, U32 = 1, I32 = 2, U64 = 3, F32 = 4, F64 = 5, I64 = 6 };

inline size_t dtype_size(uint32_t d)
{
    static const size_t s[] = { 1, 4, 4, 8, 4, 8, 8 };
    if (d > 6) throw std::runtime_error("mcb: bad dtype");
    return s[d];
}

struct array_t {
    uint32_t dtype = U8;
    std::vector<uint64_t> dims;
    std::vector<uint8_t> bytes;

    size_t count() const
    {
        size_t n = 1;
        for (uint64_t d : dims) n *= (size_t)d;
        return n;
    }
    template <typename T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
    template <typename T> T* as() { return reinterpret_cast<T*>(bytes.data()); }
};

typedef std::map<std::string, array_t> file_t;

template <typename T> struct dtype_of;
template <> struct dtype_of<uint8_t> { static const uint32_t v = U8; };
template <> struct dtype_of<uint32_t> { static const uint32_t v = U32; };
template <> struct dtype_of<int32_t> { static const uint32_t v = I32; };
template <> struct dtype_of<uint64_t> { static const uint32_t v = U64; };
template <> struct dtype_of<float> { static const uint32_t v = F32; };
template <> struct dtype_of<double> { static const uint32_t v = F64; };
template <> struct dtype_of<int64_t> { static const uint32_t v = I64; };

template <typename T>
inline void put(file_t& f, const std::string& name, const T* data, std::vector<uint64_t> dims)
{
    array_t a;
    a.dtype = dtype_of<T>::v;
    a.dims = dims;
    size_t n = a.count();
    a.bytes.resize(n * sizeof(T));
    if (n) std::memcpy(a.bytes.data(), data, n * sizeof(T));
    f[name] = std::move(a);
}

template <typename T>
inline void put(file_t& f, const std::string& name, const std::vector<T>& v)
{
    put<T>(f, name, v.data(), { (uint64_t)v.size() });
}

template <typename T>
inline void put(file_t& f, const std::string& name, const std::vector<T>& v, uint64_t cols)
{
    put<T>(f, name, v.data(), { (uint64_t)(cols ? v.size() / cols : 0), cols });
}

template <typename T> inline void put_scalar(file_t& f, const std::string& name, T v) { put<T>(f, name, &v, { 1 }); }

inline void write(const std::string& path, const file_t& f)
{
    FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp) throw std::runtime_error("mcb: cannot open for write: " + path);
    uint32_t n = (uint32_t)f.size();
    std::fwrite("MCB1", 1, 4, fp);
    std::fwrite(&n, 4, 1, fp);
    for (const auto& kv : f) {
        uint32_t nl = (uint32_t)kv.first.size();
        std::fwrite(&nl, 4, 1, fp);
        std::fwrite(kv.first.data(), 1, nl, fp);
        std::fwrite(&kv.second.dtype, 4, 1, fp);
        uint32_t nd = (uint32_t)kv.second.dims.size();
        std::fwrite(&nd, 4, 1, fp);
        if (nd) std::fwrite(kv.second.dims.data(), 8, nd, fp);
        if (!kv.second.bytes.empty()) std::fwrite(kv.second.bytes.data(), 1, kv.second.bytes.size(), fp);
    }
    std::fclose(fp);
}

inline file_t read(const std::string& path)
{
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("mcb: cannot open for read: " + path);
    auto rd = [&](void* p, size_t n) {
        if (n && std::fread(p, 1, n, fp) != n) {
            std::fclose(fp);
            throw std::runtime_error("mcb: truncated file: " + path);
        }
    };
    char magic[4];
    rd(magic, 4);
    if (std::memcmp(magic, "MCB1", 4) != 0) {
        std::fclose(fp);
        throw std::runtime_error("mcb: bad magic: " + path);
    }
    uint32_t n = 0;
    rd(&n, 4);
    file_t f;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t nl = 0;
        rd(&nl, 4);
        std::string name(nl, '\0');
        rd(&name[0], nl);
        array_t a;
        rd(&a.dtype, 4);
        uint32_t nd = 0;
        rd(&nd, 4);
        a.dims.resize(nd);
        rd(a.dims.data(), 8 * (size_t)nd);
        a.bytes.resize(a.count() * dtype_size(a.dtype));
        rd(a.bytes.data(), a.bytes.size());
        f[name] = std::move(a);
    }
    std::fclose(fp);
    return f;
}

} // namespace mcb
