// pgpu_case.hpp — case files of the compiled host driver: named int32 / int64 / float64 arrays in one flat file
// (format: piclas_b200/casefile.py), and the tables of pgpu_mesh_t / pgpu_params_t filled from them by field name.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "piclas_gpu.h"

namespace pgpu {

struct Array {
  char dtype = 'd';              // 'i' int32, 'l' int64, 'd' float64
  std::vector<int64_t> shape;
  std::vector<char> data;
  int64_t count() const { int64_t n = 1; for (int64_t s : shape) n *= s; return n; }
  template <class T> const T* as() const { return reinterpret_cast<const T*>(data.data()); }
  template <class T> T* as() { return reinterpret_cast<T*>(data.data()); }
};

inline size_t dtype_size(char c) {
  if (c == 'i') return 4;
  if (c == 'l' || c == 'd') return 8;
  throw std::runtime_error(std::string("case file: unknown dtype '") + c + "'");
}

class CaseFile {
 public:
  std::map<std::string, Array> arrays;
  std::vector<std::string> order;

  static CaseFile read(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    char magic[10];
    f.read(magic, 10);
    if (!f || std::memcmp(magic, "PGPUCASE1\n", 10) != 0) throw std::runtime_error(path + " is not a PGPUCASE1 file");
    CaseFile c;
    for (;;) {
      uint32_t nlen;
      f.read(reinterpret_cast<char*>(&nlen), 4);
      if (!f) break;
      std::string name(nlen, '\0');
      f.read(&name[0], nlen);
      Array a;
      uint32_t nd;
      f.read(&a.dtype, 1);
      f.read(reinterpret_cast<char*>(&nd), 4);
      a.shape.resize(nd);
      f.read(reinterpret_cast<char*>(a.shape.data()), 8 * nd);
      a.data.resize((size_t)a.count() * dtype_size(a.dtype));
      f.read(a.data.data(), (std::streamsize)a.data.size());
      if (!f) throw std::runtime_error(path + ": truncated record " + name);
      c.order.push_back(name);
      c.arrays.emplace(name, std::move(a));
    }
    return c;
  }

  void write(const std::string& path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    f.write("PGPUCASE1\n", 10);
    for (const std::string& name : order) {
      const Array& a = arrays.at(name);
      const uint32_t nlen = (uint32_t)name.size(), nd = (uint32_t)a.shape.size();
      f.write(reinterpret_cast<const char*>(&nlen), 4);
      f.write(name.data(), nlen);
      f.write(&a.dtype, 1);
      f.write(reinterpret_cast<const char*>(&nd), 4);
      f.write(reinterpret_cast<const char*>(a.shape.data()), 8 * nd);
      f.write(a.data.data(), (std::streamsize)a.data.size());
    }
  }

  bool has(const std::string& name) const { return arrays.count(name) != 0; }
  const Array& get(const std::string& name, char dtype) const {
    auto it = arrays.find(name);
    if (it == arrays.end()) throw std::runtime_error("case file: missing array " + name);
    if (it->second.dtype != dtype) throw std::runtime_error("case file: array " + name + " has the wrong type");
    return it->second;
  }
  template <class T> void put(const std::string& name, char dtype, std::vector<int64_t> shape, const T* src) {
    Array a;
    a.dtype = dtype;
    a.shape = std::move(shape);
    a.data.resize((size_t)a.count() * sizeof(T));
    std::memcpy(a.data.data(), src, a.data.size());
    if (!arrays.count(name)) order.push_back(name);
    arrays[name] = std::move(a);
  }
};

// The structs point into the arrays of the case file, which must outlive them (as the Fortran host's module arrays do).
inline void fill_mesh(const CaseFile& c, pgpu_mesh_t& m) {
  std::memset(&m, 0, sizeof m);
  const std::string p = "mesh.";
#define PGPU_SCALAR_I32(f) m.f = c.get(p + #f, 'i').as<int32_t>()[0];
#define PGPU_SCALAR_I64(f) m.f = c.get(p + #f, 'l').as<int64_t>()[0];
#define PGPU_SCALAR_F64(f) m.f = c.get(p + #f, 'd').as<double>()[0];
#define PGPU_ARRAY_I32(f, n) std::memcpy(m.f, c.get(p + #f, 'i').as<int32_t>(), sizeof(int32_t) * (n));
#define PGPU_ARRAY_F64(f, n) std::memcpy(m.f, c.get(p + #f, 'd').as<double>(), sizeof(double) * (n));
#define PGPU_PTR_I32(f) m.f = c.has(p + #f) ? c.get(p + #f, 'i').as<int32_t>() : nullptr;
#define PGPU_PTR_F64(f) m.f = c.has(p + #f) ? c.get(p + #f, 'd').as<double>() : nullptr;
#define PGPU_MESH_FIELDS
#include "pgpu_fields.inc"
#undef PGPU_MESH_FIELDS
}

inline void fill_params(const CaseFile& c, pgpu_params_t& m) {
  std::memset(&m, 0, sizeof m);
  const std::string p = "params.";
#define PGPU_PARAMS_FIELDS
#include "pgpu_fields.inc"
#undef PGPU_PARAMS_FIELDS
#undef PGPU_SCALAR_I32
#undef PGPU_SCALAR_I64
#undef PGPU_SCALAR_F64
#undef PGPU_ARRAY_I32
#undef PGPU_ARRAY_F64
#undef PGPU_PTR_I32
#undef PGPU_PTR_F64
}

}  // namespace pgpu
