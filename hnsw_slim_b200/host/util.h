// .fvecs / .ivecs I/O and timing for the host layer.  File format and console output follow the reference's
// include/util.h:52-80,149-168 (per row [int32 d][d x 4 bytes]; "open file error", "num:", "dim:"); the
// implementation does not: rows live in ONE contiguous block (the C ABI takes nq x dim row-major buffers, the
// reference keeps a vector per row), a file is read and written as one image instead of two stream calls per
// row, and every row header is checked against the first one (the reference trusts the file).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace hs_host {
struct FileCloser {
  void operator()(std::FILE *f) const {
    if (f) std::fclose(f);
  }
};
using File = std::unique_ptr<std::FILE, FileCloser>;

[[noreturn]] inline void die_open(const std::string &path) {      // util.h:57-60: message + exit(-1)
  std::cout << "open file error " << path << std::endl;
  std::exit(-1);
}
}  // namespace hs_host

template <typename T>
void ReadData(const std::string &file_path, std::vector<T> &results, uint32_t &num, uint32_t &dim) {
  static_assert(sizeof(T) == 4, "fvecs / ivecs hold 4-byte elements");
  hs_host::File f(std::fopen(file_path.c_str(), "rb"));
  if (!f) hs_host::die_open(file_path);
  std::fseek(f.get(), 0, SEEK_END);
  const long long fsize = std::ftell(f.get());
  std::rewind(f.get());
  uint32_t d = 0;
  if (fsize < 4 || std::fread(&d, 4, 1, f.get()) != 1 || d == 0 || d > (1u << 24)) hs_host::die_open(file_path);
  const size_t stride = (size_t)d + 1;                     // 4-byte words per row, header included
  const size_t rows = (size_t)fsize / 4 / stride;          // util.h:66: file size / (d + 1) / 4
  std::rewind(f.get());
  results.assign(rows * d, T{});
  // slabs of whole rows: read, check each header, drop it
  const size_t slab_rows = std::max<size_t>(1, (8u << 20) / (stride * 4));
  std::vector<uint32_t> slab(slab_rows * stride);
  for (size_t r0 = 0; r0 < rows; r0 += slab_rows) {
    const size_t cnt = std::min(slab_rows, rows - r0);
    if (std::fread(slab.data(), 4 * stride, cnt, f.get()) != cnt) hs_host::die_open(file_path);
    for (size_t r = 0; r < cnt; ++r) {
      const uint32_t *row = slab.data() + r * stride;
      if (row[0] != d) {
        std::cout << "open file error " << file_path << " (row " << r0 + r << " has dimension " << row[0] << ", not " << d
                  << ")" << std::endl;
        std::exit(-1);
      }
      std::memcpy(results.data() + (r0 + r) * d, row + 1, 4 * (size_t)d);
    }
  }
  num = (uint32_t)rows;
  dim = d;
  std::cout << "num: " << num << std::endl;
  std::cout << "dim: " << dim << std::endl;
}

template <typename T>
void WriteData(const std::string &file_path, const std::vector<T> &rows, uint32_t num, uint32_t dim) {
  static_assert(sizeof(T) == 4, "fvecs / ivecs hold 4-byte elements");
  hs_host::File f(std::fopen(file_path.c_str(), "wb"));
  if (!f) hs_host::die_open(file_path);
  const size_t stride = (size_t)dim + 1;
  const size_t slab_rows = std::max<size_t>(1, (8u << 20) / (stride * 4));
  std::vector<uint32_t> slab(slab_rows * stride);
  for (size_t r0 = 0; r0 < num; r0 += slab_rows) {
    const size_t cnt = std::min(slab_rows, (size_t)num - r0);
    for (size_t r = 0; r < cnt; ++r) {
      slab[r * stride] = dim;
      std::memcpy(&slab[r * stride + 1], rows.data() + (r0 + r) * dim, 4 * (size_t)dim);
    }
    if (std::fwrite(slab.data(), 4 * stride, cnt, f.get()) != cnt) hs_host::die_open(file_path);
  }
}

inline double time_cost(std::chrono::system_clock::time_point s, std::chrono::system_clock::time_point e) {
  return std::chrono::duration<double, std::milli>(e - s).count();
}
