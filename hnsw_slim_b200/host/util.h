// .fvecs / .ivecs I/O and timing, mirroring the reference's include/util.h:52-168.
// Rows are kept in ONE contiguous block (the C ABI takes nq x dim row-major buffers); the
// reference keeps a vector per row.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

// util.h:52-80: per row [int32 d][d x 4 bytes]; row count = file size / (d + 1) / 4.
// Like the reference, an unreadable file ends the process (exit(-1), util.h:57-60).
template <typename T>
void ReadData(const std::string &file_path, std::vector<T> &results, uint32_t &num, uint32_t &dim) {
  static_assert(sizeof(T) == 4, "fvecs / ivecs hold 4-byte elements");
  std::ifstream in(file_path, std::ios::binary);
  if (!in.is_open()) {
    std::cout << "open file error " << file_path << std::endl;
    exit(-1);
  }
  in.read((char *)&dim, 4);
  in.seekg(0, std::ios::end);
  const size_t fsize = (size_t)in.tellg();
  num = (uint32_t)(fsize / (dim + 1) / 4);
  results.resize((size_t)num * dim);
  in.seekg(0, std::ios::beg);
  for (size_t i = 0; i < num; ++i) {
    in.seekg(4, std::ios::cur);
    in.read((char *)(results.data() + i * dim), dim * 4);
  }
  std::cout << "num: " << num << std::endl;
  std::cout << "dim: " << dim << std::endl;
}

// util.h:149-168
template <typename T>
void WriteData(const std::string &file_path, const std::vector<T> &rows, uint32_t num, uint32_t dim) {
  std::ofstream out(file_path, std::ios::binary);
  if (!out.is_open()) {
    std::cout << "open file error " << file_path << std::endl;
    exit(-1);
  }
  for (size_t i = 0; i < num; ++i) {
    out.write((const char *)&dim, 4);
    out.write((const char *)(rows.data() + i * dim), dim * 4);
  }
}

inline double time_cost(std::chrono::system_clock::time_point s, std::chrono::system_clock::time_point e) {
  return std::chrono::duration<double, std::milli>(e - s).count();
}
