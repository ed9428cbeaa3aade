// CPU: llkv::walk_descriptor of the C++ host mirror over an in-memory "pager" (key -> blob) read from stdin.
// Input: lines "<key> <hex blob>", then "walk <descriptor_pk> <prim_type> <lower_kind> <lower_bits> <upper_kind> <upper_bits>".
// Output: the descriptor's totals and the surviving chunk pks.
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "../../rust-llkv_b200/host/llkv_gpu.hpp"

static std::string unhex(const std::string& h) {
  std::string out;
  for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back((char)std::stoi(h.substr(i, 2), nullptr, 16));
  return out;
}

int main() {
  std::map<uint64_t, std::string> pager;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream in(line);
    std::string first;
    in >> first;
    if (first == "walk") {
      uint64_t pk, lo_bits, hi_bits;
      int prim, lo_kind, hi_kind;
      in >> pk >> prim >> lo_kind >> lo_bits >> hi_kind >> hi_bits;
      llkv_range_bound lo{lo_kind, 0, lo_bits}, hi{hi_kind, 0, hi_bits};
      llkv_column_descriptor desc;
      try {
        auto metas = llkv::walk_descriptor([&](uint64_t key) -> const std::string& { return pager.at(key); }, pk, &desc, prim, &lo, &hi);
        std::printf("desc %llu %llu %u", (unsigned long long)desc.total_row_count, (unsigned long long)desc.total_chunk_count, desc.data_type_code);
        for (const auto& m : metas) std::printf(" %llu", (unsigned long long)m.chunk_pk);
        std::printf("\n");
      } catch (const std::exception& e) {
        std::printf("error %s\n", e.what());
      }
    } else {
      std::string hex;
      in >> hex;
      pager[std::stoull(first)] = unhex(hex);
    }
  }
  return 0;
}
