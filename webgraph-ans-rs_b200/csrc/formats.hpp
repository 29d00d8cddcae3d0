// On-disk formats of the ANS graph: epserde 0.6.1 framing, sux 0.4.6 Elias-Fano, and the Java/webgraph
// BVGraph reader that feeds the recompression.  Byte layouts per SURVEY.md 8c (verified on the golden
// tests/data/cnr-2000/cnr-2000.ef of the reference).
#pragma once
#include <functional>

#include "common.hpp"

namespace wga {

std::vector<uint8_t> read_whole_file(const std::string& path);
void write_whole_file(const std::string& path, const std::vector<uint8_t>& bytes);

// ---- Elias-Fano as sux::dict::EliasFano<SelectAdaptConst<BitVec,_,12,4>, BitFieldVec> ---------------
struct EliasFano {
  uint64_t n = 0, u = 0, l = 0;
  std::vector<uint64_t> low;        // BitFieldVec words: n values of l bits (+1 padding word as sux does)
  std::vector<uint64_t> high;       // BitVec words
  uint64_t high_len = 0;            // bits
  std::vector<uint64_t> inventory;  // SelectAdaptConst inventory
  std::vector<uint64_t> spill;

  // EliasFanoBuilder::new(n,u) + push* + build + map_high_bits(SelectAdaptConst::<_,_,12,4>::new)
  // (src/bvgraph/random_access.rs:225-236)
  static EliasFano build(const uint64_t* values, uint64_t n, uint64_t u);
  uint64_t get(uint64_t i) const;             // IndexedSeq::get via the inventory (bvgraph_decoder_factory.rs:49)
  void expand(std::vector<uint64_t>& out) const;  // all values, linear scan
  std::vector<uint8_t> serialize() const;     // full epserde file
  static EliasFano deserialize(const std::vector<uint8_t>& bytes);
  uint64_t payload_bytes() const { return 8 * (low.size() + high.size() + inventory.size() + spill.size()); }
};

// ---- the three files (src/bvgraph/random_access.rs:52-82, 198-221) -----------------------------------
void load_prelude(const std::string& path, Prelude& out);
void store_prelude(const std::string& path, const Prelude& p);
void load_states(const std::string& path, std::vector<uint32_t>& out);
void store_states(const std::string& path, const std::vector<uint32_t>& s);

// ---- Java/webgraph BVGraph sequential reader (.graph + .properties; big-endian, default codes) -----
struct BvProperties {
  uint64_t nodes = 0, arcs = 0, window = 7, max_ref_count = 3, min_interval_length = 4, zetak = 3;
};
BvProperties read_bv_properties(const std::string& path);
// calls sink(node, successors) for every node in order; successors ascending
void read_bvgraph(const std::string& basename,
                  const std::function<void(uint64_t, const std::vector<uint64_t>&)>& sink,
                  BvProperties* props_out = nullptr);

}  // namespace wga
