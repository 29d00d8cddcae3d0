// =============================================================================
//  wgo_io.hpp -- CPU ORACLE file readers.   TEST INFRASTRUCTURE ONLY (see wgo.hpp).
// =============================================================================
//  * Java/webgraph BVGraph `.graph` + `.properties` sequential reader (big-endian,
//    MSB-first bit order; gamma / unary / zeta_k codes).  The reference obtains this
//    from webgraph-rs + dsi-bitstream 0.4.0 (src/bvgraph/random_access.rs:101-103);
//    both are un-vendored, so the published code definitions are restated and
//    pinned by the golden tests/data/cnr-2000 files (arcs, residualarcs, .ef offsets).
//  * epserde 0.6.1 readers for `.ans` (Prelude, src/ans/mod.rs:31-54), `.states`
//    (Box<[u32]>, src/bvgraph/random_access.rs:202-204) and `.pointers`
//    (sux 0.4.6 Elias-Fano, src/bvgraph/factories/mod.rs:6-9).  Layout per
//    SURVEY.md 8c, [GOLD]-verified on cnr-2000.ef.  The oracle expands the
//    Elias-Fano by a linear scan of the high bits (no inventory use).
// =============================================================================
#pragma once
#include <fstream>
#include <sstream>

#include "wgo.hpp"

namespace wgo {

inline std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  f.seekg(0, std::ios::end);
  size_t n = (size_t)f.tellg();
  f.seekg(0);
  std::vector<uint8_t> b(n);
  f.read((char*)b.data(), (std::streamsize)n);
  return b;
}

// --------------------------------------------------------------------------- BV .graph reader
struct BitReaderBE {
  const uint8_t* p;
  size_t nbits;
  size_t pos = 0;
  BitReaderBE(const uint8_t* p, size_t nbytes) : p(p), nbits(nbytes * 8) {}
  unsigned bit() {
    if (pos >= nbits) throw std::out_of_range("bitstream EOF");
    unsigned b = (p[pos >> 3] >> (7 - (pos & 7))) & 1;
    pos++;
    return b;
  }
  uint64_t bits(unsigned n) {
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i) v = (v << 1) | bit();
    return v;
  }
  uint64_t unary() {
    uint64_t c = 0;
    while (!bit()) c++;
    return c;
  }
  uint64_t gamma() {
    unsigned l = (unsigned)unary();
    return ((1ull << l) | bits(l)) - 1;
  }
  uint64_t minimal_binary(uint64_t max) {
    unsigned l = ilog2(max);
    uint64_t lim = (1ull << (l + 1)) - max;
    uint64_t v = bits(l);
    if (v < lim) return v;
    return ((v << 1) | bit()) - lim;
  }
  uint64_t zeta(unsigned k) {
    uint64_t h = unary();
    uint64_t left = 1ull << (h * k);
    return minimal_binary((1ull << ((h + 1) * k)) - left) + left - 1;
  }
};

struct BVProperties {
  size_t nodes = 0, window = 7, max_ref_count = 3, min_interval_length = 4, zetak = 3;
  uint64_t arcs = 0;
  std::string compressionflags;
};

inline BVProperties read_properties(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  BVProperties p;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#') continue;
    auto eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string k = line.substr(0, eq), v = line.substr(eq + 1);
    while (!v.empty() && (v.back() == '\r' || v.back() == ' ')) v.pop_back();
    if (k == "nodes") p.nodes = std::stoull(v);
    else if (k == "arcs") p.arcs = std::stoull(v);
    else if (k == "windowsize") p.window = std::stoull(v);
    else if (k == "maxrefcount") p.max_ref_count = std::stoull(v);
    else if (k == "minintervallength") p.min_interval_length = std::stoull(v);
    else if (k == "zetak") p.zetak = std::stoull(v);
    else if (k == "compressionflags") p.compressionflags = v;
  }
  if (!p.compressionflags.empty())
    throw std::runtime_error("oracle BV reader supports default compression flags only");
  return p;
}

// CSR of a BV graph; also returns the bit offset of each record (to check against the golden .ef).
struct CSR {
  std::vector<uint64_t> offsets;  // n+1
  std::vector<uint64_t> succ;
  std::vector<uint64_t> bit_offsets;  // n+1
};

inline CSR read_bvgraph(const std::string& basename) {
  BVProperties pr = read_properties(basename + ".properties");
  std::vector<uint8_t> data = read_file(basename + ".graph");
  BitReaderBE br(data.data(), data.size());
  CSR g;
  g.offsets.assign(1, 0);
  const size_t w = pr.window;
  std::vector<std::vector<uint64_t>> back(w + 1);
  std::vector<uint64_t> out;
  for (size_t v = 0; v < pr.nodes; ++v) {
    g.bit_offsets.push_back(br.pos);
    out.clear();
    uint64_t d = br.gamma();
    if (d != 0) {
      uint64_t r = w != 0 ? br.unary() : 0;
      if (r != 0) {
        const auto& nb = back[(v - r) % (w + 1)];
        uint64_t nblocks = br.gamma();
        if (nblocks == 0) out = nb;
        else {
          uint64_t idx = br.gamma();
          out.insert(out.end(), nb.begin(), nb.begin() + idx);
          for (uint64_t b = 1; b < nblocks; ++b) {
            uint64_t end = idx + br.gamma() + 1;
            if (b % 2 == 0) out.insert(out.end(), nb.begin() + idx, nb.begin() + end);
            idx = end;
          }
          if ((nblocks & 1) == 0) out.insert(out.end(), nb.begin() + idx, nb.end());
        }
      }
      uint64_t left = d - out.size();
      if (left != 0 && pr.min_interval_length != 0) {
        uint64_t ni = br.gamma();
        if (ni != 0) {
          int64_t start = (int64_t)v + nat2int(br.gamma());
          uint64_t len = br.gamma() + pr.min_interval_length;
          for (uint64_t i = 0; i < len; ++i) out.push_back((uint64_t)start + i);
          start += (int64_t)len;
          for (uint64_t k = 1; k < ni; ++k) {
            start += 1 + (int64_t)br.gamma();
            len = br.gamma() + pr.min_interval_length;
            for (uint64_t i = 0; i < len; ++i) out.push_back((uint64_t)start + i);
            start += (int64_t)len;
          }
        }
      }
      left = d - out.size();
      if (left != 0) {
        uint64_t prev = (uint64_t)((int64_t)v + nat2int(br.zeta((unsigned)pr.zetak)));
        out.push_back(prev);
        for (uint64_t k = 1; k < left; ++k) {
          prev += 1 + br.zeta((unsigned)pr.zetak);
          out.push_back(prev);
        }
      }
      std::sort(out.begin(), out.end());
    }
    g.succ.insert(g.succ.end(), out.begin(), out.end());
    g.offsets.push_back(g.succ.size());
    back[v % (w + 1)] = out;
  }
  g.bit_offsets.push_back(br.pos);
  return g;
}

// --------------------------------------------------------------------------- epserde readers
struct ByteCursor {
  const std::vector<uint8_t>& b;
  size_t off = 0;
  explicit ByteCursor(const std::vector<uint8_t>& b) : b(b) {}
  template <class T>
  T get() {
    if (off + sizeof(T) > b.size()) throw std::out_of_range("epserde: truncated file");
    T v;
    std::memcpy(&v, b.data() + off, sizeof(T));
    off += sizeof(T);
    return v;
  }
  void align(size_t a) { off += (a - off % a) % a; }
  std::string header() {  // "epserde " | u16 major | u16 minor | u8 usize | u64 type_hash | u64 repr_hash | name
    if (b.size() < 37 || std::memcmp(b.data(), "epserde ", 8) != 0) throw std::runtime_error("epserde: bad magic");
    off = 8;
    get<uint16_t>(); get<uint16_t>();
    if (get<uint8_t>() != 8) throw std::runtime_error("epserde: usize != 8");
    get<uint64_t>(); get<uint64_t>();
    uint64_t n = get<uint64_t>();
    std::string name((const char*)b.data() + off, (size_t)n);
    off += (size_t)n;
    return name;
  }
  template <class T>
  std::vector<T> zero_copy_vec() {  // len | pad to align_of<T> | raw
    uint64_t n = get<uint64_t>();
    align(alignof(T));
    if (off + n * sizeof(T) > b.size()) throw std::out_of_range("epserde: truncated vector");
    std::vector<T> v((size_t)n);
    std::memcpy(v.data(), b.data() + off, (size_t)n * sizeof(T));
    off += (size_t)n * sizeof(T);
    return v;
  }
};

inline void load_prelude(const std::string& path, ANSGraph& g) {  // src/ans/mod.rs:31-54
  std::vector<uint8_t> bytes = read_file(path);
  ByteCursor c(bytes);
  c.header();
  uint64_t nt = c.get<uint64_t>();
  if (nt != COMPONENTS) throw std::runtime_error(".ans: expected 9 tables");
  for (int i = 0; i < COMPONENTS; ++i) {  // component_model4encoder.rs:37-57 field order
    g.tables[i].table = c.zero_copy_vec<EncoderModelEntry>();
    g.tables[i].frame_size = c.get<uint64_t>();
    g.tables[i].radix = c.get<uint64_t>();
    g.tables[i].fidelity = c.get<uint64_t>();
    g.tables[i].folding_threshold = c.get<uint64_t>();
    g.tables[i].folding_offset = c.get<uint64_t>();
  }
  g.stream = c.zero_copy_vec<uint16_t>();
  g.state = c.get<uint32_t>();
  g.number_of_nodes = c.get<uint64_t>();
  g.compression_window = c.get<uint64_t>();
  g.min_interval_length = c.get<uint64_t>();
  g.number_of_arcs = c.get<uint64_t>();
}

inline void load_states(const std::string& path, ANSGraph& g) {
  std::vector<uint8_t> bytes = read_file(path);
  ByteCursor c(bytes);
  c.header();
  g.states = c.zero_copy_vec<uint32_t>();
}

// Elias-Fano -> plain values, by scanning the high bits (value i = ((select1(i)-i) << l) | low[i]).
inline std::vector<uint64_t> load_elias_fano(const std::string& path) {
  std::vector<uint8_t> bytes = read_file(path);
  ByteCursor c(bytes);
  c.header();
  uint64_t n = c.get<uint64_t>();
  c.get<uint64_t>();  // u
  uint64_t l = c.get<uint64_t>();
  std::vector<uint64_t> low = c.zero_copy_vec<uint64_t>();
  uint64_t bit_width = c.get<uint64_t>();
  c.get<uint64_t>();  // mask
  c.get<uint64_t>();  // len
  if (bit_width != l) throw std::runtime_error("EF: bit_width != l");
  std::vector<uint64_t> high = c.zero_copy_vec<uint64_t>();
  c.get<uint64_t>();  // bit length
  std::vector<uint64_t> out;
  out.reserve((size_t)n);
  uint64_t i = 0;
  for (size_t wi = 0; wi < high.size() && i < n; ++wi) {
    uint64_t w = high[wi];
    while (w && i < n) {
      unsigned b = (unsigned)__builtin_ctzll(w);
      uint64_t pos = (uint64_t)wi * 64 + b;
      uint64_t hi = pos - i;
      uint64_t lo = 0;
      if (l) {
        uint64_t bp = i * l;
        size_t wd = (size_t)(bp >> 6);
        unsigned sh = (unsigned)(bp & 63);
        lo = low[wd] >> sh;
        if (sh + l > 64) lo |= low[wd + 1] << (64 - sh);
        lo &= (1ull << l) - 1;
      }
      out.push_back((hi << l) | lo);
      ++i;
      w &= w - 1;
    }
  }
  if (i != n) throw std::runtime_error("EF: fewer ones than n");
  return out;
}

inline ANSGraph load_ans_graph(const std::string& basename, bool random_access) {
  ANSGraph g;
  load_prelude(basename + ".ans", g);
  if (random_access) {
    load_states(basename + ".states", g);
    g.pointers = load_elias_fano(basename + ".pointers");
  }
  return g;
}

}  // namespace wgo
