// On-disk formats: epserde framing, Elias-Fano (.pointers), .ans, .states, and the BVGraph reader.
#include "formats.hpp"

#include <algorithm>
#include <fstream>

namespace wga {

std::vector<uint8_t> read_whole_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(WGA_E_IO, "cannot open " + path);
  f.seekg(0, std::ios::end);
  std::streamoff n = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> b((size_t)n);
  if (n) f.read((char*)b.data(), n);
  if (!f) throw Error(WGA_E_IO, "cannot read " + path);
  return b;
}

void write_whole_file(const std::string& path, const std::vector<uint8_t>& bytes) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw Error(WGA_E_IO, "Could not create " + path);
  f.write((const char*)bytes.data(), (std::streamsize)bytes.size());
  if (!f) throw Error(WGA_E_IO, "cannot write " + path);
}

namespace {

// ---------------------------------------------------------------------------------------- epserde
// "epserde " | u16 major=1 | u16 minor=1 | u8 sizeof(usize)=8 | u64 type_hash | u64 repr_hash |
// usize len + type name | value.  Everything little-endian and unaligned except zero-copy vectors,
// whose data is padded to align_of::<T>() relative to the start of the file.
struct Out {
  std::vector<uint8_t> b;
  template <class T>
  void put(T v) {
    size_t o = b.size();
    b.resize(o + sizeof(T));
    std::memcpy(b.data() + o, &v, sizeof(T));
  }
  void align(size_t a) {
    while (b.size() % a) b.push_back(0);
  }
  template <class T>
  void vec(const T* p, uint64_t n, size_t alignment) {
    put<uint64_t>(n);
    align(alignment);
    size_t o = b.size();
    b.resize(o + n * sizeof(T));
    if (n) std::memcpy(b.data() + o, p, n * sizeof(T));
  }
  void header(uint64_t type_hash, uint64_t repr_hash, const std::string& name) {
    b.insert(b.end(), {'e', 'p', 's', 'e', 'r', 'd', 'e', ' '});
    put<uint16_t>(1);
    put<uint16_t>(1);
    put<uint8_t>(8);
    put<uint64_t>(type_hash);
    put<uint64_t>(repr_hash);
    put<uint64_t>(name.size());
    b.insert(b.end(), name.begin(), name.end());
  }
};

struct In {
  const std::vector<uint8_t>& b;
  size_t off = 0;
  explicit In(const std::vector<uint8_t>& b) : b(b) {}
  template <class T>
  T get() {
    if (off + sizeof(T) > b.size()) throw Error(WGA_E_FORMAT, "epserde: truncated file");
    T v;
    std::memcpy(&v, b.data() + off, sizeof(T));
    off += sizeof(T);
    return v;
  }
  void align(size_t a) { off += (a - off % a) % a; }
  std::string header() {
    if (b.size() < 37 || std::memcmp(b.data(), "epserde ", 8) != 0) throw Error(WGA_E_FORMAT, "epserde: bad magic");
    off = 8;
    uint16_t major = get<uint16_t>();
    get<uint16_t>();
    if (major != 1) throw Error(WGA_E_FORMAT, "epserde: unsupported major version");
    if (get<uint8_t>() != 8) throw Error(WGA_E_FORMAT, "epserde: file written with a non-64-bit usize");
    get<uint64_t>();  // type hash: not checked (unpinned for Prelude / Box<[u32]>)
    get<uint64_t>();  // repr hash
    uint64_t n = get<uint64_t>();
    if (off + n > b.size()) throw Error(WGA_E_FORMAT, "epserde: truncated type name");
    std::string name((const char*)b.data() + off, (size_t)n);
    off += (size_t)n;
    return name;
  }
  template <class T>
  void vec(std::vector<T>& out, size_t alignment) {
    uint64_t n = get<uint64_t>();
    align(alignment);
    if (n > (b.size() - off) / sizeof(T)) throw Error(WGA_E_FORMAT, "epserde: truncated vector");
    out.resize((size_t)n);
    if (n) std::memcpy(out.data(), b.data() + off, (size_t)n * sizeof(T));
    off += (size_t)n * sizeof(T);
  }
};

// Type identification of the Elias-Fano alias (src/bvgraph/factories/mod.rs:6-9).  These 16 hash bytes
// and the name are those of the golden tests/data/cnr-2000/cnr-2000.ef, which is the same Rust type.
const uint64_t EF_TYPE_HASH = 0x890ce77a9258940cull;
const uint64_t EF_REPR_HASH = 0xf27f4cf54b9dc82cull;
const char* EF_TYPE_NAME =
    "sux::dict::elias_fano::EliasFano<sux::rank_sel::select_adapt_const::SelectAdaptConst<"
    "sux::bits::bit_vec::BitVec<alloc::boxed::Box<[usize]>>, alloc::boxed::Box<[usize]>, 12, 4>>";
// PARITY UNPINNED: no reference-written .ans/.states exists to read these hashes from; the real
// reference would refuse files whose hashes differ.  Our own loader does not check them.
const uint64_t UNPINNED_HASH = 0;
const char* PRELUDE_TYPE_NAME = "webgraph_ans::ans::Prelude";
const char* STATES_TYPE_NAME = "alloc::boxed::Box<[u32]>";

constexpr unsigned LOG2_ONES_PER_INVENTORY = 12;
constexpr unsigned LOG2_U64_PER_SUBINVENTORY = 4;
constexpr uint64_t ONES_PER_INVENTORY = 1ull << LOG2_ONES_PER_INVENTORY;
constexpr uint64_t U64_PER_SUB = 1ull << LOG2_U64_PER_SUBINVENTORY;
constexpr uint64_t ONES_PER_SUB16 = ONES_PER_INVENTORY / (U64_PER_SUB * 4);  // 64
constexpr uint64_t ONES_PER_SUB32 = ONES_PER_INVENTORY / (U64_PER_SUB * 2);  // 128
constexpr uint64_t INV_FLAG_U32 = 1ull << 63;  // [MEM] span >= 2^16: u32 sub-inventory (unpinned)
constexpr uint64_t INV_FLAG_U64 = 3ull << 62;  // [MEM] span >= 2^32: exact positions in the spill (unpinned)

}  // namespace

// -------------------------------------------------------------------------------------- Elias-Fano
EliasFano EliasFano::build(const uint64_t* values, uint64_t n, uint64_t u) {
  EliasFano ef;
  ef.n = n;
  ef.u = u;
  ef.l = (n && u >= n) ? (uint64_t)(63 - __builtin_clzll(u / n)) : 0;  // EliasFanoBuilder::new
  const uint64_t l = ef.l;
  ef.low.assign((n * l + 63) / 64, 0);
  ef.high_len = n + (u >> l) + 1;
  ef.high.assign((ef.high_len + 63) / 64, 0);
  uint64_t prev = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = values[i];
    if (v < prev) throw Error(WGA_E_ARG, "Elias-Fano: values must be non-decreasing");
    if (v >= u && !(v == 0 && u == 0)) throw Error(WGA_E_ARG, "Elias-Fano: value exceeds the upper bound");
    prev = v;
    if (l) {
      uint64_t lo = v & ((1ull << l) - 1);
      uint64_t bp = i * l;
      ef.low[bp >> 6] |= lo << (bp & 63);
      if ((bp & 63) + l > 64) ef.low[(bp >> 6) + 1] |= lo >> (64 - (bp & 63));
    }
    uint64_t hp = (v >> l) + i;
    ef.high[hp >> 6] |= 1ull << (hp & 63);
  }
  // SelectAdaptConst<_,_,12,4>::new : every 4096th one -> [u64 position][16 u64 sub-inventory];
  // trailing u64 = bit-vector length.  Sub-inventory = u16 offsets of every 64th one when the block
  // spans < 2^16 bits ([GOLD] on cnr-2000.ef).
  const uint64_t blocks = (n + ONES_PER_INVENTORY - 1) / ONES_PER_INVENTORY;
  ef.inventory.assign(blocks * (1 + U64_PER_SUB) + 1, 0);
  std::vector<uint64_t> blockpos;  // positions of the ones of the current block
  uint64_t one = 0;
  auto flush_block = [&](uint64_t k, uint64_t next_start) {
    if (blockpos.empty()) return;
    uint64_t base = blockpos[0];
    uint64_t span = next_start - base;
    uint64_t* inv = &ef.inventory[k * (1 + U64_PER_SUB)];
    inv[0] = base;
    if (span < (1ull << 16)) {
      for (uint64_t j = 0; j * ONES_PER_SUB16 < blockpos.size(); ++j) {
        uint64_t off = blockpos[j * ONES_PER_SUB16] - base;
        inv[1 + j / 4] |= off << (16 * (j % 4));
      }
    } else if (span < (1ull << 32)) {
      inv[0] |= INV_FLAG_U32;
      for (uint64_t j = 0; j * ONES_PER_SUB32 < blockpos.size(); ++j) {
        uint64_t off = blockpos[j * ONES_PER_SUB32] - base;
        inv[1 + j / 2] |= off << (32 * (j % 2));
      }
    } else {
      inv[0] |= INV_FLAG_U64;
      inv[1] = ef.spill.size();
      for (uint64_t p : blockpos) ef.spill.push_back(p);
    }
    blockpos.clear();
  };
  uint64_t k = 0;
  for (size_t wi = 0; wi < ef.high.size(); ++wi) {
    uint64_t w = ef.high[wi];
    while (w) {
      uint64_t pos = (uint64_t)wi * 64 + (uint64_t)__builtin_ctzll(w);
      if (one && one % ONES_PER_INVENTORY == 0) flush_block(k++, pos);
      blockpos.push_back(pos);
      ++one;
      w &= w - 1;
    }
  }
  flush_block(k, ef.high_len);
  ef.inventory.back() = ef.high_len;
  return ef;
}

uint64_t EliasFano::get(uint64_t i) const {
  if (i >= n) throw Error(WGA_E_ARG, "Elias-Fano: index out of range");
  // select1(i) through the inventory
  const uint64_t k = i >> LOG2_ONES_PER_INVENTORY;
  const uint64_t* inv = &inventory[k * (1 + U64_PER_SUB)];
  const uint64_t flags = inv[0] & INV_FLAG_U64;
  uint64_t pos = inv[0] & ~INV_FLAG_U64;
  uint64_t within = i & (ONES_PER_INVENTORY - 1);
  uint64_t residual;
  if (flags == 0) {
    uint64_t j = within / ONES_PER_SUB16;
    pos += (inv[1 + j / 4] >> (16 * (j % 4))) & 0xFFFF;
    residual = within % ONES_PER_SUB16;
  } else if (flags == INV_FLAG_U32) {
    uint64_t j = within / ONES_PER_SUB32;
    pos += (inv[1 + j / 2] >> (32 * (j % 2))) & 0xFFFFFFFFull;
    residual = within % ONES_PER_SUB32;
  } else {
    pos = spill[inv[1] + within];
    residual = 0;
  }
  // scan forward `residual` ones from pos (pos itself is a one)
  size_t wi = (size_t)(pos >> 6);
  uint64_t w = high[wi] & (~0ull << (pos & 63));
  while (true) {
    uint64_t c = (uint64_t)__builtin_popcountll(w);
    if (residual < c) break;
    residual -= c;
    w = high[++wi];
  }
  for (uint64_t r = 0; r < residual; ++r) w &= w - 1;
  uint64_t sel = (uint64_t)wi * 64 + (uint64_t)__builtin_ctzll(w);
  uint64_t lo = 0;
  if (l) {
    uint64_t bp = i * l;
    lo = low[bp >> 6] >> (bp & 63);
    if ((bp & 63) + l > 64) lo |= low[(bp >> 6) + 1] << (64 - (bp & 63));
    lo &= (1ull << l) - 1;
  }
  return ((sel - i) << l) | lo;
}

void EliasFano::expand(std::vector<uint64_t>& out) const {
  out.resize(n);
  uint64_t i = 0;
  for (size_t wi = 0; wi < high.size() && i < n; ++wi) {
    uint64_t w = high[wi];
    while (w && i < n) {
      uint64_t pos = (uint64_t)wi * 64 + (uint64_t)__builtin_ctzll(w);
      uint64_t lo = 0;
      if (l) {
        uint64_t bp = i * l;
        lo = low[bp >> 6] >> (bp & 63);
        if ((bp & 63) + l > 64) lo |= low[(bp >> 6) + 1] << (64 - (bp & 63));
        lo &= (1ull << l) - 1;
      }
      out[i] = ((pos - i) << l) | lo;
      ++i;
      w &= w - 1;
    }
  }
  if (i != n) throw Error(WGA_E_FORMAT, "Elias-Fano: high bits hold fewer ones than n");
}

std::vector<uint8_t> EliasFano::serialize() const {
  Out o;
  o.header(EF_TYPE_HASH, EF_REPR_HASH, EF_TYPE_NAME);
  o.put<uint64_t>(n);
  o.put<uint64_t>(u);
  o.put<uint64_t>(l);
  o.vec(low.data(), low.size(), 8);  // BitFieldVec { bits, bit_width, mask, len }
  o.put<uint64_t>(l);
  o.put<uint64_t>(l ? (1ull << l) - 1 : 0);
  o.put<uint64_t>(n);
  o.vec(high.data(), high.size(), 8);  // SelectAdaptConst { bits: BitVec { bits, len }, inventory, spill }
  o.put<uint64_t>(high_len);
  o.vec(inventory.data(), inventory.size(), 8);
  o.vec(spill.data(), spill.size(), 8);
  return o.b;
}

EliasFano EliasFano::deserialize(const std::vector<uint8_t>& bytes) {
  In in(bytes);
  in.header();
  EliasFano ef;
  ef.n = in.get<uint64_t>();
  ef.u = in.get<uint64_t>();
  ef.l = in.get<uint64_t>();
  in.vec(ef.low, 8);
  uint64_t bw = in.get<uint64_t>();
  in.get<uint64_t>();
  uint64_t len = in.get<uint64_t>();
  if (bw != ef.l || len != ef.n || ef.l > 63) throw Error(WGA_E_FORMAT, "Elias-Fano: inconsistent low-bits vector");
  in.vec(ef.high, 8);
  ef.high_len = in.get<uint64_t>();
  in.vec(ef.inventory, 8);
  in.vec(ef.spill, 8);
  if (ef.low.size() < (ef.n * ef.l + 63) / 64 || ef.high.size() < (ef.high_len + 63) / 64)
    throw Error(WGA_E_FORMAT, "Elias-Fano: truncated bit vectors");
  return ef;
}

// ------------------------------------------------------------------------------------ .ans / .states
void load_prelude(const std::string& path, Prelude& p) {
  std::vector<uint8_t> bytes = read_whole_file(path);
  In in(bytes);
  in.header();
  if (in.get<uint64_t>() != WGA_COMPONENTS) throw Error(WGA_E_FORMAT, ".ans: expected 9 component tables");
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    ComponentModel& t = p.tables[c];
    in.vec(t.table, 4);
    t.frame_size = in.get<uint64_t>();
    t.radix = in.get<uint64_t>();
    t.fidelity = in.get<uint64_t>();
    t.folding_threshold = in.get<uint64_t>();
    t.folding_offset = in.get<uint64_t>();
  }
  in.vec(p.stream, 2);
  p.state = in.get<uint32_t>();
  p.number_of_nodes = in.get<uint64_t>();
  p.compression_window = in.get<uint64_t>();
  p.min_interval_length = in.get<uint64_t>();
  p.number_of_arcs = in.get<uint64_t>();
}

void store_prelude(const std::string& path, const Prelude& p) {
  Out o;
  o.header(UNPINNED_HASH, UNPINNED_HASH, PRELUDE_TYPE_NAME);
  o.put<uint64_t>(WGA_COMPONENTS);
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    const ComponentModel& t = p.tables[c];
    o.vec(t.table.data(), t.table.size(), 4);
    o.put<uint64_t>(t.frame_size);
    o.put<uint64_t>(t.radix);
    o.put<uint64_t>(t.fidelity);
    o.put<uint64_t>(t.folding_threshold);
    o.put<uint64_t>(t.folding_offset);
  }
  o.vec(p.stream.data(), p.stream.size(), 2);
  o.put<uint32_t>(p.state);
  o.put<uint64_t>(p.number_of_nodes);
  o.put<uint64_t>(p.compression_window);
  o.put<uint64_t>(p.min_interval_length);
  o.put<uint64_t>(p.number_of_arcs);
  write_whole_file(path, o.b);
}

void load_states(const std::string& path, std::vector<uint32_t>& out) {
  std::vector<uint8_t> bytes = read_whole_file(path);
  In in(bytes);
  in.header();
  in.vec(out, 4);
}

void store_states(const std::string& path, const std::vector<uint32_t>& s) {
  Out o;
  o.header(UNPINNED_HASH, UNPINNED_HASH, STATES_TYPE_NAME);
  o.vec(s.data(), s.size(), 4);
  write_whole_file(path, o.b);
}

// ---------------------------------------------------------------------------------- BVGraph reader
BvProperties read_bv_properties(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error(WGA_E_IO, "cannot open " + path);
  BvProperties p;
  std::string line;
  bool have_nodes = false;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '#' || line[0] == '!') continue;
    size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string k = line.substr(0, eq), v = line.substr(eq + 1);
    auto trim = [](std::string& s) {
      while (!s.empty() && (s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
      size_t i = 0;
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
      s.erase(0, i);
    };
    trim(k);
    trim(v);
    if (k == "nodes") { p.nodes = std::stoull(v); have_nodes = true; }
    else if (k == "arcs") p.arcs = std::stoull(v);
    else if (k == "windowsize") p.window = std::stoull(v);
    else if (k == "maxrefcount") p.max_ref_count = std::stoull(v);
    else if (k == "minintervallength") p.min_interval_length = std::stoull(v);
    else if (k == "zetak") p.zetak = std::stoull(v);
    else if (k == "compressionflags" && !v.empty())
      throw Error(WGA_E_UNSUPPORTED, "BVGraph with non-default compression flags: " + v);
  }
  if (!have_nodes) throw Error(WGA_E_FORMAT, path + ": no `nodes` property");
  return p;
}

namespace {
// MSB-first bit reader over a big-endian byte stream with a 64-bit window.
struct BitsBE {
  const uint8_t* p;
  size_t nbytes;
  size_t byte = 0;    // next byte to load
  uint64_t buf = 0;   // left-aligned
  unsigned have = 0;  // valid bits in buf
  BitsBE(const uint8_t* p, size_t n) : p(p), nbytes(n) {}
  void refill() {
    while (have <= 56 && byte < nbytes) {
      buf |= (uint64_t)p[byte++] << (56 - have);
      have += 8;
    }
  }
  uint64_t bits(unsigned n) {  // n <= 57
    if (n == 0) return 0;
    if (have < n) {
      refill();
      if (have < n) throw Error(WGA_E_FORMAT, "BVGraph: bitstream ends inside a record");
    }
    uint64_t v = buf >> (64 - n);
    buf <<= n;
    have -= n;
    return v;
  }
  uint64_t unary() {
    uint64_t c = 0;
    while (true) {
      if (have == 0) {
        refill();
        if (have == 0) throw Error(WGA_E_FORMAT, "BVGraph: bitstream ends inside a record");
      }
      if (buf == 0) {
        c += have;
        have = 0;
        continue;
      }
      unsigned z = (unsigned)__builtin_clzll(buf);
      if (z >= have) {
        c += have;
        buf = 0;
        have = 0;
        continue;
      }
      c += z;
      buf <<= (z + 1);
      have -= (z + 1);
      return c;
    }
  }
  uint64_t gamma() {
    unsigned l = (unsigned)unary();
    if (l > 56) throw Error(WGA_E_FORMAT, "BVGraph: gamma code too long");
    return ((1ull << l) | bits(l)) - 1;
  }
  uint64_t minimal_binary(uint64_t max) {
    unsigned l = 63 - (unsigned)__builtin_clzll(max);
    uint64_t lim = (1ull << (l + 1)) - max;
    uint64_t v = bits(l);
    if (v < lim) return v;
    return ((v << 1) | bits(1)) - lim;
  }
  uint64_t zeta(unsigned k) {
    uint64_t h = unary();
    if ((h + 1) * k > 56) throw Error(WGA_E_FORMAT, "BVGraph: zeta code too long");
    uint64_t left = 1ull << (h * k);
    return minimal_binary((1ull << ((h + 1) * k)) - left) + left - 1;
  }
};
inline int64_t n2i(uint64_t x) { return (x & 1) ? -(int64_t)((x + 1) >> 1) : (int64_t)(x >> 1); }
}  // namespace

void read_bvgraph(const std::string& basename,
                  const std::function<void(uint64_t, const std::vector<uint64_t>&)>& sink, BvProperties* props_out) {
  BvProperties pr = read_bv_properties(basename + ".properties");
  if (props_out) *props_out = pr;
  std::vector<uint8_t> data = read_whole_file(basename + ".graph");
  BitsBE br(data.data(), data.size());
  const uint64_t w = pr.window, L = pr.min_interval_length;
  std::vector<std::vector<uint64_t>> back(w + 1);
  std::vector<uint64_t> out;
  for (uint64_t v = 0; v < pr.nodes; ++v) {
    out.clear();
    const uint64_t d = br.gamma();
    if (d) {
      const uint64_t r = w ? br.unary() : 0;
      if (r > v) throw Error(WGA_E_FORMAT, "BVGraph: reference before node 0");
      if (r) {
        const std::vector<uint64_t>& nb = back[(v - r) % (w + 1)];
        uint64_t nblocks = br.gamma();
        if (nblocks == 0) out = nb;
        else {
          uint64_t idx = br.gamma();
          if (idx > nb.size()) throw Error(WGA_E_FORMAT, "BVGraph: copy block exceeds the referenced list");
          out.insert(out.end(), nb.begin(), nb.begin() + idx);
          for (uint64_t b = 1; b < nblocks; ++b) {
            uint64_t end = idx + br.gamma() + 1;
            if (end > nb.size()) throw Error(WGA_E_FORMAT, "BVGraph: copy block exceeds the referenced list");
            if ((b & 1) == 0) out.insert(out.end(), nb.begin() + idx, nb.begin() + end);
            idx = end;
          }
          if ((nblocks & 1) == 0) out.insert(out.end(), nb.begin() + idx, nb.end());
        }
      }
      bool need_sort = !out.empty();
      if (out.size() > d) throw Error(WGA_E_FORMAT, "BVGraph: more copied arcs than the outdegree");
      if (d - out.size() != 0 && L != 0) {
        uint64_t ni = br.gamma();
        int64_t start = 0;
        for (uint64_t k = 0; k < ni; ++k) {
          start = k == 0 ? (int64_t)v + n2i(br.gamma()) : start + 1 + (int64_t)br.gamma();
          uint64_t len = br.gamma() + L;
          for (uint64_t i = 0; i < len; ++i) out.push_back((uint64_t)start + i);
          start += (int64_t)len;
        }
      }
      if (out.size() > d) throw Error(WGA_E_FORMAT, "BVGraph: more arcs than the outdegree");
      uint64_t left = d - out.size();
      if (left) {
        uint64_t prev = (uint64_t)((int64_t)v + n2i(br.zeta((unsigned)pr.zetak)));
        out.push_back(prev);
        for (uint64_t k = 1; k < left; ++k) {
          prev += 1 + br.zeta((unsigned)pr.zetak);
          out.push_back(prev);
        }
      }
      if (need_sort || L != 0) std::sort(out.begin(), out.end());
    }
    sink(v, out);
    back[v % (w + 1)] = out;
  }
}

}  // namespace wga
