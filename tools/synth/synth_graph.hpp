// Synthetic graphs of the benchmark shapes.  Every successor list is a pure function of
// (seed, node id) -- a counter-based SplitMix64 stream keyed by the node -- so node ranges can be generated
// independently and reproducibly on any number of ranks / threads.
//   kind 0  web-like   : nodes of one "host" (16 consecutive ids) share a template list and keep ~75 % of
//                        it (-> BvComp copy blocks from one of the previous nodes), link runs of consecutive
//                        ids inside nearby hosts (-> intervals) and a few far pages with power-law gaps
//                        (-> residuals); LLP-like locality.
//   kind 1  social-like: power-law out-degree (alpha ~ 2); communities of 8 consecutive accounts share part of their
//                        lists, most other targets are near the account, the rest popular or uniformly random.
//   kind 2  co-authorship: the same with cliques inside the communities and almost only near targets.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <thread>
#include <vector>

// Header-only so that the workload generator is the same code for the product's test API (wga_synth_graph), the
// benchmark and the oracle-only reference arm of bench.py, without either library linking the other.
namespace wgsynth {
namespace detail {

inline uint64_t mix(uint64_t x) {  // SplitMix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct Rng {
  uint64_t s;
  Rng(uint64_t seed, uint64_t key, uint64_t salt) : s(mix(seed ^ mix(key * 0xD6E8FEB86659FD93ull + salt))) {}
  uint64_t next() { return mix(s += 0x9E3779B97F4A7C15ull); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  uint64_t below(uint64_t n) { return n ? (uint64_t)(unit() * (double)n) : 0; }
  uint64_t geometric(double mean) {  // >= 0
    if (mean <= 0) return 0;
    double p = 1.0 / (1.0 + mean);
    double u = unit();
    return (uint64_t)(std::log(1.0 - u) / std::log(1.0 - p));
  }
  // signed gap with |gap| ~ power law (exponent ~1.3), scale-limited to span
  int64_t powerlaw_gap(double span) {
    double u = unit();
    double g = std::pow(span, u * u);  // heavy concentration near 1, tail up to span
    return (next() & 1) ? (int64_t)g : -(int64_t)g;
  }
};

constexpr uint64_t HOST = 16;

inline void web_template(uint64_t seed, uint64_t host, uint64_t N, double mean_degree, std::vector<uint64_t>& T) {
  Rng r(seed, host, 1);
  T.clear();
  const uint64_t base = host * HOST;
  // intra-host navigation: a run of consecutive pages of this host
  uint64_t run_len = 2 + r.geometric(mean_degree * 0.12);
  uint64_t run_start = base + r.below(HOST);
  for (uint64_t i = 0; i < run_len; ++i) T.push_back(run_start + i);
  // a second run in a neighbouring host
  if (r.unit() < 0.6) {
    int64_t hh = (int64_t)host + r.powerlaw_gap(64.0);
    if (hh < 0) hh = 0;
    uint64_t s = (uint64_t)hh * HOST + r.below(HOST);
    uint64_t l = 4 + r.geometric(mean_degree * 0.10);
    for (uint64_t i = 0; i < l; ++i) T.push_back(s + i);
  }
  // far pages
  uint64_t far = 1 + r.geometric(mean_degree * 0.75);
  for (uint64_t i = 0; i < far; ++i) {
    int64_t t = (int64_t)base + r.powerlaw_gap((double)N * 0.5);
    if (t < 0) t = -t;
    T.push_back((uint64_t)t);
  }
  for (auto& x : T) x = x % N;
}

inline void web_list(uint64_t seed, uint64_t v, uint64_t N, double mean_degree, std::vector<uint64_t>& out,
              std::vector<uint64_t>& T) {
  out.clear();
  Rng r(seed, v, 2);
  if (r.unit() < 0.08) return;  // dangling pages
  mean_degree *= 1.34;  // calibration: dangling pages, dropped segments and duplicates
  web_template(seed, v / HOST, N, mean_degree, T);
  std::sort(T.begin(), T.end());
  const bool faithful = r.unit() < 0.35;  // navigation pages copy the whole template
  // the others drop whole segments of the (sorted) template, as pages of one site share link blocks
  for (size_t i = 0; i < T.size();) {
    size_t seg = 1 + (size_t)r.geometric(7.0);
    bool keep = faithful || r.unit() < 0.8;
    for (size_t k = 0; k < seg && i < T.size(); ++k, ++i)
      if (keep) out.push_back(T[i]);
  }
  uint64_t own = r.geometric(mean_degree * 0.08);
  for (uint64_t i = 0; i < own; ++i) {
    int64_t t = (int64_t)v + r.powerlaw_gap((double)N * 0.25);
    if (t < 0) t = -t;
    out.push_back((uint64_t)t % N);
  }
  if (r.unit() < 0.2) {
    uint64_t s = (v + 1 + r.below(32)) % N, l = 4 + r.geometric(3.0);
    for (uint64_t i = 0; i < l && s + i < N; ++i) out.push_back(s + i);
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

// Social-like lists.  Accounts of one community (8 consecutive ids, as a locality-preserving ordering such as LLP
// leaves them) share part of their lists (-> some copying, as in co-authorship and follower graphs), most other
// targets are near the account in id space (scale independent of the graph size, so that bit/link does not grow
// with log N), the rest are popular accounts (small ids, very skewed) and uniformly random ones.
constexpr uint64_t COMMUNITY = 8;

struct SocialShape {
  double near, popular;  // fractions of the targets near the account / among the popular accounts (rest: uniform)
  double span;           // scale of the near gaps
  double shared;         // size of the community's shared list relative to the mean degree
  double keep;           // probability that a member keeps a shared target
  bool clique;           // members of a community link to each other (co-authorship)
  double degree_scale;   // calibration: duplicates removed from the lists
};
// kind 1: follower graph (twitter-2010-shaped) ; kind 2: co-authorship (dblp-2011-shaped)
constexpr SocialShape SOCIAL_SHAPES[2] = {{0.62, 0.28, 65536.0, 0.35, 0.75, false, 1.23},
                                          {0.85, 0.08, 4096.0, 0.70, 0.85, true, 1.28}};

inline uint64_t social_target(Rng& r, uint64_t v, uint64_t N, const SocialShape& sh) {
  const double c = r.unit();
  if (c < sh.near) {  // near v: heavy-tailed gap, absolute scale
    int64_t g = (int64_t)v + r.powerlaw_gap(sh.span);
    return (uint64_t)(g < 0 ? -g : g) % N;
  }
  if (c < sh.near + sh.popular) {  // popular accounts: skewed towards small ids
    const double x = r.unit();
    const double x2 = x * x;
    return (uint64_t)((double)N * x2 * x2 * x2) % N;
  }
  return r.below(N);
}

inline void social_list(uint64_t seed, uint64_t v, uint64_t N, double mean_degree, std::vector<uint64_t>& out,
                        std::vector<uint64_t>& T, const SocialShape& sh) {
  out.clear();
  Rng r(seed, v, 3);
  mean_degree *= sh.degree_scale;
  // Pareto degrees: alpha = 2 -> mean = 2*dmin ; capped so that one record cannot dominate the decode
  // E[min(Pareto(2, xm), cap)] = xm * (2 - xm / cap); 10 % of the accounts follow nobody
  double u = r.unit();
  double cap = std::min<double>((double)N * 0.02, 200000.0);
  double dmin = mean_degree / (0.9 * 1.5);
  for (int it = 0; it < 4; ++it) dmin = mean_degree / (0.9 * (1.5 - dmin / cap));
  double d = dmin / std::sqrt(1.0 - u) - dmin * 0.5;
  if (d > cap) d = cap;
  uint64_t deg = (uint64_t)d;
  if (r.unit() < 0.1) deg = 0;
  if (deg == 0) return;
  // the community's shared targets: about a third of a typical list
  const uint64_t com = v / COMMUNITY;
  Rng rc(seed, com, 4);
  T.clear();
  const uint64_t shared = 1 + rc.geometric(mean_degree * sh.shared);
  for (uint64_t i = 0; i < shared; ++i) T.push_back(social_target(rc, com * COMMUNITY, N, sh));
  if (sh.clique) {
    const uint64_t members = 2 + rc.below(COMMUNITY - 1);
    for (uint64_t i = 0; i < members; ++i) T.push_back((com * COMMUNITY + i) % N);
  }
  out.reserve(deg + 8);
  uint64_t from_template = 0;
  for (size_t i = 0; i < T.size() && from_template < deg; ++i)
    if (T[i] != v && r.unit() < sh.keep) { out.push_back(T[i]); ++from_template; }
  for (uint64_t i = from_template; i < deg; ++i) out.push_back(social_target(r, v, N, sh));
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

}  // namespace detail

inline void synth_list(int kind, uint64_t seed, uint64_t v, uint64_t N, double mean_degree, std::vector<uint64_t>& out,
                std::vector<uint64_t>& scratch) {
  if (kind == 0) detail::web_list(seed, v, N, mean_degree, out, scratch);
  else detail::social_list(seed, v, N, mean_degree, out, scratch, detail::SOCIAL_SHAPES[kind - 1]);
}

// Fills CSR for nodes [first,last). h_succ == nullptr: count only (offsets still written when given).
inline uint64_t synth_graph(int kind, uint64_t N, double mean_degree, uint64_t seed, uint64_t first, uint64_t last,
                     int threads, uint64_t* h_offsets, uint32_t* h_succ) {
  if (kind < 0 || kind > 2) throw std::invalid_argument("unknown synthetic graph kind");
  if (first > last || last > N) throw std::invalid_argument("bad node range");
  if (N >= (1ull << 32)) throw std::invalid_argument("synthetic graphs are limited to 2^32 nodes");
  if (threads < 1) threads = 1;
  const uint64_t n = last - first;
  std::vector<uint64_t> local_offsets;
  uint64_t* offs = h_offsets;
  if (!offs) {
    local_offsets.resize(n + 1);
    offs = local_offsets.data();
  }
  auto run = [&](bool fill) {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
      pool.emplace_back([&, t] {
        std::vector<uint64_t> out, scratch;
        uint64_t a = first + n * t / threads, b = first + n * (t + 1) / threads;
        for (uint64_t v = a; v < b; ++v) {
          synth_list(kind, seed, v, N, mean_degree, out, scratch);
          if (!fill) offs[v - first + 1] = out.size();
          else {
            uint32_t* dst = h_succ + offs[v - first];
            for (size_t i = 0; i < out.size(); ++i) dst[i] = (uint32_t)out[i];
          }
        }
      });
    }
    for (auto& th : pool) th.join();
  };
  run(false);  // degrees -> offsets
  offs[0] = 0;
  for (uint64_t i = 0; i < n; ++i) offs[i + 1] += offs[i];
  if (h_succ) run(true);
  return offs[n];
}

}  // namespace wgsynth
