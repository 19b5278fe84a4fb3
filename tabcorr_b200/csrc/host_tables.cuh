// tabcorr_b200 -- host-side table preparation: padded layouts, DMMA fragment streams, quadrature
// plans, tile / workspace heuristics, math-table upload, launch helpers.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"
#include "device_math.cuh"
#include "occupation.cuh"
#include "predict_kernel.cuh"
#include "leauthaud11.cuh"
#include "aux_kernels.cuh"
#include "tcgen05_contract.cuh"
#include "halo_bins.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// host-side table preparation
// ------------------------------------------------------------------------------------------
template <typename T>
int upload(const std::vector<T>& host, T** dev) {
  *dev = nullptr;
  size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
  TC_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), bytes));
  if (!host.empty())
    TC_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TC_OK;
}

struct PlanHost {
  OccPlan dev{};
  std::vector<void*> allocations;
};

struct Layout {
  bool built = false;
  bool built32 = false;   // afrag32 (3xTF32 mode) is built on first use
  bool built_tc = false;  // tcgen05 images (3xTF32 mode on the 5th-generation tensor cores)
  TcgenDev tcgen{};
  LayoutDev dev{};
  std::vector<int> row_to_pad;
  std::vector<void*> allocations;
  std::map<int, PlanHost> plans;  // by n_gauss
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Tuning knobs for experiments (tools/bench_variants.py): TC_TUNE_<NAME>=<int> in the environment.
int tune(const char* name, int fallback) {
  const char* v = std::getenv((std::string("TC_TUNE_") + name).c_str());
  return v && *v ? std::atoi(v) : fallback;
}

}  // namespace

struct tc_table {
  int device = 0;
  int mode = 0;
  int n_rows = 0, n_r = 0, n_tables = 0, n_cen = 0;
  std::vector<double> n_h, log_min, log_max, pct, dist;
  bool has_dist = false;
  std::vector<int> is_sat;
  std::vector<std::vector<double>> matrices;  // host copies, kept to build the split layout lazily
  std::map<int, std::pair<std::vector<double>, std::vector<double>>> rules;  // n_gauss -> (x, w)
  Layout layouts[2];  // [separate]
  std::mutex mutex;
};

struct tc_interp {
  int device = 0;
  InterpDev dev{};
  int sum_knots = 0;
  std::vector<void*> allocations;
};

namespace {

// Build the padded row order and the A-fragment stream of one layout.
int build_layout(tc_table* t, int separate) {
  Layout& L = t->layouts[separate];
  if (L.built) return TC_OK;
  const int N = t->n_rows, R = t->n_r, T = t->n_tables;
  const int Reff = R * T;
  const int n_cen = t->n_cen, n_sat = N - n_cen;
  // centrals first (stable); in the split layout the satellite block starts on a 16-row tile
  const int nc_pad = separate ? round_up(n_cen, 16) : n_cen;
  const int n_pad = std::max(16, round_up(nc_pad + n_sat, 16));
  L.row_to_pad.assign(N, -1);
  std::vector<int> pad_to_row(n_pad, -1);
  {
    int ic = 0, is = nc_pad;
    for (int i = 0; i < N; i++) {
      int p = t->is_sat[i] ? is++ : ic++;
      L.row_to_pad[i] = p;
      pad_to_row[p] = i;
    }
  }
  const int T16 = n_pad / 16;
  std::vector<Chunk> chunks;
  std::vector<std::vector<int>> out_lists;
  std::vector<double2> afrag;
  long long ks_per_r = 0;
  int n_parts = 0;

  if (t->mode == TC_MODE_AUTO) {
    ks_per_r = 2LL * T16 * (T16 + 1);
    afrag.assign((size_t)Reff * ks_per_r * 32 + 32, make_double2(0.0, 0.0));
    // M'[i][j] (j <= i) = M[i][j] (i == j) or 2 M[i][j]: the reference's packed prefactor sum
    // (tabcorr.py:638-642) written as a lower-triangular matrix product.
    for (int tb = 0; tb < T; tb++) {
      const double* packed = t->matrices[tb].data();
      const size_t P = (size_t)N * (N + 1) / 2;
      for (int r = 0; r < R; r++) {
        const double* m = packed + (size_t)r * P;
        double2* dst = afrag.data() + (size_t)(tb * R + r) * ks_per_r * 32;
        for (int i = 0; i < N; i++) {
          const int pi = L.row_to_pad[i];
          for (int j = 0; j <= i; j++) {
            const int pj = L.row_to_pad[j];
            double val = m[(size_t)i * (i + 1) / 2 + j];
            if (i != j) val *= 2.0;
            const int hi = std::max(pi, pj), lo = std::min(pi, pj);
            const int mt = hi / 16, rr = hi % 16, ks = lo / 4, tg = lo % 4;
            double2& d = dst[((size_t)2 * mt * (mt + 1) + ks) * 32 + (rr % 8) * 4 + tg];
            if (rr < 8) d.x = val; else d.y = val;
          }
        }
      }
    }
    // chunks: per radial bin, the tile range cut into pieces of similar cost
    const int n_comp = separate ? 3 : 1;
    out_lists.assign((size_t)Reff * n_comp, {});
    const int c16 = nc_pad / 16;  // first satellite tile (split layout)
    // chunks per draw tile: enough to balance 12 warps, few enough that slot dispatch, accumulator
    // set-up and the 96-shuffle reduction at the end of a chunk stay small beside its DMMAs
    // (measured per 1e5 draws, 72 -> this: N = 60 0.613 -> 0.600 ms, N = 120 1.58 -> 1.55,
    // N = 240 3.86 -> 3.84; N = 500 prefers 72: 16.25 vs 16.47)
    const int chunk_target = T16 > 16 ? 72 : T16 > 8 ? 48 : 40;
    int pieces = std::max(1, std::min(T16, (tune("CHUNKS", chunk_target) + Reff - 1) / Reff));
    auto add_range = [&](int r, int mt_lo, int mt_hi, int k_begin, int k_cap, int comp, int np) {
      // cost of tile mt ~ number of k-steps
      auto cost = [&](int mt) { return std::max(0, std::min(4 * (mt + 1), k_cap) - k_begin); };
      long long total = 0;
      for (int mt = mt_lo; mt < mt_hi; mt++) total += cost(mt);
      if (total == 0) return;
      np = std::max(1, std::min(np, mt_hi - mt_lo));
      long long acc = 0;
      int start = mt_lo, piece = 0;
      for (int mt = mt_lo; mt < mt_hi; mt++) {
        acc += cost(mt);
        bool last = mt == mt_hi - 1;
        if (last || acc * np >= total * (piece + 1)) {
          Chunk c{};
          c.r = r; c.mt0 = start; c.mt1 = mt + 1; c.k_begin = k_begin; c.k_cap = k_cap;
          c.part_row = n_parts++;
          chunks.push_back(c);
          out_lists[(size_t)r * n_comp + comp].push_back(c.part_row);
          start = mt + 1;
          piece++;
        }
      }
    };
    const int kinf = 1 << 28;
    for (int r = 0; r < Reff; r++) {
      if (!separate) {
        add_range(r, 0, T16, 0, kinf, 0, pieces);
      } else {
        add_range(r, 0, c16, 0, kinf, 0, pieces);              // centrals-centrals
        add_range(r, c16, T16, 0, nc_pad / 4, 1, pieces);       // centrals-satellites
        add_range(r, c16, T16, nc_pad / 4, kinf, 2, pieces);    // satellites-satellites
      }
    }
  } else {
    const int n_rt = (Reff + 15) / 16;
    ks_per_r = n_pad / 4;
    afrag.assign((size_t)n_rt * ks_per_r * 32 + 32, make_double2(0.0, 0.0));
    for (int tb = 0; tb < T; tb++) {
      const double* m = t->matrices[tb].data();
      for (int r = 0; r < R; r++) {
        const int re = tb * R + r, rt = re / 16, rr = re % 16;
        for (int i = 0; i < N; i++) {
          const int pi = L.row_to_pad[i];
          double2& d = afrag[((size_t)rt * ks_per_r + pi / 4) * 32 + (rr % 8) * 4 + pi % 4];
          if (rr < 8) d.x = m[(size_t)r * N + i]; else d.y = m[(size_t)r * N + i];
        }
      }
    }
    const int n_comp = separate ? 2 : 1;
    out_lists.assign((size_t)Reff * n_comp, {});
    const int ks_total = n_pad / 4;
    // k-ranges: split at the centrals/satellites boundary (split layout) and into pieces
    std::vector<std::pair<int, int>> segs;
    if (separate) {
      segs.push_back({0, nc_pad / 4});
      segs.push_back({nc_pad / 4, ks_total});
    } else {
      segs.push_back({0, ks_total});
    }
    // six chunks per radial tile: every chunk writes 16 scratch rows per draw tile, the DMMAs of a
    // cross table are few, and with four chunks per warp (48) the scratch traffic cost 3-6 %
    // (per 1e5 draws, 48 / 12 / 6 / 3 / 1 chunks: N = 1104 2.04 / 1.97 / 1.93 / 1.93 / 2.24 ms,
    // N = 240 0.578 / 0.549 / 0.542 / 0.546 / 0.564 ms, N = 60 unchanged; TC_TUNE_CROSS_CHUNKS)
    const int want = std::max(1, (tune("CROSS_CHUNKS", 6) + n_rt - 1) / n_rt / (int)segs.size());
    for (int rt = 0; rt < n_rt; rt++) {
      for (size_t sg = 0; sg < segs.size(); sg++) {
        const int lo = segs[sg].first, hi = segs[sg].second;
        if (hi <= lo) continue;
        const int np = std::max(1, std::min(want, (hi - lo + 7) / 8));
        for (int pc = 0; pc < np; pc++) {
          Chunk c{};
          c.r = rt;
          c.k_begin = lo + (int)((long long)(hi - lo) * pc / np);
          c.k_cap = lo + (int)((long long)(hi - lo) * (pc + 1) / np);
          if (c.k_cap <= c.k_begin) continue;
          c.part_row = n_parts;
          n_parts += 16;
          chunks.push_back(c);
          for (int rr = 0; rr < 16; rr++) {
            const int re = rt * 16 + rr;
            if (re < Reff) out_lists[(size_t)re * n_comp + sg].push_back(c.part_row + rr);
          }
        }
      }
    }
  }
  auto chunk_cost = [&](const Chunk& c) {
    if (t->mode != TC_MODE_AUTO) return (long long)(c.k_cap - c.k_begin);
    long long s = 0;
    for (int mt = c.mt0; mt < c.mt1; mt++)
      s += std::max(0, std::min(4 * (mt + 1), c.k_cap) - c.k_begin);
    return s;
  };
  // longest chunks first (only the end of a CTA's last tile is sensitive to the order)
  std::stable_sort(chunks.begin(), chunks.end(),
                   [&](const Chunk& a, const Chunk& b) { return chunk_cost(a) > chunk_cost(b); });

  std::vector<long long> cost_prefix(chunks.size() + 1, 0);
  for (size_t c = 0; c < chunks.size(); c++)
    cost_prefix[c + 1] = cost_prefix[c] + std::max<long long>(1, chunk_cost(chunks[c]));

  std::vector<int> out_ptr(out_lists.size() + 1, 0), out_parts;
  for (size_t o = 0; o < out_lists.size(); o++) {
    out_ptr[o + 1] = out_ptr[o] + (int)out_lists[o].size();
    out_parts.insert(out_parts.end(), out_lists[o].begin(), out_lists[o].end());
  }

  double2* d_afrag; Chunk* d_chunks; int *d_out_ptr, *d_out_parts, *d_pad_to_row;
  long long* d_cost_prefix;
  int rc;
  if ((rc = upload(cost_prefix, &d_cost_prefix))) return rc;
  L.allocations.push_back(d_cost_prefix);
  if ((rc = upload(afrag, &d_afrag))) return rc;
  L.allocations.push_back(d_afrag);
  if ((rc = upload(chunks, &d_chunks))) return rc;
  L.allocations.push_back(d_chunks);
  if ((rc = upload(out_ptr, &d_out_ptr))) return rc;
  L.allocations.push_back(d_out_ptr);
  if ((rc = upload(out_parts, &d_out_parts))) return rc;
  L.allocations.push_back(d_out_parts);
  if ((rc = upload(pad_to_row, &d_pad_to_row))) return rc;
  L.allocations.push_back(d_pad_to_row);

  L.dev.n_rows = N;
  L.dev.n_pad = n_pad;
  L.dev.nc_pad = nc_pad;
  L.dev.n_parts = std::max(n_parts, 1);
  L.dev.n_chunks = (int)chunks.size();
  L.dev.n_out = (int)out_lists.size();
  L.dev.ks_per_r = ks_per_r;
  L.dev.afrag = d_afrag;
  L.dev.chunks = d_chunks;
  L.dev.chunk_cost_prefix = d_cost_prefix;
  L.dev.out_ptr = d_out_ptr;
  L.dev.out_parts = d_out_parts;
  L.dev.pad_to_row = d_pad_to_row;
  L.built = true;
  return TC_OK;
}

// round an FP32 value to TF32 (10 explicit mantissa bits), ties to even
float tf32_round_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, sizeof(u));
  u += 0xfffu + ((u >> 13) & 1u);
  u &= 0xffffe000u;
  std::memcpy(&x, &u, sizeof(u));
  return x;
}

// A-fragment stream of the 3xTF32 mode: the same lower-triangular M' as build_layout, in m16n8k8
// fragments (16-row tiles x k8-steps of 8 columns), every entry split into TF32 high and low part.
int build_afrag32(tc_table* t, int separate) {
  Layout& L = t->layouts[separate];
  if (L.built32) return TC_OK;
  if (t->mode != TC_MODE_AUTO)
    return fail(TC_EUNSUPPORTED, "the 3xTF32 mode exists for auto-correlation tables only");
  const int N = t->n_rows, R = t->n_r, T = t->n_tables, Reff = R * T;
  const int T16 = L.dev.n_pad / 16;
  const long long ks8_per_r = (long long)T16 * (T16 + 1);
  std::vector<float4> frag((size_t)Reff * ks8_per_r * 64 + 64, make_float4(0.f, 0.f, 0.f, 0.f));
  const size_t P = (size_t)N * (N + 1) / 2;
  for (int tb = 0; tb < T; tb++) {
    const double* packed = t->matrices[tb].data();
    for (int r = 0; r < R; r++) {
      const double* m = packed + (size_t)r * P;
      float4* dst = frag.data() + (size_t)(tb * R + r) * ks8_per_r * 64;
      for (int i = 0; i < N; i++) {
        const int pi = L.row_to_pad[i];
        for (int j = 0; j <= i; j++) {
          const int pj = L.row_to_pad[j];
          double val = m[(size_t)i * (i + 1) / 2 + j];
          if (i != j) val *= 2.0;
          const int hi_r = std::max(pi, pj), lo_c = std::min(pi, pj);
          const int mt = hi_r / 16, rr = hi_r % 16, ks = lo_c / 8, kk = lo_c % 8;
          const int lane = (rr % 8) * 4 + (kk % 4), reg = (rr / 8) + 2 * (kk / 4);
          const float hi = tf32_round_host((float)val);
          const float lo = tf32_round_host((float)(val - (double)hi));
          float4* block = dst + ((size_t)mt * (mt + 1) + ks) * 64;
          reinterpret_cast<float*>(&block[lane])[reg] = hi;
          reinterpret_cast<float*>(&block[32 + lane])[reg] = lo;
        }
      }
    }
  }
  float4* d_frag;
  int rc;
  if ((rc = upload(frag, &d_frag))) return rc;
  L.allocations.push_back(d_frag);
  L.dev.afrag32 = d_frag;
  L.dev.ks8_per_r = ks8_per_r;
  L.built32 = true;
  return TC_OK;
}

// Is the tcgen05 contraction applicable?  Auto tables, total predictions, a draw tile of 128 x Kp
// TF32 that leaves room for the table stages in shared memory.
bool tcgen_eligible(const tc_table* t, int separate) {
  if (t->mode != TC_MODE_AUTO || separate) return false;
  if (tune("TCGEN", 1) == 0) return false;
  const int n_pad = std::max(16, round_up(t->n_rows, 16));
  return round_up(n_pad, kTcKS) <= kTcMaxKp;
}

// Table stream of the tcgen05 contraction: the dense symmetric M_r in padded row order, split into
// TF32 high and low parts, as planes of 128 rows (2 radial bins x 64 table rows) x 64 k in the
// canonical K-major UMMA layout, in the order the kernel consumes them: column block, radial-bin
// pair, {all low planes, all high planes}; two K segments form one 64 KB bulk copy.
int build_tcgen(tc_table* t) {
  Layout& L = t->layouts[0];
  if (L.built_tc) return TC_OK;
  const int N = t->n_rows, R = t->n_r, T = t->n_tables, Reff = R * T;
  const int n_pad = L.dev.n_pad;
  TcgenDev& tc = L.tcgen;
  tc.kp = round_up(n_pad, kTcKS);
  tc.n_seg = tc.kp / kTcKS;
  tc.n_pairs = (tc.n_seg + 1) / 2;
  tc.n_ib = (N + kTcNI - 1) / kTcNI;
  tc.n_reff = Reff;
  tc.n_rp = (Reff + kTcRB - 1) / kTcRB;
  tc.n_parts = tc.n_ib * kTcRB * tc.n_rp;         // [column block][radial bin]
  const size_t rp_bytes = (size_t)2 * tc.n_pairs * kTcCopyBytes;
  std::vector<uint8_t> img((size_t)tc.n_ib * tc.n_rp * rp_bytes, 0);
  const size_t P = (size_t)N * (N + 1) / 2;
  std::vector<int> pad_to_row(tc.kp, -1);
  for (int i = 0; i < N; i++) pad_to_row[L.row_to_pad[i]] = i;
  for (int ib = 0; ib < tc.n_ib; ib++) {
    for (int rp = 0; rp < tc.n_rp; rp++) {
      uint8_t* base = img.data() + ((size_t)ib * tc.n_rp + rp) * rp_bytes;
      for (int sg = 0; sg < tc.n_seg; sg++) {
        uint8_t* lo_plane = base + (size_t)(sg / 2) * kTcCopyBytes + (size_t)(sg % 2) * kTcPlaneBytes;
        uint8_t* hi_plane = lo_plane + (size_t)tc.n_pairs * kTcCopyBytes;
        for (int n = 0; n < kTcN; n++) {
          const int re = rp * kTcRB + n / kTcNI, pi = ib * kTcNI + n % kTcNI;
          if (re >= Reff || pi >= tc.kp || pad_to_row[pi] < 0) continue;
          const int i = pad_to_row[pi];
          const double* m = t->matrices[re / R].data() + (size_t)(re % R) * P;
          for (int k = 0; k < kTcKS; k++) {
            const int pj = sg * kTcKS + k;
            if (pad_to_row[pj] < 0) continue;
            const int j = pad_to_row[pj];
            const double val = i >= j ? m[(size_t)i * (i + 1) / 2 + j] : m[(size_t)j * (j + 1) / 2 + i];
            const float hi = tf32_round_host((float)val);
            const float lo = tf32_round_host((float)(val - (double)hi));
            const size_t off = canon_offset(n, k, kTcSbo);
            std::memcpy(lo_plane + off, &lo, 4);
            std::memcpy(hi_plane + off, &hi, 4);
          }
        }
      }
    }
  }
  // scratch rows per output: one per column block
  std::vector<int> out_ptr(Reff + 1, 0), out_parts;
  for (int re = 0; re < Reff; re++) {
    for (int ib = 0; ib < tc.n_ib; ib++) out_parts.push_back(ib * kTcRB * tc.n_rp + re);
    out_ptr[re + 1] = (int)out_parts.size();
  }
  uint8_t* d_img; int *d_ptr, *d_parts;
  int rc;
  if ((rc = upload(img, &d_img))) return rc;
  L.allocations.push_back(d_img);
  if ((rc = upload(out_ptr, &d_ptr))) return rc;
  L.allocations.push_back(d_ptr);
  if ((rc = upload(out_parts, &d_parts))) return rc;
  L.allocations.push_back(d_parts);
  tc.b_img = d_img;
  tc.out_ptr = d_ptr;
  tc.out_parts = d_parts;
  L.built_tc = true;
  return TC_OK;
}

// Workspace of the tcgen05 path: operand images of the draw tiles + scratch rows.
struct TcgenWorkspace {
  size_t a_bytes, c_bytes, parts_bytes, ngal_bytes, ngal_parts_bytes, total;
  long long n_tiles;
  int n_ranges_max;
};

TcgenWorkspace plan_tcgen_workspace(const tc_table* t, long long n_draws) {
  TcgenWorkspace w{};
  const int n_pad = std::max(16, round_up(t->n_rows, 16));
  const int kp = round_up(n_pad, kTcKS);
  const int n_ib = (t->n_rows + kTcNI - 1) / kTcNI;
  auto align = [](size_t b) { return (b + 255) / 256 * 256; };
  w.n_tiles = (n_draws + kTcM - 1) / kTcM;
  w.n_ranges_max = 64;
  w.a_bytes = align((size_t)w.n_tiles * kTcM * kp * 4);
  w.c_bytes = align((size_t)w.n_tiles * n_pad * kTcM * 4);
  w.parts_bytes = align((size_t)w.n_tiles * n_ib * kTcRB * ((t->n_r * t->n_tables + kTcRB - 1) / kTcRB) * kTcM * 8);
  w.ngal_bytes = align((size_t)w.n_tiles * 2 * kTcM * 8);
  w.ngal_parts_bytes = align((size_t)w.n_ranges_max * w.n_tiles * kTcM * 8);
  w.total = w.a_bytes + w.c_bytes + w.parts_bytes + w.ngal_bytes + w.ngal_parts_bytes + 256;
  return w;
}

// Quadrature plan of a layout for one Gauss-Legendre rule (tabcorr.py:543-552,568-578).
int build_plan(tc_table* t, int separate, int n_gauss) {
  Layout& L = t->layouts[separate];
  if (L.plans.count(n_gauss)) return TC_OK;
  auto rule = t->rules.find(n_gauss);
  if (rule == t->rules.end())
    return fail(TC_EINVAL, "no quadrature rule registered for n_gauss=" + std::to_string(n_gauss) +
                               " (call tc_table_plan first)");
  const std::vector<double>& x01 = rule->second.first;
  const std::vector<double>& wq = rule->second.second;
  const int N = t->n_rows, G = n_gauss, n_pad = L.dev.n_pad;

  struct Group { double lo, hi; int sat; std::vector<int> rows; };
  std::vector<Group> groups;
  for (int pass = 0; pass < 2; pass++) {  // centrals groups first
    for (int i = 0; i < N; i++) {
      if ((t->is_sat[i] != 0) != (pass == 1)) continue;
      bool placed = false;
      for (auto& gq : groups) {
        if (gq.sat == pass && gq.lo == t->log_min[i] && gq.hi == t->log_max[i] &&
            (int)gq.rows.size() < kGroupRows) {
          gq.rows.push_back(i);
          placed = true;
          break;
        }
      }
      if (!placed) groups.push_back(Group{t->log_min[i], t->log_max[i], pass, {i}});
    }
  }
  const int n_groups = (int)groups.size();
  // the kernel evaluates `unroll` nodes per iteration
  // (tried: 10 nodes in flight -- no gain, tools/bench_variants.py)
  const int unroll = tune("OCC_UNROLL", kOccUnroll) == kOccUnroll && G % kOccUnroll == 0 ? kOccUnroll : 2;
  const int GP = round_up(G, unroll);
  std::vector<double> node_logm((size_t)n_groups * GP), node_m((size_t)n_groups * GP);
  std::vector<int> grp_rows((size_t)n_groups * kGroupRows, -1), grp_is_sat(n_groups);
  std::vector<double> row_c((size_t)(n_pad + 1) * GP, 0.0), row_nh(n_pad, 0.0), row_pct(n_pad, 0.0);
  for (int q = 0; q < n_groups; q++) {
    const Group& gq = groups[q];
    grp_is_sat[q] = gq.sat;
    for (int k = 0; k < G; k++) {
      // prim_haloprop = 10**(log_min + d_log * x) and halotools' log10(prim_haloprop)
      double m = std::pow(10.0, gq.lo + (gq.hi - gq.lo) * x01[k]);
      node_m[(size_t)q * GP + k] = m;
      node_logm[(size_t)q * GP + k] = std::log10(m);
    }
    for (int k = G; k < GP; k++) {  // zero-weight padding node
      node_m[(size_t)q * GP + k] = node_m[(size_t)q * GP];
      node_logm[(size_t)q * GP + k] = node_logm[(size_t)q * GP];
    }
    for (size_t s = 0; s < gq.rows.size(); s++) {
      const int i = gq.rows[s], p = L.row_to_pad[i];
      grp_rows[(size_t)q * kGroupRows + s] = p;
      row_nh[p] = t->n_h[i];
      row_pct[p] = t->pct[i];
      const double n = t->has_dist ? t->dist[i] + 1.0 : 0.0;  // tabcorr.py:568-574
      double norm = 0.0;
      for (int k = 0; k < G; k++) norm += wq[k] * std::pow(node_m[(size_t)q * GP + k], n);
      for (int k = 0; k < G; k++)
        row_c[(size_t)p * GP + k] = wq[k] * std::pow(node_m[(size_t)q * GP + k], n) / norm;
    }
  }
  // ---- series evaluation (occupation_item_series): group centres, scaled node moments / k!,
  // number of terms per bucket of h (centrals) and y (satellites), all in long double ---------
  std::vector<double4> grp_ser(n_groups);
  // (two zero rows behind the last moment: the term loops prefetch one row ahead)
  std::vector<double2> grp_mom((size_t)(kSerMom + 2) * n_groups, make_double2(0.0, 0.0));
  std::vector<long double> cen_mom_max(kSerMom, 0.0L), sat_mom_max(kSerMom, 0.0L);
  std::vector<long double> inv_fact(kSerMom + 64, 1.0L);
  for (size_t k = 1; k < inv_fact.size(); k++) inv_fact[k] = inv_fact[k - 1] / (long double)k;
  double cen_d_max = 0.0;
  for (int q = 0; q < n_groups; q++) {
    const Group& gq = groups[q];
    const double* nodes = (gq.sat ? node_m.data() : node_logm.data()) + (size_t)q * GP;
    double lo = nodes[0], hi = nodes[0];
    for (int k = 1; k < G; k++) { lo = std::min(lo, nodes[k]); hi = std::max(hi, nodes[k]); }
    const double centre = 0.5 * (lo + hi);
    // scaled offsets s_j in [-1, 1]: centrals (log10 M - centre) / d, satellites (M / m_ref - 1) / u_max
    std::vector<long double> sj(G);
    long double scale = 0.0L;
    for (int k = 0; k < G; k++) {
      sj[k] = gq.sat ? (long double)nodes[k] / (long double)centre - 1.0L
                     : (long double)nodes[k] - (long double)centre;
      scale = std::max(scale, fabsl(sj[k]));
    }
    const double scale_d = (double)scale;             // the kernel multiplies by this double
    for (int k = 0; k < G; k++) sj[k] = scale_d > 0.0 ? sj[k] / (long double)scale_d : 0.0L;
    if (gq.sat) {
      grp_ser[q] = make_double4(centre, scale_d, lo, hi);
    } else {
      grp_ser[q] = make_double4(centre, scale_d, 0.0, 0.0);
      cen_d_max = std::max(cen_d_max, scale_d);
    }
    for (size_t s = 0; s < gq.rows.size(); s++) {
      const int p = L.row_to_pad[gq.rows[s]];
      for (int k = 0; k < kSerMom; k++) {
        long double mk = 0.0L;
        for (int j = 0; j < G; j++) mk += (long double)row_c[(size_t)p * GP + j] * powl(sj[j], k);
        auto& mx = gq.sat ? sat_mom_max : cen_mom_max;
        mx[k] = std::max(mx[k], fabsl(mk));
        double2& dst = grp_mom[(size_t)k * n_groups + q];
        (s == 0 ? dst.x : dst.y) = (double)(mk * inv_fact[k]);
      }
    }
  }
  // Terms per bucket: the smallest K whose tail bound is below 1e-14.
  //   centrals: |term k| <= 1.0865 / sqrt(pi) sqrt(2^(k-1) (k-1)!) h^k / k! max_rows |mu_k|
  //             (Cramer: |H_n(x)| exp(-x^2 / 2) <= 1.0865 sqrt(2^n n!)), h at the bucket's upper edge;
  //   satellites: |term k| <= max_{0 <= alpha <= 4} |binom(alpha, k)| y^k max_rows |nu_k|.
  // Moments beyond the table (k > kSerMaxTerms) are bounded by 1.
  std::vector<unsigned char> cen_terms(kSerBuckets, 255), sat_terms(kSerBuckets, 255);
  {
    const int n_tail = kSerMom + 60;
    const long double eps = 1e-14L;
    std::vector<long double> binom_max(n_tail, 0.0L);
    for (int ia = 0; ia <= 4000; ia++) {
      const long double alpha = kSerAlphaMax * ia / 4000.0L;
      long double c = 1.0L;
      for (int k = 1; k < n_tail; k++) {
        c *= (alpha - (k - 1)) / k;
        binom_max[k] = std::max(binom_max[k], fabsl(c) * 1.001L);   // grid spacing margin
      }
    }
    for (int i = 0; i < kSerBuckets; i++) {
      const long double h = (i + 1) / (long double)kSerCenBucket;
      const long double y = (i + 1) / (long double)kSerSatBucket;
      std::vector<long double> tc(n_tail, 0.0L), ts(n_tail, 0.0L);
      long double herm = 1.0865L / sqrtl(3.14159265358979323846264338327950288L);  // k = 1: sqrt(2^0 0!)
      for (int k = 1; k < n_tail; k++) {
        if (k > 1) herm *= sqrtl(2.0L * (k - 1));
        tc[k] = herm * powl(h, k) * inv_fact[k] * (k < kSerMom ? cen_mom_max[k] : 1.0L);
        ts[k] = binom_max[k] * powl(y, k) * (k < kSerMom ? sat_mom_max[k] : 1.0L);
      }
      long double tail_c = 0.0L, tail_s = 0.0L;
      int kc = -1, ks = -1;
      for (int k = n_tail - 1; k >= 2; k--) {   // tail = sum of the terms beyond k
        if (tail_c <= eps) kc = k;
        if (tail_s <= eps) ks = k;
        tail_c += tc[k];
        tail_s += ts[k];
      }
      if (tune("SERIES", 1) == 0) continue;   // experiments: every pair through the node path
      if (kc >= 0 && kc <= kSerMaxTerms) cen_terms[i] = (unsigned char)std::max(kc, 2);
      if (ks >= 0 && ks <= kSerMaxTerms) sat_terms[i] = (unsigned char)std::max(ks, 1);
    }
  }
  PlanHost ph;
  double *d_logm, *d_m, *d_c, *d_nh, *d_pct; int *d_rows, *d_sat;
  int rc;
  {
    double4* d_ser; double2* d_mom; unsigned char *d_ct, *d_st;
    if ((rc = upload(grp_ser, &d_ser))) return rc; ph.allocations.push_back(d_ser);
    if ((rc = upload(grp_mom, &d_mom))) return rc; ph.allocations.push_back(d_mom);
    if ((rc = upload(cen_terms, &d_ct))) return rc; ph.allocations.push_back(d_ct);
    if ((rc = upload(sat_terms, &d_st))) return rc; ph.allocations.push_back(d_st);
    ph.dev.grp_ser = d_ser;
    ph.dev.grp_mom = d_mom;
    ph.dev.cen_terms = d_ct;
    ph.dev.sat_terms = d_st;
    ph.dev.cen_d_max = cen_d_max;
  }
  if ((rc = upload(node_logm, &d_logm))) return rc; ph.allocations.push_back(d_logm);
  if ((rc = upload(node_m, &d_m))) return rc; ph.allocations.push_back(d_m);
  {
    std::vector<double> node_inv_m(node_m.size());
    for (size_t k = 0; k < node_m.size(); k++) node_inv_m[k] = 1.0 / node_m[k];
    double* d_inv;
    if ((rc = upload(node_inv_m, &d_inv))) return rc;
    ph.allocations.push_back(d_inv);
    ph.dev.node_inv_m = d_inv;
  }
  {
    // mass bins of the leauthaud11 kernel: every centrals group with the (first unused) satellites
    // group over the same mass range -- the same nodes -- then the unpaired satellites groups
    std::vector<L11Bin> l11_bins;
    std::vector<int> pad_to_ref(n_pad, -1);
    for (int i = 0; i < N; i++) pad_to_ref[L.row_to_pad[i]] = i;
    auto same_weights = [&](int q, int s) {   // row by row the same normalised weights
      for (int k = 0; k < kGroupRows; k++) {
        const int a = grp_rows[(size_t)q * kGroupRows + k], b = grp_rows[(size_t)s * kGroupRows + k];
        if ((a < 0) != (b < 0)) return false;
        if (a >= 0 && std::memcmp(&row_c[(size_t)a * GP], &row_c[(size_t)b * GP], sizeof(double) * GP))
          return false;
      }
      return true;
    };
    auto make_bin = [&](int q, int s) {
      L11Bin bin{};
      bin.cen = q;
      bin.sat = s;
      bin.same_w = q >= 0 && s >= 0 && same_weights(q, s) ? 1 : 0;
      for (int k = 0; k < 4; k++) {
        const int grp = k < 2 ? q : s;
        const int row = grp >= 0 ? grp_rows[(size_t)grp * kGroupRows + (k & 1)] : -1;
        bin.row[k] = row;
        bin.dst[k] = row >= 0 ? pad_to_ref[row] : -1;
        bin.pct[k] = row >= 0 ? row_pct[row] : 0.0;
      }
      bin.first_logm = node_logm[(size_t)(q >= 0 ? q : s) * GP];
      return bin;
    };
    std::vector<char> used(n_groups, 0);
    for (int q = 0; q < n_groups; q++) {
      if (groups[q].sat) continue;
      int mate = -1;
      for (int s = 0; s < n_groups && mate < 0; s++)
        if (groups[s].sat && !used[s] && groups[s].lo == groups[q].lo && groups[s].hi == groups[q].hi)
          mate = s;
      if (mate >= 0) used[mate] = 1;
      l11_bins.push_back(make_bin(q, mate));
    }
    for (int s = 0; s < n_groups; s++)
      if (groups[s].sat && !used[s]) l11_bins.push_back(make_bin(-1, s));
    L11Bin* d_bins;
    if ((rc = upload(l11_bins, &d_bins))) return rc;
    ph.allocations.push_back(d_bins);
    ph.dev.l11_bins = d_bins;
    ph.dev.n_l11_bins = (int)l11_bins.size();
  }
  if ((rc = upload(grp_rows, &d_rows))) return rc; ph.allocations.push_back(d_rows);
  if ((rc = upload(grp_is_sat, &d_sat))) return rc; ph.allocations.push_back(d_sat);
  if ((rc = upload(row_c, &d_c))) return rc; ph.allocations.push_back(d_c);
  if ((rc = upload(row_nh, &d_nh))) return rc; ph.allocations.push_back(d_nh);
  if ((rc = upload(row_pct, &d_pct))) return rc; ph.allocations.push_back(d_pct);
  ph.dev.n_groups = n_groups;
  ph.dev.n_cen_groups = 0;
  for (int q = 0; q < n_groups; q++) ph.dev.n_cen_groups += grp_is_sat[q] ? 0 : 1;
  ph.dev.n_gauss = G;
  ph.dev.n_gauss_pad = GP;
  ph.dev.unroll = unroll;
  ph.dev.zero_row = n_pad;
  ph.dev.node_logm = d_logm;
  ph.dev.node_m = d_m;
  ph.dev.grp_rows = d_rows;
  ph.dev.grp_is_sat = d_sat;
  ph.dev.row_c = d_c;
  ph.dev.row_nh = d_nh;
  ph.dev.row_pct = d_pct;
  L.plans[n_gauss] = ph;
  return TC_OK;
}

size_t predict_smem_bytes(int n_pad, int nt, int n_buf) {
  return ((size_t)n_buf * n_pad * 8 * nt + kTabDoubles) * sizeof(double) + sizeof(PredictCtrl);
}

// Draw-tile width (8 nt draws) and number of W buffers.  Two buffers let the occupation of the
// next tile overlap the contraction of the current one, but halve the tile width that fits in
// shared memory, and the width is what the table stream from L2 is amortised over: measured on
// B200, N = 240: 2 x 56 draws = 1 x 64 draws (4.03 ms per 1e5 draws), N = 500: 1 x 48 draws beats
// 2 x 24 draws (16.9 vs 18.1 ms).  So two buffers are used while they leave at least 40 draws;
// within a buffer count the widest tile that fits, narrower only while the batch is too small to
// give every SM a tile.
void pick_tile(int n_pad, long long n_draws, int n_sm, int* nt_out, int* n_buf_out) {
  *nt_out = 0;
  *n_buf_out = 0;
  const int forced = tune("NBUF", 0);
  auto widest = [&](int n_buf) {
    for (int nt = 8; nt >= 1; nt--)
      if (predict_smem_bytes(n_pad, nt, n_buf) <= (size_t)kSmemLimit) return nt;
    return 0;
  };
  int n_buf = widest(2) >= 5 ? 2 : 1;
  if (forced >= 1 && forced <= kMaxWBuffers && widest(forced) > 0) n_buf = forced;
  if (widest(n_buf) == 0) n_buf = 1;
  int best = 0;
  for (int nt = 8; nt >= 1; nt--) {
    if (predict_smem_bytes(n_pad, nt, n_buf) > (size_t)kSmemLimit) continue;
    best = nt;  // the largest that fits, shrinking while the grid would not fill the device
    if ((n_draws + 8 * nt - 1) / (8 * nt) >= n_sm) break;
  }
  if (best) {
    *nt_out = best;
    *n_buf_out = n_buf;
  }
}

struct Workspace {
  size_t parts_bytes, ngal_bytes, total;
  long long n_tiles;
  int nt, n_buf;
};

Workspace plan_workspace(const Layout& L, long long n_draws, int n_sm) {
  Workspace w{};
  pick_tile(L.dev.n_pad, n_draws, n_sm, &w.nt, &w.n_buf);
  if (w.nt == 0) return w;
  const int bm = 8 * w.nt;
  w.n_tiles = (n_draws + bm - 1) / bm;
  w.parts_bytes = (size_t)w.n_tiles * L.dev.n_parts * bm * sizeof(double);
  w.ngal_bytes = (size_t)w.n_tiles * 2 * bm * sizeof(double);
  w.total = w.parts_bytes + w.ngal_bytes;
  return w;
}

// Occupation items per n-tile: about `items` / nt group ranges, split between centrals and
// satellites in proportion to their groups (at least one each where the type exists).  Auto
// tables: 14 items per tile, taken in the gaps of the contraction.  Cross tables are bound by
// the occupation arithmetic itself (2 R N flops per draw), so the items are what the 12 warps
// balance: 48 smaller ones (measured: N = 240 0.84 -> 0.71 ms per 1e5 draws, N = 1104 3.12 -> 3.04).
constexpr int kOccItemsPerTile = 14;
constexpr int kOccItemsPerTileCross = 48;
// series items of cross tables: 28 (measured: N = 1104 2.68 -> 1.94 ms per 1e5 draws against 48,
// whose draw pieces shrink to one draw per item; N = 240 0.60 -> 0.57 ms)
constexpr int kOccSeriesItemsPerTileCross = 28;

void pick_ranges(const OccPlan& plan, int nt, int* n_cen, int* n_sat,
                 int items = kOccItemsPerTile) {
  const int cen = plan.n_cen_groups, sat = plan.n_groups - plan.n_cen_groups;
  const int want = std::max(2, (tune("OCC_ITEMS", items) + nt - 1) / nt);
  auto share = [&](int count) {
    if (count == 0) return 0;
    const int units = (count + 3) / 4;
    return std::max(1, std::min(units, (int)std::lround((double)want * count / (cen + sat))));
  };
  *n_cen = share(cen);
  *n_sat = share(sat);
  if (*n_cen + *n_sat == 0) *n_cen = 1;  // table without rows cannot happen; keep n_occ > 0
}

// Items of an 8-draw block in series mode (series_item): per galaxy type `pieces` draw pieces
// (a power of two <= 8) times group ranges.  A batch cuts the draws (a warp iteration is one draw
// x 32 groups, so the per-draw constants are amortised over all groups of the type); a one-draw
// call cuts the groups into ranges of 32 so that the warps of the block share the draw.
void pick_series_ranges(const OccPlan& plan, int nt, long long n_draws, int items, int* n_cen,
                        int* n_sat, int* pieces_cen, int* pieces_sat, bool latency = false) {
  const int cen = plan.n_cen_groups, sat = plan.n_groups - plan.n_cen_groups;
  const int want = std::max(2, (tune("OCC_ITEMS", items) + nt - 1) / nt);
  auto share = [&](int count, int* ranges, int* pieces) {
    *ranges = 0;
    *pieces = 1;
    if (count == 0) return;
    if (n_draws == 1) {
      *ranges = std::min(8, (count + 31) / 32);
      return;
    }
    if (latency) {
      // a few dozen draws are a latency problem: one draw and 32 groups per item, so that an
      // item is a single chain of ~150 dependent operations and the 12 warps of a block share
      // the tile (measured at N = 240: 8 / 32 draws per call 88 / 84 us with 8-draw items, 76 /
      // 78 us so -- the node items of round 1: 75 / 74 us; from 128 draws on the larger items win)
      *pieces = 8;
      *ranges = 8 * std::min(8, (count + 31) / 32);
      return;
    }
    const int s = std::max(1, (int)std::lround((double)want * count / (cen + sat)));
    int p = 1;
    while (2 * p <= std::min(8, s)) p *= 2;
    *pieces = p;
    *ranges = p;
  };
  share(cen, n_cen, pieces_cen);
  share(sat, n_sat, pieces_sat);
  if (*n_cen + *n_sat == 0) *n_cen = 1;  // table without rows cannot happen; keep n_occ > 0
}

// Coefficient tables of the occupation math (see half_erfc_neg / pow_pos), computed in long double.
int ensure_math_tables(int device) {
  static std::mutex m;
  static std::map<int, bool> done;
  std::lock_guard<std::mutex> lock(m);
  if (done[device]) return TC_OK;
  std::vector<double> tab(kTabDoubles, 0.0);
  const int n = kErfDeg + 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  tab[kErfIntervals - 1] = 1.0;  // saturated columns: 0 below the first, 1 above the last interval
  for (int i = 0; i < kErfIntervals - 2; i++) {
    const long double xc = -6.0L + 0.5L * i;  // x = xc + s / 4 with s in [-1, 1]; t = s / 2
    std::vector<long double> fs(n), sn(n);
    for (int j = 0; j < n; j++) {
      sn[j] = cosl(pi * (j + 0.5L) / n);
      fs[j] = 0.5L * erfcl(-(xc + 0.25L * sn[j]));
    }
    // Chebyshev coefficients of the interpolant, then Chebyshev -> monomial in s
    std::vector<long double> a(n, 0.0L);
    for (int k = 0; k < n; k++) {
      long double sum = 0.0L;
      for (int j = 0; j < n; j++) sum += fs[j] * cosl(k * pi * (j + 0.5L) / n);
      a[k] = (k == 0 ? 1.0L : 2.0L) * sum / n;
    }
    std::vector<long double> mono(n, 0.0L), t0(n, 0.0L), t1(n, 0.0L), t2(n, 0.0L);
    t0[0] = 1.0L;                 // T_0
    t1[1] = 1.0L;                 // T_1
    for (int d = 0; d < n; d++) mono[d] += a[0] * t0[d] + (n > 1 ? a[1] * t1[d] : 0.0L);
    for (int k = 2; k < n; k++) {  // T_k = 2 s T_{k-1} - T_{k-2}
      for (int d = 0; d < n; d++) t2[d] = (d > 0 ? 2.0L * t1[d - 1] : 0.0L) - t0[d];
      for (int d = 0; d < n; d++) mono[d] += a[k] * t2[d];
      t0 = t1;
      t1 = t2;
    }
    long double scale = 1.0L;    // s = 2 t
    for (int d = 0; d < n; d++) {
      tab[(size_t)d * kErfStride + i + 1] = (double)(mono[d] * scale);
      scale *= 2.0L;
    }
  }
  for (int i = 0; i < kLogEntries; i++) {
    const double inv_c = (double)(1.0L / (1.0L + (i + 0.5L) / kLogEntries));
    tab[kTabLog + 2 * i] = inv_c;
    tab[kTabLog + 2 * i + 1] = (double)(-logl((long double)inv_c));
  }
  for (int j = 0; j < kExpEntries; j++) tab[kTabExp + j] = (double)exp2l((long double)j / kExpEntries);
  TC_CUDA(cudaMemcpyToSymbol(g_math_tables, tab.data(), sizeof(double) * kTabDoubles));
  {
    // wide erf table (half_erfc_neg_group): degree-19 fits over centre +- 0.75, monomials in
    // t = x - centre
    std::vector<double> wide(kL11TabDoubles, 0.0);
    const int nw = kErfWDeg + 1;
    wide[kErfWPoly + 1] = 1.0;   // saturated columns: 0 below the first, 1 above the last interval
    for (int i = 0; i < kErfWPoly; i++) {
      const long double xc = -7.0L + 0.5L * i, half = (long double)kErfWHalf;
      std::vector<long double> fs(nw), a(nw, 0.0L);
      for (int j = 0; j < nw; j++) fs[j] = 0.5L * erfcl(-(xc + half * cosl(pi * (j + 0.5L) / nw)));
      for (int k = 0; k < nw; k++) {
        long double sum = 0.0L;
        for (int j = 0; j < nw; j++) sum += fs[j] * cosl(k * pi * (j + 0.5L) / nw);
        a[k] = (k == 0 ? 1.0L : 2.0L) * sum / nw;
      }
      std::vector<long double> mono(nw, 0.0L), t0(nw, 0.0L), t1(nw, 0.0L), t2(nw, 0.0L);
      t0[0] = 1.0L;
      t1[1] = 1.0L;
      for (int d = 0; d < nw; d++) mono[d] += a[0] * t0[d] + a[1] * t1[d];
      for (int k = 2; k < nw; k++) {  // T_k = 2 s T_{k-1} - T_{k-2}
        for (int d = 0; d < nw; d++) t2[d] = (d > 0 ? 2.0L * t1[d - 1] : 0.0L) - t0[d];
        for (int d = 0; d < nw; d++) mono[d] += a[k] * t2[d];
        t0 = t1;
        t1 = t2;
      }
      long double scale = 1.0L;    // s = t / half
      for (int d = 0; d < nw; d++) {
        wide[(size_t)d * kErfWStride + i + 1] = (double)(mono[d] * scale);
        scale /= half;
      }
    }
    for (int j = 0; j < kExpEntries; j++) wide[kErfWDoubles + j] = tab[kTabExp + j];
    TC_CUDA(cudaMemcpyToSymbol(g_l11_tables, wide.data(), sizeof(double) * kL11TabDoubles));
  }
  done[device] = true;
  return TC_OK;
}

int device_sms(int device, int* n_sm) {
  static std::mutex m;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lock(m);
  auto it = cache.find(device);
  if (it == cache.end()) {
    int v = 0;
    TC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    it = cache.emplace(device, v).first;
  }
  *n_sm = it->second;
  return TC_OK;
}

template <int NT, int MODE, bool SERIES>
int launch_predict_as(const PredictArgs& args, dim3 grid, size_t smem, cudaStream_t stream) {
  static std::mutex m;
  static std::map<int, bool> configured;
  int dev = 0;
  TC_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(m);
    if (!configured[dev]) {
      TC_CUDA(cudaFuncSetAttribute(predict_kernel<NT, MODE, SERIES>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
      configured[dev] = true;
    }
  }
  predict_kernel<NT, MODE, SERIES><<<grid, kThreads, smem, stream>>>(args);
  TC_CUDA(cudaGetLastError());
  return TC_OK;
}

// args.pieces_cen != 0: series items (pick_series_ranges), else node-by-node items (pick_ranges)
template <int NT, int MODE>
int launch_predict(const PredictArgs& args, dim3 grid, size_t smem, cudaStream_t stream) {
  return args.pieces_cen != 0 ? launch_predict_as<NT, MODE, true>(args, grid, smem, stream)
                              : launch_predict_as<NT, MODE, false>(args, grid, smem, stream);
}

// Output blocks (gridDim.y) and dynamic shared memory of finalize_kernel: the staging tile of a
// block (bm draws x outputs of the block) stays below 40 KB; more blocks when few tiles would
// leave SMs idle.
void finalize_grid(int n_out, int bm, long long n_tiles, int n_sm, int* fy_out, int* smem_out) {
  const int max_o = std::max(1, 40000 / (8 * bm) - 1);
  int fy = (n_out + max_o - 1) / max_o;
  if (n_tiles <= 4LL * n_sm) {
    const int outs_per_pass = std::max(1, 256 / bm);
    fy = std::max(fy, std::min(64, (n_out + outs_per_pass - 1) / outs_per_pass));
  }
  fy = std::max(1, std::min(fy, n_out));
  const int o_per_block = (n_out + fy - 1) / fy;
  *fy_out = fy;
  *smem_out = bm * (o_per_block | 1) * (int)sizeof(double);
}

// optional per-kernel timing for bench.py (tc_profile_enable / tc_profile_read)
struct Profile {
  bool enabled = false;
  bool recorded = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};
Profile g_profile;

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; (void)cudaGetLastError(); return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) { ok = false; (void)cudaGetLastError(); }
  }
  ~DeviceGuard() { if (prev >= 0) (void)cudaSetDevice(prev); }
};

}  // namespace
