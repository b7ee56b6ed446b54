// CPU emulation of the CUDA max-log-MAP schedule (srsran_b200/csrc/map_core.cuh compiled by g++): runs the
// T = N/2 "threads" of one code block phase by phase, doing the warp-shuffle lane exchange by hand.  Lets the
// checkpoint/recompute schedule and the packed arithmetic be checked against the oracle without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../srsran_b200/csrc/map_f16.cuh"

using namespace b200;

// returns 1 if (Fast16 only) any lane's range monitor could not rule out a saturation
template <class P, int L>
static int run(int N, int K, const int16_t* in, const int16_t* apr, const int16_t* par, int16_t* out, int g = 0, int* stats = nullptr)
{
  const int T = N / 2, W = K / N, S = (W + L - 1) / L;
  std::vector<std::vector<u32>> ck(T, std::vector<u32>((size_t)(S + 1) * 8));
  std::vector<MapWin<P, L, DirectSrc>> m(T);
  std::vector<u32>              st((size_t)T * 8);
  for (int j = 0; j < T; j++) {
    m[j].src     = DirectSrc{(const u32*)in, (const u32*)apr, (const u32*)par, T, j, ck[j].data(), nullptr};
    m[j].W       = W;
    m[j].begin();
  }
  auto S8 = [&](int j) -> u32(&)[8] { return *reinterpret_cast<u32(*)[8]>(&st[(size_t)j * 8]); };
  for (int j = 0; j < T; j++)
    m[j].beta_warm(S8(j));
  {
    std::vector<u32> old = st;
    for (int j = 0; j < T; j++)
      for (int s = 0; s < 8; s++) {
        u32 nx          = j + 1 < T ? old[(size_t)(j + 1) * 8 + s] : old[(size_t)j * 8 + s];
        st[(size_t)j * 8 + s] = shift_down_lanes(old[(size_t)j * 8 + s], nx);
      }
    int32_t t[8];
    tail_trellis<P>(in + K, par + K, t);
    for (int s = 0; s < 8; s++)
      st[(size_t)(T - 1) * 8 + s] = (st[(size_t)(T - 1) * 8 + s] & 0xffffu) | ((u32)(uint16_t)t[s] << 16);
  }
  for (int j = 0; j < T; j++)
    m[j].beta_main(S8(j));
  for (int j = 0; j < T; j++)
    m[j].alpha_warm(S8(j));
  {
    std::vector<u32> old = st;
    for (int j = 0; j < T; j++)
      for (int s = 0; s < 8; s++) {
        u32 pv          = j > 0 ? old[(size_t)(j - 1) * 8 + s] : old[(size_t)j * 8 + s];
        st[(size_t)j * 8 + s] = shift_up_lanes(pv, old[(size_t)j * 8 + s]);
      }
    st[0] &= 0xffff0000u;
    for (int s = 1; s < 8; s++)
      st[s] = (st[s] & 0xffff0000u) | (u32)(uint16_t)(-P::kInf);
  }
  for (int j = 0; j < T; j++) {
    u32* o   = (u32*)out;
    auto epi = [&](int p, u32 llr, u32, u32, u32) { o[p * T + j] = llr; };
    m[j].alpha_main(S8(j), epi);
  }
  int flagged = 0;
  if (P::kMonitor) {
    int mx_a = 0, mx_b = 0;
    for (int j = 0; j < T; j++) {
      const int sa[2] = {m[j].mon_a.spread_lo(), m[j].mon_a.spread_hi()}, sb[2] = {m[j].mon_b.spread_lo(), m[j].mon_b.spread_hi()};
      for (int h = 0; h < 2; h++) {
        if (!fast16_beta_ok(sb[h], g) || !fast16_alpha_ok(sa[h], sb[h], g))
          flagged = 1;
        mx_a = sa[h] > mx_a ? sa[h] : mx_a;
        mx_b = sb[h] > mx_b ? sb[h] : mx_b;
      }
      if (m[j].mon_a.ovf & 0x80008000u)
        flagged = 1;
    }
    if (stats) {
      stats[0] = mx_a;
      stats[1] = mx_b;
    }
  }
  return flagged;
}

// The arithmetic of k_map_f16 (srsran_b200/csrc/map_f16.cuh) for one code block: same sequence of packed operations,
// same tracking points, the factored LLR, the head monitor of the first four forward steps -- only the memory staging
// and the checkpoint/recompute of the beta values (deterministic, checked on the device) are replaced by a plain array.
// Returns 1 if a range monitor could not rule out a saturation.
static int run_f16(int N, int K, const int16_t* in, const int16_t* apr, const int16_t* par, int16_t* out, int g)
{
  using P = Fast16;
  const int  T = N / 2, W = K / N;
  const u32 *in32 = (const u32*)in, *apr32 = (const u32*)apr, *par32 = (const u32*)par;
  auto row = [&](int p, int j, u32& x, u32& y) {
    y = par32[p * T + j];
    x = apr32 ? P::add(apr32[p * T + j], in32[p * T + j]) : in32[p * T + j];
  };
  std::vector<u32>      st((size_t)T * 8);
  std::vector<RangeMon> mon_b(T), mon_a(T), mon_h(T);
  for (int j = 0; j < T; j++) {
    mon_b[j].reset();
    mon_a[j].reset();
    mon_h[j].reset();
  }
  auto S8 = [&](int j) -> u32(&)[8] { return *reinterpret_cast<u32(*)[8]>(&st[(size_t)j * 8]); };
  // backward warm-up
  for (int j = 0; j < T; j++) {
    u32(&s)[8] = S8(j);
    for (int i = 0; i < 8; i++)
      s[i] = splat16(-P::kInf);
    for (int k = kWinOverlap - 1; k >= 0; k--) {
      u32 x, y;
      row(k, j, x, y);
      bwd_step<P>(s, x, y, P::add(x, y));
      if ((k & 1) == 0) {
        if (k < kWinOverlap - 2)
          mon_b[j].track(s);
        if (k != 0)
          P::normalize_now(s);
      }
    }
  }
  {
    std::vector<u32> old = st;
    for (int j = 0; j < T; j++)
      for (int s = 0; s < 8; s++) {
        u32 nx                = j + 1 < T ? old[(size_t)(j + 1) * 8 + s] : old[(size_t)j * 8 + s];
        st[(size_t)j * 8 + s] = shift_down_lanes(old[(size_t)j * 8 + s], nx);
      }
    int32_t t[8];
    tail_trellis<P>(in + K, par + K, t);
    for (int s = 0; s < 8; s++)
      st[(size_t)(T - 1) * 8 + s] = (st[(size_t)(T - 1) * 8 + s] & 0xffffu) | ((u32)(uint16_t)t[s] << 16);
  }
  // backward main pass: beta[k] before normalisation
  std::vector<std::vector<u32>> beta(T, std::vector<u32>((size_t)(W + 1) * 8));
  int                           flagged = 0;
  for (int j = 0; j < T; j++) {
    u32(&s)[8] = S8(j);
    mon_b[j].track(s);
    memcpy(&beta[j][(size_t)W * 8], s, 32);
    for (int k = W - 1; k >= 0; k--) {
      u32 x, y;
      row(k, j, x, y);
      bwd_step<P>(s, x, y, P::add(x, y));
      memcpy(&beta[j][(size_t)k * 8], s, 32);
      if ((k & 1) == 0) {
        mon_b[j].track(s);
        if (k != 0)
          P::normalize_now(s);
      }
    }
    if (!fast16_beta_ok(mon_b[j].spread_lo(), g) || !fast16_beta_ok(mon_b[j].spread_hi(), g))
      flagged = 1;
  }
  // forward warm-up (normalisation follows the loop counter)
  for (int j = 0; j < T; j++) {
    u32(&s)[8] = S8(j);
    for (int i = 0; i < 8; i++)
      s[i] = splat16(-P::kInf);
    for (int kk = 0; kk < kWinOverlap; kk++) {
      u32 x, y;
      row(W - kWinOverlap + kk, j, x, y);
      fwd_step<P>(s, x, y, P::add(x, y));
      if ((kk & 1) == 0) {
        if (kk > 2)
          mon_a[j].track(s);
        if (kk != 0)
          P::normalize_now(s);
      }
    }
  }
  {
    std::vector<u32> old = st;
    for (int j = 0; j < T; j++)
      for (int s = 0; s < 8; s++) {
        u32 pv                = j > 0 ? old[(size_t)(j - 1) * 8 + s] : old[(size_t)j * 8 + s];
        st[(size_t)j * 8 + s] = shift_up_lanes(pv, old[(size_t)j * 8 + s]);
      }
    st[0] &= 0xffff0000u;
    for (int s = 1; s < 8; s++)
      st[s] = (st[s] & 0xffff0000u) | (u32)(uint16_t)(-P::kInf);
  }
  // output pass
  u32* o = (u32*)out;
  for (int j = 0; j < T; j++) {
    u32(&al)[8] = S8(j);
    mon_h[j].track(al);
    for (int p = 0; p < W; p++) {
      u32 x, y, b[8];
      row(p, j, x, y);
      memcpy(b, &beta[j][(size_t)(p + 1) * 8], 32);
      const u32 xy  = P::add(x, y);
      const u32 llr = llr_factored<P>(al, b, x, y, xy, mon_a[j]);
      fwd_step<P>(al, x, y, xy);
      if ((p & 1) == 0) {
        mon_a[j].track(al);
        if (p != 0)
          P::normalize_now(al);
      }
      o[p * T + j] = llr;
      if (p == 3) { // what was tracked so far belongs to the head monitor
        mon_h[j].hi = p_max(mon_h[j].hi, mon_a[j].hi);
        mon_h[j].lo = p_min(mon_h[j].lo, mon_a[j].lo);
        mon_a[j].hi = 0;
        mon_a[j].lo = 0;
      }
    }
    const bool bad = !fast16_alpha_ok(mon_a[j].spread_lo(), mon_b[j].spread_lo(), g) || !fast16_alpha_ok(mon_a[j].spread_hi(), mon_b[j].spread_hi(), g) ||
                     !fast16_alpha_ok(mon_h[j].spread_lo(), mon_b[j].spread_lo(), g) || !fast16_alpha_ok(mon_h[j].spread_hi(), mon_b[j].spread_hi(), g) ||
                     ((mon_a[j].ovf | mon_h[j].ovf) & 0x80008000u) != 0;
    if (bad)
      flagged = 1;
  }
  return flagged;
}
extern "C" int emul_map_f16(int N, int K, const int16_t* in, const int16_t* apr, const int16_t* par, int16_t* out, int g)
{
  return run_f16(N, K, in, apr, par, out, g);
}

// Fast16 with monitoring: returns the flag; g = bound on every |branch metric|
extern "C" int emul_map_fast16(int N, int K, const int16_t* in, const int16_t* apr, const int16_t* par, int16_t* out, int g, int* stats)
{
  return run<Fast16, 16>(N, K, in, apr, par, out, g, stats);
}

extern "C" int emul_map_win(int bits, int N, int L, int K, const int16_t* in, const int16_t* apr, const int16_t* par, int16_t* out)
{
  if (bits == 16 && L == 16) run<Sat16, 16>(N, K, in, apr, par, out);
  else if (bits == 16 && L == 8) run<Sat16, 8>(N, K, in, apr, par, out);
  else if (bits == 16 && L == 5) run<Sat16, 5>(N, K, in, apr, par, out);
  else if (bits == 8 && L == 16) run<Sat8, 16>(N, K, in, apr, par, out);
  else if (bits == 8 && L == 8) run<Sat8, 8>(N, K, in, apr, par, out);
  else return -1;
  return 0;
}

// packed glue subtract, exported for a direct check of Sat8::glue_sub / Sat16::glue_sub
extern "C" uint32_t emul_glue_sub(int bits, uint32_t a, uint32_t b, int sat_lo, int sat_hi)
{
  return bits == 16 ? Sat16::glue_sub(a, b, sat_lo, sat_hi) : Sat8::glue_sub(a, b, sat_lo, sat_hi);
}

// Sat8F (arith.cuh): the cheaper step rules for normalised input metrics.  Random trials; returns the number of mismatches:
//   * input metrics normalised (normalize_max: every metric in [-128, 0], the maximum 0): bwd_step / fwd_step / fwd_step_llr
//     with Sat8F give the integers of Sat8;
//   * arbitrary int8-range input metrics: the Sat8F step followed by fix() gives the integers of the Sat8 step.
extern "C" int emul_sat8f_check(uint32_t seed, int n_trials)
{
  uint32_t r   = seed * 2654435761u + 12345u;
  auto     rnd = [&](int lo, int hi) { // uniform in [lo, hi]
    r = r * 1664525u + 1013904223u;
    return lo + (int)((r >> 8) % (uint32_t)(hi - lo + 1));
  };
  auto edge = [&](int lo, int hi) { // the ends of the range come up often
    const int c = rnd(0, 9);
    return c == 0 ? lo : c == 1 ? hi : c == 2 ? lo + 1 : c == 3 ? hi - 1 : rnd(lo, hi);
  };
  int bad = 0;
  for (int it = 0; it < n_trials; it++) {
    const bool norm_in = (it & 1) == 0;
    u32        o[8], b[8];
    for (int i = 0; i < 8; i++) {
      o[i] = norm_in ? pack16(edge(-128, 0), edge(-128, 0)) : pack16(edge(-128, 127), edge(-128, 127));
      b[i] = pack16(edge(-128, 127), edge(-128, 127));
    }
    if (norm_in)
      o[rnd(0, 7)] = 0; // (both halves: the maximum of a normalised vector is 0; the other entries are <= 0 either way)
    const u32 x = pack16(edge(-128, 127), edge(-128, 127)), y = pack16(edge(-128, 127), edge(-128, 127));
    const u32 xy = Sat8::add(x, y);
    for (int kind = 0; kind < 2; kind++) {
      u32 a1[8], a2[8];
      for (int i = 0; i < 8; i++)
        a1[i] = a2[i] = o[i];
      if (kind == 0) {
        bwd_step<Sat8>(a1, x, y, xy);
        bwd_step<Sat8F>(a2, x, y, xy);
      } else {
        fwd_step<Sat8>(a1, x, y, xy);
        fwd_step<Sat8F>(a2, x, y, xy);
      }
      if (!norm_in)
        Sat8F::fix(a2);
      for (int i = 0; i < 8; i++)
        bad += a1[i] != a2[i];
    }
    if (norm_in) {
      u32 a1[8], a2[8];
      for (int i = 0; i < 8; i++)
        a1[i] = a2[i] = o[i];
      RangeMon  nomon;
      const u32 l1 = fwd_step_llr<Sat8>(a1, b, x, y, xy, nomon), l2 = fwd_step_llr<Sat8F>(a2, b, x, y, xy, nomon);
      bad += l1 != l2;
      for (int i = 0; i < 8; i++)
        bad += a1[i] != a2[i];
    }
  }
  return bad;
}
