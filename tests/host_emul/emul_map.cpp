// CPU emulation of the CUDA max-log-MAP schedule (srsran_b200/csrc/map_core.cuh compiled by g++): runs the
// T = N/2 "threads" of one code block phase by phase, doing the warp-shuffle lane exchange by hand.  Lets the
// checkpoint/recompute schedule and the packed arithmetic be checked against the oracle without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../srsran_b200/csrc/map_core.cuh"

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
