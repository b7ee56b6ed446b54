// Packed fixed-point arithmetic policies for the max-log-MAP kernels.
//
// Every trellis metric register holds TWO sub-block lanes of one code block as int16x2 (lanes 2j, 2j+1 of the
// reference's SIMD vector, which are adjacent int16 elements of its lane-interleaved arrays).  A policy fixes
// what the reference's simd_add / simd_sub / simd_max / normalize mean for one decoder family:
//
//   Sat16  : tdec_win{sse16,avx16}  -- _mm*_adds/subs/max_epi16, INF 10000, normalise by state 0 every 2nd step
//            (turbodecoder_win.h:43-56, 75-79, 148-151, 480-498)
//   Sat8   : tdec_win{sse8,avx8}    -- _mm*_adds/subs/max_epi8 carried in int16 containers (sm_100a has no
//            packed 8x4 min/max/saturating-add hardware; VIADD/VIMNMX.S16x2 + clamp to [-128,127] is exact),
//            INF 0, normalise by the max every step, output >>1  (turbodecoder_win.h:170-186, 213-293)
//
// The header is also compiled by g++ (tests/host_emul) with scalar emulations of the same operations.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

typedef uint32_t u32;

// ---------------------------------------------------------------- scalar helpers on int16x2
B200_HD int32_t lo16(u32 v) { return (int32_t)(int16_t)(uint16_t)(v & 0xffffu); }
B200_HD int32_t hi16(u32 v) { return (int32_t)(int16_t)(uint16_t)(v >> 16); }
B200_HD u32 pack16(int32_t lo, int32_t hi) { return ((u32)(uint16_t)lo) | (((u32)(uint16_t)hi) << 16); }
B200_HD int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------- packed primitives
#if defined(__CUDA_ARCH__)
B200_HD u32 p_add_wrap(u32 a, u32 b) { return __vadd2(a, b); }
B200_HD u32 p_sub_wrap(u32 a, u32 b) { return __vsub2(a, b); }
B200_HD u32 p_add_sat(u32 a, u32 b) { return __vaddss2(a, b); }
B200_HD u32 p_sub_sat(u32 a, u32 b) { return __vsubss2(a, b); }
B200_HD u32 p_max(u32 a, u32 b) { return __vmaxs2(a, b); }
B200_HD u32 p_min(u32 a, u32 b) { return __vmins2(a, b); }
B200_HD u32 p_addmax(u32 a, u32 b, u32 c) { return __viaddmax_s16x2(a, b, c); } // max(a+b (wrapping), c)
B200_HD u32 p_max3(u32 a, u32 b, u32 c) { return __vimax3_s16x2(a, b, c); }
B200_HD u32 p_min3(u32 a, u32 b, u32 c) { return __vimin3_s16x2(a, b, c); }
#else
B200_HD u32 p_add_wrap(u32 a, u32 b) { return pack16(lo16(a) + lo16(b), hi16(a) + hi16(b)); }
B200_HD u32 p_sub_wrap(u32 a, u32 b) { return pack16(lo16(a) - lo16(b), hi16(a) - hi16(b)); }
B200_HD u32 p_add_sat(u32 a, u32 b)
{
  return pack16(clampi(lo16(a) + lo16(b), -32768, 32767), clampi(hi16(a) + hi16(b), -32768, 32767));
}
B200_HD u32 p_sub_sat(u32 a, u32 b)
{
  return pack16(clampi(lo16(a) - lo16(b), -32768, 32767), clampi(hi16(a) - hi16(b), -32768, 32767));
}
B200_HD u32 p_max(u32 a, u32 b)
{
  return pack16(lo16(a) > lo16(b) ? lo16(a) : lo16(b), hi16(a) > hi16(b) ? hi16(a) : hi16(b));
}
B200_HD u32 p_min(u32 a, u32 b)
{
  return pack16(lo16(a) < lo16(b) ? lo16(a) : lo16(b), hi16(a) < hi16(b) ? hi16(a) : hi16(b));
}
B200_HD u32 p_addmax(u32 a, u32 b, u32 c) { return p_max(p_add_wrap(a, b), c); }
B200_HD u32 p_max3(u32 a, u32 b, u32 c) { return p_max(p_max(a, b), c); }
B200_HD u32 p_min3(u32 a, u32 b, u32 c) { return p_min(p_min(a, b), c); }
#endif

B200_HD u32 splat16(int32_t v) { return pack16(v, v); }

// ---------------------------------------------------------------- policies
struct Sat16 {
  static constexpr int  kBits    = 16;
  static constexpr int  kInf     = 10000;
  static constexpr bool kNormMax = false;
  static constexpr bool kMonitor = false;
  B200_HD static u32 add(u32 a, u32 b) { return p_add_sat(a, b); }
  B200_HD static u32 sub(u32 a, u32 b) { return p_sub_sat(a, b); }
  B200_HD static u32 max(u32 a, u32 b) { return p_max(a, b); }
  // max(a + b, c) with the reference's saturating add
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return p_max(p_add_sat(a, b), c); }
  // max(a + b, c + d), both sums saturating
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return p_max(p_add_sat(a, b), p_add_sat(c, d)); }
  // running maximum of a-posteriori sums (fwd_step_llr): first term, further terms, final value
  B200_HD static u32 sum0(u32 a, u32 b) { return add(a, b); }
  B200_HD static u32 summax(u32 a, u32 b, u32 c) { return addmax(a, b, c); }
  B200_HD static u32 sumfin(u32 m) { return m; }
  B200_HD static u32 cand(u32 a, u32 g) { return add(a, g); } // path metric + branch metric (fwd_step_llr)
  static constexpr bool kNeedsFix = false;
  B200_HD static void fix(u32 (&)[8]) {}
  static constexpr int kNormPeriod = 2;
  B200_HD static void normalize_now(u32 (&o)[8])
  {
#pragma unroll
    for (int i = 1; i < 8; i++)
      o[i] = p_sub_sat(o[i], o[0]);
    o[0] = 0;
  }
  // normalize(k) of turbodecoder_win.h:480-498 (period 2, by state 0)
  B200_HD static void normalize(uint32_t k, u32 (&o)[8])
  {
    if ((k & 1u) == 0 && k != 0)
      normalize_now(o);
  }
  B200_HD static u32 out(u32 llr) { return llr; }
  // scalar tail-trellis adder (turbodecoder_win.h:470-478): plain C int16 addition -> wraps
  B200_HD static int32_t tail_add(int32_t a, int32_t b) { return (int32_t)(int16_t)(uint16_t)(a + b); }
  // srslte_vec_sub_sss: wrapping (vector_simd.c:132-160, simd.h:1592-1604)
  B200_HD static u32 glue_sub(u32 a, u32 b, bool, bool) { return p_sub_wrap(a, b); }
};

// Fast16: the SAME int16 decoder computed with sm_100a's native packed instructions only (VIADD.16x2,
// VIMNMX.S16x2, VIADDMNMX.S16x2 -- wrapping adds).  Wrapping and saturating arithmetic agree as long as no
// intermediate leaves the int16 range; MapWin tracks the range of the path metrics while it runs (kMonitor) and
// the kernel REPLAYS a code block with the exact Sat16 policy whenever the tracked ranges cannot rule a
// saturation out.  The results delivered are therefore always those of the saturating reference decoder.
struct Fast16 {
  static constexpr int  kBits    = 16;
  static constexpr int  kInf     = 10000;
  static constexpr bool kNormMax = false;
  static constexpr bool kMonitor = true;
  B200_HD static u32 add(u32 a, u32 b) { return p_add_wrap(a, b); }
  B200_HD static u32 sub(u32 a, u32 b) { return p_sub_wrap(a, b); }
  B200_HD static u32 max(u32 a, u32 b) { return p_max(a, b); }
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return p_addmax(a, b, c); }
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return p_addmax(a, b, p_add_wrap(c, d)); }
  // running maximum of a-posteriori sums (fwd_step_llr): first term, further terms, final value
  B200_HD static u32 sum0(u32 a, u32 b) { return add(a, b); }
  B200_HD static u32 summax(u32 a, u32 b, u32 c) { return addmax(a, b, c); }
  B200_HD static u32 sumfin(u32 m) { return m; }
  B200_HD static u32 cand(u32 a, u32 g) { return add(a, g); } // path metric + branch metric (fwd_step_llr)
  static constexpr bool kNeedsFix = false;
  B200_HD static void fix(u32 (&)[8]) {}
  static constexpr int kNormPeriod = 2;
  B200_HD static void normalize_now(u32 (&o)[8])
  {
#pragma unroll
    for (int i = 1; i < 8; i++)
      o[i] = p_sub_wrap(o[i], o[0]);
    o[0] = 0;
  }
  B200_HD static void normalize(uint32_t k, u32 (&o)[8])
  {
    if ((k & 1u) == 0 && k != 0)
      normalize_now(o);
  }
  B200_HD static u32 out(u32 llr) { return llr; }
  B200_HD static int32_t tail_add(int32_t a, int32_t b) { return (int32_t)(int16_t)(uint16_t)(a + b); }
  B200_HD static u32 glue_sub(u32 a, u32 b, bool, bool) { return p_sub_wrap(a, b); }
};

struct Sat8 {
  static constexpr int  kBits    = 8;
  static constexpr int  kInf     = 0;
  static constexpr bool kNormMax = true;
  static constexpr bool kMonitor = false;
  B200_HD static u32 clamp8(u32 v) { return p_min(p_max(v, 0xff80ff80u), 0x007f007fu); }
  B200_HD static u32 add(u32 a, u32 b) { return p_min(p_addmax(a, b, 0xff80ff80u), 0x007f007fu); }
  B200_HD static u32 sub(u32 a, u32 b) { return clamp8(p_sub_wrap(a, b)); }
  B200_HD static u32 max(u32 a, u32 b) { return p_max(a, b); }
  // max(clamp8(a + b), c) for c already in [-128, 127] (every metric is): the lower clamp is absorbed by the max with c
  // and the upper clamp commutes with it, so one fused add-max and one min give the same integers as add() + max()
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return p_min(p_addmax(a, b, c), 0x007f007fu); }
  // max(clamp8(a + b), clamp8(c + d)) = clamp8(max(a + b, c + d)) (the clamp is monotone; the int16 containers hold the
  // unclamped sums exactly): one wrapping add on the add-type pipe, a fused add-max and one clamp
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return clamp8(p_addmax(a, b, p_add_wrap(c, d))); }
  // max_i clamp8(b_i + c_i) = clamp8(max_i (b_i + c_i)): the clamp is monotone and the int16 containers hold the sum of two
  // int8-range values exactly, so the chain of a-posteriori sums runs unclamped (one fused add-max per term instead of
  // add-max + min) and is clamped once at the end
  B200_HD static u32 sum0(u32 a, u32 b) { return p_add_wrap(a, b); }
  B200_HD static u32 summax(u32 a, u32 b, u32 c) { return p_addmax(a, b, c); }
  B200_HD static u32 sumfin(u32 m) { return clamp8(m); }
  B200_HD static u32 cand(u32 a, u32 g) { return add(a, g); }
  static constexpr bool kNeedsFix = false;
  B200_HD static void fix(u32 (&)[8]) {}
  static constexpr int kNormPeriod = 1;
  B200_HD static void normalize_now(u32 (&o)[8])
  {
    u32 m = p_max3(o[0], o[1], o[2]);
    m     = p_max3(m, o[3], o[4]);
    m     = p_max3(m, o[5], o[6]);
    m     = p_max(m, o[7]);
    // o[i] - m lies in [-255, 0]: only the lower bound of the saturating subtract can act
    const u32 nm = p_sub_wrap(0u, m);
#pragma unroll
    for (int i = 0; i < 8; i++)
      o[i] = p_addmax(o[i], nm, 0xff80ff80u);
  }
  // normalize_max, period 1 (turbodecoder_win.h:180-181, 483-490)
  B200_HD static void normalize(uint32_t k, u32 (&o)[8])
  {
    if (k != 0)
      normalize_now(o);
  }
  // divide_output: per-element arithmetic >> 1 (turbodecoder_win.h:188-193, 811-813)
  // (both halves at once: bits 14..0 of each half come from the logical shift, the sign bit is put back)
  B200_HD static u32 out(u32 llr) { return ((llr >> 1) & 0x7fff7fffu) | (llr & 0x80008000u); }
  // int16 z = x + y; z > 127 ? 127 : (int8_t) z   (saturates upwards only)
  B200_HD static int32_t tail_add(int32_t a, int32_t b)
  {
    int32_t z = a + b;
    return z > 127 ? 127 : (int32_t)(int8_t)(uint8_t)z;
  }
  // srslte_vec_sub_bbb on the AVX2 build: _mm256_subs_epi8 for array index < floor(K/32)*32, plain C
  // (wrapping) for the rest (vector_simd.c:162-190); chosen per element
  B200_HD static u32 glue_sub(u32 a, u32 b, bool sat_lo, bool sat_hi)
  {
    const u32 d = p_sub_wrap(a, b);
    const u32 s = clamp8(d);
    const u32 w = p_sub_wrap(((d & 0x00ff00ffu) ^ 0x00800080u), 0x00800080u); // sign-extend the low byte
    return (sat_lo ? (s & 0xffffu) : (w & 0xffffu)) | (sat_hi ? (s & 0xffff0000u) : (w & 0xffff0000u));
  }
};

// Sat8 for trellis steps whose INPUT path metrics are normalised (normalize_max just ran: every metric <= 0, turbodecoder_win.h
// :483-490 -- with period 1 that is every step of the int8 decoders but a handful).  A metric <= 0 plus a branch metric <= 127
// cannot exceed 127, so the upper half of every saturation is dead and only the lower one is left:
//   max(sat(a + g), c)          = max(a + g, c)                  (c >= -128 absorbs the lower clamp)       1 max-type op
//   max(sat(a + y), sat(c + x)) = max3(a + y, c + x, -128)       (two adds on the add-type pipe)           1 max-type op
//   sat(a + g)                  = max(a + g, -128)                                                         1 max-type op
// against 2 / 3 / 2 with both clamps: 8 instead of 20 max-type instructions per recursion step, 12 fewer per output step (the
// max-type pipe is what bounds the int8 kernel).  The few steps whose input is NOT normalised -- the first step after a state
// was handed over or loaded un-normalised, the step after the un-normalised step 0 -- are repaired exactly by fix(): the
// containers are int16, so the un-clamped sums are exact and min(., 127) afterwards is the saturated result.
struct Sat8F : Sat8 {
  static constexpr bool kNeedsFix = true;
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return p_addmax(a, b, c); }
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return p_max3(p_add_wrap(a, b), p_add_wrap(c, d), 0xff80ff80u); }
  B200_HD static u32 cand(u32 a, u32 g) { return p_addmax(a, g, 0xff80ff80u); }
  B200_HD static void fix(u32 (&o)[8])
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
      o[i] = p_min(o[i], 0x007f007fu);
  }
};
// the policy the recursion steps of a kernel run with, given the policy of its arithmetic
template <class P>
struct StepPolicy {
  using type = P;
};
template <>
struct StepPolicy<Sat8> {
  using type = Sat8F;
};

} // namespace b200
