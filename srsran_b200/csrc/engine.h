// engine.h -- internal C++ interface of the B200 turbo-decode engine (one instance = one GPU + one stream).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/srslte_b200/batch.h"
#include "lte_tables.h"

namespace b200 {

struct CbDev;
struct CbState;
struct TbDev;
struct TbResult;
struct Softbuffer;

void        set_error(const std::string& s);
const char* last_error();

template <typename T>
struct DevBuf { // grow-only device buffer
  T*     ptr = nullptr;
  size_t cap = 0;
  int    reserve(size_t n);
  void   release();
};
template <typename T>
struct PinBuf { // grow-only pinned host buffer
  T*     ptr = nullptr;
  size_t cap = 0;
  int    reserve(size_t n);
  void   release();
};

struct DecSel {
  uint32_t lanes; // 0 = generic decoder
  uint32_t bits;  // arithmetic width
  bool     in_sb; // input arrives in lane layout
};

struct Plan; // defined in engine.cu (needs the device descriptor types)
struct LaunchState;

class Engine
{
public:
  static int create(Engine** out, int device);
  ~Engine();

  int submit_cb_batch(const srslte_b200_cb_batch_t* cfg, const void* llr, uint8_t* out, uint32_t flags);
  int submit_tb_batch(srslte_b200_tb_t* tbs, uint32_t nof_tb, int is8, uint32_t max_iterations, uint32_t flags);
  int wait();
  int   timer_start();   // records a CUDA event on the engine's stream (after draining it)
  float timer_stop_ms(); // records the closing event, waits for it, returns the device time between the two

  int softbuffer_create(Softbuffer** out, uint32_t max_cb);
  int demod_descramble(const srslte_b200_demod_t* cws, uint32_t nof_cw, int is8, uint32_t flags);
  int encode_tbs(srslte_b200_enc_t* tbs, uint32_t nof_tb, uint32_t flags);
  int ulsch_deinterleave(const srslte_b200_ulsch_t* tbs, uint32_t nof_tb, uint32_t flags);
  // host-pointer compatibility operations behind the drop-in srslte_* symbols (api.inc)
  int tdec_step(uint32_t K, uint32_t in_bits, uint32_t dec_type, bool force_not_sb, const void* input, uint32_t n_done, uint32_t n_more,
                uint8_t* out);
  int dematch_host(const void* in, void* out, uint32_t E, uint32_t cb_idx, uint32_t rv, uint32_t lanes, int bits);
  int crc_host(const uint8_t* data, uint32_t nbytes, const uint64_t* table, int order, uint32_t poly, uint64_t* out);

  // single code block sessions used by the srslte_tdec_* drop-in symbols (api.cu)
  int select_decoder(uint32_t K, uint32_t in_bits, uint32_t dec_type, bool force_not_sb, DecSel* s);
  void fill_geometry(CbDev* d, uint32_t K, const DecSel& s);
  int run(Plan& p);        // build_plan + launch_plan
  int build_plan(Plan& p);
  int launch_plan();
  int      debug_read_plane(uint32_t cb, uint32_t plane, int16_t* out, uint32_t n);
  uint32_t map_seg_len() const; // trellis steps per beta segment of the MAP kernel variant in use

  int          device  = 0;
  int          num_sms = 0;
  cudaStream_t stream  = nullptr;

  float    last_gpu_ms       = 0;
  float    last_map_ms       = 0;
  uint32_t last_launches     = 0;
  uint32_t last_map_launches = 0;

  // device tables
  DevBuf<uint16_t> d_qpp, d_rm;
  uint32_t         qpp_off[4][kNofCbSizes];
  uint32_t         rm_off[4][kNofCbSizes];
  uint32_t         rm_start[4][kNofCbSizes][4];

  // per-batch device state
  DevBuf<CbDev>    d_cbs;
  DevBuf<CbState>  d_state;
  DevBuf<TbDev>    d_tbs;
  DevBuf<TbResult> d_res;
  DevBuf<int16_t>  d_ws, d_tails, d_sb;
  DevBuf<uint8_t>  d_cbout, d_in, d_tbout;
  DevBuf<int>      d_lists, d_gmax;
  bool             opt_fast16 = true; // try the native packed-instruction path first (exact replay on range alarm)
  bool             opt_latency = true; // small batches: the 4-warp latency-shaped MAP kernel (map_lat.cuh) instead of one warp per group
  DevBuf<uint32_t> d_genbeta, d_counters, d_ckscratch, d_crctab;
  DevBuf<int>      d_parked; // fused kernel: groups whose blocks the Fast16 monitor parked for the exact-arithmetic launch
  DevBuf<int32_t>  d_stat;    // per code block: difficulty estimate of k_cb_stat
  DevBuf<int32_t>  d_scanacc; // ... and the range-monitor accumulators of their groups
  DevBuf<int32_t>  d_scan;   // time-parallel latency kernels (map_scan.cuh): transfer matrices, boundary states, monitor records
  int              opt_scan_cpg = 0;        // k_scan_fused: CTAs per group (0: chosen from the shape)
  bool             opt_scan_launch = false; // the per-half-iteration pair k_scan_mat + k_scan_out where k_scan_fused does not apply (transport blocks): off, k_map_lat's fewer launches win there
  bool             opt_scan_fused = true; // ... and, for a batch that is one run_all class, all half-iterations in ONE cooperative launch (k_scan_fused)
  bool             opt_scan = true;  // int16 classes of at most kScanMaxGroups groups: k_scan_mat + k_scan_out instead of k_map_lat
  DevBuf<int>      d_queue;  // fused kernel, time-sliced classes: groups handed back by their warp + one parked-list flag per group
  int              opt_fused_warps = 0;  // warps per CTA of the fused kernel (0: chosen per batch)
  int              opt_l2_persist = 0;   // experiment: 1 = L2 access policy window (persisting) over the checkpoint scratch, 2 = and no per-instruction hints on them
  int              opt_fused_slice = 21; // classes with CRC early stop: half-iterations per visit of a group, 10 x first + later (0: a group stays with its warp)
  bool             opt_fused = true; // large batches: one persistent launch per decoder class (map_fused.cuh)
  bool             opt_gen_fused = true; // generic decoder, K <= kGenFusedMaxK: one CTA per pair of blocks for all half-iterations (map_gen_fused.cuh)
  int              opt_auto_group = 1;   // transport-block batches without caller hints: blocks of a size ordered by a noise estimate taken from their e-bits (k_cb_stat / k_regroup)
  int              opt_fused_spread = 1; // fused kernel: smaller CTAs when a class has fewer groups than 12 per SM, so every SM gets some
  PinBuf<uint32_t> h_counters;
  uint32_t         last_redo = 0, last_half_iter = 0;
  PinBuf<uint8_t>  h_stage_in, h_stage_out, h_desc, h_tmaps;
  DevBuf<uint8_t>  d_tmaps;
  DevBuf<uint8_t>  d_dm_in, d_dm_out, d_dm_desc; // soft-demodulation front end: staged symbols + sequences, LLRs, descriptors
  PinBuf<uint8_t>  h_dm_desc, h_dm_out;
  DevBuf<float>    d_dm_csimax; // per codeword: max of its channel state information (csi_correction)
  DevBuf<uint8_t>  d_enc_in, d_enc_out, d_enc_desc; // transmit mirror: payloads, packed e-bits, descriptors
  PinBuf<uint8_t>  h_enc_desc, h_enc_out;
  DevBuf<uint8_t>  d_ul_in, d_ul_out, d_ul_uci, d_ul_desc; // PUSCH pre-steps: staged q_bits, g_bits, ACK/RI/CQI LLRs, descriptors
  PinBuf<uint8_t>  h_ul_out, h_ul_uci, h_ul_desc;
  PinBuf<TbResult> h_res;
  PinBuf<CbState>  h_state;

  std::vector<float> tb_hints; // difficulty hints for the next transport-block batch (srslte_b200_set_tb_hints)

private:
  Engine() {}
  int build_tables();
  int map_event_pair(cudaEvent_t* a, cudaEvent_t* b);
  int finish_timing();

  cudaEvent_t              ev_begin = nullptr, ev_end = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t              ev_desc = nullptr; // last upload of a front-end descriptor array from its pinned staging buffer
  std::vector<cudaEvent_t> map_events;
  size_t                   n_map_events_used = 0;

  enum { PENDING_NONE = 0, PENDING_CB, PENDING_TB };
  int               pending = PENDING_NONE;
  Plan*             plan_ptr = nullptr;
  LaunchState*      ls_ptr   = nullptr;
  std::vector<uint8_t> cache_key; // identity of the batch *ls_ptr / *plan_ptr describe (empty: not reusable)
  uint8_t*          cb_out_host  = nullptr;
  size_t            cb_out_bytes = 0;
  srslte_b200_tb_t* tb_user   = nullptr;
  uint32_t          tb_user_n = 0;
  uint32_t          tb_flags  = 0;
  // UCI LLRs of a srslte_b200_ulsch_deinterleave(..., SRSLTE_B200_UCI_DEFERRED): copied out of h_ul_uci by the next wait()
  struct UciCopy {
    void*  dst;
    size_t src_off, bytes;
  };
  std::vector<UciCopy> uci_deferred;
  int                  flush_uci(); // synchronises the stream when copies are outstanding
  struct H2dCopy {
    const uint8_t* src;
    size_t         dst, bytes;
  };
  std::vector<H2dCopy> tb_h2d;         // host-to-device copies of the batch's e-bits (replayed when the plan is reused)
  std::vector<int32_t> tb_invalid_ret; // return codes of the transport blocks the planner rejected
  size_t               tb_out_total = 0;
  int                  finish_tb_submit(uint32_t flags);
  std::vector<int>    tb_map;
  std::vector<size_t> tb_out_off;
};

void softbuffer_reset(Softbuffer* s);
void softbuffer_set_crc(Softbuffer* s, const bool* cb_crc, uint32_t n);
void softbuffer_free(Softbuffer* s);

} // namespace b200
