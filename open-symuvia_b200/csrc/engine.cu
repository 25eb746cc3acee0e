// engine.cu — B200 (sm_100a) engine of SymuVia's per-time-step microscopic vehicle update and its C-ABI
// (include/symgpu.h).  One symgpu_ctx = one Reseau: static network tables and the vehicle
// struct-of-arrays live in HBM; one call to symgpu_step() replaces Reseau::SimuTrafic phases
// InitVehicules .. SupprimeVehicules (symuvia/src/reseau.cpp:2250-2258, :14152-14470).
//
// Step pipeline (step stream + two side streams joined by events; one host synchronisation per step, on the relaxation's verdict):
//   k_signals      A8   per stop line: colour + green fraction of this step (ControleurDeFeux.cpp:96-148)
//   k_mrg_rank     A10  per merge point: priority levels inside signalised junctions (CarrefourAFeuxEx.cpp:1386-1593)
//   k_reset/k_prepare/scan/k_scatter/k_lane_rank/k_lane_links  A2/A3: counting sort by lane, in-lane order by (position, id),
//                       own-lane leader links, lane tail                  (reseau.cpp:3901-3976, :3583-3820)
//   k_lc_*         A9   lane changing, only when the scenario enables it  (reseau.cpp:4219-4330)
//   k_context      A3/A5/A12 per vehicle: context lanes, leader, deltaN, free-flow distance -> 256-byte row (rows.cuh)
//   k_seg, k_cyc_* A4   segments of the leader graph, loops cut where the reference's recursion cuts them (vehicule.cpp:1201-1216)
//   k_relax        A4-A7 chaotic relaxation of phase B to its unique fixed point (flow.cuh)
//   k_place        A4   placement on the context lanes                    (vehicule.cpp:1367-1527)
//   k_entries      N1   first waiting vehicle of each origin lane         (usage/SymuViaTripNode.cpp:1642-1851)
//   k_mrg_probe .. k_mrg_apply  A10 merge points in Convergent id order   (merge.cuh; convergent.cpp:333-990)
//   k_final        A13  finalisation, link counters, exits                (vehicule.cpp:4254-4583, reseau.cpp:3984)
//   k_mrg_rotate/sensors/sensors_sum  A10 merge sensors, beside the next step's lane ordering (reseau.cpp:2432-2600)
// There is no CPU fallback: without a CUDA device symgpu_create() fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <climits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/symgpu.h"
#include "engine_internal.h"
#include "flatnet.h"
#include "micro.cuh"
#include "xmlnet.h"

using namespace symb200;

static thread_local std::string g_err;
#define CK(call)                                                                                         \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return SYMGPU_E_CUDA; } } while (0)

__constant__ DevParams cP;
__device__ __forceinline__ const DevParams& cP_ref() { return cP; }
#include "lanechange.cuh"
#include "merge.cuh"
#include "flow.cuh"

// ------------------------------------------------------------------------------------------------
// device tables for signals
struct DevSig { const int *ctrl_seq_off, *ctrl_nseq; const double* ctrl_cycle; const int *seq_sig_off, *seq_nsig;
  const double* seq_len; const int *sig_in, *sig_out; const double *sig_green, *sig_amber, *sig_delay; };

// ControleurDeFeux::GetCouleurFeux (ControleurDeFeux.cpp:258-307) + PlanDeFeux::GetSequence (:1209-1228) +
// SignalActif::GetCouleurFeux / GetRapVertOrange (:1517-1558).  0 red, 1 green, 2 amber.
__device__ inline int couleur(const DevSig& S, int ctrl, double inst, int te, int ts, double dt, double* prop) {
  int nsec = (int)inst;
  int ntmp = nsec % (int)S.ctrl_cycle[ctrl];
  double prev = 0, tcur = 0; int seq = -1;
  for (int q = S.ctrl_seq_off[ctrl]; q < S.ctrl_seq_off[ctrl] + S.ctrl_nseq[ctrl]; q++) {
    if (prev + S.seq_len[q] > ntmp) { tcur = ntmp - prev; seq = q; break; }
    prev += S.seq_len[q];
  }
  if (seq < 0) return 1;
  int g = -1;
  for (int k = S.seq_sig_off[seq]; k < S.seq_sig_off[seq] + S.seq_nsig[seq]; k++) if (S.sig_in[k] == te && S.sig_out[k] == ts) { g = k; break; }
  if (g < 0) return 0;
  double delay = S.sig_delay[g], green = S.sig_green[g], amber = S.sig_amber[g];
  int col;
  if (tcur < delay) col = 0; else if (tcur < delay + green) col = 1;
  else if (delay + green <= tcur && tcur < delay + green + amber) col = 2; else col = 0;
  if (prop) {
    if (col != 0) *prop = fmin(1.0, (tcur - delay) / dt);
    else if (tcur < delay) *prop = 1 - fmin(1.0, tcur / dt);
    else *prop = 1 - fmin(1.0, (tcur - (delay + green + amber) / dt));
  }
  return col;
}
// ControleurDeFeux::GetStatutCouleurFeux (ControleurDeFeux.cpp:96-148), once per stop line and step
__global__ void k_signals(DevSig S, const int* sl_i, int n_sl, double inst, int* col_out, double* prop_out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sl) return;
  int ctrl = sl_i[6 * s], te = sl_i[6 * s + 1], ts = sl_i[6 * s + 3];
  double dt = cP.dt, p = 0;
  int col = couleur(S, ctrl, inst, te, ts, dt, &p);
  int pcol = (inst - dt > 0) ? couleur(S, ctrl, inst - dt, te, ts, dt, nullptr) : 1;
  int res;
  if (col == 0 && pcol == 0) { p = 0; res = 0; }
  else if ((col == 1 || col == 2) && (pcol == 1 || pcol == 2)) { p = 1; res = col; }
  else if (col == 1 && pcol == 0) res = 1;
  else if (col == 0 && (pcol == 1 || pcol == 2)) { p = 1; res = 0; }
  else res = 0;
  col_out[s] = res; prop_out[s] = p;
}

// ------------------------------------------------------------------------------------------------
// start of a step: per-lane counters, lane tails, step statistics and the relaxation's control block in one launch
__global__ void k_reset(int L, int* lane_count, int* lane_fill, int* lane_tail, int* stat, int* flow_ctl, int ctl_words) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) { lane_count[i] = 0; lane_fill[i] = 0; lane_tail[i] = -1; }
  if (i < 16) stat[i] = 0;
  if (i < ctl_words) flow_ctl[i] = 0;
}
__global__ void k_prepare(DevVeh V, int n, int* lane_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = V.status[i];
  V.computed[i] = 0;
  if (st == ST_ALIVE || st == ST_NEW) atomicAdd(&lane_count[V.lane[i]], 1);
}

// exclusive scan of int32 (three launches: block scan, scan of block sums, add)
constexpr int kScanB = 1024;
__global__ void k_scan_block(const int* in, int* out, int* bsum, int n) {
  __shared__ int sh[kScanB];
  int i = blockIdx.x * kScanB + threadIdx.x;
  int v = i < n ? in[i] : 0;
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < kScanB; off <<= 1) {
    int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) out[i] = sh[threadIdx.x] - v;
  if (threadIdx.x == kScanB - 1) bsum[blockIdx.x] = sh[threadIdx.x];
}
__global__ void k_scan_sums(int* bsum, int nb) {      // single block, serial over chunks of 1024
  __shared__ int sh[kScanB];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kScanB) {
    int i = base + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kScanB; off <<= 1) {
      int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) bsum[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[kScanB - 1];
    __syncthreads();
  }
}
__global__ void k_scan_add(int* out, const int* bsum, int n) {
  int i = blockIdx.x * kScanB + threadIdx.x;
  if (i < n) out[i] += bsum[blockIdx.x];
}

__global__ void k_scatter(DevVeh V, int n, const int* lane_start, int* lane_fill, int* uorder, double* ukey, int* uid, int* ulane) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = V.status[i];
  if (st != ST_ALIVE && st != ST_NEW) return;
  int ln = V.lane[i];
  int slot = lane_start[ln] + atomicAdd(&lane_fill[ln], 1);
  uorder[slot] = i; ukey[slot] = snap_pos(V.pos[i]); uid[slot] = V.id[i]; ulane[slot] = ln;     // the lane's slice, in arrival order, with its sort keys
}
// context memos (micro.cuh: CtxCache) of the vehicles k_final listed (they changed lane or are new): runs on a side stream beside the
// first kernels of the next step
__global__ void k_ctx_fill(DevNet N, DevVeh V, const int* cc_list, const int* cc_cnt) {
  const int cnt = *cc_cnt;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x) ctx_cache_fill(N, V, cc_list[t]);
}

// in-lane order by (start-of-step position, id) + own-lane leader links + lane tail.
// Replaces the per-query linear scans of Reseau::GetPrevVehicule (reseau.cpp:3650-3718) and
// GetNearAvalVehicule (:3583-3640) over the id-ordered lane sets.
// Two fully parallel kernels, one thread per vehicle (lanes are short: a vehicle's rank is a scan of its lane's slice, which its
// lane mates read too and L1 serves):
//   k_lane_rank   rank of (position, id) among the lane's keys -> position in the lane order; keys and lane stored in that order
//   k_lane_links  nearest vehicle ahead with the reference's rule for equal positions, lane tail
__global__ void k_lane_rank(int n_sorted, const int* lane_start, const int* lane_count, const int* uorder, const double* ukey, const int* uid, const int* ulane,
                            int* order, double* skey, int* sid, int* slane, int* posidx) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_sorted) return;
  const int ln = ulane[u]; const int s = lane_start[ln], n = lane_count[ln];
  const double k = ukey[u]; const int id = uid[u];
  int r = 0;
  for (int a = 0; a < n; a++) { double ka = ukey[s + a]; r += (ka < k || (ka == k && uid[s + a] < id)) ? 1 : 0; }
  const int p = s + r; const int x = uorder[u];
  order[p] = x; skey[p] = k; sid[p] = id; slane[p] = ln; posidx[x] = p;
}
__global__ void k_lane_links(DevVeh V, int n_sorted, const int* lane_start, const int* lane_count, const int* order, const double* skey, const int* sid,
                             const int* slane, int* lane_tail, long long* counters) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  const int ln = slane[p]; const int s = lane_start[ln], n = lane_count[ln]; const int a = p - s;
  // Equal positions: Reseau::GetLastVehicule (reseau.cpp:14041-14081) returns the first candidate (id order) that is not
  // the leader of another candidate; the candidates' leaders are read from their previous-step context (a candidate met
  // while it has already rebuilt its context this step is counted in SYMGPU_CNT_TIES but not replayed).
  auto pick = [&](int b0, int b1) -> int {          // group = order[s+b0 .. s+b1) at one position, id-ascending
    if (b1 - b0 == 1) return order[s + b0];
    for (int i = b0; i < b1; i++) {
      int idi = sid[s + i]; bool fol = false;
      for (int j = b0; j < b1 && !fol; j++) if (j != i && V.prev_leader[order[s + j]] == idi) fol = true;
      if (!fol) return order[s + i];
    }
    return order[s + b1 - 1];
  };
  const int x = order[p]; const double px = skey[p];
  int ties = 0;
  if (px >= 0 && (a == 0 || skey[p - 1] < 0)) {     // first vehicle with a position: its group holds the lane's tail (GetLastVehicule)
    int b1 = a + 1; while (b1 < n && skey[s + b1] == px) b1++;
    if (b1 - a > 1) ties++;
    lane_tail[ln] = pick(a, b1);
  }
  const double own = px <= SYM_UNDEF ? 0.0 : px;
  const bool eq_ok = V.status[x] == ST_NEW;         // GetPos(0) == UNDEF only before the first computation (reseau.cpp:3695)
  int ld = -1;
  for (int b = a + 1; b < n; b++) {
    double pb = skey[s + b];
    if (pb > own || (eq_ok && pb == own)) {
      int b1 = b + 1; while (b1 < n && skey[s + b1] == pb) b1++;
      if (b1 - b > 1) ties++;
      ld = pick(b, b1);
      break;
    }
  }
  V.ahead[x] = ld;
  if (ties) atomicAdd((unsigned long long*)&counters[SYMGPU_CNT_TIES], (unsigned long long)ties);
}

// one thread per position of the lane order: neighbouring threads are neighbours on a lane, so they walk the same context
// lanes, stop lines and movements (little divergence, table reads shared through L1) and their rows are written coalesced
#ifndef SYM_CTX_BLOCKS
#define SYM_CTX_BLOCKS 8   // 128 threads x 8 blocks = 64 registers.  With the per-vehicle context memo the kernel moves 870 B per vehicle at ~60 % of the HBM
                           // peak, so spill traffic costs more than the lost warps (C4, 256-thread blocks at 32 / 40 / 48 / 64 / 80 registers: 0.227 / 0.200 / 0.193 / 0.184 / 0.20 ms; 128 x 8: 0.184)
#endif
#ifndef SYM_CTX_THREADS
#define SYM_CTX_THREADS 128
#endif
__global__ void __launch_bounds__(SYM_CTX_THREADS, SYM_CTX_BLOCKS) k_context(DevNet N, DevVeh V, DevLaneDyn LD, int n_sorted, const int* order, const int* lane_tail, double* rows, unsigned* draws, int* overflow,
                                                                               const DevMrg* M, double inst) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  int i = order[p];
  int d = context_one(cP, N, V, lane_tail, i, V.ahead[i], overflow, rows, p, &LD, M, inst);
  V.vit_n[i] = V.vit[i];                           // initial guess of the relaxation (flow.cuh)
  if (d) atomicAdd(draws, (unsigned)d);
}

// Origins, after the main update: ready vehicle inserted? then try the first waiting vehicle of the lane
// (SymuViaTripNode::InsertionVehEnAttente, usage/SymuViaTripNode.cpp:1725-1812).  One thread per origin lane.
struct EntryDesc { int lane, pret_slot, pool_slot, dst_slot; };
__device__ inline void copy_row(const DevVeh& V, int src, int dst) {
  V.id[dst] = V.id[src]; V.cls[dst] = V.cls[src]; V.route[dst] = V.route[src]; V.route_pos[dst] = V.route_pos[src];
  V.lane[dst] = V.lane[src]; V.lane_n[dst] = V.lane_n[src]; V.prev_leader[dst] = V.prev_leader[src]; V.flags[dst] = V.flags[src];
  V.oflags[dst] = V.oflags[src]; V.vdes[dst] = V.vdes[src];
  for (int k = 0; k < kMemo; k++) V.memo[dst * kMemo + k] = V.memo[src * kMemo + k];
  V.pos[dst] = V.pos[src]; V.vit[dst] = V.vit[src]; V.acc[dst] = V.acc[src];
  V.pos_n[dst] = V.pos_n[src]; V.vit_n[dst] = V.vit_n[src]; V.acc_n[dst] = V.acc_n[src];
  V.dn[dst] = V.dn[src]; V.cpos[dst] = V.cpos[src]; V.tcre[dst] = V.tcre[src]; V.dstp[dst] = V.dstp[src];
  V.hent[dst] = V.hent[src]; V.tsort[dst] = V.tsort[src];
  V.ctx_leader[dst] = V.ctx_leader[src];
}
__global__ void k_entries(DevNet N, DevVeh V, DevLaneDyn LD, const EntryDesc* desc, int n_desc, int* result, double inst,
                          unsigned* draws, int* overflow, const int* guard) {
  if (guard && *guard == 0) return;              // launched ahead of the convergence check (step_back): nothing to do yet
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_desc) return;
  EntryDesc d = desc[e];
  int res = 0;                                   // bit0: ready vehicle inserted, bit1: waiting vehicle inserted
  bool pret_ok = true;
  if (d.pret_slot >= 0) { pret_ok = V.pos_n[d.pret_slot] > SYM_UNDEF; if (pret_ok) res |= 1; }
  if (pret_ok && d.pool_slot >= 0) {
    int w = d.pool_slot;
    int own_leader = d.pret_slot >= 0 ? d.pret_slot : LD.tail[d.lane];
    int dr = context_one(cP, N, V, LD.tail, w, own_leader, overflow);
    if (dr) atomicAdd(draws, (unsigned)dr);
    follow_one(cP, N, V, LD, w, true, inst, V.cpos[w]);
    if (V.pos_n[w] > SYM_UNDEF) {
      copy_row(V, w, d.dst_slot);
      V.status[d.dst_slot] = ST_NEW; V.status[w] = ST_DEAD; V.computed[d.dst_slot] = 1;
      res |= 2;
    } else {                                      // :1776-1783 the vehicle stays at the head of the queue
      V.pos_n[w] = SYM_UNDEF; V.vit_n[w] = cP.cls[V.cls[w]].vx; V.lane_n[w] = d.lane; V.tcre[w] = 0.0;
    }
  }
  result[e] = res;
}

// Vehicule::FinCalculTrafic (vehicule.cpp:4254-4583) + Reseau::SupprimeVehicules (reseau.cpp:3984-4216)
struct ExitRow { int slot, id; double inst; };
__global__ void k_final(DevNet N, DevVeh V, int n, int ncls, double dt, int* entre, int* sorti, ExitRow* exited, int* n_exited,
                        int* exit_winner, int* n_alive, int* err, const int* guard, int* cc_list, int* cc_cnt) {
  if (guard && *guard == 0) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool stale = false, alive = false;
  int st = i < n ? V.status[i] : ST_DEAD;
  if (st == ST_ALIVE || st == ST_NEW) {
  if (!V.computed[i]) atomicAdd(err, 1);
  else {
  int of = V.oflags[i];
  int lane1 = V.lane[i], lane0 = V.lane_n[i];
  int link1 = (of & O_LINK1NULL) ? -1 : N.lane_link[lane1];
  int link0 = N.lane_link[lane0];
  if (link1 < 0) V.dstp[i] = V.pos_n[i]; else V.dstp[i] += V.vit_n[i] * dt;              // :4328-4333
  int c = V.cls[i];
  if (link1 != link0) {                                                                    // :4337-4442
    if (link1 >= 0 && N.link_brick[link1] < 0) atomicAdd(&sorti[link1 * ncls + c], 1);
    if (N.link_brick[link0] < 0) atomicAdd(&entre[link0 * ncls + c], 1);
    if (N.link_brick[link0] < 0) { int k = route_index(N, V.route[i], V.route_pos[i], link0); if (k >= 0) V.route_pos[i] = k; }
  }
  if (of & O_ENDLEG) {                                                                     // TripNode::VehiculeEnter TripNode.cpp:167-177
    int k = atomicAdd(n_exited, 1);
    exited[k].slot = i; exited[k].id = V.id[i]; exited[k].inst = V.tsort[i];
    atomicMax(&exit_winner[lane0], i);
    V.status[i] = ST_DEAD;
  } else {
    V.status[i] = ST_ALIVE;
    alive = true;
    if (cc_list) { int id = V.id[i];                                                       // the vehicle's context memo is for another lane (or another vehicle):
      stale = V.cc_lane[i] != lane0 || V.cc_id[i] != id;                                   // listed for k_ctx_fill, key set now so that it is listed once
      if (stale) { V.cc_lane[i] = lane0; V.cc_id[i] = id; } }
  }
  }
  }
  { const unsigned ma = __ballot_sync(0xffffffffu, alive);                                 // vehicles left on the network: one atomic per warp
    if (ma && (threadIdx.x & 31) == 0) atomicAdd(n_alive, __popc(ma)); }
  if (cc_list) {                                                                           // one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, stale);
    if (m) { const int lane = threadIdx.x & 31; int base = 0;
      if (lane == 0) base = atomicAdd(cc_cnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (stale) cc_list[base + __popc(m & ((1u << lane) - 1))] = i; }
  }
}
// Sortie::VehiculeEnter + GetNextEnterInstant (sortie.cpp:60-130): the last vehicle of the list that left through
// the lane sets the earliest exit instant of the next one.
__global__ void k_exit_lanes(DevVeh V, const int* exit_lanes, const double* exit_cap, const int* exit_nl, int n_exit_lanes,
                             int* exit_winner, double* next_sortie, double inst, double dt, double duree, const int* guard) {
  if (guard && *guard == 0) return;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_exit_lanes) return;
  int ln = exit_lanes[e];
  int w = exit_winner[ln];
  if (w < 0) return;
  exit_winner[ln] = -1;
  double cap = exit_cap[e];
  double r;
  if (cap < 0 || cap >= 9999) r = inst;
  else {
    double cum = 0, cump = 0, ti = V.tsort[w];
    while (cum < 1 && ti <= duree) { cump = cum; cum += cap * dt / exit_nl[e]; ti += dt; }
    r = cum < 1 ? -1.0 : (ti - dt) + (1 - cump) / (cum - cump) * dt;
  }
  next_sortie[ln] = r;
}

// Trajectory rows for the writers (Vehicule::SortieTrafic vehicule.cpp:1993-2176): alive vehicles in m_LstVehicles order,
// packed contiguously so that one asynchronous copy per field brings them to the host.
__global__ void k_alive_flags(const int* status, int n, int* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = status[i] == ST_ALIVE;
}
__global__ void k_pack_traj(DevVeh V, int n, const int* flag, const int* off, int* o_id, int* o_lane, double* o_pos, double* o_vit, double* o_acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  int k = off[i];
  o_id[k] = V.id[i]; o_lane[k] = V.lane[i]; o_pos[k] = V.pos[i]; o_vit[k] = V.vit[i]; o_acc[k] = V.acc[i];
}

// ---- partitioned runs ------------------------------------------------------------------------------------------------
// tails of the lanes this GPU owns and another GPU looks into (first vehicle a follower arriving from upstream would meet)
__global__ void k_mp_pack_tails(DevVeh V, const int* export_lanes, int n, const int* lane_tail, double* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int t = lane_tail[export_lanes[k]];
  double* o = out + 4 * (size_t)k;
  if (t < 0) { o[0] = __hiloint2double(0, -1); o[1] = 0; o[2] = 0; o[3] = 0; }
  else { o[0] = __hiloint2double(V.cls[t], V.id[t]); o[1] = V.pos[t]; o[2] = V.vit[t]; o[3] = 0; }
}
// Tail exchange over peer memory (NVLink): the pack kernel stores every record straight into the halo area of the GPU that looks into
// the lane (exp_dst[k], opened through CUDA IPC) and the last block to finish publishes the step's sequence number in that GPU's flag
// slot (system-scope fence by every thread before the block counts itself done, release store of the flag) — no collective, no host.
__global__ void k_mp_pack_tails_p2p(DevVeh V, const int* export_lanes, int n, const int* lane_tail, double* const* exp_dst, unsigned* done_ctr,
                                    unsigned long long* const* flag_dst, int world, unsigned long long seq) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    int t = lane_tail[export_lanes[k]];
    double* o = exp_dst[k];
    if (t < 0) { o[0] = __hiloint2double(0, -1); o[1] = 0; o[2] = 0; o[3] = 0; }
    else { o[0] = __hiloint2double(V.cls[t], V.id[t]); o[1] = V.pos[t]; o[2] = V.vit[t]; o[3] = 0; }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(done_ctr, 1u);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      __threadfence_system();
      for (int q = 0; q < world; q++) if (flag_dst[q]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_dst[q]), "l"(seq) : "memory");
    }
  }
}
// the GPU that looks into the lanes: waits until every peer it has a halo from has published this step (bounded: ~2 s, then an error)
__global__ void k_mp_wait_tails(const unsigned long long* flags, const int* halo_off, int world, unsigned long long seq, int* err) {
  int q = threadIdx.x;
  if (q >= world || halo_off[q + 1] == halo_off[q]) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + q) : "memory");
    if (v >= seq) break;
    if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1); break; }
  }
}
__global__ void k_mp_unpack_ghosts(DevVeh V, const int* halo_lanes, int n, int ghost_base, int* lane_tail, const double* in) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* o = in + 4 * (size_t)k;
  int id = __double2loint(o[0]); int g = ghost_base + k; int ln = halo_lanes[k];
  if (id < 0) { V.status[g] = ST_DEAD; lane_tail[ln] = -1; return; }
  V.status[g] = ST_GHOST; V.id[g] = id; V.cls[g] = __double2hiint(o[0]); V.lane[g] = ln; V.pos[g] = o[1]; V.vit[g] = o[2];
  lane_tail[ln] = g;
}
// vehicles that ended the step on a lane owned by another GPU leave with their whole state (16 doubles per row)
constexpr int kMigD = 16;
__global__ void k_mp_pack_migrants(DevVeh V, int n, const int* lane_owner, int rank, int cap_per_peer, int* count, double* out, int* free_slots, int free_base, int* free_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || V.status[i] != ST_ALIVE) return;
  int q = lane_owner[V.lane[i]];
  if (q == rank) return;
  int k = atomicAdd(&count[q], 1);
  if (k < cap_per_peer) {
    double* o = out + ((size_t)q * cap_per_peer + k) * kMigD;
    o[0] = __hiloint2double(V.cls[i], V.id[i]); o[1] = __hiloint2double(V.route_pos[i], V.route[i]); o[2] = __hiloint2double(V.prev_leader[i], V.lane[i]);
    o[3] = __hiloint2double(V.oflags[i], V.flags[i]); o[4] = __hiloint2double(V.memo[i * kMemo + 1], V.memo[i * kMemo]); o[5] = __hiloint2double(V.memo[i * kMemo + 3], V.memo[i * kMemo + 2]);
    o[6] = V.pos[i]; o[7] = V.vit[i]; o[8] = V.acc[i]; o[9] = V.dn[i]; o[10] = V.cpos[i]; o[11] = V.tcre[i]; o[12] = V.dstp[i]; o[13] = V.hent[i]; o[14] = V.tsort[i]; o[15] = 0;
  }
  V.status[i] = ST_DEAD;
  free_slots[free_base + atomicAdd(free_cnt, 1)] = i;          // the slot is reused by a later arrival
}
// the number of rows of every per-peer region rides in the spare double of its first row, so the receiver needs no separate count exchange
__global__ void k_mp_row_counts(const int* count, double* out, int cap_per_peer, int world) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < world) out[(size_t)q * cap_per_peer * kMigD + 15] = (double)count[q];
}
__global__ void k_mp_unpack_migrants(DevVeh V, int slot0, int n, const double* in, const int* free_slots, int free_top, int n_reuse) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* o = in + (size_t)k * kMigD; int i = k < n_reuse ? free_slots[free_top - 1 - k] : slot0 + (k - n_reuse);
  V.status[i] = ST_ALIVE; V.id[i] = __double2loint(o[0]); V.cls[i] = __double2hiint(o[0]); V.route[i] = __double2loint(o[1]); V.route_pos[i] = __double2hiint(o[1]);
  V.lane[i] = __double2loint(o[2]); V.lane_n[i] = V.lane[i]; V.prev_leader[i] = __double2hiint(o[2]); V.flags[i] = __double2loint(o[3]); V.oflags[i] = __double2hiint(o[3]);
  V.memo[i * kMemo] = __double2loint(o[4]); V.memo[i * kMemo + 1] = __double2hiint(o[4]); V.memo[i * kMemo + 2] = __double2loint(o[5]); V.memo[i * kMemo + 3] = __double2hiint(o[5]);
  V.pos[i] = o[6]; V.vit[i] = o[7]; V.acc[i] = o[8]; V.pos_n[i] = o[6]; V.vit_n[i] = o[7]; V.acc_n[i] = o[8];
  V.dn[i] = o[9]; V.cpos[i] = o[10]; V.tcre[i] = o[11]; V.dstp[i] = o[12]; V.hent[i] = o[13]; V.tsort[i] = o[14];
}
__global__ void k_mp_drop_foreign(DevVeh V, int n, const int* lane_owner, int rank, int* kept) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || V.status[i] != ST_ALIVE) return;
  if (lane_owner[V.lane[i]] != rank) V.status[i] = ST_DEAD; else atomicAdd(kept, 1);
}

// ================================================================================================
// host side
// ================================================================================================
constexpr int kTrajBufs = 4;   // caller buffers for trajectory rows: with three in rotation the consumer lags two steps and a copy never has to end inside its step
namespace {

// RandManager (RandManager.cpp:5-30): boost::mt19937 + uniform_int<>(0, 0x7FFE)
struct HostRng {
  uint32_t mt[624]; int idx = 624; unsigned count = 0;
  void seed(uint32_t s) { mt[0] = s; for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i; idx = 624; count = 0; }
  uint32_t raw() {
    if (idx >= 624) {
      for (int i = 0; i < 624; i++) { uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
      idx = 0; }
    uint32_t y = mt[idx++]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; return y;
  }
  int next() { count++; const uint32_t bucket = 0xFFFFFFFFu / 32767u; for (;;) { uint32_t r = raw() / bucket; if (r <= 32766u) return (int)r; } }
  void skip(unsigned n) { for (unsigned i = 0; i < n; i++) next(); }
};

struct HostVeh { int id, cls, route, lane; double tcre, hent; };   // a generated vehicle not yet uploaded
struct EntryLane { int lane; int pret_slot = -1; int pool_slot = -1; std::deque<HostVeh> queue; int pool_base; };
struct HostEntry { int node, link; std::vector<std::pair<double, double>> demand; std::vector<int> dest_exit; std::vector<double> dest_coeff;
  std::vector<std::vector<std::pair<int, double>>> dest_routes; std::vector<double> lane_coefs, type_coefs; bool exponential = false;
  double crit = 0; std::vector<EntryLane> lanes; };

template <class T> struct DBuf {
  T* p = nullptr; size_t n = 0;
  DBuf() = default; DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { free(); }                               // every device buffer of a context goes with it (symgpu_destroy)
  cudaError_t alloc(size_t count) { n = count; if (!count) { p = nullptr; return cudaSuccess; } cudaError_t e = cudaMalloc(&p, count * sizeof(T)); if (e == cudaSuccess) e = cudaMemset(p, 0, count * sizeof(T)); return e; }
  cudaError_t upload(const std::vector<T>& h) { cudaError_t e = alloc(h.size()); if (e != cudaSuccess || h.empty()) return e; return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); }
  void free() { if (p) cudaFree(p); p = nullptr; }
};

}  // namespace

struct ProfRec { int kid; cudaEvent_t a, b; };
struct symgpu_ctx {
  FlatNet net; DevParams P; int device = 0; cudaStream_t stream = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double t = 0; int n_alive = 0, n_slots = 0, cap = 0, pool_base = 0, n_pool = 0; int stamp = 0;
  double last_ms = 0; long long launches = 0, passes = 0, created = 0;
  HostRng rng; unsigned dev_draws_seen = 0;
  // static tables
  DBuf<int> lane_link, lane_num, lane_prev, lane_flags, succ_off, succ_key, succ_next, sl_off, sl_i, link_first, link_nl, link_up,
      link_type, link_brick, route_off, route_links, ctrl_seq_off, ctrl_nseq, seq_sig_off, seq_nsig, sig_in, sig_out, exit_lanes, exit_nl;
  DBuf<double> lane_len, sl_pos, link_len, link_vreg, ctrl_cycle, seq_len, sig_green, sig_amber, sig_delay, exit_cap;
  // per step
  DBuf<int> sl_col; DBuf<double> sl_prop;
  DBuf<int> lane_count, lane_start, lane_fill, lane_tail, exit_winner, bsum, order, entre, sorti, stat;
  DBuf<int> f_posidx, f_isfront, f_seg_len, f_seg_dep, f_segpos, f_minslot, f_cut, f_fronts, f_ulist, rl_lead, f_A, f_A2, f_Mn, f_Mn2, f_oncyc, f_cid;
  DBuf<double> rows, b_goal; DBuf<int> b_flags;
  DBuf<unsigned long long> f_rootkey; DBuf<FlowCtl> f_ctl; DBuf<double> vl_used; DevFlow df; FlowCtl* h_ctl = nullptr; FlowCtl* h_ctl_seed = nullptr;
  DBuf<double> next_sortie; DBuf<long long> counters; DBuf<unsigned> draws; DBuf<ExitRow> exited; DBuf<EntryDesc> edesc; DBuf<int> eresult;
  // vehicles
  DBuf<int> v_vdes, lc_voies_ok, lc_chgt, lc_delta_next, lc_link_head, lc_nchg, link_down, lc_tl_pre, lc_rfol, lc_rtgt, lc_had_mand; DBuf<unsigned char> lc_kind, lc_code; DBuf<double> lc_phi, u_key, s_key; DBuf<int> lc_moved, u_order, u_id, u_lane, s_id, s_lane; bool lc_seq = false; int side_mask = 31; DBuf<DevRng> lc_rng; DevLc dlc; bool chgtvoie = false;
  DBuf<int> v_status, v_id, v_cls, v_route, v_rpos, v_lane, v_lane_n, v_pl, v_flags, v_oflags, v_memo, v_ahead, v_computed, v_ctx_nl,
      v_ctx_lane, v_ctx_leader, v_ctx_flags;
  DBuf<double> v_pos, v_vit, v_acc, v_pos_n, v_vit_n, v_acc_n, v_dn, v_cpos, v_tcre, v_dstp, v_hent, v_tsort, v_ctx_gap, v_ctx_dn0, v_ctx_dfree;
  DBuf<int> cc_lane, cc_id, cc_list, cc_cnt; DBuf<CtxCache> cc; bool use_cc = true, cc_fill = false /* k_final listed vehicles for k_ctx_fill */, aux_wait = false /* the step has work on the side stream */, dbg_sync = false, dbg_failed = false, coop_loops = true; cudaStream_t aux = nullptr, aux2 = nullptr; cudaEvent_t cc_go = nullptr, cc_done = nullptr, ev_fork = nullptr, ev_rng = nullptr, ev_dirty = nullptr, ev_sens = nullptr, ev_eval = nullptr; bool rng_wait = false, sens_wait = false;   // per-vehicle memo of the context walks (micro.cuh: CtxCache)
  DevNet dn; DevVeh dv; DevLaneDyn dl; DevSig ds;
  // merge points (merge.cuh): static tables, per-point state, sensors, per-step records; the device copy of the struct for the kernels that take a pointer
  struct Mrg { bool on = false; int n_upd = 0, n_rows = 0; DevMrg h; DBuf<DevMrg> d;
    DBuf<int> link_cvg, cvg_node, cvg_down, cvg_pt0, cvg_npt, cvg_am0, cvg_nam, cvg_win, cvg_sens0, cvg_am, pt_cvg, pt_down, pt_vam0, pt_nvam, pt_mode, pt_pm0, vam_lane, vam_ref, vam_niv,
        pm_vam, pm_sl, pm_occ0, pm_occ, lane_pt, lane_vam, sens_link, sens_cnt0, sens_win, cnt, lsens0, lsens, lcam0, lcam, lcav0, lcav, lmv0, lmv_ent, lmv_exit, fm0, fm_exit, fm_tl0, fm_ntl, fm_tl,
        node_type, pt_rand, pt_free, cand_up, own_min, own_max, st, inv, ninv, pre, kmax, decide, ndraw, flist, fsorted, rbuf, res, dlist, nused, nvoie, ulanes, sens_row0, row_tot, row_sens, flag, out;
    DBuf<double> cvg_gamma, cvg_mu, cvg_poscpt, cvg_tf, sens_pos, pt_last, proba, mrgx; DBuf<unsigned> seen; DBuf<unsigned long long> kf, kfo, kbsum, mark; } mg;
  bool host_rng_fresh = true;                                      // the host copy of the random stream equals the device's
  std::vector<signed char> cls_of_id;                               // vehicle type by vehicle id (output writers)
  std::vector<HostEntry> entries; std::vector<ExitRow> last_exited; int last_id = 0;
  struct ApiVeh { int id, cls, entry, dest, lane; double dbt; };
  std::vector<ApiVeh> api_queue;                                   // Reseau::m_VehiclesToCreate (reseau.cpp:2233-2240)
  // trajectory output: packed device rows + side stream + caller buffers registered for direct DMA (two in flight)
  DBuf<int> t_flag, t_off, t_bsum, t_id, t_lane; DBuf<double> t_pos, t_vit, t_acc; cudaStream_t side = nullptr; cudaEvent_t t_ready = nullptr, t_packed = nullptr; bool pack_wait = false; int traj_req = -1, traj_req_n = 0, traj_req_rows = 0; double traj_host_us = 0; long traj_host_n = 0; cudaEvent_t t_done[kTrajBufs] = {}, t_begin[kTrajBufs] = {}; double t_copy_ms[kTrajBufs] = {};
  struct TrajBuf { int cap = 0; int* id = nullptr; int* lane = nullptr; double* pos = nullptr; double* vit = nullptr; double* acc = nullptr; bool registered = false; } tbuf[kTrajBufs];
  int profile = 0; bool prof_open = false; std::vector<ProfRec> prof_recs; std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  long long lane_changes = 0; int n_sms = 148; int relax_batch = kRelaxBatch; bool ext_stream = false;
  double k_ms[16] = {0}; long long k_launches[16] = {0}; long long veh_steps = 0;
  struct StepScratch { std::vector<EntryDesc> desc; std::vector<EntryLane*> desc_lane; int n = 0, n_sorted = 0; bool has_entries = false; } ss;
  // partitioned runs (one process per GPU): lane ownership, halo (ghost tail vehicles) and migration lists
  struct Mp { bool on = false; int rank = 0, world = 1, ghost_base = 0, n_halo = 0, n_export = 0; std::vector<int> lane_owner, halo_lanes, halo_off, export_lanes, export_off;
    DBuf<int> d_lane_owner, d_halo_lanes, d_export_lanes, d_export_peer, d_mig_count, d_free; int free_count = 0; int mig_cap = 0; std::vector<int> h_mig_count; unsigned draw_delta = 0;
    // tail exchange over peer memory: halo area (n_halo records of 4 doubles) + one flag per peer, exported through CUDA IPC
    struct P2p { bool on = false; DBuf<unsigned long long> area; std::vector<void*> peer; DBuf<double*> exp_dst; DBuf<unsigned long long*> flag_dst; DBuf<unsigned> done; DBuf<int> d_halo_off, err;
      unsigned long long seq = 0; } p2p; } mp;
  int* h_stat = nullptr;   // pinned: [0..3] stat, [4] n_exited, [5] n_alive, [6] err, [7] draws, [8] first
};

static void fill_dev_structs(symgpu_ctx* c) {
  DevNet& n = c->dn;
  n.lane_link = c->lane_link.p; n.lane_num = c->lane_num.p; n.lane_prev = c->lane_prev.p; n.lane_flags = c->lane_flags.p; n.lane_len = c->lane_len.p;
  n.succ_off = c->succ_off.p; n.succ_key = c->succ_key.p; n.succ_next = c->succ_next.p;
  n.sl_off = c->sl_off.p; n.sl_i = c->sl_i.p; n.sl_pos = c->sl_pos.p;
  n.link_first = c->link_first.p; n.link_nl = c->link_nl.p; n.link_up = c->link_up.p; n.link_type = c->link_type.p; n.link_brick = c->link_brick.p;
  n.link_len = c->link_len.p; n.link_vreg = c->link_vreg.p; n.route_off = c->route_off.p; n.route_links = c->route_links.p;
  n.sl_col = c->sl_col.p; n.sl_prop = c->sl_prop.p; n.n_lanes = c->net.n_lanes(); n.n_links = c->net.n_links(); n.n_sl = c->net.n_sl();
  DevVeh& v = c->dv;
  v.status = c->v_status.p; v.id = c->v_id.p; v.cls = c->v_cls.p; v.route = c->v_route.p; v.route_pos = c->v_rpos.p; v.lane = c->v_lane.p;
  v.vdes = c->v_vdes.p; n.link_down = c->link_down.p; v.cc_lane = c->cc_lane.p; v.cc_id = c->cc_id.p; v.cc = c->cc.p;
  v.lane_n = c->v_lane_n.p; v.prev_leader = c->v_pl.p; v.flags = c->v_flags.p; v.oflags = c->v_oflags.p; v.memo = c->v_memo.p;
  v.pos = c->v_pos.p; v.vit = c->v_vit.p; v.acc = c->v_acc.p; v.pos_n = c->v_pos_n.p; v.vit_n = c->v_vit_n.p; v.acc_n = c->v_acc_n.p;
  v.dn = c->v_dn.p; v.cpos = c->v_cpos.p; v.tcre = c->v_tcre.p; v.dstp = c->v_dstp.p; v.hent = c->v_hent.p; v.tsort = c->v_tsort.p;
  v.ahead = c->v_ahead.p; v.computed = c->v_computed.p; v.ctx_nl = c->v_ctx_nl.p; v.ctx_lane = c->v_ctx_lane.p; v.ctx_leader = c->v_ctx_leader.p;
  v.ctx_flags = c->v_ctx_flags.p; v.ctx_gap = c->v_ctx_gap.p; v.ctx_dn0 = c->v_ctx_dn0.p; v.ctx_dfree = c->v_ctx_dfree.p;
  DevLaneDyn& l = c->dl;
  l.count = c->lane_count.p; l.start = c->lane_start.p; l.fill = c->lane_fill.p; l.tail = c->lane_tail.p; l.head_done = nullptr;
  l.next_sortie = c->next_sortie.p; l.exit_winner = c->exit_winner.p;
  DevFlow& fl = c->df;
  fl.order = c->order.p; fl.posidx = c->f_posidx.p; fl.isfront = c->f_isfront.p; fl.seg_len = c->f_seg_len.p; fl.seg_dep = c->f_seg_dep.p; fl.segpos = c->f_segpos.p;
  fl.minslot = c->f_minslot.p; fl.cut = c->f_cut.p; fl.fronts = c->f_fronts.p; fl.ulist = c->f_ulist.p;
  fl.ctl = c->f_ctl.p; fl.A = c->f_A.p; fl.A2 = c->f_A2.p; fl.Mn = c->f_Mn.p; fl.Mn2 = c->f_Mn2.p; fl.oncyc = c->f_oncyc.p; fl.cid = c->f_cid.p; fl.rootkey = c->f_rootkey.p;
  DevLc& lc = c->dlc;
  lc.order = c->order.p; lc.lane_start = c->lane_start.p; lc.lane_count = c->lane_count.p; lc.vdes = c->v_vdes.p; lc.voies_ok = c->lc_voies_ok.p; lc.chgt = c->lc_chgt.p;
  lc.delta_next = c->lc_delta_next.p; lc.link_delta_head = c->lc_link_head.p; lc.rng = c->lc_rng.p; lc.n_changes = c->lc_nchg.p;
  lc.moved = c->lc_moved.p; lc.posidx = c->f_posidx.p; lc.spos = c->s_key.p; lc.counters = c->counters.p; lc.kind = c->lc_kind.p; lc.code = c->lc_code.p; lc.tl_pre = c->lc_tl_pre.p; lc.rfol = c->lc_rfol.p; lc.rtgt = c->lc_rtgt.p; lc.had_mand = c->lc_had_mand.p; lc.phi = c->lc_phi.p;
  lc.dst_chgtvoie = c->net.params[SYMFLAT_P_DSTCHGT]; lc.dst_force = c->net.params[SYMFLAT_P_DST_FORCE] > 0 ? c->net.params[SYMFLAT_P_DST_FORCE] : -1.0;
  lc.phi_force = c->net.params[SYMFLAT_P_PHI_FORCE] > 0 ? c->net.params[SYMFLAT_P_PHI_FORCE] : 1.0; lc.ordre = 1;
  DevSig& s = c->ds;
  s.ctrl_seq_off = c->ctrl_seq_off.p; s.ctrl_nseq = c->ctrl_nseq.p; s.ctrl_cycle = c->ctrl_cycle.p; s.seq_sig_off = c->seq_sig_off.p; s.seq_nsig = c->seq_nsig.p;
  s.seq_len = c->seq_len.p; s.sig_in = c->sig_in.p; s.sig_out = c->sig_out.p; s.sig_green = c->sig_green.p; s.sig_amber = c->sig_amber.p; s.sig_delay = c->sig_delay.p;
}

template <class T> static std::vector<T> col(const std::vector<T>& v, int stride, int k) {
  std::vector<T> o(v.size() / stride);
  for (size_t i = 0; i < o.size(); i++) o[i] = v[i * stride + k];
  return o;
}

struct HostRow { int status, id, cls, route, rpos, lane; double pos, vit, acc, tcre, hent; };
static int upload_rows(symgpu_ctx* c, int slot0, const std::vector<HostRow>& rows) {
  size_t n = rows.size();
  if (!n) return SYMGPU_OK;
  std::vector<int> st(n), id(n), cl(n), ro(n), rp(n), la(n), pl(n, -1), fl(n, 0), memo(n * kMemo, -1);
  std::vector<double> po(n), vi(n), ac(n), dnv(n, 1.0), z(n, 0.0), tc(n), he(n);
  for (size_t i = 0; i < n; i++) { const HostRow& r = rows[i]; st[i] = r.status; id[i] = r.id; cl[i] = r.cls; ro[i] = r.route; rp[i] = r.rpos; la[i] = r.lane;
    po[i] = r.pos; vi[i] = r.vit; ac[i] = r.acc; tc[i] = r.tcre; he[i] = r.hent; }
  auto I = [&](DBuf<int>& b, const std::vector<int>& h, int w = 1) { return cudaMemcpyAsync(b.p + (size_t)slot0 * w, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream); };
  auto D = [&](DBuf<double>& b, const std::vector<double>& h) { return cudaMemcpyAsync(b.p + slot0, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream); };
  CK(I(c->v_status, st)); CK(I(c->v_id, id)); CK(I(c->v_cls, cl)); CK(I(c->v_route, ro)); CK(I(c->v_rpos, rp)); CK(I(c->v_lane, la)); CK(I(c->v_lane_n, la));
  CK(I(c->v_vdes, pl)); CK(I(c->v_pl, pl)); CK(I(c->v_flags, fl)); CK(I(c->v_oflags, fl)); CK(I(c->v_memo, memo, kMemo));
  CK(D(c->v_pos, po)); CK(D(c->v_vit, vi)); CK(D(c->v_acc, ac)); CK(D(c->v_pos_n, po)); CK(D(c->v_vit_n, vi)); CK(D(c->v_acc_n, ac));
  CK(D(c->v_dn, dnv)); CK(D(c->v_cpos, z)); CK(D(c->v_tcre, tc)); CK(D(c->v_dstp, z)); CK(D(c->v_hent, he)); CK(D(c->v_tsort, z));
  CK(cudaStreamSynchronize(c->stream));     // host vectors go out of scope
  return SYMGPU_OK;
}

static int build_merge(symgpu_ctx* c);
static double max_debit_max(const symgpu_ctx* c);
extern "C" const char* symgpu_last_error(void) { return g_err.c_str(); }
const symb200::FlatNet& symgpu_internal_net(symgpu_ctx* c) { return c->net; }
int symgpu_internal_class_of(symgpu_ctx* c, int id) { return id >= 0 && id < (int)c->cls_of_id.size() ? c->cls_of_id[id] : -1; }
static void note_class(symgpu_ctx* c, int id, int cls) { if (id < 0) return; if (id >= (int)c->cls_of_id.size()) c->cls_of_id.resize(std::max<size_t>((size_t)id + 1, c->cls_of_id.size() * 2), -1); c->cls_of_id[id] = (signed char)cls; }
int symgpu_internal_create_vehicle(symgpu_ctx* c, int cls, int entry, int dest, int lane, double dbt) {
  int id = c->last_id++;                                             // IncLastIdVeh at request time
  c->api_queue.push_back({id, cls, entry, dest, lane, dbt});
  return id;
}

extern "C" int symgpu_create(const char* path, int device, symgpu_ctx** out) {
  if (!out || !path) return SYMGPU_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { g_err = "no usable CUDA device (this library has no CPU path)"; return SYMGPU_E_NODEVICE; }
  CK(cudaSetDevice(device));
  auto c = new symgpu_ctx(); c->device = device;
  { const size_t pl = strlen(path); const bool flat = pl > 8 && !strcmp(path + pl - 8, ".symflat");
    // anything else is read as a ROOT_SYMUBRUIT document (xmlnet.h: the subset of SURVEY.md Appendix B) and flattened in memory
    if (flat ? !c->net.load(path) : !flat_from_xml(path, c->net, c->net.error)) { g_err = c->net.error; delete c; return SYMGPU_E_LOAD; } }
  FlatNet& f = c->net;
  const int T = f.n_cls(), L = f.n_lanes(), K = f.n_links();
  if (T > kMaxCls) { g_err = "too many vehicle classes"; delete c; return SYMGPU_E_UNSUPPORTED; }
  // every (lane, next link) movement must have a single target lane of rate >= 1 (see micro.cuh next_voie)
  for (int l = 0; l < L; l++) for (int m = f.succ_off[l]; m < f.succ_off[l + 1]; m++) {
    if (f.succ_f[m] < 1.0) { g_err = "movement with taux_utilisation < 1: multi-lane next-lane draws are not supported on the device yet"; delete c; return SYMGPU_E_UNSUPPORTED; }
    for (int m2 = m + 1; m2 < f.succ_off[l + 1]; m2++) if (f.succ_i[3 * m2] == f.succ_i[3 * m]) { g_err = "movement with several target lanes is not supported on the device yet"; delete c; return SYMGPU_E_UNSUPPORTED; }
  }
  c->chgtvoie = f.params[SYMFLAT_P_CHGTVOIE] != 0;
  DevParams& P = c->P; memset(&P, 0, sizeof(P));
  P.dt = f.params[SYMFLAT_P_DT]; P.relax = f.params[SYMFLAT_P_RELAX]; P.duree = f.params[SYMFLAT_P_NSTEPS] * P.dt; P.bevel = 0;
  P.accbornee = f.params[SYMFLAT_P_ACCBORNEE] != 0; P.ncls = T; P.min_kv = 999;
  for (int i = 0; i < T; i++) { const double* r = &f.cls[i * SYMFLAT_CLS_STRIDE]; DevCls& k = P.cls[i];
    k.w = r[0]; k.kx = r[1]; k.vx = r[2]; k.esp = r[3]; k.decel = r[4]; k.tau_lc = r[5]; k.pi_rab = r[6]; k.law = (int)r[7]; k.nplage = std::min(4, (int)r[8]);
    for (int j = 0; j < k.nplage; j++) { k.ax[j] = r[9 + 2 * j]; k.vs[j] = r[10 + 2 * j]; }
    double kv = -k.kx * k.w / (k.vx - k.w); if (kv < P.min_kv) P.min_kv = kv; }     // Reseau::GetMinKv reseau.cpp:9719-9737
  for (int i = 0; i < T; i++) P.maxinf[i] = max_influence(P, P.cls[i]);
  CK(cudaMemcpyToSymbol(cP, &P, sizeof(P)));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
  CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->t_ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->t_packed, cudaEventDisableTiming));
  for (int k = 0; k < kTrajBufs; k++) { CK(cudaEventCreate(&c->t_done[k])); CK(cudaEventCreate(&c->t_begin[k])); }
  CK(cudaMallocHost(&c->h_stat, 16 * sizeof(int)));
  // static tables
  CK(c->lane_link.upload(col(f.lane_i, 4, 0))); CK(c->lane_num.upload(col(f.lane_i, 4, 1))); CK(c->lane_prev.upload(col(f.lane_i, 4, 2)));
  CK(c->lane_flags.upload(col(f.lane_i, 4, 3))); CK(c->lane_len.upload(f.lane_f));
  CK(c->succ_off.upload(f.succ_off)); CK(c->succ_key.upload(col(f.succ_i, 3, 0))); CK(c->succ_next.upload(col(f.succ_i, 3, 1)));
  CK(c->sl_off.upload(f.sl_off)); CK(c->sl_i.upload(f.sl_i)); CK(c->sl_pos.upload(f.sl_f));
  CK(c->link_first.upload(col(f.link_i, 8, 0))); CK(c->link_nl.upload(col(f.link_i, 8, 1))); CK(c->link_up.upload(col(f.link_i, 8, 2)));
  CK(c->link_down.upload(col(f.link_i, 8, 3)));
  CK(c->link_type.upload(col(f.link_i, 8, 4))); CK(c->link_brick.upload(col(f.link_i, 8, 5)));
  CK(c->link_len.upload(col(f.link_f, 2, 0))); CK(c->link_vreg.upload(col(f.link_f, 2, 1)));
  CK(c->route_off.upload(f.route_off)); CK(c->route_links.upload(f.route_links));
  CK(c->ctrl_seq_off.upload(col(f.ctrl_i, 2, 0))); CK(c->ctrl_nseq.upload(col(f.ctrl_i, 2, 1))); CK(c->ctrl_cycle.upload(f.ctrl_f));
  CK(c->seq_sig_off.upload(col(f.seq_i, 2, 0))); CK(c->seq_nsig.upload(col(f.seq_i, 2, 1))); CK(c->seq_len.upload(f.seq_f));
  CK(c->sig_in.upload(col(f.sig_i, 2, 0))); CK(c->sig_out.upload(col(f.sig_i, 2, 1)));
  CK(c->sig_green.upload(col(f.sig_f, 3, 0))); CK(c->sig_amber.upload(col(f.sig_f, 3, 1))); CK(c->sig_delay.upload(col(f.sig_f, 3, 2)));
  CK(c->sl_col.alloc(f.n_sl())); CK(c->sl_prop.alloc(f.n_sl()));
  // exits
  { std::vector<int> el, enl; std::vector<double> ec;
    for (int x = 0; x < f.n_exits(); x++) { int link = f.exit_i[2 * x + 1]; for (int k = 0; k < f.link_i[8 * link + 1]; k++) { el.push_back(f.link_i[8 * link] + k); enl.push_back(f.link_i[8 * link + 1]); ec.push_back(f.exit_f[x]); } }
    CK(c->exit_lanes.upload(el)); CK(c->exit_nl.upload(enl)); CK(c->exit_cap.upload(ec)); }
  // origins
  int n_entry_lanes = 0;
  for (int e = 0; e < f.n_entries(); e++) { const int* r = &f.entry_i[8 * e]; HostEntry h; h.node = r[0]; h.link = r[1];
    for (int d = r[2]; d < r[2] + r[3]; d++) h.demand.push_back({f.dem_f[2 * d], f.dem_f[2 * d + 1]});
    for (int d = r[4]; d < r[4] + r[5]; d++) { h.dest_exit.push_back(f.dest_i[3 * d]); h.dest_coeff.push_back(f.dest_f[d]); std::vector<std::pair<int, double>> rr;
      for (int q = f.dest_i[3 * d + 1]; q < f.dest_i[3 * d + 1] + f.dest_i[3 * d + 2]; q++) rr.push_back({f.droute_i[q], f.droute_f[q]}); h.dest_routes.push_back(rr); }
    int nl = f.link_i[8 * h.link + 1]; h.exponential = (r[7] >> 30) & 1; int tco = r[7] & 0x3fffffff;
    for (int k = 0; k < nl; k++) h.lane_coefs.push_back(f.lanecoef_f[r[6] + k]);
    for (int k = 0; k < T; k++) h.type_coefs.push_back(f.typecoef_f[tco + k]);
    for (int k = 0; k < nl; k++) { EntryLane el; el.lane = f.link_i[8 * h.link] + k; h.lanes.push_back(el); n_entry_lanes++; }
    c->entries.push_back(std::move(h)); }
  // capacity: initial vehicles + demand budget
  double budget = 0;
  for (auto& e : c->entries) { double tt = 0; for (auto& d : e.demand) { double dur = std::min(d.first, std::max(0.0, P.duree - tt)); budget += dur * d.second; tt += d.first; } }
  c->n_pool = n_entry_lanes;
  size_t cap = (size_t)f.n_init() + (size_t)(budget * 1.25) + 2 * (size_t)(P.duree / P.dt) * (size_t)n_entry_lanes / 4 + 4096 + n_entry_lanes;
  if (const char* env = getenv("SYMGPU_CAPACITY")) cap = std::max(cap, (size_t)atoll(env));
  c->cap = (int)cap; c->pool_base = c->cap - c->n_pool;
  const size_t ghost_reserve = std::min<size_t>((size_t)L, 65536);   // slots past `cap` for tail vehicles of lanes owned by other GPUs
  c->mp.ghost_base = c->cap; cap += ghost_reserve;
  { int q = 0; for (auto& e : c->entries) for (auto& l : e.lanes) l.pool_base = c->pool_base + q++; }
  // dynamic lane / link state
  CK(c->lane_count.alloc(L)); CK(c->lane_start.alloc(L)); CK(c->lane_fill.alloc(L)); CK(c->lane_tail.alloc(L));
  CK(c->exit_winner.alloc(L)); CK(cudaMemset(c->exit_winner.p, 0xff, L * sizeof(int))); CK(c->next_sortie.alloc(L));
  CK(c->bsum.alloc((L + kScanB - 1) / kScanB + 1)); CK(c->order.alloc(cap)); CK(c->entre.alloc((size_t)K * T)); CK(c->sorti.alloc((size_t)K * T));
  CK(c->stat.alloc(16)); CK(c->counters.alloc(16)); CK(c->draws.alloc(1));
  for (DBuf<int>* b : {&c->f_posidx, &c->f_isfront, &c->f_seg_len, &c->f_seg_dep, &c->f_segpos, &c->f_minslot, &c->f_cut,
                       &c->f_fronts, &c->f_ulist, &c->rl_lead, &c->f_A, &c->f_A2, &c->f_Mn, &c->f_Mn2, &c->f_oncyc, &c->f_cid}) CK(b->alloc(cap));
  CK(c->t_flag.alloc(cap)); CK(c->t_off.alloc(cap)); CK(c->t_bsum.alloc(cap / kScanB + 2)); CK(c->t_id.alloc(cap)); CK(c->t_lane.alloc(cap));
  CK(c->t_pos.alloc(cap)); CK(c->t_vit.alloc(cap)); CK(c->t_acc.alloc(cap));
  CK(c->f_rootkey.alloc(cap)); CK(c->rows.alloc(((size_t)cap + 32) * kRowD)); CK(c->b_goal.alloc(cap)); CK(c->b_flags.alloc(cap)); CK(c->f_ctl.alloc(1)); CK(cudaMallocHost(&c->h_ctl, sizeof(FlowCtl)));
  CK(c->vl_used.alloc(cap)); CK(c->u_key.alloc(cap)); CK(c->s_key.alloc(cap)); for (DBuf<int>* b : {&c->u_order, &c->u_id, &c->u_lane, &c->s_id, &c->s_lane}) CK(b->alloc(cap)); CK(cudaMallocHost(&c->h_ctl_seed, sizeof(FlowCtl))); memset(c->h_ctl_seed, 0, sizeof(FlowCtl)); c->h_ctl_seed->changed[0] = 1;
  { int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)); c->n_sms = std::max(1, sms); }
  CK(c->exited.alloc(cap)); CK(c->edesc.alloc(std::max(1, n_entry_lanes))); CK(c->eresult.alloc(std::max(1, n_entry_lanes)));
  // vehicles
  for (DBuf<int>* b : {&c->v_status, &c->v_id, &c->v_cls, &c->v_route, &c->v_rpos, &c->v_lane, &c->v_lane_n, &c->v_pl, &c->v_flags, &c->v_oflags, &c->v_ahead,
                       &c->v_computed, &c->v_ctx_nl, &c->v_ctx_leader, &c->v_ctx_flags}) CK(b->alloc(cap));
  CK(c->v_memo.alloc(cap * kMemo)); CK(c->v_ctx_lane.alloc(cap * kMaxCtx));
  { const char* e = getenv("SYMGPU_COOP_LOOPS"); int co = 0; cudaDeviceGetAttribute(&co, cudaDevAttrCooperativeLaunch, device); c->coop_loops = co && !(e && !strcmp(e, "0")); }
  { const char* e = getenv("SYMGPU_DEBUG_SYNC"); c->dbg_sync = e && !strcmp(e, "1"); }
  { const char* e = getenv("SYMGPU_SIDE"); c->side_mask = e ? atoi(e) : 31; }                            // experiments: bit 0 signals, 1 early stream advance, 2 shared points, 3 sensors, 4 clean points on side streams
  { const char* e = getenv("SYMGPU_CTXCACHE"); c->use_cc = !(e && !strcmp(e, "0")); }                    // SYMGPU_CTXCACHE=0: every vehicle walks the tables every step (cross-check)
  if (c->use_cc) { CK(c->cc_lane.alloc(cap)); CK(cudaMemset(c->cc_lane.p, 0xff, cap * sizeof(int))); CK(c->cc_id.alloc(cap)); CK(cudaMemset(c->cc_id.p, 0xff, cap * sizeof(int)));
    CK(c->cc_list.alloc(cap)); CK(c->cc_cnt.alloc(1)); CK(c->cc.alloc(cap)); CK(cudaMemset(c->cc.p, 0xff, cap * sizeof(CtxCache))); }   // key (-1, -1): matches no vehicle
  CK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->aux2, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&c->ev_fork, &c->ev_rng, &c->ev_dirty, &c->ev_sens, &c->ev_eval}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->cc_go, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->cc_done, cudaEventDisableTiming));
  CK(c->v_vdes.alloc(cap)); CK(cudaMemset(c->v_vdes.p, 0xff, cap * sizeof(int))); CK(c->lc_voies_ok.alloc(cap)); CK(c->lc_chgt.alloc(cap)); CK(c->lc_delta_next.alloc(cap));
  CK(c->lc_link_head.alloc(K)); CK(c->lc_nchg.alloc(2)); CK(c->lc_rng.alloc(1));
  if (c->chgtvoie) { CK(c->lc_kind.alloc(2 * cap)); CK(c->lc_code.alloc(3 * cap)); CK(c->lc_tl_pre.alloc(3 * cap)); CK(c->lc_rfol.alloc(3 * cap)); CK(c->lc_rtgt.alloc(3 * cap));
    CK(c->lc_phi.alloc(3 * cap)); CK(c->lc_had_mand.alloc(cap)); CK(c->lc_moved.alloc(cap)); const char* e = getenv("SYMGPU_LC"); c->lc_seq = e && !strcmp(e, "seq"); }
  for (DBuf<double>* b : {&c->v_pos, &c->v_vit, &c->v_acc, &c->v_pos_n, &c->v_vit_n, &c->v_acc_n, &c->v_dn, &c->v_cpos, &c->v_tcre, &c->v_dstp, &c->v_hent,
                          &c->v_tsort, &c->v_ctx_gap, &c->v_ctx_dn0, &c->v_ctx_dfree}) CK(b->alloc(cap));
  fill_dev_structs(c);
  { int rc = build_merge(c); if (rc != SYMGPU_OK) { delete c; return rc; } }
  c->rng.seed((unsigned)f.params[SYMFLAT_P_SEED]);                    // reseau.cpp:983-995
  CK(cudaMemcpy(c->lc_rng.p, &c->rng, sizeof(DevRng), cudaMemcpyHostToDevice));   // the stream lives on the device; the host copy is refreshed when the origins draw
  // initial vehicles (reseau.cpp:1293-1340): already on the network, state at index 0
  { std::vector<HostRow> rows(f.n_init());
    for (int v = 0; v < f.n_init(); v++) { HostRow& r = rows[v]; r.status = ST_ALIVE; r.id = f.iv_id.empty() ? c->last_id++ : f.iv_id[v]; if (!f.iv_id.empty()) c->last_id = std::max(c->last_id, r.id + 1); r.cls = f.iv_i[4 * v]; r.lane = f.iv_i[4 * v + 1];
      r.route = f.iv_i[4 * v + 2]; r.rpos = f.iv_i[4 * v + 3]; r.pos = f.iv_f[3 * v]; r.vit = f.iv_f[3 * v + 1]; r.acc = f.iv_f[3 * v + 2]; r.tcre = 0; r.hent = 0; note_class(c, r.id, r.cls); }
    int rc = upload_rows(c, 0, rows); if (rc != SYMGPU_OK) { delete c; return rc; }
    c->n_slots = c->n_alive = f.n_init(); }
  *out = c;
  return SYMGPU_OK;
}

extern "C" void symgpu_destroy(symgpu_ctx* c) {
  if (c) for (void* p : c->mp.p2p.peer) if (p) cudaIpcCloseMemHandle(p);
  if (c && c->traj_host_n && getenv("SYMGPU_TIMING")) fprintf(stderr, "symgpu: trajectory enqueue %.1f us of host time per call (%ld calls)\n", c->traj_host_us / c->traj_host_n, c->traj_host_n);
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (DBuf<int>* b : {&c->lane_link, &c->lane_num, &c->lane_prev, &c->lane_flags, &c->succ_off, &c->succ_key, &c->succ_next, &c->sl_off, &c->sl_i, &c->link_first,
                       &c->link_nl, &c->link_up, &c->link_type, &c->link_brick, &c->route_off, &c->route_links, &c->ctrl_seq_off, &c->ctrl_nseq, &c->seq_sig_off,
                       &c->seq_nsig, &c->sig_in, &c->sig_out, &c->exit_lanes, &c->exit_nl, &c->sl_col, &c->lane_count, &c->lane_start, &c->lane_fill, &c->lane_tail,
                       &c->exit_winner, &c->bsum, &c->order, &c->entre, &c->sorti, &c->stat, &c->eresult, &c->v_status, &c->v_id,
                       &c->f_posidx, &c->f_isfront, &c->f_seg_len, &c->f_seg_dep, &c->f_segpos, &c->f_minslot, &c->f_cut,
                       &c->f_fronts, &c->f_ulist, &c->rl_lead, &c->f_A, &c->f_A2, &c->f_Mn, &c->f_Mn2, &c->f_oncyc, &c->f_cid,
                       &c->v_cls, &c->v_route, &c->v_rpos, &c->v_lane, &c->v_lane_n, &c->v_pl, &c->v_flags, &c->v_oflags, &c->v_memo, &c->v_ahead, &c->v_computed,
                       &c->v_ctx_nl, &c->v_ctx_lane, &c->v_ctx_leader, &c->v_ctx_flags}) b->free();
  for (DBuf<double>* b : {&c->lane_len, &c->sl_pos, &c->link_len, &c->link_vreg, &c->ctrl_cycle, &c->seq_len, &c->sig_green, &c->sig_amber, &c->sig_delay, &c->exit_cap,
                          &c->sl_prop, &c->next_sortie, &c->v_pos, &c->v_vit, &c->v_acc, &c->v_pos_n, &c->v_vit_n, &c->v_acc_n, &c->v_dn, &c->v_cpos, &c->v_tcre,
                          &c->v_dstp, &c->v_hent, &c->v_tsort, &c->v_ctx_gap, &c->v_ctx_dn0, &c->v_ctx_dfree}) b->free();
  for (int k = 0; k < kTrajBufs; k++) if (c->tbuf[k].registered) { cudaHostUnregister(c->tbuf[k].id); cudaHostUnregister(c->tbuf[k].lane); cudaHostUnregister(c->tbuf[k].pos); cudaHostUnregister(c->tbuf[k].vit); cudaHostUnregister(c->tbuf[k].acc); }
  c->t_flag.free(); c->t_off.free(); c->t_bsum.free(); c->t_id.free(); c->t_lane.free(); c->t_pos.free(); c->t_vit.free(); c->t_acc.free();
  if (c->aux2) cudaStreamDestroy(c->aux2); for (cudaEvent_t e : {c->ev_fork, c->ev_rng, c->ev_dirty, c->ev_sens, c->ev_eval}) if (e) cudaEventDestroy(e);
  if (c->aux) cudaStreamDestroy(c->aux); if (c->cc_go) cudaEventDestroy(c->cc_go); if (c->cc_done) cudaEventDestroy(c->cc_done);
  if (c->side) cudaStreamDestroy(c->side); if (c->t_ready) cudaEventDestroy(c->t_ready); if (c->t_packed) cudaEventDestroy(c->t_packed); for (int k = 0; k < kTrajBufs; k++) { if (c->t_done[k]) cudaEventDestroy(c->t_done[k]); if (c->t_begin[k]) cudaEventDestroy(c->t_begin[k]); }
  c->counters.free(); c->draws.free(); c->exited.free(); c->edesc.free(); c->f_rootkey.free(); c->f_ctl.free(); c->rows.free(); c->b_goal.free(); c->b_flags.free();
  if (c->h_ctl) cudaFreeHost(c->h_ctl);
  if (c->h_ctl_seed) cudaFreeHost(c->h_ctl_seed);
  c->vl_used.free(); c->u_key.free(); c->s_key.free(); for (DBuf<int>* b : {&c->u_order, &c->u_id, &c->u_lane, &c->s_id, &c->s_lane}) b->free();
  if (c->h_stat) cudaFreeHost(c->h_stat);
  if (c->ev0) cudaEventDestroy(c->ev0); if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream && !c->ext_stream) cudaStreamDestroy(c->stream);
  delete c;
}

// ---- origins (host): usage/SymuViaTripNode.cpp:847-955, :212-325, :331-840 ---------------------------
static double demande_value(const HostEntry& e, double inst) {
  double acc = 0; for (auto& d : e.demand) { acc += d.first; if (inst <= acc) return d.second; }
  return e.demand.empty() ? 0.0 : e.demand.back().second;
}
static void update_crit(symgpu_ctx* c, HostEntry& e, double inst, bool due_to_creation, double max_debit) {
  double dt = c->P.dt;
  if (!e.exponential) { e.crit += demande_value(e, inst) * dt; return; }                           // :905-911
  if (fabs(demande_value(e, inst) - demande_value(e, inst - dt)) > DBL_EPSILON || fabs(inst - dt) < DBL_EPSILON || due_to_creation) {
    double niveau = demande_value(e, inst); if (niveau > max_debit) niveau = max_debit;
    double r = 0; while (r <= 0) r = (double)c->rng.next() / 32767.0;
    e.crit = niveau > 0 ? inst - log(r) * (1 - niveau / max_debit) / niveau + 1 / max_debit : DBL_MAX;   // :935-940
  }
}
static double max_debit_max(const symgpu_ctx* c) { double m = 0; for (int i = 0; i < c->P.ncls; i++) { const DevCls& k = c->P.cls[i]; double q = k.w * k.kx / ((k.w / k.vx) - 1); if (q > m) m = q; } return m; }


// ---- merge points: tables of merge.cuh from the SYMFLAT1 sections (Convergent::Init / FinInit convergent.cpp:208-255, the loader's
// CarrefourAFeuxEx::Init :489-919) --------------------------------------------------------------------------------------------
static int build_merge(symgpu_ctx* c) {
  const FlatNet& f = c->net; auto& g = c->mg; DevMrg& h = g.h; memset(&h, 0, sizeof(h));
  const int K = f.n_links(), L = f.n_lanes(), T = f.n_cls(), nc = f.n_cvg(), np = f.n_pt();
  g.on = nc > 0;
  std::vector<int> link_cvg(K, -1), cvg_node(nc), cvg_down(nc), cvg_pt0(nc), cvg_npt(nc), cvg_am0(nc), cvg_nam(nc), cvg_win(nc), cvg_sens0(nc), cvg_am;
  std::vector<double> cvg_gamma(nc), cvg_mu(nc), cvg_poscpt(nc), sens_pos; std::vector<int> sens_link, sens_cnt0, sens_win, row_sens, sens_row0;
  size_t cnt_total = 0;
  for (int k = 0; k < nc; k++) { const int* r = &f.cvg_i[8 * k]; cvg_node[k] = r[0]; cvg_down[k] = r[1]; cvg_pt0[k] = r[2]; cvg_npt[k] = r[3]; cvg_am0[k] = r[4]; cvg_nam[k] = r[5];
    cvg_gamma[k] = f.cvg_f[4 * k + 1]; cvg_mu[k] = f.cvg_f[4 * k + 2]; cvg_poscpt[k] = f.cvg_f[4 * k + 3];
    cvg_win[k] = (int)(f.cvg_f[4 * k] / c->P.dt);                                          // PonctualSensor::InitData sensors/PonctualSensor.cpp:131
    if (cvg_win[k] < 1) { g_err = "merge sensor window shorter than one time step"; return SYMGPU_E_UNSUPPORTED; }
    cvg_sens0[k] = (int)sens_link.size();
    auto add = [&](int link, double pos) { const int nl = f.link_i[8 * link + 1]; const int s = (int)sens_link.size();
      sens_link.push_back(link); sens_pos.push_back(pos); sens_win.push_back(cvg_win[k]); sens_cnt0.push_back((int)cnt_total); sens_row0.push_back((int)row_sens.size() / 2);
      for (int q = 0; q < T * nl; q++) { row_sens.push_back(s); row_sens.push_back(q); }
      cnt_total += (size_t)T * nl * cvg_win[k]; };
    add(cvg_down[k], cvg_poscpt[k]);                                                       // "CptTav" (FinInit :239)
    for (int a = r[4]; a < r[4] + r[5]; a++) { const int am = f.cvgam_i[a]; add(am, f.link_f[2 * am] - 5); link_cvg[am] = k; }   // "CptAm" at length - 5 (:247)
  }
  cvg_am = f.cvgam_i;
  if (cnt_total > 0x7fffffffULL) { g_err = "merge sensor storage too large"; return SYMGPU_E_CAPACITY; }
  std::vector<int> lsens0(K + 1, 0), lsens(sens_link.size());
  for (int s : sens_link) lsens0[s + 1]++;
  for (int k = 0; k < K; k++) lsens0[k + 1] += lsens0[k];
  { std::vector<int> fill(lsens0.begin(), lsens0.end() - 1); for (size_t s = 0; s < sens_link.size(); s++) lsens[fill[sens_link[s]]++] = (int)s; }
  // points and their upstream lanes
  std::vector<int> pt_cvg(np), pt_down(np), pt_vam0(np), pt_nvam(np), pt_mode(np, 2), vam_lane(f.ptam_i.size() / 2), vam_ref(vam_lane.size()), lane_pt(L, -1), lane_vam(L, -1);
  for (int p = 0; p < np; p++) { pt_cvg[p] = f.pt_i[4 * p]; pt_down[p] = f.pt_i[4 * p + 1]; pt_vam0[p] = f.pt_i[4 * p + 2]; pt_nvam[p] = f.pt_i[4 * p + 3]; }
  for (size_t a = 0; a < vam_lane.size(); a++) { vam_lane[a] = f.ptam_i[2 * a]; vam_ref[a] = f.ptam_i[2 * a + 1]; }
  for (int k = 0; k < nc; k++) for (int p = cvg_pt0[k]; p < cvg_pt0[k] + cvg_npt[k]; p++) for (int a = pt_vam0[p]; a < pt_vam0[p] + pt_nvam[p]; a++)
    if (lane_pt[vam_lane[a]] < 0) { lane_pt[vam_lane[a]] = p; lane_vam[vam_lane[a]] = a; }
  // junction movements through every point whose priorities follow the signals (CarrefourAFeuxEx::UpdateConvergents :1386-1593)
  const int nm = (int)f.cafmvt_i.size() / 5;
  std::vector<int> pt_pm0(np + 1, 0), pm_vam, pm_sl, pm_occ0(1, 0), pm_occ;
  std::vector<std::vector<int>> node_mvts(f.n_nodes());
  for (int m = 0; m < nm; m++) node_mvts[f.cafmvt_i[5 * m]].push_back(m);
  for (int k = 0; k < nc; k++) { const int node = cvg_node[k]; const bool ctl = node >= 0 && f.node_i[4 * node + 1] >= 0;
    bool stop = !ctl;
    for (int p = cvg_pt0[k]; p < cvg_pt0[k] + cvg_npt[k]; p++) {
      if (stop) { pt_mode[p] = 2; continue; }
      if (pt_nvam[p] == 0) { pt_mode[p] = 2; stop = true; continue; }                       // :1428-1429 leaves the loop over the points
      if (pt_nvam[p] == 1) { pt_mode[p] = 1; stop = true; continue; }                       // :1431-1435
      pt_mode[p] = 0;
    } }
  std::unordered_map<unsigned long long, int> sl_of;                                       // (entry link, exit link) -> first stop line of the movement
  for (int s2 = f.n_sl() - 1; s2 >= 0; s2--) sl_of[((unsigned long long)(unsigned)f.sl_i[6 * s2 + 1] << 32) | (unsigned)f.sl_i[6 * s2 + 3]] = s2;
  for (int p = 0; p < np; p++) { pt_pm0[p] = (int)pm_vam.size();
    if (pt_mode[p] == 0) { const int node = cvg_node[pt_cvg[p]]; const int ctrl = f.node_i[4 * node + 1];
      for (int m : node_mvts[node]) { const int* r = &f.cafmvt_i[5 * m]; const int ent = r[1], ex = r[2];
        int tm = -1;                                                                       // first internal link of the movement that leads into the point
        for (int q = r[3]; q < r[3] + r[4] && tm < 0; q++) for (int a = pt_vam0[p]; a < pt_vam0[p] + pt_nvam[p]; a++) if (f.lane_i[4 * vam_lane[a]] == f.cafmvt_links[q]) { tm = q; break; }
        if (tm < 0) continue;
        int sl = -1; { auto it = sl_of.find(((unsigned long long)(unsigned)ent << 32) | (unsigned)ex); if (it != sl_of.end() && f.sl_i[6 * it->second] == ctrl) sl = it->second; }
        if (sl < 0) { g_err = "merge point inside a signalised junction: movement without a stop line (signal colour unknown)"; return SYMGPU_E_UNSUPPORTED; }
        for (int q = r[3]; q < r[3] + r[4]; q++) for (int a = pt_vam0[p]; a < pt_vam0[p] + pt_nvam[p]; a++) if (f.lane_i[4 * vam_lane[a]] == f.cafmvt_links[q]) {
          pm_vam.push_back(a); pm_sl.push_back(sl);
          for (int q2 = r[3]; q2 <= tm; q2++) pm_occ.push_back(f.link_i[8 * f.cafmvt_links[q2]]);
          pm_occ0.push_back((int)pm_occ.size()); break; }
      } } }
  pt_pm0[np] = (int)pm_vam.size();
  // movements an internal link belongs to; first movement (lane-pair order) between an entry link and an exit link
  std::vector<int> lmv0(K + 1, 0), lmv_ent, lmv_exit, fm0(K + 1, 0), fm_exit, fm_tl0, fm_ntl, fm_tl;
  { std::vector<std::vector<std::pair<int, int>>> per(K);
    for (int m = 0; m < nm; m++) { const int* r = &f.cafmvt_i[5 * m]; for (int q = r[3]; q < r[3] + r[4]; q++) per[f.cafmvt_links[q]].push_back({r[1], r[2]}); }
    for (int k = 0; k < K; k++) { for (auto& e : per[k]) { lmv_ent.push_back(e.first); lmv_exit.push_back(e.second); } lmv0[k + 1] = (int)lmv_ent.size(); }
    struct Fm { int ex, a, b, m; };
    std::vector<std::vector<Fm>> pe(K);
    for (int m = 0; m < nm; m++) { const int* r = &f.cafmvt_i[5 * m]; if (r[4] <= 0) continue;
      const int first_il = f.cafmvt_links[r[3]], last_il = f.cafmvt_links[r[3] + r[4] - 1];
      const int ent_lane = f.lane_i[4 * f.link_i[8 * first_il] + 2]; const int a = ent_lane >= 0 ? f.lane_i[4 * ent_lane + 1] : 0;
      const int ll = f.link_i[8 * last_il]; int b = 0; for (int q = f.succ_off[ll]; q < f.succ_off[ll + 1]; q++) if (f.succ_i[3 * q] == r[2]) { b = f.lane_i[4 * f.succ_i[3 * q + 1] + 1]; break; }
      auto& v = pe[r[1]]; bool seen = false;
      for (auto& e : v) if (e.ex == r[2]) { seen = true; if (a < e.a || (a == e.a && b < e.b)) { e.a = a; e.b = b; e.m = m; } }
      if (!seen) v.push_back({r[2], a, b, m}); }
    for (int k = 0; k < K; k++) { for (auto& e : pe[k]) { const int* r = &f.cafmvt_i[5 * e.m]; fm_exit.push_back(e.ex); fm_tl0.push_back((int)fm_tl.size()); fm_ntl.push_back(r[4]);
        for (int q = r[3]; q < r[3] + r[4]; q++) fm_tl.push_back(f.cafmvt_links[q]); }
      fm0[k + 1] = (int)fm_exit.size(); } }
  std::vector<int> node_type(f.n_nodes()); for (int n = 0; n < f.n_nodes(); n++) node_type[n] = f.node_i[4 * n];
  auto one = [](std::vector<int>& v) { if (v.empty()) v.push_back(0); };
  one(cvg_am); one(pm_vam); one(pm_sl); one(pm_occ); one(lmv_ent); one(lmv_exit); one(fm_exit); one(fm_tl0); one(fm_ntl); one(fm_tl); one(lsens); one(row_sens);
  std::vector<int> lcam_l = f.lcam_links, lcav_l = f.lcav_links; one(lcam_l); one(lcav_l);
  CK(g.link_cvg.upload(link_cvg)); CK(g.cvg_node.upload(cvg_node)); CK(g.cvg_down.upload(cvg_down)); CK(g.cvg_pt0.upload(cvg_pt0)); CK(g.cvg_npt.upload(cvg_npt)); CK(g.cvg_am0.upload(cvg_am0));
  CK(g.cvg_nam.upload(cvg_nam)); CK(g.cvg_win.upload(cvg_win)); CK(g.cvg_sens0.upload(cvg_sens0)); CK(g.cvg_am.upload(cvg_am)); CK(g.cvg_gamma.upload(cvg_gamma)); CK(g.cvg_mu.upload(cvg_mu));
  CK(g.cvg_poscpt.upload(cvg_poscpt)); CK(g.cvg_tf.upload(f.cvg_tf)); CK(g.pt_cvg.upload(pt_cvg)); CK(g.pt_down.upload(pt_down)); CK(g.pt_vam0.upload(pt_vam0)); CK(g.pt_nvam.upload(pt_nvam));
  CK(g.pt_mode.upload(pt_mode)); CK(g.pt_pm0.upload(pt_pm0)); CK(g.vam_lane.upload(vam_lane)); CK(g.vam_ref.upload(vam_ref)); CK(g.vam_niv.upload(vam_ref)); CK(g.pm_vam.upload(pm_vam)); CK(g.pm_sl.upload(pm_sl));
  CK(g.pm_occ0.upload(pm_occ0)); CK(g.pm_occ.upload(pm_occ)); CK(g.lane_pt.upload(lane_pt)); CK(g.lane_vam.upload(lane_vam)); CK(g.sens_link.upload(sens_link)); CK(g.sens_cnt0.upload(sens_cnt0));
  CK(g.sens_win.upload(sens_win)); { if (sens_row0.empty()) sens_row0.push_back(0); CK(g.sens_row0.upload(sens_row0)); CK(g.row_tot.alloc(std::max<size_t>(1, row_sens.size() / 2))); } CK(g.sens_pos.upload(sens_pos)); CK(g.cnt.alloc(std::max<size_t>(1, cnt_total))); CK(g.lsens0.upload(lsens0)); CK(g.lsens.upload(lsens));
  CK(g.lcam0.upload(f.lcam_off)); CK(g.lcam.upload(lcam_l)); CK(g.lcav0.upload(f.lcav_off)); CK(g.lcav.upload(lcav_l)); CK(g.lmv0.upload(lmv0)); CK(g.lmv_ent.upload(lmv_ent)); CK(g.lmv_exit.upload(lmv_exit));
  CK(g.fm0.upload(fm0)); CK(g.fm_exit.upload(fm_exit)); CK(g.fm_tl0.upload(fm_tl0)); CK(g.fm_ntl.upload(fm_ntl)); CK(g.fm_tl.upload(fm_tl)); CK(g.node_type.upload(node_type)); CK(g.row_sens.upload(row_sens));
  g.n_rows = (int)row_sens.size() / 2; if (sens_link.empty()) g.n_rows = 0;
  const size_t cap = c->v_status.n;
  CK(g.pt_last.alloc(std::max(1, np))); CK(g.pt_rand.alloc(std::max(1, np))); CK(g.pt_free.alloc(std::max(1, np))); CK(g.cand_up.alloc(L)); CK(cudaMemset(g.cand_up.p, 0xff, L * sizeof(int)));
  { std::vector<int> mx(cap, INT_MAX), mn(cap, -1); CK(g.own_min.upload(mx)); CK(g.own_max.upload(mn)); }
  for (DBuf<int>* b : {&g.st, &g.ninv, &g.pre, &g.kmax, &g.decide, &g.ndraw, &g.flist, &g.fsorted, &g.res, &g.dlist}) CK(b->alloc(std::max(1, np)));
  CK(g.mark.alloc(cap));
  CK(g.rbuf.alloc(1 << 20)); CK(g.kbsum.alloc(np / 1024 + 3)); CK(g.kf.alloc(std::max(1, np))); CK(g.kfo.alloc(std::max(1, np)));
  CK(g.inv.alloc((size_t)std::max(1, np) * kMrgInv)); CK(g.proba.alloc(std::max(1, np))); CK(g.nused.alloc(cap)); CK(g.ulanes.alloc(cap * 3)); CK(g.nvoie.alloc(cap)); CK(cudaMemset(g.nvoie.p, 0xff, cap * sizeof(int)));
  CK(g.mrgx.alloc(((size_t)cap + 32) * kMrgX)); CK(g.flag.alloc(1)); CK(g.out.alloc(4)); CK(g.seen.alloc(1));
  h.n_cvg = nc; h.n_pt = np; h.n_sens = (int)sens_link.size(); h.n_cls = T;
  h.max_debit_max = max_debit_max(c); h.max_vit_max = 0; for (int i = 0; i < T; i++) h.max_vit_max = std::max(h.max_vit_max, c->P.cls[i].vx); h.vx0 = c->P.cls[0].vx;
  h.link_cvg = g.link_cvg.p; h.cvg_node = g.cvg_node.p; h.cvg_down = g.cvg_down.p; h.cvg_pt0 = g.cvg_pt0.p; h.cvg_npt = g.cvg_npt.p; h.cvg_am0 = g.cvg_am0.p; h.cvg_nam = g.cvg_nam.p; h.cvg_win = g.cvg_win.p;
  h.cvg_sens0 = g.cvg_sens0.p; h.cvg_am = g.cvg_am.p; h.cvg_gamma = g.cvg_gamma.p; h.cvg_mu = g.cvg_mu.p; h.cvg_poscpt = g.cvg_poscpt.p; h.cvg_tf = g.cvg_tf.p;
  h.pt_cvg = g.pt_cvg.p; h.pt_down = g.pt_down.p; h.pt_vam0 = g.pt_vam0.p; h.pt_nvam = g.pt_nvam.p; h.pt_mode = g.pt_mode.p; h.pt_pm0 = g.pt_pm0.p; h.vam_lane = g.vam_lane.p; h.vam_ref = g.vam_ref.p; h.vam_niv = g.vam_niv.p;
  h.pm_vam = g.pm_vam.p; h.pm_sl = g.pm_sl.p; h.pm_occ0 = g.pm_occ0.p; h.pm_occ = g.pm_occ.p; h.lane_pt = g.lane_pt.p; h.lane_vam = g.lane_vam.p;
  h.sens_link = g.sens_link.p; h.sens_cnt0 = g.sens_cnt0.p; h.sens_pos = g.sens_pos.p; h.cnt = g.cnt.p; h.sens_win = g.sens_win.p; h.sens_row0 = g.sens_row0.p; h.row_tot = g.row_tot.p; h.lsens0 = g.lsens0.p; h.lsens = g.lsens.p;
  h.lcam0 = g.lcam0.p; h.lcam = g.lcam.p; h.lcav0 = g.lcav0.p; h.lcav = g.lcav.p; h.lmv0 = g.lmv0.p; h.lmv_ent = g.lmv_ent.p; h.lmv_exit = g.lmv_exit.p;
  h.fm0 = g.fm0.p; h.fm_exit = g.fm_exit.p; h.fm_tl0 = g.fm_tl0.p; h.fm_ntl = g.fm_ntl.p; h.fm_tl = g.fm_tl.p; h.node_type = g.node_type.p;
  h.pt_last = g.pt_last.p; h.pt_rand = g.pt_rand.p; h.pt_free = g.pt_free.p; h.cand_up = g.cand_up.p; h.own_min = g.own_min.p; h.own_max = g.own_max.p;
  h.st = g.st.p; h.inv = g.inv.p; h.ninv = g.ninv.p; h.pre = g.pre.p; h.kmax = g.kmax.p; h.decide = g.decide.p; h.ndraw = g.ndraw.p; h.flist = g.flist.p; h.proba = g.proba.p; h.kf = g.kf.p; h.kfo = g.kfo.p; h.fsorted = g.fsorted.p; h.rbuf = g.rbuf.p; h.rbuf_cap = 1 << 20; h.res = g.res.p; h.dlist = g.dlist.p; h.n_dirty = g.flag.p; h.mark = g.mark.p; h.sens_go = g.out.p + 1; h.eval_done = g.out.p + 2;
  h.nused = g.nused.p; h.nvoie = g.nvoie.p; h.ulanes = g.ulanes.p; h.mrgx = g.mrgx.p; h.rng = c->lc_rng.p; h.draws = c->draws.p; h.counters = c->counters.p;
  std::vector<DevMrg> one_h(1, h); CK(g.d.upload(one_h));
  return SYMGPU_OK;
}

// SymCreateVehicle path (reseau.cpp:13351 -> :2233-2240 -> GenerateVehicle with type, lane and destination imposed):
// only the itinerary (:515), law-coefficient and Jorge draws remain; the vehicle joins m_LstVehicles BEFORE
// InitVehicules, so it is shifted once (state at index 1, position still undefined) like an older ready vehicle.
// partitioned runs: every rank replays the creation of ALL origins (same seed, same draws, same vehicle ids) and keeps the
// vehicles whose entry lane it owns, so the global random sequence and the ids need no exchange
static inline bool owns_lane(symgpu_ctx* c, int lane) { return !c->mp.on || c->mp.lane_owner[lane] == c->mp.rank; }

static void create_api_vehicles(symgpu_ctx* c, std::vector<HostRow>& new_rows, std::vector<EntryLane*>& new_row_lane) {
  const double dt = c->P.dt, inst = c->t;
  for (auto& a : c->api_queue) {
    HostEntry& e = c->entries[a.entry];
    int nl = (int)e.lanes.size(); int nv = a.lane;
    if (nv < 0) { double r = c->rng.next() / 32767.0, sum = 0; for (nv = 0; nv < nl; nv++) { double cf = e.lane_coefs[nv]; if (cf <= 0) continue; sum += cf; if (sum >= r) break; } if (nv >= nl) nv = nl - 1; }
    if (c->P.cls[a.cls].law != 0) c->rng.next();
    int route = e.dest_routes[a.dest].back().first;
    { double r = c->rng.next() / 32767.0, sum = 0; for (auto& q : e.dest_routes[a.dest]) { sum += q.second; if (r <= sum) { route = q.first; break; } } }
    c->rng.next();
    EntryLane& el = e.lanes[nv];
    if (!owns_lane(c, el.lane)) continue;
    c->created++;
    HostVeh hv{a.id, a.cls, route, el.lane, a.dbt, inst - dt + a.dbt}; note_class(c, a.id, a.cls);
    if (el.pret_slot != -1 || !el.queue.empty() || el.pool_slot != -1) el.queue.push_back(hv);
    else { HostRow r{ST_ALIVE, a.id, a.cls, route, 0, el.lane, SYM_UNDEF, c->P.cls[a.cls].vx, 0.0, a.dbt, hv.hent}; new_rows.push_back(r); new_row_lane.push_back(&el);
      el.pret_slot = -2; /* ready vehicle registered, slot assigned at upload */ }
  }
  c->api_queue.clear();
}

// returns rows to append to m_LstVehicles this step (ready vehicles), fills waiting queues
static void create_vehicles(symgpu_ctx* c, std::vector<HostRow>& new_rows, std::vector<EntryLane*>& new_row_lane) {
  const double dt = c->P.dt, inst = c->t; const FlatNet& f = c->net; const double mdm = max_debit_max(c);
  for (auto& e : c->entries) {
    for (;;) {                                                                                      // CreationVehicules :212-325
      bool creation; double tcre = 0;
      if (e.exponential) { creation = e.crit <= inst; if (creation) { tcre = e.crit - (inst - dt); if (inst - dt + tcre >= 0) update_crit(c, e, inst - dt + tcre, true, mdm); } }
      else { creation = e.crit >= 1; if (creation) { double q = demande_value(e, inst); if (q != 0.0) { tcre = dt - (e.crit - 1) * dt / (q * dt); e.crit -= 1; } } }
      if (!creation) break;
      // GenerateVehicle :331-840 — draw order: lane (:1966), type (RepartitionTypeVehicule.cpp:92), [law coefficient
      // CarFollowingFactory.cpp:99], destination (:1288), itinerary (:515), Jorge class (vehicule.cpp:749)
      int nl = (int)e.lanes.size(); int nv = 0;
      { double r = c->rng.next() / 32767.0, sum = 0; for (; nv < nl; nv++) { double cf = e.lane_coefs[nv]; if (cf <= 0) continue; sum += cf; if (sum >= r) break; } if (nv >= nl) nv = nl - 1; }
      int tv = -1;
      { double r = c->rng.next() / 32767.0, sum = 0; int lastnn = -1; for (int i = 0; i < c->P.ncls; i++) { double cf = e.type_coefs[i]; if (cf <= 0) continue; lastnn = i; sum += cf; if (r <= sum) { tv = i; break; } } if (tv < 0) tv = lastnn; }
      if (c->P.cls[tv].law != 0) c->rng.next();
      int id = c->last_id++;
      int d = -1;
      { double r = (double)c->rng.next() / 32767.0, sum = 0; for (size_t i = 0; i < e.dest_coeff.size(); i++) { if (e.dest_coeff[i] <= 0) continue; sum += e.dest_coeff[i]; if (i == e.dest_coeff.size() - 1) sum = 1; if (r <= sum) { d = (int)i; break; } } }
      if (d < 0) continue;
      int route = e.dest_routes[d].back().first;
      { double r = c->rng.next() / 32767.0, sum = 0; for (auto& q : e.dest_routes[d]) { sum += q.second; if (r <= sum) { route = q.first; break; } } }
      c->rng.next();                                                                                // TirageJorge
      EntryLane& el = e.lanes[nv];
      if (!owns_lane(c, el.lane)) continue;
      c->created++;
      HostVeh hv{id, tv, route, el.lane, tcre, inst - dt + tcre}; note_class(c, id, tv);
      (void)f;
      if (el.pret_slot != -1 || !el.queue.empty() || el.pool_slot != -1) el.queue.push_back(hv);   // :794-800
      else { HostRow r{ST_NEW, id, tv, route, 0, el.lane, SYM_UNDEF, c->P.cls[tv].vx, 0.0, tcre, hv.hent}; new_rows.push_back(r); new_row_lane.push_back(&el);
        el.pret_slot = -2; /* ready vehicle registered, slot assigned at upload */ }
    }
  }
}

// per-kernel device timing (CUDA events on the step stream), enabled by symgpu_profile(); bench.py uses it for the roofline
enum { KID_SIGNALS = 0, KID_PREPARE, KID_SCAN, KID_SCATTER, KID_SORT, KID_CONTEXT, KID_LC, KID_SEG, KID_FOLLOW, KID_LOOP, KID_PLACE, KID_ENTRIES, KID_FINAL, KID_MERGE, KID_COUNT };
static const char* kKernelNames[KID_COUNT] = {"k_signals", "k_prepare", "k_scan", "k_scatter", "k_lane_order", "k_context", "k_lane_change", "k_seg", "k_relax",
                                              "k_cyc_loops", "k_place", "k_entries", "k_final", "k_merge"};
static void prof_begin(symgpu_ctx* c, int kid);
static void prof_end(symgpu_ctx* c);
static void dbg_sync(symgpu_ctx* c, int line);
static int traj_enqueue(symgpu_ctx* c);
#define LAUNCH(c) do { (c)->launches++; if ((c)->dbg_sync) dbg_sync(c, __LINE__); } while (0)
// profile 1: every kernel group is bracketed by events (per-kernel breakdown); 2: only the relaxation sweeps (the dominant kernel,
// timed inside the measured region without the other groups' event records)
#define KBEGIN(c, kid) do { if ((c)->profile == 1 || ((c)->profile == 2 && (kid) == KID_FOLLOW)) { prof_begin(c, kid); (c)->prof_open = true; } } while (0)
#define KEND(c) do { LAUNCH(c); if ((c)->prof_open) { prof_end(c); (c)->prof_open = false; } } while (0)
#define KEND0(c) do { if ((c)->prof_open) { prof_end(c); (c)->prof_open = false; } } while (0)

static cudaEvent_t prof_event(symgpu_ctx* c) {
  if (c->ev_used == c->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); }
  return c->ev_pool[c->ev_used++];
}
static void prof_begin(symgpu_ctx* c, int kid) { ProfRec r{kid, prof_event(c), prof_event(c)}; cudaEventRecord(r.a, c->stream); c->prof_recs.push_back(r); }
static void prof_end(symgpu_ctx* c) { cudaEventRecord(c->prof_recs.back().b, c->stream); }
static void prof_collect(symgpu_ctx* c) {
  for (auto& r : c->prof_recs) { float ms = 0; if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->k_ms[r.kid] += ms; c->k_launches[r.kid]++; } }
  c->prof_recs.clear(); c->ev_used = 0;
}

// first half of a step: origins (host), signals, lane membership and in-lane order.  In a partitioned run the tails of
// the boundary lanes are exchanged between the two halves.
// SYMGPU_DEBUG_SYNC=1: wait for the device after every launch and name the line of the first one that failed (bring-up aid)
static void dbg_sync(symgpu_ctx* c, int line) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess && !c->dbg_failed) { c->dbg_failed = true; fprintf(stderr, "symgpu: launch before engine.cu:%d failed: %s\n", line, cudaGetErrorString(e)); }
}
static int step_front(symgpu_ctx* c) {
  const int L = c->net.n_lanes();
  cudaStream_t s = c->stream;
  const bool has_entries = !c->entries.empty();
  auto& desc = c->ss.desc; auto& desc_lane = c->ss.desc_lane; desc.clear(); desc_lane.clear();
  c->t += c->P.dt;                                                                                  // reseau.cpp:2105
  // origins: criterion update (reseau.cpp:2219-2223) and creation (:2255); the device draws of the previous step
  // come first in the global random sequence
  std::vector<HostRow> new_rows; std::vector<EntryLane*> new_row_lane;
  if ((has_entries || !c->api_queue.empty()) && !c->mp.on && !c->host_rng_fresh) {      // the origins draw on the host: bring the stream back (2.5 KB, the device is idle)
    CK(cudaMemcpy(&c->rng, c->lc_rng.p, sizeof(DevRng), cudaMemcpyDeviceToHost)); c->host_rng_fresh = true; }
  const unsigned host_count0 = c->rng.count;
  if (has_entries) {
    const double mdm = max_debit_max(c);
    for (auto& e : c->entries) update_crit(c, e, c->t, false, mdm);
    create_api_vehicles(c, new_rows, new_row_lane);
    create_vehicles(c, new_rows, new_row_lane);
    if (c->n_slots + (int)new_rows.size() + c->n_pool > c->pool_base) { g_err = "vehicle capacity exhausted (set SYMGPU_CAPACITY)"; return SYMGPU_E_CAPACITY; }
    for (size_t i = 0; i < new_rows.size(); i++) new_row_lane[i]->pret_slot = c->n_slots + (int)i;
    int rc = upload_rows(c, c->n_slots, new_rows); if (rc != SYMGPU_OK) return rc;
    c->n_slots += (int)new_rows.size();
    // queue heads move to their pool slot (they keep their state between attempts)
    std::vector<HostRow> one(1);
    for (auto& e : c->entries) for (auto& el : e.lanes) {
      if (el.pool_slot < 0 && !el.queue.empty()) {
        HostVeh hv = el.queue.front(); el.queue.pop_front();
        one[0] = HostRow{ST_POOL, hv.id, hv.cls, hv.route, 0, hv.lane, SYM_UNDEF, c->P.cls[hv.cls].vx, 0.0, hv.tcre, hv.hent};
        rc = upload_rows(c, el.pool_base, one); if (rc != SYMGPU_OK) return rc;
        el.pool_slot = el.pool_base;
      }
      if (el.pret_slot >= 0 || el.pool_slot >= 0) {
        EntryDesc d{el.lane, el.pret_slot, el.pool_slot, -1};
        if (el.pool_slot >= 0) d.dst_slot = c->n_slots++;
        desc.push_back(d); desc_lane.push_back(&el);
      }
    }
    if (!desc.empty()) CK(cudaMemcpyAsync(c->edesc.p, desc.data(), desc.size() * sizeof(EntryDesc), cudaMemcpyHostToDevice, s));
  }
  if (!c->mp.on && c->rng.count != host_count0) CK(cudaMemcpyAsync(c->lc_rng.p, &c->rng, sizeof(DevRng), cudaMemcpyHostToDevice, s));   // ... and forth when it drew
  const int n = c->n_slots;
  const int n_sorted = c->n_alive + (int)new_rows.size();   // vehicles that are members of a lane this step
  const int B = 256;
  auto G = [&](int cnt) { return (cnt + B - 1) / B; };
  k_reset<<<G(std::max(L, 32)), B, 0, s>>>(L, c->lane_count.p, c->lane_fill.p, c->lane_tail.p, c->stat.p, (int*)c->f_ctl.p, (int)(sizeof(FlowCtl) / sizeof(int))); LAUNCH(c);
  if (n > 0) { KBEGIN(c, KID_PREPARE); k_prepare<<<G(n), B, 0, s>>>(c->dv, n, c->lane_count.p); KEND(c); }
  // side stream, beside this step's lane ordering: the signal states of the step (read from k_context on) and the context memos
  // of the vehicles the last step listed; step_back waits for both (the lane-change phase too: it moves vehicles to other lanes)
  const bool side_sig = c->net.n_sl() > 0 && c->profile != 1 && (c->side_mask & 1);                 // per-kernel timing (profile) keeps everything on one stream
  if (side_sig || c->cc_fill) {
    if (c->mp.on || c->ext_stream) { CK(cudaEventRecord(c->cc_go, s)); CK(cudaStreamWaitEvent(c->aux, c->cc_go, 0)); }   // else: the host waited for the last step, nothing is in flight
    if (side_sig) { k_signals<<<G(c->net.n_sl()), B, 0, c->aux>>>(c->ds, c->sl_i.p, c->net.n_sl(), c->t, c->sl_col.p, c->sl_prop.p); LAUNCH(c); }
    if (c->cc_fill) { k_ctx_fill<<<c->n_sms * 16, 64, 0, c->aux>>>(c->dn, c->dv, c->cc_list.p, c->cc_cnt.p); LAUNCH(c); c->cc_fill = false; }
    CK(cudaEventRecord(c->cc_done, c->aux)); c->aux_wait = true;
    if (c->chgtvoie) { CK(cudaStreamWaitEvent(s, c->cc_done, 0)); c->aux_wait = false; }
  }
  if (c->net.n_sl() && !side_sig) { KBEGIN(c, KID_SIGNALS); k_signals<<<G(c->net.n_sl()), B, 0, s>>>(c->ds, c->sl_i.p, c->net.n_sl(), c->t, c->sl_col.p, c->sl_prop.p); KEND(c); }
  if (n > 0) {
    int nb = (L + kScanB - 1) / kScanB;
    KBEGIN(c, KID_SCAN); k_scan_block<<<nb, kScanB, 0, s>>>(c->lane_count.p, c->lane_start.p, c->bsum.p, L); LAUNCH(c);
    if (nb > 1) { k_scan_sums<<<1, kScanB, 0, s>>>(c->bsum.p, nb); LAUNCH(c); k_scan_add<<<nb, kScanB, 0, s>>>(c->lane_start.p, c->bsum.p, L); }
    KEND(c);
    KBEGIN(c, KID_SCATTER); k_scatter<<<G(n), B, 0, s>>>(c->dv, n, c->lane_start.p, c->lane_fill.p, c->u_order.p, c->u_key.p, c->u_id.p, c->u_lane.p); KEND(c);
    KBEGIN(c, KID_SORT);
      if (n_sorted > 0) { k_lane_rank<<<G(n_sorted), B, 0, s>>>(n_sorted, c->lane_start.p, c->lane_count.p, c->u_order.p, c->u_key.p, c->u_id.p, c->u_lane.p, c->order.p, c->s_key.p, c->s_id.p, c->s_lane.p, c->f_posidx.p); LAUNCH(c);
        k_lane_links<<<G(n_sorted), B, 0, s>>>(c->dv, n_sorted, c->lane_start.p, c->lane_count.p, c->order.p, c->s_key.p, c->s_id.p, c->s_lane.p, c->lane_tail.p, c->counters.p); } KEND(c);
  }
  // lane changing (reseau.cpp:14181): sequential acceptance tests on the device random stream, then the lane order is rebuilt
  if (c->chgtvoie && n > 0) {
    if (c->sens_wait) { CK(cudaStreamWaitEvent(s, c->ev_sens, 0)); c->sens_wait = false; }   // lane changing rewrites what the sensor kernel of the last step reads
    if (c->pack_wait) { CK(cudaStreamWaitEvent(s, c->t_packed, 0)); c->pack_wait = false; }
    const int K = c->net.n_links();
    KBEGIN(c, KID_LC);
    k_lc_prepare<<<G(n), B, 0, s>>>(c->dn, c->dv, c->dlc, n); LAUNCH(c);
    CK(cudaMemsetAsync(c->lc_link_head.p, 0xff, K * sizeof(int), s)); CK(cudaMemsetAsync(c->lc_nchg.p, 0, 2 * sizeof(int), s));
    if (c->lc_seq) { k_lc_sequential<<<1, 32, 0, s>>>(c->dn, c->dv, c->dlc, n); LAUNCH(c); }
    else { k_lc_eval<<<G(n), B, 0, s>>>(c->dn, c->dv, c->dlc, n); LAUNCH(c); k_lc_walk<<<1, kLcBlock, 0, s>>>(c->dn, c->dv, c->dlc, n); LAUNCH(c);
      k_rng_skip<<<1, 256, 0, s>>>(c->lc_rng.p, &c->lc_nchg.p[1], 0); LAUNCH(c); }   // the walk drew through a look-ahead buffer: advance the stream by what it consumed
    KEND(c);
    int nchg2[2] = {0, 0};
    CK(cudaMemcpyAsync(nchg2, c->lc_nchg.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int nchg = nchg2[0];
    c->lane_changes += nchg;
    if (nchg > 0) {
      CK(cudaMemsetAsync(c->lane_count.p, 0, L * sizeof(int), s)); CK(cudaMemsetAsync(c->lane_fill.p, 0, L * sizeof(int), s));
      k_prepare<<<G(n), B, 0, s>>>(c->dv, n, c->lane_count.p); LAUNCH(c);
      int nb = (L + kScanB - 1) / kScanB;
      k_scan_block<<<nb, kScanB, 0, s>>>(c->lane_count.p, c->lane_start.p, c->bsum.p, L); LAUNCH(c);
      if (nb > 1) { k_scan_sums<<<1, kScanB, 0, s>>>(c->bsum.p, nb); LAUNCH(c); k_scan_add<<<nb, kScanB, 0, s>>>(c->lane_start.p, c->bsum.p, L); LAUNCH(c); }
      k_scatter<<<G(n), B, 0, s>>>(c->dv, n, c->lane_start.p, c->lane_fill.p, c->u_order.p, c->u_key.p, c->u_id.p, c->u_lane.p); LAUNCH(c);
      CK(cudaMemsetAsync(c->lane_tail.p, 0xff, L * sizeof(int), s));
      if (n_sorted > 0) { k_lane_rank<<<G(n_sorted), B, 0, s>>>(n_sorted, c->lane_start.p, c->lane_count.p, c->u_order.p, c->u_key.p, c->u_id.p, c->u_lane.p, c->order.p, c->s_key.p, c->s_id.p, c->s_lane.p, c->f_posidx.p); LAUNCH(c);
        k_lane_links<<<G(n_sorted), B, 0, s>>>(c->dv, n_sorted, c->lane_start.p, c->lane_count.p, c->order.p, c->s_key.p, c->s_id.p, c->s_lane.p, c->lane_tail.p, c->counters.p); } LAUNCH(c);
    }
  }
  c->ss.n = n; c->ss.n_sorted = n_sorted; c->ss.has_entries = has_entries;
  return SYMGPU_OK;
}

// second half: context, dependency-ordered car following, placement, origins' queues, finalisation
static int step_back(symgpu_ctx* c) {
  const int L = c->net.n_lanes(), T = c->net.n_cls();
  cudaStream_t s = c->stream;
  auto& desc = c->ss.desc; auto& desc_lane = c->ss.desc_lane;
  const int n = c->ss.n, n_sorted = c->ss.n_sorted; const bool has_entries = c->ss.has_entries;
  const int B = 256;
  auto G = [&](int cnt) { return (cnt + B - 1) / B; };
  const int by_id = c->mp.on ? 1 : 0;
  if (c->aux_wait) { CK(cudaStreamWaitEvent(s, c->cc_done, 0)); c->aux_wait = false; }
  if (c->sens_wait) { CK(cudaStreamWaitEvent(s, c->ev_sens, 0)); c->sens_wait = false; }   // the merge sensors of the last step were counted beside this step's lane ordering
  if (c->mg.on) {                                                         // merge points: candidates of the step, priority levels from this step's signal colours
    KBEGIN(c, KID_MERGE); CK(cudaMemsetAsync(c->mg.cand_up.p, 0xff, L * sizeof(int), s));
    k_mrg_rank<<<G(c->mg.h.n_pt), B, 0, s>>>(c->mg.h, c->sl_col.p, c->lane_count.p); KEND(c); }
  if (n > 0) {
    if (n_sorted > 0) { KBEGIN(c, KID_CONTEXT);
      k_context<<<(n_sorted + SYM_CTX_THREADS - 1) / SYM_CTX_THREADS, SYM_CTX_THREADS, 0, s>>>(c->dn, c->dv, c->dl, n_sorted, c->order.p, c->lane_tail.p, c->rows.p, c->draws.p, (int*)&c->counters.p[SYMGPU_CNT_CTX_OVERFLOW],
                                                                                             c->mg.on ? c->mg.d.p : nullptr, c->t); KEND(c);
      if (c->mg.on && !has_entries && c->profile != 1 && (c->side_mask & 2)) {                     // this step's count-only draws are final: the stream passes them beside the relaxation
        CK(cudaEventRecord(c->ev_fork, s)); CK(cudaStreamWaitEvent(c->aux2, c->ev_fork, 0));
        k_rng_ctx<<<1, 256, 0, c->aux2>>>(c->lc_rng.p, c->draws.p, c->mg.seen.p, nullptr); LAUNCH(c);
        CK(cudaEventRecord(c->ev_rng, c->aux2)); c->rng_wait = true; } }
    // dependency-ordered follow phase (flow.cuh): segments -> dataflow rounds -> loop cutting when a round stalls
    if (n_sorted > 0) {
    KBEGIN(c, KID_SEG); k_seg<<<G(n_sorted), B, 0, s>>>(c->dv, c->df, n_sorted, c->s_lane.p, c->lane_count.p, c->lane_start.p, by_id); KEND(c);
    // leader loops: always looked for (the segment graph is small), cut where the reference's recursion would cut them
    { int Klev = 1; while ((1LL << (2 * Klev)) * kCycWalk < (long long)std::max(4, n_sorted)) Klev++;   // 4^K * walk >= longest possible chain
      Klev++;
      const int Gs = std::min(G(n_sorted), c->n_sms * 8), Gu = std::min(G(n_sorted), c->n_sms);
      KBEGIN(c, KID_LOOP);
      k_cyc_walk<<<Gs, B, 0, s>>>(c->df); LAUNCH(c);
      if (c->coop_loops) {                                                  // jump rounds + mark + cut in one cooperative launch
        DevVeh dv = c->dv; DevFlow df = c->df; double* rp = c->rows.p; long long* cp = c->counters.p; int kl = Klev, bi = by_id;
        void* args[] = {&dv, &df, &rp, &cp, &kl, &bi};
        CK(cudaLaunchCooperativeKernel((void*)k_cyc_rest, dim3(c->n_sms), dim3(256), args, 0, s)); LAUNCH(c);
      } else {
      int flip = 0;
      for (int k = 0; k < Klev; k++) { k_cyc_jump<<<Gu, B, 0, s>>>(c->df, flip); LAUNCH(c); flip ^= 1; }
      k_cyc_mark<<<Gu, B, 0, s>>>(c->dv, c->df, flip, by_id); LAUNCH(c);
      k_cyc_break<<<Gu, B, 0, s>>>(c->dv, c->df, c->rows.p, c->counters.p);
      }
      KEND(c); }
    }
  }
  // relaxation sweeps until one changes nothing (flow.cuh), then placement / origins' queues / finalisation.  The first batch of
  // sweeps is sized by what the previous step needed; the kernels after it are launched at once, guarded by the device-side
  // verdict of that batch (k_relax_done), so the common case has ONE host synchronisation per step.  If the batch did not reach
  // the fixed point the guarded kernels did nothing: more sweeps follow, then the same kernels unguarded.
  std::vector<int> eres(desc.size());
  auto sweeps = [&](int nb, bool first) {
    int* ch = c->f_ctl.p->changed;                                       // device address arithmetic only
    if (!first) cudaMemcpyAsync(ch, c->h_ctl_seed->changed, sizeof(int) * (kRelaxBatch + 1), cudaMemcpyHostToDevice, s);
    for (int j = 0; j < nb; j++) {
      k_relax<<<(n_sorted + SYM_RELAX_THREADS - 1) / SYM_RELAX_THREADS, SYM_RELAX_THREADS, 0, s>>>(c->dn, c->dv, c->dl, c->rows.p, c->b_goal.p, c->b_flags.p, c->vl_used.p, c->rl_lead.p, n_sorted, first && j == 0, ch + j, ch + j + 1, c->t, c->mg.on ? c->mg.d.p : nullptr, c->mg.mrgx.p); LAUNCH(c);
    }
    k_relax_done<<<1, 32, 0, s>>>(c->f_ctl.p, nb); LAUNCH(c);
    cudaMemcpyAsync(c->h_ctl, c->f_ctl.p, sizeof(FlowCtl), cudaMemcpyDeviceToHost, s);
  };
  auto finish = [&](const int* guard) {
    if (c->pack_wait) { cudaStreamWaitEvent(s, c->t_packed, 0); c->pack_wait = false; }      // trajectory rows of the last step are packed: status may change
    if (n > 0) {
      if (n_sorted > 0) { KBEGIN(c, KID_PLACE); k_place<<<G(n_sorted), B, 0, s>>>(c->dn, c->dv, c->rows.p, c->b_goal.p, c->b_flags.p, n_sorted, c->t, guard, c->mg.on ? c->mg.d.p : nullptr); KEND(c); }
      if (!desc.empty()) { KBEGIN(c, KID_ENTRIES); k_entries<<<G((int)desc.size()), B, 0, s>>>(c->dn, c->dv, c->dl, c->edesc.p, (int)desc.size(), c->eresult.p, c->t, c->draws.p,
                                                                      (int*)&c->counters.p[SYMGPU_CNT_CTX_OVERFLOW], guard); KEND(c); }
      // merge points (merge.cuh), after the origins' queues like reseau.cpp:14290-14314; the stream first passes this step's count-only draws
      if (c->mg.on && n_sorted > 0) { auto& g = c->mg; const int Gp = G(g.h.n_pt);
        KBEGIN(c, KID_MERGE); cudaMemsetAsync(g.flag.p, 0, sizeof(int), s);
        k_mrg_probe<<<Gp, B, 0, s>>>(c->dn, c->dv, g.h, c->order.p, c->s_key.p, c->lane_start.p, c->lane_count.p, c->f_posidx.p, c->rows.p, c->t, g.n_upd, guard); LAUNCH(c);
        k_mrg_classify<<<Gp, B, 0, s>>>(g.h, guard); LAUNCH(c);
        const bool side = c->profile != 1 && (c->side_mask & 4);                    // the points that share vehicles are settled beside the clean ones
        if (side) { cudaEventRecord(c->ev_fork, s); cudaStreamWaitEvent(c->aux2, c->ev_fork, 0); }
        k_mrg_dirty<<<1, 256, 0, side ? c->aux2 : s>>>(c->dn, c->dv, g.h, c->order.p, c->s_key.p, c->lane_start.p, c->lane_count.p, c->f_posidx.p, c->rows.p, c->t, g.n_upd, guard); LAUNCH(c);
        if (side) cudaEventRecord(c->ev_dirty, c->aux2);
        // the clean points that need no value of the stream run beside everything that follows (their draw counts are known from the probe): the
        // ordered walk and the clean points it decides touch other vehicles; its literal visits wait for this kernel on the device (eval_done)
        const bool eside = side && (c->side_mask & 16); cudaStream_t qe = eside ? c->aux : s;
        if (eside) cudaStreamWaitEvent(qe, c->ev_fork, 0);
        k_mrg_eval<<<Gp, B, 0, qe>>>(c->dn, c->dv, g.h, c->order.p, c->s_key.p, c->lane_start.p, c->lane_count.p, c->f_posidx.p, c->rows.p, c->t, g.n_upd, guard); LAUNCH(c);
        if (eside) cudaEventRecord(c->ev_eval, qe);
        if (side) cudaStreamWaitEvent(s, c->ev_dirty, 0);
        { const int np = g.h.n_pt, nbk = (np + 1023) / 1024;                              // offset of every point's draws in the stream and rank of the visited points: one scan
          k_scan64_block<<<nbk, 1024, 0, s>>>(g.kf.p, g.kfo.p, g.kbsum.p, np); LAUNCH(c);
          k_scan64_sums<<<1, 32, 0, s>>>(g.kbsum.p, nbk); LAUNCH(c);
          if (nbk > 1) { k_scan64_add<<<nbk, 1024, 0, s>>>(g.kfo.p, g.kbsum.p, np); LAUNCH(c); }
          k_mrg_list<<<Gp, B, 0, s>>>(g.h); LAUNCH(c); }
        if (c->rng_wait) { cudaStreamWaitEvent(s, c->ev_rng, 0); c->rng_wait = false; }
        else { k_rng_ctx<<<1, 256, 0, s>>>(c->lc_rng.p, c->draws.p, g.seen.p, guard); LAUNCH(c); }
        k_mrg_walk<<<1, 256, 0, s>>>(c->dn, c->dv, g.h, c->order.p, c->s_key.p, c->lane_start.p, c->lane_count.p, c->f_posidx.p, c->rows.p, c->t, g.n_upd, g.kbsum.p + (g.h.n_pt + 1023) / 1024, g.out.p, guard, Gp); LAUNCH(c);
        k_mrg_apply<<<Gp, B, 0, s>>>(c->dn, c->dv, g.h, c->order.p, c->s_key.p, c->lane_start.p, c->lane_count.p, c->f_posidx.p, c->rows.p, c->t, g.n_upd, guard); LAUNCH(c);
        if (eside) cudaStreamWaitEvent(s, c->ev_eval, 0);
        k_mrg_tally<<<Gp, B, 0, s>>>(g.h, g.n_upd, guard); KEND(c); }
      else if (!c->mp.on) { k_rng_ctx<<<1, 256, 0, s>>>(c->lc_rng.p, c->draws.p, c->mg.seen.p, guard); LAUNCH(c); }
      if (c->use_cc) cudaMemsetAsync(c->cc_cnt.p, 0, sizeof(int), s);
      KBEGIN(c, KID_FINAL); k_final<<<G(n), B, 0, s>>>(c->dn, c->dv, n, T, c->P.dt, c->entre.p, c->sorti.p, c->exited.p, &c->stat.p[4], c->exit_winner.p, &c->stat.p[5], &c->stat.p[6], guard, c->cc_list.p, c->cc_cnt.p); LAUNCH(c);
      if (c->exit_lanes.n) { k_exit_lanes<<<G((int)c->exit_lanes.n), B, 0, s>>>(c->dv, c->exit_lanes.p, c->exit_cap.p, c->exit_nl.p, (int)c->exit_lanes.n, c->exit_winner.p,
                                                                                c->next_sortie.p, c->t, c->P.dt, c->P.duree, guard); }
      KEND(c);
      if (c->mg.on && c->mg.h.n_sens > 0) { auto& g = c->mg;                  // merge sensors (reseau.cpp:14425-14458): rotate the windows, count the crossings of the step
        const bool side = c->profile != 1 && (c->side_mask & 8); cudaStream_t q = side ? c->aux2 : s;       // beside the next step's lane ordering (step_back waits for ev_sens)
        KBEGIN(c, KID_MERGE); if (side) { cudaEventRecord(c->ev_fork, s); cudaStreamWaitEvent(q, c->ev_fork, 0); }
        k_mrg_rotate<<<G(g.n_rows), B, 0, q>>>(g.h, g.n_upd + 1, g.n_rows, g.row_sens.p); LAUNCH(c);
        k_mrg_sensors<<<G(n), B, 0, q>>>(c->dn, c->dv, g.h, n, g.n_upd + 1); LAUNCH(c);
        k_mrg_sensors_sum<<<G(g.n_rows), B, 0, q>>>(g.h, g.n_upd + 1, g.n_rows, g.row_sens.p); LAUNCH(c);
        if (side) { cudaEventRecord(c->ev_sens, q); c->sens_wait = true; } KEND0(c); }
    }
    cudaMemcpyAsync(c->h_stat + 4, &c->stat.p[4], 3 * sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(c->h_stat + 7, c->draws.p, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    if (!desc.empty()) cudaMemcpyAsync(eres.data(), c->eresult.p, desc.size() * sizeof(int), cudaMemcpyDeviceToHost, s);
  };
  auto tally = [&](int nb, bool first_batch) -> bool {                   // counts the sweeps that changed something; true when one changed nothing
    bool done = false; int used = 0;
    for (int j = 1; j <= nb && !done; j++) { used = j; if (c->h_ctl->changed[j] == 0) done = true; else c->passes++; }
    if (done && first_batch) c->relax_batch = std::min(kRelaxBatch, std::max(3, used + 2)); else if (!done) c->relax_batch = kRelaxBatch;   // one sweep of margin: a batch that falls short costs a host round trip
    return done;
  };
  const bool relax = n > 0 && n_sorted > 0;
  if (relax) {
    const int nb = c->relax_batch;
    KBEGIN(c, KID_FOLLOW); sweeps(nb, true); KEND0(c);
    finish(&c->f_ctl.p->converged);
    CK(cudaStreamSynchronize(s));
    if (!tally(nb, true)) {
      for (int sweep = nb;;) {
        KBEGIN(c, KID_FOLLOW); sweeps(kRelaxBatch, false); KEND0(c);
        CK(cudaStreamSynchronize(s));
        sweep += kRelaxBatch;
        if (tally(kRelaxBatch, false)) break;
        if (sweep > n_sorted + kRelaxBatch) { g_err = "relaxation does not terminate (leader loop left uncut)"; return SYMGPU_E_CUDA; }
      }
      finish(nullptr);
      CK(cudaStreamSynchronize(s));
    }
  } else { finish(nullptr); CK(cudaStreamSynchronize(s)); }
  CK(cudaGetLastError());
  if (c->profile) prof_collect(c);
  if (c->h_stat[6]) { g_err = "internal error: vehicles left uncomputed"; return SYMGPU_E_CUDA; }
  c->n_alive = n > 0 ? c->h_stat[5] : 0; c->cc_fill = c->use_cc && n > 0;
  c->veh_steps += c->n_alive;
  int nex = n > 0 ? c->h_stat[4] : 0;
  c->last_exited.resize(nex);
  if (nex) { CK(cudaMemcpy(c->last_exited.data(), c->exited.p, nex * sizeof(ExitRow), cudaMemcpyDeviceToHost));
    std::sort(c->last_exited.begin(), c->last_exited.end(), [&](const ExitRow& a, const ExitRow& b) { return c->mp.on ? a.id < b.id : a.slot < b.slot; }); }
  // global random sequence: account for the draws made on the device during this step
  unsigned dd = (unsigned)c->h_stat[7];
  if (c->mp.on) {                                                              // partitioned runs: the host owns the stream (no device draws with values there)
    if (has_entries) c->mp.draw_delta = dd - c->dev_draws_seen;               // committed with the other ranks' draws (symgpu_mp_commit_draws)
    else c->rng.count += dd - c->dev_draws_seen; }
  else c->host_rng_fresh = false;                                              // the stream advanced on the device (next-lane memos, lane changes, merge points)
  c->dev_draws_seen = dd; c->mg.n_upd++;
  for (size_t i = 0; i < desc.size(); i++) { EntryLane* el = desc_lane[i];
    if (eres[i] & 1) el->pret_slot = -1;
    if (eres[i] & 2) el->pool_slot = -1; }
  // swap start/end-of-step buffers
  std::swap(c->v_pos.p, c->v_pos_n.p); std::swap(c->v_vit.p, c->v_vit_n.p); std::swap(c->v_acc.p, c->v_acc_n.p); std::swap(c->v_lane.p, c->v_lane_n.p);
  fill_dev_structs(c);
  return c->n_alive;
}
static int do_step(symgpu_ctx* c) {
  if (c->mp.on) { g_err = "partitioned context: drive it with symgpu_mp_step_front / _back / _import"; return SYMGPU_E_ARG; }
  int rc = step_front(c);
  if (rc != SYMGPU_OK) return rc;
  return step_back(c);
}

// ---- partitioned runs: one process per GPU, the caller moves the buffers between ranks (NCCL via torch.distributed) ----------
// node_part[n_nodes]: owning rank of every node.  A lane belongs to the rank of the node it leads to (internal junction lanes:
// their junction; exit stubs: the junction they leave), so a follower and the junction it crosses are on one GPU.
extern "C" int symgpu_mp_configure(symgpu_ctx* c, int rank, int world, const int* node_part) {
  if (!c || !node_part || world < 1 || rank < 0 || rank >= world) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  if (c->chgtvoie) { g_err = "partitioned runs do not support lane changing yet (global random stream)"; return SYMGPU_E_UNSUPPORTED; }
  if (c->mg.on) { g_err = "partitioned runs do not support merge points yet (their draws are ordered over the whole network)"; return SYMGPU_E_UNSUPPORTED; }
  const FlatNet& f = c->net; const int L = f.n_lanes();
  auto& m = c->mp; m.on = true; m.rank = rank; m.world = world;
  m.lane_owner.resize(L);
  for (int l = 0; l < L; l++) { int link = f.lane_i[4 * l]; int brick = f.link_i[8 * link + 5]; int up = f.link_i[8 * link + 2], down = f.link_i[8 * link + 3];
    int ta = f.link_i[8 * link + 4]; int node = brick >= 0 ? brick : ((ta == 'S' || ta == 'P') ? up : down); m.lane_owner[l] = node_part[node]; }
  // halo of rank r: lanes owned by someone else that an owned lane leads into; export of r to q: lanes of r in q's halo
  std::vector<std::vector<int>> halo(world), expo(world);
  std::vector<char> seen(L, 0);
  for (int l = 0; l < L; l++) for (int k = f.succ_off[l]; k < f.succ_off[l + 1]; k++) { int nx = f.succ_i[3 * k + 1]; int a = m.lane_owner[l], b = m.lane_owner[nx];
    if (a == b) continue;
    if (a == rank && !(seen[nx] & 1)) { seen[nx] |= 1; halo[b].push_back(nx); }
    if (b == rank && !(seen[nx] & 2)) { /* per peer dedupe below */ } }
  for (int q = 0; q < world; q++) { if (q == rank) continue; std::vector<char> s2(L, 0);
    for (int l = 0; l < L; l++) if (m.lane_owner[l] == q) for (int k = f.succ_off[l]; k < f.succ_off[l + 1]; k++) { int nx = f.succ_i[3 * k + 1];
      if (m.lane_owner[nx] == rank && !s2[nx]) { s2[nx] = 1; expo[q].push_back(nx); } } }
  m.halo_lanes.clear(); m.export_lanes.clear(); m.halo_off.assign(world + 1, 0); m.export_off.assign(world + 1, 0);
  std::vector<int> export_peer;
  for (int q = 0; q < world; q++) { std::sort(halo[q].begin(), halo[q].end()); std::sort(expo[q].begin(), expo[q].end());
    m.halo_lanes.insert(m.halo_lanes.end(), halo[q].begin(), halo[q].end()); m.halo_off[q + 1] = (int)m.halo_lanes.size();
    m.export_lanes.insert(m.export_lanes.end(), expo[q].begin(), expo[q].end()); m.export_off[q + 1] = (int)m.export_lanes.size(); }
  m.n_halo = (int)m.halo_lanes.size(); m.n_export = (int)m.export_lanes.size();
  if ((size_t)m.n_halo > std::min<size_t>((size_t)L, 65536)) { g_err = "halo larger than the ghost reserve"; return SYMGPU_E_CAPACITY; }
  CK(m.d_lane_owner.upload(m.lane_owner)); CK(m.d_halo_lanes.upload(m.halo_lanes)); CK(m.d_export_lanes.upload(m.export_lanes)); CK(m.d_mig_count.alloc(world + 1)); CK(m.d_free.alloc(c->cap));
  m.h_mig_count.assign(world, 0);
  // keep only the vehicles on owned lanes
  CK(cudaMemset(c->stat.p, 0, sizeof(int)));
  if (c->n_slots) { k_mp_drop_foreign<<<(c->n_slots + 255) / 256, 256, 0, c->stream>>>(c->dv, c->n_slots, m.d_lane_owner.p, rank, c->stat.p); LAUNCH(c); }
  CK(cudaMemcpyAsync(c->h_stat, c->stat.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
  c->n_alive = c->h_stat[0];
  return c->n_alive;
}
// entries (tail records of 4 doubles) this rank sends to / receives from every peer each step
// tail exchange over peer memory, set-up: (1) every rank allocates its halo area + flags and hands out a CUDA IPC handle (64 bytes);
// (2) the caller gathers the handles and, per peer q, where this rank's records start inside q's area (q's halo offset of this rank)
// and q's halo size; the engine opens the peers' areas and precomputes the destination of every exported record.
extern "C" int symgpu_mp_p2p_handle(symgpu_ctx* c, void* handle64) {
  if (!c || !c->mp.on || !handle64) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  auto& m = c->mp; auto& pp = m.p2p;
  CK(pp.area.alloc(4 * (size_t)m.n_halo + m.world + 1));
  cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, pp.area.p));
  static_assert(sizeof(h) == 64, "CUDA IPC handle is 64 bytes");
  memcpy(handle64, &h, 64);
  return m.n_halo;
}
extern "C" int symgpu_mp_p2p_connect(symgpu_ctx* c, const void* handles, const int* peer_halo_off, const int* peer_n_halo) {
  if (!c || !c->mp.on || !handles || !peer_halo_off || !peer_n_halo || !c->mp.p2p.area.p) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  auto& m = c->mp; auto& pp = m.p2p;
  pp.peer.assign(m.world, nullptr);
  std::vector<double*> dst(std::max(1, m.n_export), nullptr); std::vector<unsigned long long*> fl(m.world, nullptr);
  for (int q = 0; q < m.world; q++) {
    const int ne = m.export_off[q + 1] - m.export_off[q];
    if (q == m.rank || ne == 0) continue;
    cudaIpcMemHandle_t h; memcpy(&h, (const char*)handles + 64 * (size_t)q, 64);
    CK(cudaIpcOpenMemHandle(&pp.peer[q], h, cudaIpcMemLazyEnablePeerAccess));
    double* base = (double*)pp.peer[q];
    for (int j = 0; j < ne; j++) dst[m.export_off[q] + j] = base + 4 * ((size_t)peer_halo_off[q] + j);
    fl[q] = (unsigned long long*)pp.peer[q] + 4 * (size_t)peer_n_halo[q] + m.rank;
  }
  CK(pp.exp_dst.upload(dst)); CK(pp.flag_dst.upload(fl)); CK(pp.done.alloc(1)); CK(pp.err.alloc(1)); CK(pp.d_halo_off.upload(m.halo_off));
  pp.on = true;
  return SYMGPU_OK;
}
extern "C" int symgpu_mp_sizes(symgpu_ctx* c, int* export_cnt, int* halo_cnt) {
  if (!c || !c->mp.on) return SYMGPU_E_ARG;
  for (int q = 0; q < c->mp.world; q++) { if (export_cnt) export_cnt[q] = c->mp.export_off[q + 1] - c->mp.export_off[q]; if (halo_cnt) halo_cnt[q] = c->mp.halo_off[q + 1] - c->mp.halo_off[q]; }
  return SYMGPU_OK;
}
// first half of the step; writes the tails of the exported lanes (n_export x 4 doubles, peers in rank order) to device memory `send_tails`
extern "C" int symgpu_mp_step_front(symgpu_ctx* c, double* send_tails) {
  if (!c || !c->mp.on) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev0, c->stream));
  int rc = step_front(c); if (rc != SYMGPU_OK) return rc;
  if (c->mp.n_export) {
    if (c->ss.n == 0) CK(cudaMemsetAsync(c->lane_tail.p, 0xff, c->net.n_lanes() * sizeof(int), c->stream));
    if (c->mp.p2p.on) { auto& pp = c->mp.p2p; pp.seq++;
      k_mp_pack_tails_p2p<<<(c->mp.n_export + 255) / 256, 256, 0, c->stream>>>(c->dv, c->mp.d_export_lanes.p, c->mp.n_export, c->lane_tail.p, pp.exp_dst.p, pp.done.p, pp.flag_dst.p, c->mp.world, pp.seq); LAUNCH(c); }
    else { k_mp_pack_tails<<<(c->mp.n_export + 255) / 256, 256, 0, c->stream>>>(c->dv, c->mp.d_export_lanes.p, c->mp.n_export, c->lane_tail.p, send_tails); LAUNCH(c); } }
  else if (c->mp.p2p.on) c->mp.p2p.seq++;
  if (!c->ext_stream) CK(cudaStreamSynchronize(c->stream));   // on the caller's stream the collective that follows is ordered after the pack
  return SYMGPU_OK;
}
// second half: `recv_tails` (n_halo x 4 doubles, peers in rank order) become ghost leaders; then the rest of the step; vehicles that
// ended on a foreign lane are packed into `send_rows` (world regions of cap_per_peer rows of 16 doubles), counts in send_counts[world]
extern "C" int symgpu_mp_step_back(symgpu_ctx* c, const double* recv_tails, double* send_rows, int cap_per_peer, int* send_counts) {
  if (!c || !c->mp.on || !send_counts) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  auto& m = c->mp;
  if (c->ss.n == 0) CK(cudaMemsetAsync(c->lane_tail.p, 0xff, c->net.n_lanes() * sizeof(int), c->stream));
  if (m.n_halo && m.p2p.on) {
    k_mp_wait_tails<<<1, 32, 0, c->stream>>>(m.p2p.area.p + 4 * (size_t)m.n_halo, m.p2p.d_halo_off.p, m.world, m.p2p.seq, m.p2p.err.p); LAUNCH(c);
    recv_tails = (const double*)m.p2p.area.p; }
  if (m.n_halo) { k_mp_unpack_ghosts<<<(m.n_halo + 255) / 256, 256, 0, c->stream>>>(c->dv, m.d_halo_lanes.p, m.n_halo, m.ghost_base, c->lane_tail.p, recv_tails); LAUNCH(c); }
  int r = step_back(c); if (r < 0) return r;
  if (m.p2p.on) { int e = 0; CK(cudaMemcpy(&e, m.p2p.err.p, sizeof(int), cudaMemcpyDeviceToHost)); if (e) { g_err = "tail exchange over peer memory timed out"; return SYMGPU_E_CUDA; } }
  CK(cudaMemsetAsync(m.d_mig_count.p, 0, (m.world + 1) * sizeof(int), c->stream));
  if (c->n_slots) { k_mp_pack_migrants<<<(c->n_slots + 255) / 256, 256, 0, c->stream>>>(c->dv, c->n_slots, m.d_lane_owner.p, m.rank, cap_per_peer, m.d_mig_count.p, send_rows,
                                                                                  m.d_free.p, m.free_count, &m.d_mig_count.p[m.world]); LAUNCH(c); }
  k_mp_row_counts<<<1, 64, 0, c->stream>>>(m.d_mig_count.p, send_rows, cap_per_peer, m.world); LAUNCH(c);
  CK(cudaMemcpyAsync(m.h_mig_count.data(), m.d_mig_count.p, m.world * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int out = 0;
  for (int q = 0; q < m.world; q++) { send_counts[q] = m.h_mig_count[q]; out += m.h_mig_count[q]; if (m.h_mig_count[q] > cap_per_peer) { g_err = "migration buffer too small"; return SYMGPU_E_CAPACITY; } }
  c->n_alive -= out; m.free_count += out;
  return c->n_alive;
}
// origins in a partitioned run: the draws the device made this step (next-lane memo) are part of the one global random sequence;
// the caller sums symgpu_mp_local_draws over the ranks and hands the total to every rank before the next step
// run on the caller's stream (e.g. torch's current stream, on which its NCCL collectives are ordered): the exchanges of a partitioned
// step then need no host synchronisation between the engine's kernels and the collectives
extern "C" int symgpu_set_stream(symgpu_ctx* c, void* stream) {
  if (!c) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->stream && !c->ext_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)stream; c->ext_stream = true;
  return SYMGPU_OK;
}
extern "C" int symgpu_mp_has_origins(symgpu_ctx* c) { return c ? (c->entries.empty() ? 0 : 1) : SYMGPU_E_ARG; }
extern "C" long symgpu_mp_local_draws(symgpu_ctx* c) { return c && c->mp.on ? (long)c->mp.draw_delta : (long)SYMGPU_E_ARG; }
extern "C" int symgpu_mp_commit_draws(symgpu_ctx* c, long total) {
  if (!c || !c->mp.on || total < 0) return SYMGPU_E_ARG;
  c->rng.skip((unsigned)total); c->mp.draw_delta = 0;
  return SYMGPU_OK;
}
// vehicles arriving from other ranks (n_rows x 16 doubles in device memory) join the end of the local list
// arrivals straight from the per-peer regions of the fixed-size all_to_all (the row count of a region rides in the spare double of
// its first row, k_mp_row_counts): one strided copy brings the counts to the host, one unpack launch per non-empty region
extern "C" int symgpu_mp_import_regions(symgpu_ctx* c, const double* recv_regions, int cap_per_peer) {
  if (!c || !c->mp.on || !recv_regions || cap_per_peer <= 0) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  const int W = c->mp.world;
  std::vector<double> cnt(W, 0.0);
  CK(cudaMemcpy2DAsync(cnt.data(), sizeof(double), recv_regions + 15, (size_t)cap_per_peer * kMigD * sizeof(double), sizeof(double), W, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int q = 0; q < W; q++) {
    if (q == c->mp.rank) continue;
    const int n_rows = (int)cnt[q];
    if (n_rows <= 0) continue;
    if (n_rows > cap_per_peer) { g_err = "migration region overflow"; return SYMGPU_E_CAPACITY; }
    const int reuse = std::min(n_rows, c->mp.free_count);
    if (c->n_slots + (n_rows - reuse) > c->pool_base) { g_err = "vehicle capacity exhausted (set SYMGPU_CAPACITY)"; return SYMGPU_E_CAPACITY; }
    k_mp_unpack_migrants<<<(n_rows + 255) / 256, 256, 0, c->stream>>>(c->dv, c->n_slots, n_rows, recv_regions + (size_t)q * cap_per_peer * kMigD, c->mp.d_free.p, c->mp.free_count, reuse); LAUNCH(c);
    c->mp.free_count -= reuse; c->n_slots += n_rows - reuse; c->n_alive += n_rows;
  }
  CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->last_ms = ms;
  return c->n_alive;
}
extern "C" int symgpu_mp_import(symgpu_ctx* c, const double* recv_rows, int n_rows) {
  if (!c || !c->mp.on || n_rows < 0) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  const int reuse = std::min(n_rows, c->mp.free_count);
  if (c->n_slots + (n_rows - reuse) > c->pool_base) { g_err = "vehicle capacity exhausted (set SYMGPU_CAPACITY)"; return SYMGPU_E_CAPACITY; }
  if (n_rows) { k_mp_unpack_migrants<<<(n_rows + 255) / 256, 256, 0, c->stream>>>(c->dv, c->n_slots, n_rows, recv_rows, c->mp.d_free.p, c->mp.free_count, reuse); LAUNCH(c); }
  c->mp.free_count -= reuse; c->n_slots += n_rows - reuse; c->n_alive += n_rows;
  CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->last_ms = ms;
  return c->n_alive;
}

extern "C" int symgpu_step(symgpu_ctx* c) {
  if (!c) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev0, c->stream));
  int r = do_step(c);
  if (r < 0) return r;
  CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->last_ms = ms;
  return r;
}
extern "C" int symgpu_run(symgpu_ctx* c, int n_steps) {
  if (!c) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev0, c->stream));
  int r = c->n_alive;
  for (int k = 0; k < n_steps; k++) { r = do_step(c); if (r < 0) return r; }
  CK(cudaEventRecord(c->ev1, c->stream)); CK(cudaEventSynchronize(c->ev1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->last_ms = ms;
  return r;
}

// list-order gather on the host (m_LstVehicles order = slot order)
static int gather(symgpu_ctx* c, std::vector<int>& slots) {
  std::vector<int> st(c->n_slots);
  if (c->n_slots) CK(cudaMemcpy(st.data(), c->v_status.p, c->n_slots * sizeof(int), cudaMemcpyDeviceToHost));
  slots.clear();
  for (int i = 0; i < c->n_slots; i++) if (st[i] == ST_ALIVE) slots.push_back(i);
  return SYMGPU_OK;
}
template <class T> static int fetch(symgpu_ctx* c, const T* dev, const std::vector<int>& slots, T* out, int cap, int width = 1) {
  if (!out) return SYMGPU_OK;
  std::vector<T> h((size_t)c->n_slots * width);
  if (c->n_slots) CK(cudaMemcpy(h.data(), dev, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < slots.size() && (int)k < cap; k++) out[k] = h[(size_t)slots[k] * width];
  return SYMGPU_OK;
}
extern "C" int symgpu_get_state(symgpu_ctx* c, int cap, int* id, int* lane, int* route_pos, int* flags, double* pos, double* vit, double* acc,
                                double* deltaN, int* leader_id) {
  if (!c) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  std::vector<int> slots; int rc = gather(c, slots); if (rc) return rc;
  if ((rc = fetch(c, c->v_id.p, slots, id, cap))) return rc;
  if ((rc = fetch(c, c->v_lane.p, slots, lane, cap))) return rc;
  if ((rc = fetch(c, c->v_rpos.p, slots, route_pos, cap))) return rc;
  if ((rc = fetch(c, c->v_pos.p, slots, pos, cap))) return rc;
  if ((rc = fetch(c, c->v_vit.p, slots, vit, cap))) return rc;
  if ((rc = fetch(c, c->v_acc.p, slots, acc, cap))) return rc;
  if ((rc = fetch(c, c->v_dn.p, slots, deltaN, cap))) return rc;
  if ((rc = fetch(c, c->v_pl.p, slots, leader_id, cap))) return rc;
  if (flags) {
    std::vector<int> of(slots.size()), cl(slots.size());
    if ((rc = fetch(c, c->v_oflags.p, slots, of.data(), (int)of.size()))) return rc;
    if ((rc = fetch(c, c->v_cls.p, slots, cl.data(), (int)cl.size()))) return rc;
    for (size_t k = 0; k < slots.size() && (int)k < cap; k++) flags[k] = (of[k] & 7) | (cl[k] << 8);
  }
  return (int)slots.size();
}
// Caller buffers for trajectory rows: page-locked in place so the device copies straight into them.
extern "C" int symgpu_traj_bind(symgpu_ctx* c, int which, int cap, int* id, int* lane, double* pos, double* vit, double* acc) {
  if (!c || which < 0 || which >= kTrajBufs || cap <= 0 || !id || !lane || !pos || !vit || !acc) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  auto& b = c->tbuf[which];
  if (b.registered) { cudaHostUnregister(b.id); cudaHostUnregister(b.lane); cudaHostUnregister(b.pos); cudaHostUnregister(b.vit); cudaHostUnregister(b.acc); b.registered = false; }
  b.cap = cap; b.id = id; b.lane = lane; b.pos = pos; b.vit = vit; b.acc = acc;
  bool ok = cudaHostRegister(id, (size_t)cap * 4, cudaHostRegisterDefault) == cudaSuccess && cudaHostRegister(lane, (size_t)cap * 4, cudaHostRegisterDefault) == cudaSuccess &&
            cudaHostRegister(pos, (size_t)cap * 8, cudaHostRegisterDefault) == cudaSuccess && cudaHostRegister(vit, (size_t)cap * 8, cudaHostRegisterDefault) == cudaSuccess &&
            cudaHostRegister(acc, (size_t)cap * 8, cudaHostRegisterDefault) == cudaSuccess;
  cudaGetLastError();
  b.registered = ok;          // not registered (already pinned by the caller, or refused): the copies still work, staged by the driver
  return SYMGPU_OK;
}
// One step, then pack + copy the trajectory rows to bound buffer `which` on the side stream (returns rows; the rows are
// valid after symgpu_traj_wait(which)).  The next step can run while the copy is in flight.
// pack + copy the CURRENT trajectory rows to bound buffer `which` on the side stream (no step): partitioned runs call it
// after symgpu_mp_import
// pack + copies of the rows requested by symgpu_traj_copy_async, on the side stream beside the next step.  (Enqueuing them later —
// from the next step's step_front, or behind its k_context launch, with two or three caller buffers — was measured: 0.95 / 0.97 /
// 1.02 ms per step end to end against 0.95 here; what the copy costs the stepping loop is device-side, not the ~30 us of API calls.)
static int traj_enqueue(symgpu_ctx* c) {
  const int which = c->traj_req; if (which < 0) return SYMGPU_OK;
  struct Timer { symgpu_ctx* c; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Timer() { c->traj_host_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); c->traj_host_n++; } } timer{c};
  c->traj_req = -1;
  auto& b = c->tbuf[which];
  const int n = c->traj_req_n; const int rows = c->traj_req_rows;
  if (n > 0) {
    // the side stream only reads the start-of-step arrays of the next step (id, lane, pos, vit, acc) and the status, which that step
    // leaves alone until its lane-change phase / k_final — they wait for t_packed
    cudaStream_t s = c->side; const int B = 256; int G = (n + B - 1) / B; int nb = (n + kScanB - 1) / kScanB;
    if (c->mp.on || c->ext_stream) { CK(cudaEventRecord(c->t_ready, c->stream)); CK(cudaStreamWaitEvent(s, c->t_ready, 0)); }   // else the host waited for the step
    k_alive_flags<<<G, B, 0, s>>>(c->v_status.p, n, c->t_flag.p); LAUNCH(c);
    k_scan_block<<<nb, kScanB, 0, s>>>(c->t_flag.p, c->t_off.p, c->t_bsum.p, n); LAUNCH(c);
    if (nb > 1) { k_scan_sums<<<1, kScanB, 0, s>>>(c->t_bsum.p, nb); LAUNCH(c); k_scan_add<<<nb, kScanB, 0, s>>>(c->t_off.p, c->t_bsum.p, n); LAUNCH(c); }
    k_pack_traj<<<G, B, 0, s>>>(c->dv, n, c->t_flag.p, c->t_off.p, c->t_id.p, c->t_lane.p, c->t_pos.p, c->t_vit.p, c->t_acc.p); LAUNCH(c);
    CK(cudaEventRecord(c->t_packed, s)); c->pack_wait = true;
    CK(cudaEventRecord(c->t_begin[which], c->side));
    CK(cudaMemcpyAsync(b.id, c->t_id.p, (size_t)rows * 4, cudaMemcpyDeviceToHost, c->side));
    CK(cudaMemcpyAsync(b.lane, c->t_lane.p, (size_t)rows * 4, cudaMemcpyDeviceToHost, c->side));
    CK(cudaMemcpyAsync(b.pos, c->t_pos.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, c->side));
    CK(cudaMemcpyAsync(b.vit, c->t_vit.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, c->side));
    CK(cudaMemcpyAsync(b.acc, c->t_acc.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, c->side));
  }
  CK(cudaEventRecord(c->t_done[which], c->side));
  return SYMGPU_OK;
}
extern "C" int symgpu_traj_copy_async(symgpu_ctx* c, int which) {
  if (!c || which < 0 || which >= kTrajBufs || !c->tbuf[which].id) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  if (c->traj_req >= 0) { int rc = traj_enqueue(c); if (rc != SYMGPU_OK) return rc; }
  const int rows = c->n_alive;
  if (rows > c->tbuf[which].cap) { g_err = "trajectory buffer too small"; return SYMGPU_E_CAPACITY; }
  c->traj_req = which; c->traj_req_n = c->n_slots; c->traj_req_rows = rows;
  { int rc = traj_enqueue(c); if (rc != SYMGPU_OK) return rc; }
  return rows;
}
extern "C" int symgpu_step_traj_async(symgpu_ctx* c, int which) {
  if (!c || which < 0 || which >= kTrajBufs || !c->tbuf[which].id) return SYMGPU_E_ARG;
  int r = symgpu_step(c);
  if (r < 0) return r;
  return symgpu_traj_copy_async(c, which);
}
extern "C" int symgpu_traj_wait(symgpu_ctx* c, int which) {
  if (!c || which < 0 || which >= kTrajBufs) return SYMGPU_E_ARG;
  if (c->traj_req == which) { int rc = traj_enqueue(c); if (rc != SYMGPU_OK) return rc; }
  CK(cudaEventSynchronize(c->t_done[which]));
  float ms = 0; if (cudaEventElapsedTime(&ms, c->t_begin[which], c->t_done[which]) == cudaSuccess) c->t_copy_ms[which] = ms; else cudaGetLastError();
  return SYMGPU_OK;
}
// Synchronous convenience form: step, copy the rows into the caller's (pageable or pinned) buffers, return rows.
extern "C" int symgpu_step_traj(symgpu_ctx* c, int cap, int* id, int* lane, double* pos, double* vit, double* acc) {
  if (!c) return SYMGPU_E_ARG;
  auto saved = c->tbuf[0];
  c->tbuf[0] = symgpu_ctx::TrajBuf(); c->tbuf[0].cap = cap; c->tbuf[0].id = id; c->tbuf[0].lane = lane; c->tbuf[0].pos = pos; c->tbuf[0].vit = vit; c->tbuf[0].acc = acc;
  int r = symgpu_step_traj_async(c, 0);
  if (r >= 0) { int w = symgpu_traj_wait(c, 0); if (w < 0) r = w; }
  c->tbuf[0] = saved;
  return r;
}
extern "C" double symgpu_traj_copy_ms(symgpu_ctx* c, int which) { return c && which >= 0 && which < kTrajBufs ? c->t_copy_ms[which] : 0.0; }
extern "C" int symgpu_profile(symgpu_ctx* c, int enable) { if (!c) return SYMGPU_E_ARG; c->profile = enable; for (int k = 0; k < 16; k++) { c->k_ms[k] = 0; c->k_launches[k] = 0; } return SYMGPU_OK; }
extern "C" int symgpu_kernel_count(void) { return KID_COUNT; }
extern "C" const char* symgpu_kernel_name(int k) { return k >= 0 && k < KID_COUNT ? kKernelNames[k] : ""; }
extern "C" double symgpu_kernel_ms(symgpu_ctx* c, int k) { return c && k >= 0 && k < KID_COUNT ? c->k_ms[k] : 0.0; }
extern "C" long symgpu_kernel_launches(symgpu_ctx* c, int k) { return c && k >= 0 && k < KID_COUNT ? (long)c->k_launches[k] : 0; }
extern "C" long long symgpu_vehicle_steps(symgpu_ctx* c) { return c ? c->veh_steps : 0; }
extern "C" int symgpu_num_vehicles(symgpu_ctx* c) { return c ? c->n_alive : SYMGPU_E_ARG; }
extern "C" double symgpu_time(symgpu_ctx* c) { return c ? c->t : 0.0; }
extern "C" unsigned symgpu_rng_count(symgpu_ctx* c) {
  if (!c) return 0;
  if (!c->mp.on && !c->host_rng_fresh) { cudaSetDevice(c->device); if (cudaMemcpy(&c->rng, c->lc_rng.p, sizeof(DevRng), cudaMemcpyDeviceToHost) == cudaSuccess) c->host_rng_fresh = true; }
  return c->rng.count;
}
// merge points: last passage instant and draw counter of every point, in Liste_convergents order (test access)
extern "C" int symgpu_get_merge_state(symgpu_ctx* c, int cap, double* inst_last, int* rand_numbers) {
  if (!c) return SYMGPU_E_ARG;
  const int np = c->mg.on ? c->mg.h.n_pt : 0; const int m = std::min(np, cap);
  CK(cudaSetDevice(c->device));
  if (m > 0 && inst_last) CK(cudaMemcpy(inst_last, c->mg.pt_last.p, m * sizeof(double), cudaMemcpyDeviceToHost));
  if (m > 0 && rand_numbers) CK(cudaMemcpy(rand_numbers, c->mg.pt_rand.p, m * sizeof(int), cudaMemcpyDeviceToHost));
  return np;
}
extern "C" double symgpu_last_step_ms(symgpu_ctx* c) { return c ? c->last_ms : 0.0; }
extern "C" int symgpu_num_links(symgpu_ctx* c) { return c ? c->net.n_links() : SYMGPU_E_ARG; }
extern "C" int symgpu_num_classes(symgpu_ctx* c) { return c ? c->net.n_cls() : SYMGPU_E_ARG; }
extern "C" int symgpu_waiting(symgpu_ctx* c) { if (!c) return SYMGPU_E_ARG; int n = 0; for (auto& e : c->entries) for (auto& l : e.lanes) n += (int)l.queue.size() + (l.pool_slot >= 0); return n; }
extern "C" long symgpu_counter(symgpu_ctx* c, int which) {
  if (!c) return SYMGPU_E_ARG;
  if (which == SYMGPU_CNT_CREATED) return (long)c->created;
  if (which == SYMGPU_CNT_PASSES) return (long)c->passes;
  if (which == SYMGPU_CNT_LAUNCHES) return (long)c->launches;
  if (which == SYMGPU_CNT_SLOTS) return c->n_slots;
  if (which == SYMGPU_CNT_LANE_CHANGES) return (long)c->lane_changes;
  long long h[16];
  cudaSetDevice(c->device);
  if (cudaMemcpy(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return SYMGPU_E_CUDA;
  if (which == SYMGPU_CNT_CTX_OVERFLOW) return (long)(h[which] & 0xffffffff);
  return which >= 0 && which < 16 ? (long)h[which] : SYMGPU_E_ARG;
}
extern "C" int symgpu_get_link_counts(symgpu_ctx* c, int* entre, int* sorti) {
  if (!c) return SYMGPU_E_ARG;
  CK(cudaSetDevice(c->device));
  size_t n = (size_t)c->net.n_links() * c->net.n_cls();
  if (entre) CK(cudaMemcpy(entre, c->entre.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  if (sorti) CK(cudaMemcpy(sorti, c->sorti.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  return (int)n;
}
extern "C" int symgpu_get_exited(symgpu_ctx* c, int cap, int* id, double* instant) {
  if (!c) return SYMGPU_E_ARG;
  int n = (int)c->last_exited.size();
  for (int k = 0; k < n && k < cap; k++) { if (id) id[k] = c->last_exited[k].id; if (instant) instant[k] = c->last_exited[k].inst; }
  return n;
}
