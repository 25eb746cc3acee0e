// lanechange.cuh — lane changing.
//
// Reseau::ChangementDeVoie (reseau.cpp:4219-4330) is a sequential loop over m_LstVehicles: every acceptance test draws from the
// one global random stream (AbstractCarFollowing.cpp:324) and a successful change moves the vehicle at once (vehicule.cpp:2674),
// which changes the tests — and the number of draws — of the vehicles that come later in the list.  Bit-exact lane assignment
// therefore needs the reference's order.  The device keeps that order exactly but only serialises what is order-dependent:
//   k_lc_prepare   per vehicle, parallel: CalculVoiesPossibles / CalculVoieDesiree (vehicule.cpp:4749-5038)
//   k_lc_eval      per vehicle, parallel, no side effects: everything CalculChangementVoie and TestLaneChange do up to the
//                  random draw — neighbours on both lanes, supply/demand, the acceptance probability phi — stored as one
//                  record per (vehicle, pass, candidate lane).  A record is valid as long as nothing changed on its link.
//   k_lc_walk      one block: thread 0 walks the events in list order through the records (skipping, 256 at a time, the
//                  events that neither draw nor change anything), draws from MT19937 on the device, commits the accepted
//                  changes; after every commit the block re-evaluates, in parallel, the later records of the link concerned
//                  (and of the links leading into it).  Events with the reference's rare side effects on failure (the
//                  0.5 m move-back, the lane swap of two stopped vehicles) run the literal sequential code.
//   k_lc_sequential  the literal one-thread version (SYMGPU_LC=seq), kept as the device-side cross-check.
#pragma once
#include "micro.cuh"

namespace symb200 {

struct DevRng { unsigned mt[624]; int idx; unsigned count; };

__device__ inline unsigned rng_raw(DevRng* r) {                        // boost::mt19937
  if (r->idx >= 624) {
    for (int i = 0; i < 624; i++) { unsigned y = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
      r->mt[i] = r->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
    r->idx = 0; }
  unsigned y = r->mt[r->idx++]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; return y;
}
__device__ inline int rng_next(DevRng* r) {                            // RandManager::myRand RandManager.cpp:26-30
  r->count++;
  const unsigned bucket = 0xFFFFFFFFu / 32767u;
  for (;;) { unsigned v = rng_raw(r) / bucket; if (v <= 32766u) return (int)v; }
}

// advances a state held in shared memory: a twist of the 624 words has four dependency phases (words 0..226 only need old words,
// 227..453 need the new 0..226, 454..622 the new 227..453, word 623 the new 0 and 396); the tempered outputs of a block of words are
// produced in parallel and a rejected value (4 in 2^32) sends that block through a serial compaction.
struct MtShared { unsigned mt[624]; unsigned nw[624]; int idx; unsigned count; };
__device__ inline void mt_twist(MtShared& S) {
  const int tid = threadIdx.x;
  auto f = [&](const unsigned* lo, const unsigned* hi, const unsigned* far) { const unsigned y = (*lo & 0x80000000u) | (*hi & 0x7fffffffu); return *far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); };
  if (tid < 227) S.nw[tid] = f(&S.mt[tid], &S.mt[tid + 1], &S.mt[tid + 397]);
  __syncthreads();
  if (tid < 227) S.nw[227 + tid] = f(&S.mt[227 + tid], &S.mt[228 + tid], &S.nw[tid]);
  __syncthreads();
  if (tid < 169) S.nw[454 + tid] = f(&S.mt[454 + tid], &S.mt[455 + tid], &S.nw[227 + tid]);
  __syncthreads();
  if (tid == 0) S.nw[623] = f(&S.mt[623], &S.nw[0], &S.nw[396]);
  __syncthreads();
  for (int i = tid; i < 624; i += blockDim.x) S.mt[i] = S.nw[i];
  __syncthreads();
}
// consume `n` accepted draws; out != nullptr: their values are stored (out[0..n))
__device__ inline void rng_advance(MtShared& S, unsigned n, int* out) {
  const int tid = threadIdx.x; const unsigned bucket = 0xFFFFFFFFu / 32767u;
  unsigned left = n, written = 0; int idx = S.idx;                        // uniform over the block
  while (left > 0) {
    if (idx >= 624) { mt_twist(S); idx = 0; }
    const unsigned m = min(left, (unsigned)(624 - idx));
    int rej = 0; unsigned vals[3]; int nv = 0;
    for (unsigned k = tid; k < m; k += blockDim.x) { unsigned y = S.mt[idx + k]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
      const unsigned v = y / bucket; vals[nv++] = v; if (v > 32766u) rej++; }
    const int nrej = __syncthreads_count(rej);
    if (nrej == 0) { if (out) { int q = 0; for (unsigned k = tid; k < m; k += blockDim.x) out[written + k] = (int)vals[q++]; } written += m; left -= m; }
    else { if (out && tid == 0) { unsigned w = written; for (unsigned k = 0; k < m; k++) { unsigned y = S.mt[idx + k]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
          const unsigned v = y / bucket; if (v <= 32766u) out[w++] = (int)v; } }
      written += m - nrej; left -= m - nrej; }
    idx += (int)m;
  }
  __syncthreads();
  if (tid == 0) { S.idx = idx; S.count += n; }
  __syncthreads();
}
__device__ inline void mt_load(MtShared& S, const DevRng* r) { for (int k = threadIdx.x; k < 624; k += blockDim.x) S.mt[k] = r->mt[k]; if (threadIdx.x == 0) { S.idx = r->idx; S.count = r->count; } __syncthreads(); }
__device__ inline void mt_store(const MtShared& S, DevRng* r) { __syncthreads(); for (int k = threadIdx.x; k < 624; k += blockDim.x) r->mt[k] = S.mt[k]; if (threadIdx.x == 0) { r->idx = S.idx; r->count = S.count; } }
// one draw by a single thread from a state in shared memory (the rare case where the look-ahead buffer runs dry inside an event)
__device__ inline int mt_next_serial(MtShared& S) {
  const unsigned bucket = 0xFFFFFFFFu / 32767u;
  for (;;) {
    if (S.idx >= 624) { for (int i = 0; i < 624; i++) { unsigned y = (S.mt[i] & 0x80000000u) | (S.mt[(i + 1) % 624] & 0x7fffffffu); S.mt[i] = S.mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); } S.idx = 0; }
    unsigned y = S.mt[S.idx++]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    const unsigned v = y / bucket; if (v <= 32766u) { S.count++; return (int)v; }
  }
}

struct DevLc {
  const int *order, *lane_start, *lane_count;
  int *vdes, *voies_ok, *chgt, *delta_next, *link_delta_head, *moved; const int* posidx; const double* spos;   // spos: the lane order's sort keys (k_lane_rank)
  DevRng* rng; int* n_changes;      // n_changes[1]: draws consumed by k_lc_walk (the host advances its own stream)
  double dst_chgtvoie, dst_force, phi_force; int ordre;
  // records of k_lc_eval: stage 0 = mandatory pass, 1 / 2 = first / second candidate lane of the discretionary pass
  unsigned char *kind, *code; int *tl_pre, *rfol, *rtgt, *had_mand; double* phi;
  long long* counters;
};
enum : int { LC_FAIL = 0, LC_TEST = 1, LC_SLOW = 2 };

// Vehicule::CalculVoiesPossibles + CalculVoieDesiree (vehicule.cpp:4749-5038), phase 6 of the step, per vehicle
__global__ void k_lc_prepare(DevNet N, DevVeh V, DevLc C, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  C.chgt[i] = 0; C.delta_next[i] = -1; C.moved[i] = 0;
  if (V.status[i] != ST_ALIVE) return;                                // vehicles created this step have no link at index 1 (reseau.cpp:3941)
  int lane = V.lane[i]; int link = N.lane_link[lane]; int nl = N.link_nl[link]; int first = N.link_first[link];
  double p1 = snap_pos(V.pos[i]);
  int mask = 0;
  bool zone = p1 > N.lane_len[lane] - C.dst_chgtvoie;
  int ta = N.link_type[link];
  if (zone && ta != 0 && N.link_brick[link] < 0) {
    if (ta == 'S' || ta == 'P') mask = (1 << nl) - 1;
    else {
      // next link of the itinerary (Vehicule::CalculNextTuyau vehicule.cpp:2764-2826)
      int route = V.route[i]; int o = N.route_off[route], e = N.route_off[route + 1];
      int k = route_index(N, route, V.route_pos[i], link); int nxt = (k >= 0 && o + k + 1 < e) ? N.route_links[o + k + 1] : -1;
      for (int q = 0; q < nl; q++) { int ln = first + q; bool ok = false;
        if (nxt >= 0) for (int m = N.succ_off[ln]; m < N.succ_off[ln + 1]; m++) if (N.succ_key[m] == nxt) { ok = true; break; }
        if (ok && !(N.lane_flags[ln] & 1)) mask |= 1 << q; }
    }
  } else for (int q = 0; q < nl; q++) if (!(N.lane_flags[first + q] & 1)) mask |= 1 << q;
  C.voies_ok[i] = mask;
  if (C.vdes[i] == -1) {
    int nV = N.lane_num[lane];
    if (!((mask >> nV) & 1)) {
      int d = -1;
      if (nV - 1 >= 0 && ((mask >> (nV - 1)) & 1)) d = nV - 1;
      else if (nV + 1 < nl && ((mask >> (nV + 1)) & 1)) d = nV + 1;
      else { for (int k = nV - 2; k >= 0 && d < 0; k--) if ((mask >> k) & 1) d = nV - 1;
             for (int k = nV + 2; k < nl && d < 0; k++) if ((mask >> k) & 1) d = nV + 1; }
      C.vdes[i] = d;
    }
  }
}

// ---- the sequential part (one thread) ----------------------------------------------------------------------------------
// rb..srng: k_lc_walk draws through a buffer of upcoming values kept in shared memory (all null elsewhere: direct draws)
constexpr int kSnapLanes = 4, kSnapCap = 96;
// the vehicles currently on one link, lane by lane, copied to shared memory by k_lc_walk around a commit: the neighbour queries of that
// link then scan this copy (one 16-byte record per vehicle) instead of chasing the lane slices through global memory
struct SnapE { double p; int v; int id; };
struct LcSnap { int link, lane0, nl, over; int cnt[kSnapLanes]; double len[kSnapLanes]; SnapE e[kSnapLanes][kSnapCap]; int prev[kSnapLanes][kSnapCap]; };
#ifdef SYM_LC_TIMING
__device__ long long g_lcprof2[12];
#define LCQ(i) do { long long t_ = clock64(); g_lcprof2[i] += t_ - tq_; tq_ = t_; } while (0)
#define LCQ0 long long tq_ = clock64()
#else
#define LCQ(i) do { } while (0)
#define LCQ0 do { } while (0)
#endif
struct LcView { const DevNet& N; const DevVeh& V; const DevLc& C; int* rb = nullptr; int* rb_pos = nullptr; int* rb_fill = nullptr; int* rb_base = nullptr; MtShared* srng = nullptr; LcSnap* snap = nullptr; };
constexpr int kRb = 2048;
// make at least `need` upcoming draws available in the buffer (one thread)
__device__ inline void lc_fill(const LcView& W, int need) {
  int pos = *W.rb_pos, fill = *W.rb_fill;
  if (fill - pos >= need) return;
  for (int k = pos; k < fill; k++) W.rb[k - pos] = W.rb[k];
  *W.rb_base += pos; fill -= pos; pos = 0;
  while (fill < kRb) W.rb[fill++] = mt_next_serial(*W.srng);
  *W.rb_pos = pos; *W.rb_fill = fill;
}
// the same by the whole block (call sites are uniform; the buffer words were last written before a barrier)
__device__ inline void lc_fill_block(const LcView& W, int need) {
  const int pos = *W.rb_pos, fill = *W.rb_fill;
  if (fill - pos >= need) return;
  const int rem = fill - pos; int keep[(kRb + 255) / 256]; int nk = 0;
  for (int k = threadIdx.x; k < rem; k += blockDim.x) keep[nk++] = W.rb[pos + k];
  __syncthreads();
  nk = 0; for (int k = threadIdx.x; k < rem; k += blockDim.x) W.rb[k] = keep[nk++];
  __syncthreads();
  rng_advance(*W.srng, (unsigned)(kRb - rem), W.rb + rem);
  if (threadIdx.x == 0) { *W.rb_base += pos; *W.rb_pos = 0; *W.rb_fill = kRb; }
  __syncthreads();
}
__device__ inline int lc_draw(const LcView& W) {                       // RandManager::myRand in list order
  if (!W.rb) return rng_next(W.C.rng);
  lc_fill(W, 1);
  return W.rb[(*W.rb_pos)++];
}
constexpr int kTie = 8;

// The vehicles currently on `lane` = the lane's slice of the start-of-phase order (sorted by position; `spos` keeps the sort
// keys), minus those that left it or were displaced since (`moved`), plus the link's list of displaced vehicles.  Queries
// bracket the slice by binary search and only visit the few entries that can matter; the displaced list is short.
__device__ __forceinline__ bool slice_ok(const LcView& W, int lane, int y) { return W.V.lane[y] == lane && !W.C.moved[y] && W.V.status[y] == ST_ALIVE; }   // vehicles created this step join their lane set later (vehicule.cpp:1224)
template <class F> __device__ inline void for_slice(const LcView& W, int lane, int q0, int q1, F f) {
  for (int q = q0; q < q1; q++) { int y = W.C.order[q]; if (slice_ok(W, lane, y)) f(y); }
}
template <class F> __device__ inline void for_moved(const LcView& W, int lane, F f) {
  for (int y = W.C.link_delta_head[W.N.lane_link[lane]]; y >= 0; y = W.C.delta_next[y]) if (W.V.lane[y] == lane) f(y);
}
__device__ inline int slice_lb(const LcView& W, int lane, double pos) {             // first order position of the lane with key >= pos
  int lo = W.C.lane_start[lane], hi = lo + W.C.lane_count[lane];
  while (lo < hi) { int m = (lo + hi) >> 1; if (W.C.spos[m] < pos) lo = m + 1; else hi = m; }
  return lo;
}
__device__ inline int slice_ub(const LcView& W, int lane, double pos) {             // first order position of the lane with key > pos
  int lo = W.C.lane_start[lane], hi = lo + W.C.lane_count[lane];
  while (lo < hi) { int m = (lo + hi) >> 1; if (W.C.spos[m] <= pos) lo = m + 1; else hi = m; }
  return lo;
}
// [q0, end of the group of equal keys that holds the first usable entry at or after q0)
__device__ inline int slice_group_fwd(const LcView& W, int lane, int q0, int excl) {
  const int e = W.C.lane_start[lane] + W.C.lane_count[lane];
  int q = q0; while (q < e) { int y = W.C.order[q]; if (y != excl && slice_ok(W, lane, y)) break; q++; }
  if (q >= e) return e;
  const double k = W.C.spos[q]; while (q < e && W.C.spos[q] == k) q++;
  return q;
}
// [start of the group of equal keys that holds the last usable entry before q1, q1)
__device__ inline int slice_group_bwd(const LcView& W, int lane, int q1, int excl) {
  const int b = W.C.lane_start[lane];
  int q = q1 - 1; while (q >= b) { int y = W.C.order[q]; if (y != excl && slice_ok(W, lane, y)) break; q--; }
  if (q < b) return b;
  const double k = W.C.spos[q]; while (q >= b && W.C.spos[q] == k) q--;
  return q + 1;
}
__device__ inline void lc_mark_moved(const LcView& W, int v) {                       // v no longer sits where the sorted slice says
  W.C.moved[v] = 1;
  const int link = W.N.lane_link[W.V.lane[v]];
  if (W.C.delta_next[v] == -1 && W.C.link_delta_head[link] != v) { W.C.delta_next[v] = W.C.link_delta_head[link]; W.C.link_delta_head[link] = v; }
}
__device__ inline double p1of(const LcView& W, int y) { return snap_pos(W.V.pos[y]); }
// Reseau::GetLastVehicule (reseau.cpp:14041-14081) over candidates collected in id order
__device__ __noinline__ int pick_last(const LcView& W, int* c, int n) {
  if (n == 0) return -1;
  if (n == 1) return c[0];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (W.V.id[c[j]] < W.V.id[c[i]]) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { bool fol = false; for (int j = 0; j < n && !fol; j++) if (i != j && W.V.prev_leader[c[j]] == W.V.id[c[i]]) fol = true; if (!fol) return c[i]; }
  return c[n - 1];
}
// Reseau::GetFirstVehicule(list) (reseau.cpp:14084-14128)
__device__ __noinline__ int pick_first(const LcView& W, int* c, int n) {
  if (n == 0) return -1;
  if (n == 1) return c[0];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (W.V.id[c[j]] < W.V.id[c[i]]) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { int ld = W.V.prev_leader[c[i]]; if (ld < 0) return c[i];
    bool found = false; for (int j = 0; j < n && !found; j++) if (i != j && W.V.id[c[j]] == ld) found = true; if (!found) return c[i]; }
  return c[n - 1];
}
// ---- the same queries over the shared-memory copy of a link (LcSnap): every vehicle of the lane is visited, the filters are the same.
// First pass: the extreme position and how many vehicles share it (registers only); the tie rules run only when several do.
__device__ __forceinline__ bool snap_has(const LcView& W, int lane) { return W.snap && W.snap->link >= 0 && (unsigned)(lane - W.snap->lane0) < (unsigned)W.snap->nl; }
__device__ __noinline__ int snap_pick_last(const LcSnap& S, int li, int* c, int n) {
  const SnapE* e = S.e[li]; const int* pr = S.prev[li];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (e[c[j]].id < e[c[i]].id) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { bool fol = false; for (int j = 0; j < n && !fol; j++) if (i != j && pr[c[j]] == e[c[i]].id) fol = true; if (!fol) return e[c[i]].v; }
  return e[c[n - 1]].v;
}
__device__ __noinline__ int snap_pick_first(const LcSnap& S, int li, int* c, int n) {
  const SnapE* e = S.e[li]; const int* pr = S.prev[li];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (e[c[j]].id < e[c[i]].id) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { int ld = pr[c[i]]; if (ld < 0) return e[c[i]].v;
    bool found = false; for (int j = 0; j < n && !found; j++) if (i != j && e[c[j]].id == ld) found = true; if (!found) return e[c[i]].v; }
  return e[c[n - 1]].v;
}
// vehicles of the lane at position `at`, except `excl` (at most kTie, in copy order)
__device__ __noinline__ int snap_ties(const LcSnap& S, int li, double at, int excl, int* c) {
  const SnapE* e = S.e[li]; const int cnt = S.cnt[li]; int n = 0;
  for (int i = 0; i < cnt; i++) if (e[i].p == at && e[i].v != excl && n < kTie) c[n++] = i;
  return n;
}
// The scans keep no branch inside the loop (selects on the extreme position, then a count of the vehicles standing there) and read
// the records four at a time before using them: the reads overlap and two threads of a warp scanning different lanes do not take turns.
template <class F> __device__ __forceinline__ void snap_scan(const SnapE* e, int cnt, F f) {
  for (int i0 = 0; i0 < cnt; i0 += 4) {
    SnapE x[4];
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = e[min(i0 + j, cnt - 1)];
#pragma unroll
    for (int j = 0; j < 4; j++) f(x[j], i0 + j < cnt);
  }
}
__device__ __forceinline__ int snap_near_aval(const LcView& W, int lane, double pos, int excl) {
  const LcSnap& S = *W.snap; const int li = lane - S.lane0; const SnapE* e = S.e[li]; const int cnt = S.cnt[li];
  const double bound = S.len[li] + 0.00001; double pv = DBL_MAX; int bv = -1;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { const bool better = in && x.v != excl && x.p >= pos && x.p <= bound && x.p < pv; pv = better ? x.p : pv; bv = better ? x.v : bv; });
  if (bv < 0) return -1;
  int nt = 0;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { nt += (in && x.p == pv && x.v != excl) ? 1 : 0; });
  if (nt <= 1) return bv;
  int c[kTie]; const int n = snap_ties(S, li, pv, excl, c);
  return snap_pick_last(S, li, c, n);
}
__device__ __forceinline__ int snap_near_amont(const LcView& W, int lane, double pos, int excl) {
  const LcSnap& S = *W.snap; const int li = lane - S.lane0; const SnapE* e = S.e[li]; const int cnt = S.cnt[li];
  double pt = -DBL_MAX; int bv = -1;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { const bool better = in && x.v != excl && pos > x.p && x.p >= -99999 && x.p > pt; pt = better ? x.p : pt; bv = better ? x.v : bv; });
  if (bv < 0 || !((pos - pt) < 9999)) return -1;
  int nt = 0;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { nt += (in && x.p == pt && x.v != excl) ? 1 : 0; });
  if (nt <= 1) return bv;
  int c[kTie]; const int n = snap_ties(S, li, pt, excl, c);
  return snap_pick_first(S, li, c, n);
}
__device__ __forceinline__ int snap_first_vehicule(const LcView& W, int lane) {               // the largest position (>= 0), the largest id among equals
  const LcSnap& S = *W.snap; const int li = lane - S.lane0; const SnapE* e = S.e[li]; const int cnt = S.cnt[li];
  int r = -1; double p = -1; int rid = -1;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { const bool better = in && x.p >= 0 && (x.p > p || (x.p == p && x.id > rid)); p = better ? x.p : p; r = better ? x.v : r; rid = better ? x.id : rid; });
  return r;
}
__device__ __forceinline__ int snap_prev_vehicule(const LcView& W, int v, int lane, double own) {
  const LcSnap& S = *W.snap; const int li = lane - S.lane0; const SnapE* e = S.e[li]; const int cnt = S.cnt[li];
  const double bound = S.len[li] + 1; double pv = DBL_MAX; int bv = -1;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { const bool better = in && x.v != v && x.p > own && x.p <= bound && x.p < pv; pv = better ? x.p : pv; bv = better ? x.v : bv; });
  if (bv < 0) return -1;
  int nt = 0;
  snap_scan(e, cnt, [&](const SnapE& x, bool in) { nt += (in && x.p == pv && x.v != v) ? 1 : 0; });
  if (nt <= 1) return bv;
  int c[kTie]; const int n = snap_ties(S, li, pv, v, c);
  return snap_pick_last(S, li, c, n);
}
__device__ inline void snap_update(const LcView& W, int v, int old_lane, int lane, double pos) {            // v moved from old_lane (one thread)
  if (!snap_has(W, lane)) return;
  LcSnap& S = *W.snap; const int ln = lane - S.lane0, li = old_lane - S.lane0;
  if ((unsigned)li < (unsigned)S.nl) { const int cnt = S.cnt[li]; int i = -1;
    for (int k = 0; k < cnt; k++) i = S.e[li][k].v == v ? k : i;
    if (i >= 0) {
      SnapE x = S.e[li][i]; const int pr = S.prev[li][i]; x.p = snap_pos(pos);
      if (li == ln) { S.e[li][i] = x; return; }
      const int last = --S.cnt[li]; S.e[li][i] = S.e[li][last]; S.prev[li][i] = S.prev[li][last];
      if (S.cnt[ln] >= kSnapCap) { S.link = -1; return; }
      const int k = S.cnt[ln]++; S.e[ln][k] = x; S.prev[ln][k] = pr; return; } }
  S.link = -1;                                                                               // not in the copy: the copy cannot be trusted
}
__device__ __noinline__ int near_aval(const LcView& W, int lane, double pos, int excl) {           // reseau.cpp:3583-3640
  if (snap_has(W, lane)) return snap_near_aval(W, lane, pos, excl);
  int c[kTie]; int n = 0; double pv = W.N.lane_len[lane] + 0.00001;
  auto f = [&](int y) { double p = p1of(W, y); if (p >= pos && p <= pv && y != excl) { if (p == pv) { if (n < kTie) c[n++] = y; } else { n = 1; c[0] = y; pv = p; } } };
  const int q0 = slice_lb(W, lane, pos);
  for_slice(W, lane, q0, slice_group_fwd(W, lane, q0, excl), f); for_moved(W, lane, f);
  return pick_last(W, c, n);
}
__device__ __noinline__ int near_amont(const LcView& W, int lane, double pos, int excl) {          // reseau.cpp:3374-3436 (first stage; no previous lane in scope)
  if (snap_has(W, lane)) return snap_near_amont(W, lane, pos, excl);
  int c[kTie]; int n = 0; double pt = -99999;
  auto f = [&](int y) { double p = p1of(W, y); if (pos > p && p >= pt && y != excl) { if (p == pt) { if (n < kTie) c[n++] = y; } else { n = 1; c[0] = y; pt = p; } } };
  const int q1 = slice_lb(W, lane, pos);
  for_slice(W, lane, slice_group_bwd(W, lane, q1, excl), q1, f); for_moved(W, lane, f);
  if (n > 0 && (pos - pt) < 9999) return pick_first(W, c, n);
  return -1;
}
__device__ __noinline__ int first_vehicule(const LcView& W, int lane) {                            // reseau.cpp:3255-3295 (last maximum in id order)
  if (snap_has(W, lane)) return snap_first_vehicule(W, lane);
  int r = -1; double p = 0; int rid = -1;
  auto f = [&](int y) { double q = p1of(W, y); if (q > p || (q == p && (r < 0 || W.V.id[y] > rid))) { if (q >= p) { r = y; p = q; rid = W.V.id[y]; } } };
  const int q1 = W.C.lane_start[lane] + W.C.lane_count[lane];
  for_slice(W, lane, slice_group_bwd(W, lane, q1, -1), q1, f); for_moved(W, lane, f);
  return r;
}
__device__ __noinline__ int prev_vehicule(const LcView& W, int v) {                                // reseau.cpp:3650-3718
  int lane = W.V.lane[v];
  if (snap_has(W, lane)) return snap_prev_vehicule(W, v, lane, p1of(W, v));
  int c[kTie]; int n = 0; double pv = W.N.lane_len[lane] + 1; double own = p1of(W, v);
  auto f = [&](int y) { double p = p1of(W, y); if (p > own && p <= pv && y != v) { if (p == pv) { if (n < kTie) c[n++] = y; } else { n = 1; c[0] = y; pv = p; } } };
  const int q0 = slice_ub(W, lane, own);
  for_slice(W, lane, q0, slice_group_fwd(W, lane, q0, v), f); for_moved(W, lane, f);
  return pick_last(W, c, n);
}
__device__ inline double dist_veh(const LcView& W, int a, int b) {                           // reseau.cpp:4446-4560
  if (a < 0 || b < 0) return 9999;
  int la = W.N.lane_link[W.V.lane[a]], lb = W.N.lane_link[W.V.lane[b]];
  if (la == lb) return fabs(p1of(W, b) - p1of(W, a));
  if (W.N.link_down[la] == W.N.link_up[lb]) return W.N.lane_len[W.V.lane[a]] - p1of(W, a) + p1of(W, b);
  if (W.N.link_down[lb] == W.N.link_up[la]) return W.N.lane_len[W.V.lane[b]] - p1of(W, b) + p1of(W, a);
  return 9999;
}
__device__ inline double vit_eq(const DevCls& c, double conc) {                              // DiagFonda.cpp:116-172, :221-276
  double kc = c.w * c.kx / (c.w - c.vx); double debit;
  if (conc <= kc) { debit = c.vx * conc; if (fabs(debit) < DBL_EPSILON) debit = 0; }
  else { debit = c.w * (conc - c.kx); if (fabs(debit) < DBL_EPSILON) debit = 0; }
  if (fabs(debit) < DBL_EPSILON && fabs(conc) < DBL_EPSILON) return 0;
  double vit = conc < DBL_EPSILON ? c.vx : debit / conc;
  if (fabs(vit) < DBL_EPSILON) vit = 0;
  return vit;
}
__device__ inline double debit_max(const DevCls& c) { return c.w * c.kx / ((c.w / c.vx) - 1); }
__device__ inline double vreg_act(const LcView& W, int y) { return W.N.link_vreg[W.N.lane_link[W.V.lane[y]]]; }   // GetVitRegTuyAct (set at the end of the previous step)

// the lane's memoised next-lane draw (Vehicule::GetTirage) consumed through CalculNextVoie
__device__ inline int next_voie_draw(const LcView& W, int v, int lane) {
  int tl; int nx = next_voie(W.N, W.V.route[v], W.V.route_pos[v], lane, &tl);
  if (tl >= 0 && memo_touch(&W.V.memo[v * kMemo], tl)) lc_draw(W);
  return nx;
}
// everything of TestLaneChange before the draw; false = refused without a draw
__device__ __noinline__ bool lane_change_phi(const DevParams& P, const LcView& W, int v, double pv, int fol, int lead, int lead_orig, int tgt, bool force, double* phi_out) {   // AbstractCarFollowing.cpp:61-323
  const DevCls& c = P.cls[W.V.cls[v]]; double kc = c.w * c.kx / (c.w - c.vx);
  int lane = W.V.lane[v]; int link = W.N.lane_link[lane];
  double si;
  if (lead_orig >= 0) si = p1of(W, lead_orig) - pv;
  else if (force) si = fmax(1 / c.kx, W.N.lane_len[lane] - pv);
  else si = 1 / kc;
  double sf = (fol >= 0 && lead >= 0) ? dist_veh(W, lead, fol) : 1 / kc;
  double vit = fmin(vreg_act(W, v), c.vx);
  double dem_orig = fmin(vit / si, debit_max(c));
  double dem_dest;
  if (fol >= 0) { const DevCls& cf = P.cls[W.V.cls[fol]]; dem_dest = fmin(cf.vx / sf, debit_max(cf)); }
  else dem_dest = fmin(vit / sf, debit_max(c));
  double off_dest;
  if (lead >= 0) { const DevCls& cl = P.cls[W.V.cls[lead]]; off_dest = fmin((cl.kx - (1 / sf)) * (-1) * cl.w, debit_max(cl)); }
  else off_dest = debit_max(c);
  double pi;
  if (force) pi = 1 / P.dt;
  else {
    double vtgt;
    if (sf > 0) { const DevCls& d = lead >= 0 ? P.cls[W.V.cls[lead]] : c; vtgt = vit_eq(d, 1 / sf); if (lead >= 0) vtgt = W.V.vit[lead]; } else vtgt = 0;
    vtgt = fmin(vtgt, W.N.link_vreg[W.N.lane_link[tgt]]);
    if (fol < 0 && lead >= 0) { vtgt = fmax(vtgt, W.V.vit[lead]); vtgt = fmin(vtgt, vreg_act(W, lead)); }
    vtgt = fmin(vtgt, c.vx);
    double vorg = DBL_MAX;
    if (si > 0) { const DevCls& d = lead_orig >= 0 ? P.cls[W.V.cls[lead_orig]] : c; vorg = vit_eq(d, 1 / si);
      if (c.law == 0 && (W.V.flags[v] & F_HASCTX) && W.V.dn[v] < 1 && lead_orig >= 0) vorg = W.V.vit[lead_orig]; }
    else vorg = lead_orig >= 0 ? W.V.vit[lead_orig] : W.V.vit[v];
    vorg = fmin(vorg, W.N.link_vreg[link]); vorg = fmin(vorg, c.vx);
    pi = fmax(0., vtgt - vorg) / (W.N.link_vreg[link] * c.tau_lc);
    if (pi < 0.000001) return false;
  }
  double phi;
  if (force && W.N.link_len[link] - pv < W.C.dst_force) phi = W.C.phi_force;
  else {
    if (fol < 0 || lead < 0) dem_dest = 0;
    if (dem_dest > 0) phi = fmin(1., off_dest / dem_dest) * pi * dem_orig / vit * P.dt * si;
    else phi = pi * dem_orig / vit * P.dt * si;
  }
  *phi_out = phi;
  return true;
}
__device__ inline bool test_lane_change(const DevParams& P, const LcView& W, int v, int fol, int lead, int lead_orig, int tgt, bool force) {
  double phi;
  if (!lane_change_phi(P, W, v, p1of(W, v), fol, lead, lead_orig, tgt, force, &phi)) return false;
  double r = lc_draw(W) / 32767.0;                                   // AbstractCarFollowing.cpp:324-328
  return r < phi;
}
__device__ __noinline__ void changement_voie(const LcView& W, int v, int tgt, int fol) {            // vehicule.cpp:2657-2711
  LCQ0;
  W.C.chgt[v] = 1;
  int old = W.V.lane[v]; int link = W.N.lane_link[old];
  double p = p1of(W, v) * W.N.lane_len[tgt] / W.N.lane_len[old];
  W.V.lane[v] = tgt;
  if (fol >= 0 && W.N.lane_link[W.V.lane[fol]] == link) p = fmax(p, p1of(W, fol));
  W.V.pos[v] = p;
  LCQ(2);
  lc_mark_moved(W, v);
  LCQ(3);
  snap_update(W, v, old, tgt, p);
  LCQ(4);
  W.C.vdes[v] = -1;
  next_voie_draw(W, v, tgt);
  LCQ(5);
  atomicAdd(W.C.n_changes, 1);
}
__device__ __noinline__ bool calcul_changement_voie(const DevParams& P, const LcView& W, int v, int tgt_num, bool force) {   // vehicule.cpp:2248-2457
  double p1 = p1of(W, v);
  if (p1 <= SYM_UNDEF) return false;
  if (W.C.chgt[v]) return false;
  int lane = W.V.lane[v]; int link = W.N.lane_link[lane]; bool desired = W.C.vdes[v] != -1; bool oblig = W.N.lane_flags[lane] & 1;
  if (first_vehicule(W, lane) == v && !oblig && !desired) return false;
  int tgt = W.N.link_first[link] + tgt_num;
  double ratio = W.N.lane_len[tgt] / W.N.lane_len[lane];
  bool move_back = false;
  int same = near_aval(W, tgt, ratio * p1 - 0.01, v);
  if (same >= 0 && fabs(ratio * p1 - p1of(W, same)) < 0.01) {
    if (fabs(p1 - W.N.lane_len[lane]) < 0.01) {
      if (force && W.C.vdes[same] == W.N.lane_num[lane] && W.C.vdes[v] == W.N.lane_num[W.V.lane[same]] && W.V.vit[v] < 0.01 && W.V.vit[same] < 0.01) {
        changement_voie(W, v, tgt, -1); changement_voie(W, same, lane, -1); return true;
      }
      W.V.pos[v] = p1 - 0.5; p1 = p1 - 0.5; move_back = true; lc_mark_moved(W, v);
    } else return false;
  }
  double q = ratio * p1 + 0.001;
  int fol = near_amont(W, tgt, q, v);
  int lead = near_aval(W, tgt, q, v);
  if (lead < 0 && (oblig || desired)) { int nx = next_voie_draw(W, v, tgt); if (nx >= 0) lead = near_aval(W, nx, 0, -1); }
  int lead_orig = prev_vehicule(W, v);
  if (lead_orig < 0 && !oblig && !desired && !force) return false;   // vehicule.cpp:2440-2442 (the reference does not undo the move-back here)
  if (test_lane_change(P, W, v, fol, lead, lead_orig, tgt, force)) { changement_voie(W, v, tgt, fol); return true; }
  if (move_back) W.V.pos[v] = p1 + 0.5;
  return false;
}
// ---- parallel evaluation / ordered commit -------------------------------------------------------------------------------
// CalculChangementVoie without side effects: what the event (v, candidate lane) will do when its turn comes, provided
// nothing changes on its link before.
#ifdef SYM_LC_TIMING
__device__ long long g_lcprof[12];
#define LCP(i) do { if (W.snap) { long long t_ = clock64(); atomicAdd((unsigned long long*)&g_lcprof[i], (unsigned long long)(t_ - tp_)); tp_ = t_; } } while (0)
#else
#define LCP(i) do { } while (0)
#endif
__device__ __noinline__ void lc_eval_stage(const DevParams& P, const LcView& W, int v, int tgt_num, bool force, int slot) {
  unsigned char code = LC_FAIL; int tl_pre = -1, rf = -1, rt = -1; double phi = 0;
#ifdef SYM_LC_TIMING
  long long tp_ = clock64(); if (W.snap) atomicAdd((unsigned long long*)&g_lcprof[11], 1ULL);
#endif
  do {
    double p1 = p1of(W, v);
    if (p1 <= SYM_UNDEF || W.C.chgt[v]) break;
    int lane = W.V.lane[v]; int link = W.N.lane_link[lane]; bool desired = W.C.vdes[v] != -1; bool oblig = W.N.lane_flags[lane] & 1;
    LCP(0);
    if (first_vehicule(W, lane) == v && !oblig && !desired) break;
    LCP(1);
    int tgt = W.N.link_first[link] + tgt_num;
    double ratio = W.N.lane_len[tgt] / W.N.lane_len[lane];
    LCP(2);
    int same = near_aval(W, tgt, ratio * p1 - 0.01, v);
    if (same >= 0 && fabs(ratio * p1 - p1of(W, same)) < 0.01) {
      if (fabs(p1 - W.N.lane_len[lane]) < 0.01) code = LC_SLOW;       // move-back / swap: literal code at commit time
      break;
    }
    LCP(3);
    double q = ratio * p1 + 0.001;
    int fol = near_amont(W, tgt, q, v);
    int lead = near_aval(W, tgt, q, v);
    LCP(4);
    if (lead < 0 && (oblig || desired)) {
      int tl; int nx = next_voie(W.N, W.V.route[v], W.V.route_pos[v], tgt, &tl);
      tl_pre = tl;
      if (nx >= 0) lead = near_aval(W, nx, 0, -1);
    }
    LCP(5);
    int lead_orig = prev_vehicule(W, v);
    LCP(6);
    if (lead_orig < 0 && !oblig && !desired && !force) break;
    if (!lane_change_phi(P, W, v, p1, fol, lead, lead_orig, tgt, force, &phi)) break;
    LCP(7);
    code = LC_TEST; rf = fol; rt = tgt;
  } while (0);
  W.C.code[slot] = code; W.C.tl_pre[slot] = tl_pre; W.C.rfol[slot] = rf; W.C.rtgt[slot] = rt; W.C.phi[slot] = phi;
  LCP(8);
}
// records of vehicle v for the passes >= from_pass; n = number of slots.  part 0: the mandatory stage, 1 / 2: first / second candidate lane
// of the discretionary pass (independent of each other: the walk spreads them over threads); lc_eval_kind then classifies the events.
// returns what lc_eval_kind needs to know about the part: bit 0 = the part's pass applies to the vehicle, bit 1 = its record draws or tests
__device__ inline int lc_eval_part(const DevParams& P, const LcView& W, int v, int n, int from_pass, int part) {
  const bool alive = W.V.status[v] == ST_ALIVE;
  if (part == 0) { if (!(from_pass == 0 && alive && W.C.vdes[v] != -1)) return 0; lc_eval_stage(P, W, v, W.C.vdes[v], true, v); return 1 | ((W.C.code[v] != LC_FAIL || W.C.tl_pre[v] >= 0) ? 2 : 0); }
  if (!(alive && !W.C.chgt[v] && !W.C.had_mand[v])) return 0;
  const int lane = W.V.lane[v]; const int link = W.N.lane_link[lane]; const int nl = W.N.link_nl[link];
  if (nl < 2) return 0;
  const int nV = W.N.lane_num[lane]; int nVF = nV + 1, nVS = nV - 1;
  if (W.C.ordre == 2) { nVF = nV - 1; nVS = nV + 1; }
  const int tn = part == 1 ? nVF : nVS; const int slot = part == 1 ? n + v : 2 * n + v;
  W.C.code[slot] = LC_FAIL; W.C.tl_pre[slot] = -1;
  if (tn >= 0 && tn < nl && ((W.C.voies_ok[v] >> tn) & 1)) lc_eval_stage(P, W, v, tn, false, slot);
  return 1 | ((W.C.code[slot] != LC_FAIL || W.C.tl_pre[slot] >= 0) ? 2 : 0);
}
// event kinds from the three parts' answers (r0 only when from_pass == 0)
__device__ inline void lc_eval_kind(const LcView& W, int v, int n, int from_pass, int r0, int r1, int r2) {
#ifdef SYM_LC_NOKIND
  return;
#endif
  if (from_pass == 0) W.C.kind[v] = (r0 & 1) ? ((r0 & 2) ? 1 : 2) : 0;       // 2: only marks the vehicle as handled (m_bDejaCalcule)
  W.C.kind[n + v] = ((r1 | r2) & 2) ? 1 : 0;
}
__device__ inline void lc_eval_vehicle(const DevParams& P, const LcView& W, int v, int n, int from_pass) {
  const int r0 = lc_eval_part(P, W, v, n, from_pass, 0), r1 = lc_eval_part(P, W, v, n, from_pass, 1), r2 = lc_eval_part(P, W, v, n, from_pass, 2);
  lc_eval_kind(W, v, n, from_pass, r0, r1, r2);
}
__global__ void k_lc_eval(DevNet N, DevVeh V, DevLc C, int n) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  LcView W{N, V, C};
  C.had_mand[v] = (V.status[v] == ST_ALIVE && C.vdes[v] != -1) ? 1 : 0;
  lc_eval_vehicle(cP_ref(), W, v, n, 0);
}
// what a commit invalidates: the records of the vehicles whose neighbour queries can reach the moved vehicle's old or new
// place, i.e. those between its strict neighbours on the lane it left and on the lane it joined (normalised abscissa x =
// pos / lane length, which the lane-to-lane projection of vehicule.cpp:2300 preserves); `full`: the whole link;
// `preds`: also the links leading into this one (their vehicles look at the first vehicle of the next lane, :2408-2420).
struct LcDirty { bool dirty, full, preds; double xlo, xhi; };
__device__ __noinline__ void lc_window(const LcView& W, int lane, double pos, int excl, LcDirty* d) {
  double lo = -1, hi = -1;
  auto g = [&](double p) { if (p < pos) { if (lo < 0 || p > lo) lo = p; } else if (p > pos) { if (hi < 0 || p < hi) hi = p; } };
  if (snap_has(W, lane)) { const LcSnap& S = *W.snap; const int li = lane - S.lane0; const SnapE* e = S.e[li]; const int cnt = S.cnt[li]; double l2 = -1, h2 = DBL_MAX;
    snap_scan(e, cnt, [&](const SnapE& x, bool in0) { const bool in = in0 && x.v != excl; l2 = (in && x.p < pos && x.p > l2) ? x.p : l2; h2 = (in && x.p > pos && x.p < h2) ? x.p : h2; });
    lo = l2; hi = h2 == DBL_MAX ? -1 : h2; }
  else {
    auto f = [&](int y) { if (y == excl) return; g(p1of(W, y)); };
    const int qa = slice_lb(W, lane, pos), qb = slice_ub(W, lane, pos);
    for_slice(W, lane, slice_group_bwd(W, lane, qa, excl), qa, f); for_slice(W, lane, qb, slice_group_fwd(W, lane, qb, excl), f); for_moved(W, lane, f);
  }
  const double len = W.N.lane_len[lane];
  if (lo < 0) { d->preds = true; d->xlo = -1; } else d->xlo = fmin(d->xlo, lo / len);
  d->xhi = hi < 0 ? 2.0 : fmax(d->xhi, hi / len);
}
// one stage of an event at its turn (thread 0 of k_lc_walk)
__device__ inline bool lc_run_stage(const DevParams& P, const LcView& W, int v, int slot, int tgt_num, bool force, LcDirty* d) {
  LCQ0;
  const int code = W.C.code[slot];
  if (code == LC_SLOW) { d->dirty = d->full = d->preds = true; return calcul_changement_voie(P, W, v, tgt_num, force); }
  const int tl = W.C.tl_pre[slot];
  if (tl >= 0 && memo_touch(&W.V.memo[v * kMemo], tl)) lc_draw(W);
  if (code != LC_TEST) { LCQ(0); return false; }
  double r = lc_draw(W) / 32767.0;
  if (!(r < W.C.phi[slot])) { LCQ(0); return false; }
  d->dirty = true;
  LCQ(0);
  lc_window(W, W.V.lane[v], p1of(W, v), v, d);
  LCQ(1);
  changement_voie(W, v, W.C.rtgt[slot], W.C.rfol[slot]);
#ifdef SYM_LC_TIMING
  tq_ = clock64();
#endif
  lc_window(W, W.V.lane[v], p1of(W, v), v, d);
  LCQ(6);
  return true;
}
constexpr int kLcBlock = 256;
// one event at its turn, literally (thread 0): the stop event of a chunk — an accepted change or an event with side effects
__device__ __noinline__ void lc_run_event(const DevParams& P, const LcView& W, int ev, int n, LcDirty* d) {
  const DevNet& N = W.N; const DevVeh& V = W.V; const DevLc& C = W.C;
  const int pass = ev >= n; const int v = pass ? ev - n : ev;
  if (!pass) {                                                        // mandatory pass (reseau.cpp:4238-4262)
    if (C.vdes[v] != -1) { bool ok = lc_run_stage(P, W, v, v, C.vdes[v], true, d); C.chgt[v] = 1; if (ok) C.vdes[v] = -1; }
  } else if (!C.chgt[v]) {                                            // discretionary pass (:4264-4327)
    const int link = N.lane_link[V.lane[v]];
    const int nV = N.lane_num[V.lane[v]]; int nVF = nV + 1, nVS = nV - 1;
    if (C.ordre == 2) { nVF = nV - 1; nVS = nV + 1; }
    bool ok = lc_run_stage(P, W, v, n + v, nVF, false, d);
    if (!ok && !d->dirty) ok = lc_run_stage(P, W, v, 2 * n + v, nVS, false, d);
    else if (!ok && d->dirty) {                                       // the first candidate left side effects: the second one is evaluated on the spot
      const int nl = N.link_nl[link];
      if (nVS >= 0 && nVS < nl && ((C.voies_ok[v] >> nVS) & 1)) calcul_changement_voie(P, W, v, nVS, false);
    }
  }
}
// copies the vehicles currently on link K into the shared-memory snapshot (whole block); the snapshot stays disabled when the link holds
// more vehicles than it has room for
__device__ __noinline__ void snap_build(const LcView& W, LcSnap* S, int K) {
  const int tid = threadIdx.x;
  const int l0 = W.N.link_first[K], nl = W.N.link_nl[K], l1 = l0 + nl - 1;
  if (tid == 0) { S->link = -1; S->lane0 = l0; S->nl = nl; S->over = nl > kSnapLanes; }
  if (tid < kSnapLanes) { S->cnt[tid] = 0; S->len[tid] = tid < nl ? W.N.lane_len[l0 + tid] : 0.0; }
  __syncthreads();
  if (nl <= kSnapLanes) {
    const int a = W.C.lane_start[l0], b = W.C.lane_start[l1] + W.C.lane_count[l1];
    auto put = [&](int y) { const int li = W.V.lane[y] - l0; if ((unsigned)li >= (unsigned)nl) { S->over = 1; return; }
      const int i = atomicAdd(&S->cnt[li], 1); if (i >= kSnapCap) { S->over = 1; return; }
      S->e[li][i] = SnapE{p1of(W, y), y, W.V.id[y]}; S->prev[li][i] = W.V.prev_leader[y]; };
    for (int q = a + tid; q < b; q += blockDim.x) { const int y = W.C.order[q]; if (!W.C.moved[y] && W.V.status[y] == ST_ALIVE) put(y); }
    if (tid == 0) for (int y = W.C.link_delta_head[K]; y >= 0; y = W.C.delta_next[y]) put(y);
  }
  __syncthreads();
  if (tid == 0) { if (!S->over) S->link = K; else for (int li = 0; li < kSnapLanes; li++) S->cnt[li] = 0; }
  __syncthreads();
}
constexpr int kLcList = 1024;
__global__ void __launch_bounds__(kLcBlock) k_lc_walk(DevNet N, DevVeh V, DevLc C, int n) {
  const DevParams& P = cP_ref();
  __shared__ MtShared srng; __shared__ LcSnap snap; __shared__ int rb[kRb]; __shared__ int rb_pos, rb_fill, rb_base;
  __shared__ int s_wsum[kLcBlock / 32], s_stop, s_stop_off, s_total, s_dirty_link, s_full, s_preds, s_nlist, s_list[kLcList]; __shared__ unsigned char s_res[3 * kLcList]; __shared__ double s_xlo, s_xhi;
  LcView W{N, V, C, rb, &rb_pos, &rb_fill, &rb_base, &srng, &snap};
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5; const int E = 2 * n;
  mt_load(srng, C.rng);
  if (tid == 0) { rb_pos = rb_fill = rb_base = 0; snap.link = -1; snap.nl = 0; }
  __syncthreads();
  int base = 0; long long walked = 0, reevals = 0;
#ifdef SYM_LC_TIMING
  long long tA = 0, tB = 0, tC = 0, t0 = clock64(), tF = 0, tS = 0, tW = 0, tP = 0, tK = 0, tP0 = 0, tK0 = 0; int nchunk = 0, nstop = 0, ndirty = 0, npreds = 0, nsnap = 0;
#endif
  while (base < E) {
#ifdef SYM_LC_TIMING
    long long c0 = clock64();
#endif
    lc_fill_block(W, 4 * kLcBlock + 8);
    if (tid == 0) s_stop = 0x7fffffff;
    __syncthreads();
#ifdef SYM_LC_TIMING
    tF += clock64() - c0;
#endif
    // every thread: its event under the assumption that all tests before it (and its own) are refused
    const int e = base + tid;
    int cnt = 0, nt = 0, toff[2]; double tphi[2]; bool hard = false, live = false;
    int memo_l[kMemo]; int v = 0, pass = 0;
    if (e < E && C.kind[e] == 1) {
      pass = e >= n; v = pass ? e - n : e;
      int slots[2], ns = 0;
      if (!pass) { if (C.vdes[v] != -1) slots[ns++] = v; }
      else if (!C.chgt[v]) { slots[ns++] = n + v; slots[ns++] = 2 * n + v; }
      live = ns > 0;
      for (int k = 0; k < kMemo; k++) memo_l[k] = V.memo[v * kMemo + k];
      for (int q = 0; q < ns && !hard; q++) {
        const int code = C.code[slots[q]];
        if (code == LC_SLOW) { hard = true; break; }
        const int tl = C.tl_pre[slots[q]];
        if (tl >= 0) cnt += memo_touch(memo_l, tl);
        if (code == LC_TEST) { toff[nt] = cnt; tphi[nt] = C.phi[slots[q]]; nt++; cnt++; }
      }
    }
    // exclusive prefix sum of the draw counts over the chunk
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    int woff = 0; for (int w = 0; w < wid; w++) woff += s_wsum[w];
    const int off = woff + inc - cnt;
    if (tid == kLcBlock - 1) s_total = woff + inc;
    bool stop = hard;
    for (int t = 0; t < nt && !stop; t++) { double r = rb[rb_pos + off + toff[t]] / 32767.0; if (r < tphi[t]) stop = true; }
    if (live && stop) atomicMin(&s_stop, e);
    __syncthreads();
    const int st = s_stop;
    if (live && e < st) {                                             // refused without side effects: keep the memo, mark the vehicle handled
      for (int k = 0; k < kMemo; k++) V.memo[v * kMemo + k] = memo_l[k];
      if (!pass) C.chgt[v] = 1;
      walked++;
    }
    if (e == st) s_stop_off = off;
    __syncthreads();
#ifdef SYM_LC_TIMING
    long long c1 = clock64(); tA += c1 - c0; nchunk++;
#endif
    if (st == 0x7fffffff) { if (tid == 0) rb_pos += s_total; base += kLcBlock; __syncthreads(); continue; }
    // the stop event: an accepted change or an event with side effects.  Its link is copied to shared memory first (unless it needs the literal code)
    const int from_pass = st >= n; const int vcur = from_pass ? st - n : st;
    const int K0 = N.lane_link[V.lane[vcur]];
    const bool slow = from_pass ? (C.code[n + vcur] == LC_SLOW || C.code[2 * n + vcur] == LC_SLOW) : C.code[vcur] == LC_SLOW;
    if (!slow) snap_build(W, &snap, K0); else { if (tid == 0) snap.link = -1; __syncthreads(); }
#ifdef SYM_LC_TIMING
    tS += clock64() - c1; nsnap += snap.cnt[0] + snap.cnt[1] + snap.cnt[2] + snap.cnt[3];
#endif
    if (tid == 0) {
      s_dirty_link = -1;
      rb_pos += s_stop_off;
      LcDirty d{false, false, false, 3.0, -1.0};
      lc_run_event(P, W, st, n, &d);
      if (d.dirty) { s_dirty_link = K0; s_full = d.full; s_preds = d.preds; s_xlo = d.xlo; s_xhi = d.xhi; }
      s_nlist = 0;
    }
    __syncthreads();
#ifdef SYM_LC_TIMING
    long long c2 = clock64(); tB += c2 - c1;
#endif
    base = st + 1;
    const int K = s_dirty_link;
#ifdef SYM_LC_TIMING
    nstop++; if (K >= 0) { ndirty++; if (s_preds) npreds++; }
#endif
    if (K >= 0) {                                                     // re-evaluate the later events of link K (window) and, if needed, of the links that lead into it
      if (snap.link != K) snap_build(W, &snap, K);                    // (the copy followed the change; rebuilt after the literal code)
      const int upn = N.link_up[K]; const bool full = s_full, preds = s_preds;
      const double margin = 0.05 / N.link_len[K]; const double xlo = s_xlo - margin, xhi = s_xhi + margin;
      // which vehicles, from which pass on: entry = slot * 2 + (1 when only the discretionary pass is left)
      auto want = [&](int y) { int fp; if (from_pass == 0) fp = y > vcur ? 0 : 1; else if (y > vcur) fp = 1; else return;
        const int i = atomicAdd(&s_nlist, 1); if (i < kLcList) s_list[i] = y * 2 + fp; };
      if (snap.link == K) { for (int idx = tid; idx < snap.nl * kSnapCap; idx += kLcBlock) { const int li = idx / kSnapCap, i = idx % kSnapCap; if (i >= snap.cnt[li]) continue;
          if (!full) { const double x = snap.e[li][i].p / snap.len[li]; if (x < xlo || x > xhi) continue; } want(snap.e[li][i].v); } }
      else { const int l0 = N.link_first[K], l1 = l0 + N.link_nl[K] - 1; const int a = C.lane_start[l0], b = C.lane_start[l1] + C.lane_count[l1];
        for (int q = a + tid; q < b; q += kLcBlock) { const int y = C.order[q]; if (!full) { const double x = p1of(W, y) / N.lane_len[V.lane[y]]; if (x < xlo || x > xhi) continue; } want(y); } }
      if (preds) for (int k = tid; k < N.n_links; k += kLcBlock) { if (k == K || N.link_down[k] != upn) continue;
          const int l0 = N.link_first[k], l1 = l0 + N.link_nl[k] - 1; const int a = C.lane_start[l0], b = C.lane_start[l1] + C.lane_count[l1];
          for (int q = a; q < b; q++) want(C.order[q]); }
      __syncthreads();
#ifdef SYM_LC_TIMING
      long long c3 = clock64(); tW += c3 - c2;
#endif
      const int nl = s_nlist;
      if (nl <= kLcList) {
        // one stage per thread, spread over the warps first: the stages of a window take different branches
        for (int w = (tid & 31) * (kLcBlock / 32) + (tid >> 5); w < 3 * nl; w += kLcBlock) { const int ent = s_list[w / 3]; s_res[w] = (unsigned char)lc_eval_part(P, W, ent >> 1, n, ent & 1, w % 3); }
#ifdef SYM_LC_TIMING
        tP0 += clock64() - c3;
#endif
        __syncthreads();
#ifdef SYM_LC_TIMING
        long long c4 = clock64(); tP += c4 - c3;
#endif
        for (int w = tid; w < nl; w += kLcBlock) { const int ent = s_list[w]; lc_eval_kind(W, ent >> 1, n, ent & 1, s_res[3 * w], s_res[3 * w + 1], s_res[3 * w + 2]); }
        if (tid == 0) reevals += nl;
#ifdef SYM_LC_TIMING
        tK0 += clock64() - c4; __syncthreads(); tK += clock64() - c4;
#endif
      } else {                                                        // more than the list holds: the plain loops
        for (int k = preds ? 0 : K; k < (preds ? N.n_links : K + 1); k++) {
          if (k != K && N.link_down[k] != upn) continue;
          const int l0 = N.link_first[k], l1 = l0 + N.link_nl[k] - 1; const int a = C.lane_start[l0], b = C.lane_start[l1] + C.lane_count[l1];
          for (int q = a + tid; q < b; q += kLcBlock) {
            const int y = C.order[q];
            if (k == K && !full) { const double x = p1of(W, y) / N.lane_len[V.lane[y]]; if (x < xlo || x > xhi) continue; }
            if (from_pass == 0) lc_eval_vehicle(P, W, y, n, y > vcur ? 0 : 1);
            else if (y > vcur) lc_eval_vehicle(P, W, y, n, 1);
            else continue;
            reevals++;
          }
        }
      }
    }
    __syncthreads();
#ifdef SYM_LC_TIMING
    tC += clock64() - c2;
#endif
  }
#ifdef SYM_LC_TIMING
  if (tid == 0) { printf("stage prof (thread 0, %lld stages):", g_lcprof[11]); for (int i = 0; i < 9; i++) { printf(" %lld", g_lcprof[i]); g_lcprof[i] = 0; } g_lcprof[11] = 0; printf("\n"); }
  if (tid == 0) { printf("event prof:"); for (int i = 0; i < 7; i++) { printf(" %lld", g_lcprof2[i]); g_lcprof2[i] = 0; } printf("\n"); }
  if (tid == 0) printf("lc_walk E %d chunks %d stops %d dirty %d preds %d | total %lld eval %lld (fill %lld) event %lld (snap %lld, %d entries) reeval %lld (collect %lld parts %lld [own %lld] kinds %lld [own %lld]) cycles\n", E, nchunk, nstop, ndirty, npreds, clock64() - t0, tA, tF, tB, tS, nsnap, tC, tW, tP, tP0, tK, tK0);
#endif
  if (walked) atomicAdd((unsigned long long*)&C.counters[SYMGPU_CNT_LC_EVENTS], (unsigned long long)walked);
  if (reevals) atomicAdd((unsigned long long*)&C.counters[SYMGPU_CNT_LC_REEVAL], (unsigned long long)reevals);
  if (tid == 0) C.n_changes[1] = rb_base + rb_pos;
}
// Reseau::ChangementDeVoie (reseau.cpp:4219-4330): one thread, m_LstVehicles (slot) order
__global__ void k_lc_sequential(DevNet N, DevVeh V, DevLc C, int n) {
  if (blockIdx.x || threadIdx.x) return;
  const DevParams& P = cP_ref();
  LcView W{N, V, C};
  for (int v = 0; v < n; v++) {                                       // mandatory pass
    if (V.status[v] != ST_ALIVE) continue;
    if (C.vdes[v] != -1) { bool ok = calcul_changement_voie(P, W, v, C.vdes[v], true); C.chgt[v] = 1; if (ok) C.vdes[v] = -1; }
  }
  for (int v = 0; v < n; v++) {                                       // discretionary pass
    int st = V.status[v]; if (st != ST_ALIVE) continue;               // needs a link at index 1
    if (C.chgt[v]) continue;
    int lane = V.lane[v]; int link = N.lane_link[lane]; int nl = N.link_nl[link];
    if (nl < 2) continue;
    int nV = N.lane_num[lane]; int nVF = nV + 1, nVS = nV - 1;
    if (C.ordre == 2) { nVF = nV - 1; nVS = nV + 1; }
    int mask = C.voies_ok[v];
    if (nVF >= 0 && nVF < nl && ((mask >> nVF) & 1) && calcul_changement_voie(P, W, v, nVF, false)) continue;
    if (nVS >= 0 && nVS < nl && ((mask >> nVS) & 1) && calcul_changement_voie(P, W, v, nVS, false)) continue;
  }
}

}  // namespace symb200
