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
struct LcView { const DevNet& N; const DevVeh& V; const DevLc& C; int* rb = nullptr; int* rb_pos = nullptr; int* rb_fill = nullptr; int* rb_base = nullptr; DevRng* srng = nullptr; };
constexpr int kRb = 2048;
// make at least `need` upcoming draws available in the buffer (one thread)
__device__ inline void lc_fill(const LcView& W, int need) {
  int pos = *W.rb_pos, fill = *W.rb_fill;
  if (fill - pos >= need) return;
  for (int k = pos; k < fill; k++) W.rb[k - pos] = W.rb[k];
  *W.rb_base += pos; fill -= pos; pos = 0;
  while (fill < kRb) W.rb[fill++] = rng_next(W.srng);
  *W.rb_pos = pos; *W.rb_fill = fill;
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
__device__ inline int pick_last(const LcView& W, int* c, int n) {
  if (n == 0) return -1;
  if (n == 1) return c[0];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (W.V.id[c[j]] < W.V.id[c[i]]) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { bool fol = false; for (int j = 0; j < n && !fol; j++) if (i != j && W.V.prev_leader[c[j]] == W.V.id[c[i]]) fol = true; if (!fol) return c[i]; }
  return c[n - 1];
}
// Reseau::GetFirstVehicule(list) (reseau.cpp:14084-14128)
__device__ inline int pick_first(const LcView& W, int* c, int n) {
  if (n == 0) return -1;
  if (n == 1) return c[0];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (W.V.id[c[j]] < W.V.id[c[i]]) { int t = c[i]; c[i] = c[j]; c[j] = t; }
  for (int i = 0; i < n; i++) { int ld = W.V.prev_leader[c[i]]; if (ld < 0) return c[i];
    bool found = false; for (int j = 0; j < n && !found; j++) if (i != j && W.V.id[c[j]] == ld) found = true; if (!found) return c[i]; }
  return c[n - 1];
}
__device__ inline int near_aval(const LcView& W, int lane, double pos, int excl) {           // reseau.cpp:3583-3640
  int c[kTie]; int n = 0; double pv = W.N.lane_len[lane] + 0.00001;
  auto f = [&](int y) { double p = p1of(W, y); if (p >= pos && p <= pv && y != excl) { if (p == pv) { if (n < kTie) c[n++] = y; } else { n = 1; c[0] = y; pv = p; } } };
  const int q0 = slice_lb(W, lane, pos);
  for_slice(W, lane, q0, slice_group_fwd(W, lane, q0, excl), f); for_moved(W, lane, f);
  return pick_last(W, c, n);
}
__device__ inline int near_amont(const LcView& W, int lane, double pos, int excl) {          // reseau.cpp:3374-3436 (first stage; no previous lane in scope)
  int c[kTie]; int n = 0; double pt = -99999;
  auto f = [&](int y) { double p = p1of(W, y); if (pos > p && p >= pt && y != excl) { if (p == pt) { if (n < kTie) c[n++] = y; } else { n = 1; c[0] = y; pt = p; } } };
  const int q1 = slice_lb(W, lane, pos);
  for_slice(W, lane, slice_group_bwd(W, lane, q1, excl), q1, f); for_moved(W, lane, f);
  if (n > 0 && (pos - pt) < 9999) return pick_first(W, c, n);
  return -1;
}
__device__ inline int first_vehicule(const LcView& W, int lane) {                            // reseau.cpp:3255-3295 (last maximum in id order)
  int r = -1; double p = 0; int rid = -1;
  auto f = [&](int y) { double q = p1of(W, y); if (q > p || (q == p && (r < 0 || W.V.id[y] > rid))) { if (q >= p) { r = y; p = q; rid = W.V.id[y]; } } };
  const int q1 = W.C.lane_start[lane] + W.C.lane_count[lane];
  for_slice(W, lane, slice_group_bwd(W, lane, q1, -1), q1, f); for_moved(W, lane, f);
  return r;
}
__device__ inline int prev_vehicule(const LcView& W, int v) {                                // reseau.cpp:3650-3718
  int lane = W.V.lane[v]; int c[kTie]; int n = 0; double pv = W.N.lane_len[lane] + 1; double own = p1of(W, v);
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
__device__ inline bool lane_change_phi(const DevParams& P, const LcView& W, int v, double pv, int fol, int lead, int lead_orig, int tgt, bool force, double* phi_out) {   // AbstractCarFollowing.cpp:61-323
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
__device__ inline void changement_voie(const LcView& W, int v, int tgt, int fol) {            // vehicule.cpp:2657-2711
  W.C.chgt[v] = 1;
  int old = W.V.lane[v]; int link = W.N.lane_link[old];
  double p = p1of(W, v) * W.N.lane_len[tgt] / W.N.lane_len[old];
  W.V.lane[v] = tgt;
  if (fol >= 0 && W.N.lane_link[W.V.lane[fol]] == link) p = fmax(p, p1of(W, fol));
  W.V.pos[v] = p;
  lc_mark_moved(W, v);
  W.C.vdes[v] = -1;
  next_voie_draw(W, v, tgt);
  atomicAdd(W.C.n_changes, 1);
}
__device__ inline bool calcul_changement_voie(const DevParams& P, const LcView& W, int v, int tgt_num, bool force) {   // vehicule.cpp:2248-2457
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
__device__ inline void lc_eval_stage(const DevParams& P, const LcView& W, int v, int tgt_num, bool force, int slot) {
  unsigned char code = LC_FAIL; int tl_pre = -1, rf = -1, rt = -1; double phi = 0;
  do {
    double p1 = p1of(W, v);
    if (p1 <= SYM_UNDEF || W.C.chgt[v]) break;
    int lane = W.V.lane[v]; int link = W.N.lane_link[lane]; bool desired = W.C.vdes[v] != -1; bool oblig = W.N.lane_flags[lane] & 1;
    if (first_vehicule(W, lane) == v && !oblig && !desired) break;
    int tgt = W.N.link_first[link] + tgt_num;
    double ratio = W.N.lane_len[tgt] / W.N.lane_len[lane];
    int same = near_aval(W, tgt, ratio * p1 - 0.01, v);
    if (same >= 0 && fabs(ratio * p1 - p1of(W, same)) < 0.01) {
      if (fabs(p1 - W.N.lane_len[lane]) < 0.01) code = LC_SLOW;       // move-back / swap: literal code at commit time
      break;
    }
    double q = ratio * p1 + 0.001;
    int fol = near_amont(W, tgt, q, v);
    int lead = near_aval(W, tgt, q, v);
    if (lead < 0 && (oblig || desired)) {
      int tl; int nx = next_voie(W.N, W.V.route[v], W.V.route_pos[v], tgt, &tl);
      tl_pre = tl;
      if (nx >= 0) lead = near_aval(W, nx, 0, -1);
    }
    int lead_orig = prev_vehicule(W, v);
    if (lead_orig < 0 && !oblig && !desired && !force) break;
    if (!lane_change_phi(P, W, v, p1, fol, lead, lead_orig, tgt, force, &phi)) break;
    code = LC_TEST; rf = fol; rt = tgt;
  } while (0);
  W.C.code[slot] = code; W.C.tl_pre[slot] = tl_pre; W.C.rfol[slot] = rf; W.C.rtgt[slot] = rt; W.C.phi[slot] = phi;
}
// records of vehicle v for the passes >= from_pass; n = number of slots
__device__ inline void lc_eval_vehicle(const DevParams& P, const LcView& W, int v, int n, int from_pass) {
  const bool alive = W.V.status[v] == ST_ALIVE;
  if (from_pass == 0) {
    unsigned char k = 0;
    if (alive && W.C.vdes[v] != -1) { lc_eval_stage(P, W, v, W.C.vdes[v], true, v); k = (W.C.code[v] != LC_FAIL || W.C.tl_pre[v] >= 0) ? 1 : 2; }
    W.C.kind[v] = k;                                                  // 2: only marks the vehicle as handled (m_bDejaCalcule)
  }
  unsigned char k = 0;
  if (alive && !W.C.chgt[v] && !W.C.had_mand[v]) {
    int lane = W.V.lane[v]; int link = W.N.lane_link[lane]; int nl = W.N.link_nl[link];
    if (nl >= 2) {
      int nV = W.N.lane_num[lane]; int nVF = nV + 1, nVS = nV - 1;
      if (W.C.ordre == 2) { nVF = nV - 1; nVS = nV + 1; }
      int mask = W.C.voies_ok[v];
      W.C.code[n + v] = LC_FAIL; W.C.code[2 * n + v] = LC_FAIL; W.C.tl_pre[n + v] = -1; W.C.tl_pre[2 * n + v] = -1;
      if (nVF >= 0 && nVF < nl && ((mask >> nVF) & 1)) lc_eval_stage(P, W, v, nVF, false, n + v);
      if (nVS >= 0 && nVS < nl && ((mask >> nVS) & 1)) lc_eval_stage(P, W, v, nVS, false, 2 * n + v);
      if (W.C.code[n + v] != LC_FAIL || W.C.tl_pre[n + v] >= 0 || W.C.code[2 * n + v] != LC_FAIL || W.C.tl_pre[2 * n + v] >= 0) k = 1;
    }
  }
  W.C.kind[n + v] = k;
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
__device__ inline void lc_window(const LcView& W, int lane, double pos, int excl, LcDirty* d) {
  double lo = -1, hi = -1;
  auto f = [&](int y) { if (y == excl) return; double p = p1of(W, y); if (p < pos) { if (lo < 0 || p > lo) lo = p; } else if (p > pos) { if (hi < 0 || p < hi) hi = p; } };
  const int qa = slice_lb(W, lane, pos), qb = slice_ub(W, lane, pos);
  for_slice(W, lane, slice_group_bwd(W, lane, qa, excl), qa, f); for_slice(W, lane, qb, slice_group_fwd(W, lane, qb, excl), f); for_moved(W, lane, f);
  const double len = W.N.lane_len[lane];
  if (lo < 0) { d->preds = true; d->xlo = -1; } else d->xlo = fmin(d->xlo, lo / len);
  d->xhi = hi < 0 ? 2.0 : fmax(d->xhi, hi / len);
}
// one stage of an event at its turn (thread 0 of k_lc_walk)
__device__ inline bool lc_run_stage(const DevParams& P, const LcView& W, int v, int slot, int tgt_num, bool force, LcDirty* d) {
  const int code = W.C.code[slot];
  if (code == LC_SLOW) { d->dirty = d->full = d->preds = true; return calcul_changement_voie(P, W, v, tgt_num, force); }
  const int tl = W.C.tl_pre[slot];
  if (tl >= 0 && memo_touch(&W.V.memo[v * kMemo], tl)) lc_draw(W);
  if (code != LC_TEST) return false;
  double r = lc_draw(W) / 32767.0;
  if (!(r < W.C.phi[slot])) return false;
  d->dirty = true;
  lc_window(W, W.V.lane[v], p1of(W, v), v, d);
  changement_voie(W, v, W.C.rtgt[slot], W.C.rfol[slot]);
  lc_window(W, W.V.lane[v], p1of(W, v), v, d);
  return true;
}
constexpr int kLcBlock = 256;
// one event at its turn, literally (thread 0): the stop event of a chunk — an accepted change or an event with side effects
__device__ inline void lc_run_event(const DevParams& P, const LcView& W, int ev, int n, LcDirty* d) {
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
__global__ void __launch_bounds__(kLcBlock) k_lc_walk(DevNet N, DevVeh V, DevLc C, int n) {
  const DevParams& P = cP_ref();
  __shared__ DevRng srng; __shared__ int rb[kRb]; __shared__ int rb_pos, rb_fill, rb_base;
  __shared__ int s_wsum[kLcBlock / 32], s_stop, s_stop_off, s_total, s_dirty_link, s_full, s_preds; __shared__ double s_xlo, s_xhi;
  LcView W{N, V, C, rb, &rb_pos, &rb_fill, &rb_base, &srng};
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5; const int E = 2 * n;
  for (int k = tid; k < 624; k += kLcBlock) srng.mt[k] = C.rng->mt[k];
  if (tid == 0) { srng.idx = C.rng->idx; srng.count = 0; rb_pos = rb_fill = rb_base = 0; }
  __syncthreads();
  int base = 0; long long walked = 0;
  while (base < E) {
    if (tid == 0) { lc_fill(W, 4 * kLcBlock + 8); s_stop = 0x7fffffff; }
    __syncthreads();
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
    if (tid == 0) {
      s_dirty_link = -1;
      if (st == 0x7fffffff) rb_pos += s_total;
      else {
        rb_pos += s_stop_off;
        LcDirty d{false, false, false, 3.0, -1.0};
        const int vv = st >= n ? st - n : st; const int link = N.lane_link[V.lane[vv]];
        lc_run_event(P, W, st, n, &d);
        if (d.dirty) { s_dirty_link = link; s_full = d.full; s_preds = d.preds; s_xlo = d.xlo; s_xhi = d.xhi; }
      }
    }
    __syncthreads();
    if (st == 0x7fffffff) { base += kLcBlock; continue; }
    base = st + 1;
    const int K = s_dirty_link;
    if (K >= 0) {                                                     // re-evaluate the later events of link K (window) and, if needed, of the links that lead into it
      const int from_pass = st >= n; const int vcur = from_pass ? st - n : st;
      const int upn = N.link_up[K]; const bool full = s_full, preds = s_preds;
      const double margin = 0.05 / N.link_len[K]; const double xlo = s_xlo - margin, xhi = s_xhi + margin;
      for (int k = preds ? 0 : K; k < (preds ? N.n_links : K + 1); k++) {
        if (k != K && N.link_down[k] != upn) continue;
        const int l0 = N.link_first[k], l1 = l0 + N.link_nl[k] - 1;
        const int a = C.lane_start[l0], b = C.lane_start[l1] + C.lane_count[l1];
        for (int q = a + tid; q < b; q += kLcBlock) {
          const int y = C.order[q];
          if (k == K && !full) { const double x = p1of(W, y) / N.lane_len[V.lane[y]]; if (x < xlo || x > xhi) continue; }
          if (from_pass == 0) lc_eval_vehicle(P, W, y, n, y > vcur ? 0 : 1);
          else if (y > vcur) lc_eval_vehicle(P, W, y, n, 1);
          else continue;
          atomicAdd((unsigned long long*)&C.counters[SYMGPU_CNT_LC_REEVAL], 1ULL);
        }
      }
    }
    __syncthreads();
  }
  if (walked) atomicAdd((unsigned long long*)&C.counters[SYMGPU_CNT_LC_EVENTS], (unsigned long long)walked);
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
