// merge.cuh — merge points: Convergent / PointDeConvergence on the device.
//
// Reference: PointDeConvergence::CalculInsertion (convergent.cpp:333-990), Convergent::ComputeTraffic / CalculVitesseLeader
// (:1121-1168, :1368-1403), AbstractCarFollowing::TestConvergentInsertion (carFollowing/AbstractCarFollowing.cpp:336-664),
// CarrefourAFeuxEx::UpdateConvergents (CarrefourAFeuxEx.cpp:1386-1593), the merge sensors (reseau.cpp:2432-2600,
// sensors/SensorsManager.cpp:462-700) and the leader-speed correction of NewellCarFollowing.cpp:266-298.
//
// In the reference the merge points are post-processed one after the other (Liste_convergents is a std::map: id order) after
// every vehicle has been computed; a point only reads and rewrites a handful of vehicles (the vehicle waiting on a non-priority
// branch, the follower on the priority branch, the leader downstream), and some points draw from the global random stream.
// On the device:
//   k_mrg_rank    per point, start of the step: priority levels of the points inside signalised junctions from this step's
//                 signal colours and the occupancy of the internal lanes at the end of the previous step
//   (phase A/B)   rows.cuh: the speed imposed on the head of a non-priority branch and the pessimistic leader speed
//   k_mrg_probe   per point, parallel, read-only: which vehicles the point would touch -> ownership marks (min / max point), its count-only
//                 draws (exact: the probe remembers the memo entries it has counted), whether its decision needs VALUES of the stream
//   k_mrg_classify  points whose vehicles no other point touches are "clean", the others "dirty"
//   k_mrg_eval    per clean point that needs no value, parallel, on a side stream beside everything below (other vehicles): processed at once
//   k_mrg_dirty   one block, beside k_mrg_eval: the dirty points in rounds that respect the point order wherever two of them share a vehicle
//   k_scan64 / k_mrg_list  stream offset of every point from the draw counts, in point order; compact list of the points to visit
//   k_mrg_walk    one block, points in id order: the draws of the clean points that need values (congested branch,
//                 AbstractCarFollowing.cpp:357-375) against a sliding window of generated values, and the few dirty points left, literally
//   k_mrg_apply   per point, parallel: the clean points whose decision the walk has just made
//   k_mrg_tally   after the join: marks back to idle, insertion / refusal counts, go-ahead for the sensors
//   k_mrg_rotate / k_mrg_sensors / k_mrg_sensors_sum  per vehicle, end of the step, beside the next step's lane ordering: merge sensor counts
// Every arithmetic expression is the reference's, in its order (fp64, no contraction).
#pragma once
#include "lanechange.cuh"
#include "micro.cuh"
#include "rows.cuh"

namespace symb200 {

constexpr int kMrgInv = 12;       // vehicles one merge point may touch (waiting candidates, followers, leader, upstream memo misses)
constexpr int kMrgLanes = kMaxCtx;
enum : int { MS_NONE = 0, MS_DONE = 1, MS_VALUES = 2, MS_DIRTY = 3, MS_OVERFLOW = 4, MS_SETTLED = 5 };   // MS_SETTLED: a shared point k_mrg_dirty has processed (k_mrg_eval, which may run beside it, must not take it for a clean one)
enum : int { MM_PROBE = 0, MM_EVAL = 1, MM_DECIDED = 2, MM_SERIAL = 3 };

struct DevMrg {
  int n_cvg, n_pt, n_sens, n_cls;
  double max_debit_max, max_vit_max, vx0;
  const int* link_cvg;                                                     // link -> Convergent at its downstream end (-1)
  const int *cvg_node, *cvg_down, *cvg_pt0, *cvg_npt, *cvg_am0, *cvg_nam, *cvg_win, *cvg_sens0, *cvg_am;
  const double *cvg_gamma, *cvg_mu, *cvg_poscpt, *cvg_tf;
  const int *pt_cvg, *pt_down, *pt_vam0, *pt_nvam, *pt_mode, *pt_pm0;
  const int *vam_lane, *vam_ref; int* vam_niv;
  const int *pm_vam, *pm_sl, *pm_occ0, *pm_occ;                            // movements of the junction that pass through a point
  const int *lane_pt, *lane_vam;                                           // upstream lane -> its point (first in cvg order) and entry
  const int *sens_link, *sens_cnt0; const double* sens_pos; int* cnt;      // ponctual sensors; cnt: [type][lane][window] ring per sensor
  const int *sens_win; const int* sens_row0; int* row_tot;                 // row_tot[sens_row0[s] + type * n_lanes + lane]: the window sum the next step reads (k_mrg_sensors_sum)
  const int *lsens0, *lsens;                                               // link -> sensors (Tuyau::getListeCapteursConvergents)
  const int *lcam0, *lcam, *lcav0, *lcav;                                  // connections: m_LstTuyAm / m_LstTuyAv
  const int *lmv0, *lmv_ent, *lmv_exit;                                    // internal link -> (entry link, exit link) of its movements
  const int *fm0, *fm_exit, *fm_tl0, *fm_ntl, *fm_tl;                      // entry link -> per exit link: internal links of the FIRST movement
  const int* node_type;
  double* pt_last; int *pt_rand, *pt_free;                                 // m_dbInstLastPassage, m_nRandNumbers, m_bFreeFlow
  int* cand_up;                                                            // upstream lane -> first vehicle that crossed it entirely this step
  int *own_min, *own_max;                                                  // per vehicle slot: lowest / highest point that touches it
  int *st, *inv, *ninv, *pre, *kmax, *decide, *ndraw, *flist; double* proba;   // per point, this step; flist: the points the walk must visit (unordered)
  int* res;                                                               // per point: 1 inserted, 2 refused this step (summed by k_mrg_apply)
  int* sens_go;                                                           // update count the sensor kernel may run for (set by k_mrg_tally once the step is certain)
  int* eval_done;                                                         // blocks of k_mrg_eval that have finished this step (the walk's literal visits wait for all of them)
  int *dlist, *n_dirty; unsigned long long* mark;                         // points that share a vehicle with another point; per vehicle: (round << 32 | ~point) of the lowest unfinished point touching it
  unsigned long long *kf, *kfo; int *fsorted, *rbuf; int rbuf_cap;                       // kf: per point, draws known before the walk (low 40 bits) | visit flag (bit 40); kfo: its exclusive scan = offset in the stream | rank in the visit list
  int *nused, *nvoie, *ulanes;                                             // per vehicle: intermediate lanes used this step (count, first three of them); m_pNextVoie (sticky)
  double* mrgx;                                                            // per order position: phase A -> phase B side record (6 doubles)
  DevRng* rng; unsigned* draws;                                            // the stream (device copy) and the cumulative count of count-only draws
  long long* counters;
};
enum { MX_OWN = 0, MX_POSL, MX_LENL, MX_BORNE, MX_REPL, MX_VUNC, kMrgX = 6 };

struct MrgView { const DevParams& P; const DevNet& N; const DevVeh& V; const DevMrg& M; const int* order; const double* skey; const int* lane_start;
                 const int* lane_count; const int* posidx; const double* rows; double inst; int n_upd; };

__device__ __forceinline__ double min_std(double a, double b) { return (b < a) ? b : a; }      // std::min<double>, also for NaN operands
__device__ __forceinline__ double mkey(const MrgView& W, int q) { double k = W.skey[q]; return k <= SYM_UNDEF ? 0.0 : k; }
__device__ __forceinline__ double mpos1(const MrgView& W, int x) { double p = snap_pos(W.V.pos[x]); return p <= SYM_UNDEF ? 0.0 : p; }
__device__ __forceinline__ int mlink1(const MrgView& W, int x) { return (W.V.oflags[x] & O_LINK1NULL) ? -1 : W.N.lane_link[W.V.lane[x]]; }
__device__ __forceinline__ double cvg_tf(const MrgView& W, int c, int k) { return fmax(W.M.cvg_tf[c * W.M.n_cls + k], 1 / debit_max(W.P.cls[k])); }

// Reseau::GetFirstVehicule / GetLastVehicule over a group of equal positions [g0, g1) of the lane order (id ascending), with the
// leaders of THIS step's contexts (every vehicle has been computed when the merge points are processed)
__device__ inline int grp_first(const MrgView& W, int g0, int g1) {                            // reseau.cpp:14084-14128
  if (g1 - g0 == 1) return W.order[g0];
  for (int i = g0; i < g1; i++) { int ld = W.V.ctx_leader[W.order[i]]; if (ld < 0) return W.order[i];
    bool found = false; for (int j = g0; j < g1 && !found; j++) if (j != i && W.order[j] == ld) found = true;
    if (!found) return W.order[i]; }
  return W.order[g1 - 1];
}
__device__ inline int grp_last(const MrgView& W, int g0, int g1) {                             // reseau.cpp:14041-14081
  if (g1 - g0 == 1) return W.order[g0];
  for (int i = g0; i < g1; i++) { bool fol = false;
    for (int j = g0; j < g1 && !fol; j++) if (j != i && W.V.ctx_leader[W.order[j]] == W.order[i]) fol = true;
    if (!fol) return W.order[i]; }
  return W.order[g1 - 1];
}
// largest key < pos (strict) or <= pos: [g0, g1) of that group, false when none
__device__ inline bool grp_below(const MrgView& W, int lane, double pos, bool strict, int* g0, int* g1, double* key) {
  const int s = W.lane_start[lane]; int lo = s, hi = s + W.lane_count[lane];
  while (lo < hi) { int m = (lo + hi) >> 1; double k = mkey(W, m); if (strict ? (k < pos) : (k <= pos)) lo = m + 1; else hi = m; }
  if (lo == s) return false;
  const double k = mkey(W, lo - 1); int a = lo - 1; while (a - 1 >= s && mkey(W, a - 1) == k) a--;
  *g0 = a; *g1 = lo; *key = k; return true;
}
// Reseau::GetNearAmontVehicule reseau.cpp:3374-3520
__device__ inline int near_amont_m(const MrgView& W, int lane, double pos, double dstmax) {
  if (lane < 0) return -1;
  const int link0 = W.N.lane_link[lane];
  if (dstmax < 0) dstmax = 9999;
  int g0, g1; double pt;
  bool have = grp_below(W, lane, pos, true, &g0, &g1, &pt);
  if (dstmax > 0 && have && (pos - pt) < dstmax) return grp_first(W, g0, g1);
  dstmax -= pos;
  if (have) return -1;                                                                     // :3469-3473 with dbPosTmp reset to -99999
  const int pv = W.N.lane_prev[lane];
  if (pv >= 0 && dstmax > 0) {
    pos = W.N.lane_len[pv] + 0.001;
    if (grp_below(W, pv, pos, false, &g0, &g1, &pt)) return (pos - pt) < dstmax ? grp_first(W, g0, g1) : -1;
  }
  dstmax -= pos;
  if (dstmax > 0 && W.M.lcam0[link0 + 1] - W.M.lcam0[link0] == 1) {
    const int ul = W.M.lcam[W.M.lcam0[link0]];
    if (W.N.link_nl[ul] == 1) { const int up = W.N.link_first[ul]; pos = W.N.lane_len[up] + 0.001;
      if (grp_below(W, up, pos, false, &g0, &g1, &pt)) return grp_first(W, g0, g1); }
  }
  return -1;
}
// Reseau::GetNearAvalVehicule reseau.cpp:3583-3640
__device__ inline int near_aval_m(const MrgView& W, int lane, double pos) {
  if (lane < 0) return -1;
  const int s = W.lane_start[lane], e = s + W.lane_count[lane]; int lo = s, hi = e;
  while (lo < hi) { int m = (lo + hi) >> 1; if (mkey(W, m) < pos) lo = m + 1; else hi = m; }
  if (lo == e) return -1;
  const double k = mkey(W, lo); if (!(k <= W.N.lane_len[lane] + 0.00001)) return -1;
  int b = lo + 1; while (b < e && mkey(W, b) == k) b++;
  return grp_last(W, lo, b);
}
// nearest vehicle with pos >= 0 on a lane (second and third stage of GetPrevVehicule, reseau.cpp:3740-3815), excluding `excl`
__device__ inline int first_from_zero(const MrgView& W, int lane, int excl) {
  const int s = W.lane_start[lane], e = s + W.lane_count[lane];
  int a = s; while (a < e && (W.order[a] == excl || !(mkey(W, a) >= 0))) a++;
  if (a == e || !(mkey(W, a) <= W.N.lane_len[lane] + 1)) return -1;
  const double k = mkey(W, a); int c[kTie]; int n = 0;
  for (int b = a; b < e && mkey(W, b) == k; b++) if (W.order[b] != excl && n < kTie) c[n++] = W.order[b];
  if (n == 1) return c[0];
  for (int i = 0; i < n; i++) { bool fol = false; for (int j = 0; j < n && !fol; j++) if (j != i && W.V.ctx_leader[c[j]] == c[i]) fol = true; if (!fol) return c[i]; }
  return c[n - 1];
}
// the context lanes of a vehicle (row of phase A, or the per-vehicle arrays for long contexts and the origins' queue heads)
__device__ inline int ctx_lanes_of(const MrgView& W, int x, int* lanes) {
  const int st = W.V.status[x];
  if (st == ST_ALIVE || st == ST_NEW) { const int p = W.posidx[x];
    if (W.order[p] == x) { const Row r = row_at(W.rows, p); const int nl = (hi32(r[R_I0]) >> 8) & 15;
      if (nl <= kRowLanes) { for (int k = 0; k < nl; k++) lanes[k] = lo32(r[R_REC + kRecD * k + 4]); return nl; } } }
  const int nl = W.V.ctx_nl[x]; for (int k = 0; k < nl; k++) lanes[k] = W.V.ctx_lane[x * kMaxCtx + k];
  return nl;
}
// Vehicule::GetDstParcourueEx vehicule.cpp:3879-3915 (m_LstUsedLanes = the first `nused` context lanes after the start lane)
__device__ inline double dst_ex(const MrgView& W, int x) {
  const int l1 = W.V.lane[x], l0 = W.V.lane_n[x]; const int k1 = mlink1(W, x), k0 = W.N.lane_link[l0];
  if (k0 == k1) return W.V.pos_n[x] - mpos1(W, x);
  if (k1 < 0 && l1 == l0) return W.V.pos_n[x];
  double d = W.N.lane_len[l1] - mpos1(W, x);
  const int nu = W.M.nused[x];
  if (nu > 0) { int lanes[kMrgLanes]; ctx_lanes_of(W, x, lanes); for (int k = 1; k <= nu; k++) if (l0 != lanes[k]) d += W.N.lane_len[lanes[k]]; }
  return d + W.V.pos_n[x];
}
// Vehicule::UpdateLstUsedLanes vehicule.cpp:1751-1775 after the end lane has been rewritten
__device__ inline void update_used(const MrgView& W, int x) {
  int nu = W.M.nused[x]; if (nu <= 0) return;
  int lanes[kMrgLanes]; ctx_lanes_of(W, x, lanes);
  const int l0 = W.V.lane_n[x];
  for (int k = 1; k <= nu; k++) if (lanes[k] == l0) { nu = k - 1; break; }
  if (W.N.lane_link[l0] == mlink1(W, x)) nu = 0;
  W.M.nused[x] = nu;
}
__device__ inline void set_acc_m(const MrgView& W, int x, double a) { if (fabs(a) < 0.000001) a = 0; W.V.acc_n[x] = a; }
// a vehicle that does not insert waits at the end of its upstream lane (convergent.cpp:405-412, :433-438, :819-827, :972-987)
__device__ inline void immobilise(const MrgView& W, int x, int lane, bool acc_sic) {
  W.V.pos_n[x] = W.N.lane_len[lane] - 0.0001; W.V.lane_n[x] = lane;
  W.V.vit_n[x] = dst_ex(W, x) / W.P.dt;
  set_acc_m(W, x, acc_sic ? W.V.vit[x] / W.P.dt : (W.V.vit_n[x] - W.V.vit[x]) / W.P.dt);
  update_used(W, x);
}
// first index k in [k0, re) with route_links[k] == link, or re.  The route entries are read four at a time: the reads do not depend on
// each other, only the early exit does (a merge point is a chain of dependent reads; this one was a quarter of it)
__device__ inline int route_find(const MrgView& W, int k0, int re, int link) {
  const int* rl = W.N.route_links;
  for (int k = k0; k < re; k += 4) {
    const int a = rl[k], b = rl[min(k + 1, re - 1)], c = rl[min(k + 2, re - 1)], d = rl[min(k + 3, re - 1)];
    if (a == link) return k;
    if (b == link && k + 1 < re) return k + 1;
    if (c == link && k + 2 < re) return k + 2;
    if (d == link && k + 3 < re) return k + 3;
  }
  return re;
}
// Vehicule::IsItineraire vehicule.cpp:4060-4160
__device__ inline bool is_itineraire(const MrgView& W, int x, int link) {
  const int ro = W.N.route_off[W.V.route[x]], re = W.N.route_off[W.V.route[x] + 1];
  if (route_find(W, ro, re, link) < re) return true;
  const int br = W.N.link_brick[link];
  if (br >= 0) for (int k0 = ro; k0 + 1 < re; k0 += 4) {                                     // the link is inside a junction: a movement of the route goes through it
    int rk[5], dn[4];
    for (int j = 0; j < 5; j++) rk[j] = W.N.route_links[min(k0 + j, re - 1)];
    for (int j = 0; j < 4; j++) dn[j] = W.N.link_down[rk[j]];
    for (int j = 0; j < 4 && k0 + j + 1 < re; j++) if (dn[j] == br)
      for (int m = W.M.lmv0[link]; m < W.M.lmv0[link + 1]; m++) if (W.M.lmv_ent[m] == rk[j] && W.M.lmv_exit[m] == rk[j + 1]) return true;
  }
  return false;
}
// Vehicule::CalculPosition vehicule.cpp:4615-4760 over Reseau::ComputeItineraireComplet (reseau.cpp:14002-14032)
__device__ inline double calcul_position(const MrgView& W, int x, double tr, int lane) {
  double p = mpos1(W, x) + W.V.vit_n[x] * tr;
  const int l1 = W.V.lane[x]; const int k1 = W.N.lane_link[l1], kt = W.N.lane_link[lane];
  if (kt == k1) return p;
  const int ro = W.N.route_off[W.V.route[x]], re = W.N.route_off[W.V.route[x] + 1];
  // The complete itinerary (links and the internal links of the first movement between two of them) is walked from its start to the vehicle's
  // link k1, then on to kt.  Before k1 is met nothing is accumulated, so the walk may start where k1 first appears: at its first index in the
  // route when it is a link of the route, or at the first pair (entry link, exit link) of the route that one of its movements joins when it is
  // inside a junction (the internal links of other pairs cannot be k1).
  int ks = re;
  if (W.N.link_brick[k1] < 0) ks = route_find(W, ro, re, k1);
  else {
    for (int k = ro + 1; k < re && ks == re; k++) { const int prev = W.N.route_links[k - 1], cur = W.N.route_links[k];
      for (int m = W.M.lmv0[k1]; m < W.M.lmv0[k1 + 1]; m++) if (W.M.lmv_ent[m] == prev && W.M.lmv_exit[m] == cur) { ks = k; break; } }
  }
  bool started = false;
  for (int k = ks; k < re; k++) {
    const int cur = W.N.route_links[k];
    if (k > ro) {                                                                          // internal links of the first movement between route[k-1] and route[k]
      const int prev = W.N.route_links[k - 1];
      for (int f = W.M.fm0[prev]; f < W.M.fm0[prev + 1]; f++) if (W.M.fm_exit[f] == cur) {
        for (int t = W.M.fm_tl0[f]; t < W.M.fm_tl0[f] + W.M.fm_ntl[f]; t++) { const int il = W.M.fm_tl[t];
          if (started) { if (il == kt) return p; p -= W.N.link_len[il]; }
          else if (il == k1) { started = true; p = -(W.N.lane_len[l1] - p); } }
        break; }
    }
    if (started) { if (cur == kt) return p; p -= W.N.link_len[cur]; }
    else if (cur == k1) { started = true; p = -(W.N.lane_len[l1] - p); }
  }
  return 9999;
}
struct MrgAcc { int inv[kMrgInv]; int ninv; int ndraw; bool probe; bool overflow; int pmv[6], pmt[6]; int npm = 0; };   // pmv / pmt: memo entries a probe has already counted as drawn (a probe writes nothing)
__device__ inline void touch(MrgAcc& A, int x) { if (x < 0) return; for (int k = 0; k < A.ninv; k++) if (A.inv[k] == x) return; if (A.ninv < kMrgInv) A.inv[A.ninv++] = x; else A.overflow = true; }
// Vehicule::CalculNextVoie with its memoised draw (GetTirage, vehicule.cpp:1779-1795).  probe: nothing is written, a miss is
// reported through *miss.  Returns the next lane; *ndraw is incremented when a draw is consumed.
__device__ inline int next_voie_m(const MrgView& W, int x, int lane, bool probe, int* ndraw, bool* miss, MrgAcc* acc = nullptr) {
  int tl; const int nx = next_voie(W.N, W.V.route[x], W.V.route_pos[x], lane, &tl);
  if (tl >= 0) { int* memo = &W.V.memo[x * kMemo]; bool hit = false; for (int k = 0; k < kMemo; k++) if (memo[k] == tl) hit = true;
    if (!hit && probe && acc) {                                                            // the evaluation would have memoised it at the first miss
      for (int k = 0; k < acc->npm; k++) if (acc->pmv[k] == x && acc->pmt[k] == tl) hit = true;
      if (!hit) { if (acc->npm < 6) { acc->pmv[acc->npm] = x; acc->pmt[acc->npm] = tl; acc->npm++; } else acc->overflow = true; } }
    if (!hit) { (*ndraw)++; if (miss) *miss = true; if (!probe) memo_touch(memo, tl); } }
  return nx;
}
// Reseau::GetVehiculeAmont reseau.cpp:9638-9690: depth-first over the upstream connections, first lane of every link
__device__ inline int vehicule_amont(const MrgView& W, MrgAcc& A, int link, double dst) {
  struct Fr { int link, it; double cum; } stk[6]; int sp = 0;
  stk[0] = {link, -1, 0.0};
  while (sp >= 0) {
    Fr& f = stk[sp];
    if (f.it < 0) {                                                                        // entering the link
      const int v = near_amont_m(W, W.N.link_first[f.link], W.N.link_len[f.link], -1.0);
      if (v >= 0) {                                                                        // found: unwind through the callers' tests (:9679-9683)
        int res = v; int lvl = sp - 1;
        while (lvl >= 0 && res >= 0) { const Fr& c = stk[lvl]; bool ok = false;
          if (dst > c.cum + W.N.link_len[c.link] - mpos1(W, res)) { bool miss = false; int nd = 0;
            const int nx = next_voie_m(W, res, W.V.lane[res], A.probe, &nd, &miss, &A); A.ndraw += nd; if (miss) touch(A, res);
            ok = nx == W.N.link_first[c.link]; }
          if (ok) lvl--; else break; }
        if (lvl < 0) return res;
        sp = lvl; continue;                                                                // the caller at `lvl` goes on with its next upstream link
      }
      const int nt = W.M.node_type[W.N.link_up[f.link]];
      if (dst < f.cum || nt == 'E' || nt == 'S' || nt == 'P' || sp == 5) { sp--; continue; }
      f.it = 0;
    }
    const int a0 = W.M.lcam0[f.link];
    if (a0 + f.it >= W.M.lcam0[f.link + 1]) { sp--; continue; }
    const int am = W.M.lcam[a0 + f.it]; f.it++;
    stk[sp + 1] = {am, -1, f.cum + W.N.link_len[f.link]}; sp++;
  }
  return -1;
}
// sensors: SensorsManager::GetDebitMoyen / GetNbVehTotal / GetRatioTypeVehicule (sensors/SensorsManager.cpp:551-700).
// The reference shifts its per-step arrays; here a ring: logical slot j of update count n sits at (max(0, n - W) + j) % W.
__device__ inline int capteur_of(const MrgView& W, int c, int link) {                          // Convergent::GetCapteur :1000-1014
  if (link == W.M.cvg_down[c]) return W.M.cvg_sens0[c];
  for (int k = 0; k <= W.M.cvg_nam[c]; k++) if (W.M.sens_link[W.M.cvg_sens0[c] + k] == link) return W.M.cvg_sens0[c] + k;
  return -1;
}
__device__ inline int nb_inst_of(int n_upd, int win) { return n_upd == 0 ? -1 : min(n_upd - 1, win - 1); }
__device__ inline double cpt_somme(const MrgView& W, int s, int nvoie, int type) {
  const int link = W.M.sens_link[s]; const int nl = W.N.link_nl[link]; if (nvoie < 0 || nvoie >= nl) return 0;
  double sum = 0;                                                                           // window sums are kept per (sensor, type, lane): counts are integers, any order of addition is exact
  for (int t = 0; t < W.M.n_cls; t++) if (type < 0 || t == type) sum += (double)W.M.row_tot[W.M.sens_row0[s] + t * nl + nvoie];
  return sum;
}
__device__ inline double debit_moyen(const MrgView& W, int s, int nvoie) { if (s < 0) return 0; const int win = W.M.sens_win[s];
  return cpt_somme(W, s, nvoie, -1) / (min(nb_inst_of(W.n_upd, win), win) * W.P.dt); }
__device__ inline double ratio_type(const MrgView& W, int s, int nvoie, int type) { if (s < 0) return 0; const double tot = cpt_somme(W, s, nvoie, -1);
  return tot > 0 ? cpt_somme(W, s, nvoie, type) / tot : 0; }

// AbstractCarFollowing::Compute on a post-processing context (leader already computed, no leader corrections):
// NewellCarFollowing.cpp:53-136, :204-398 ; GippsCarFollowing.cpp:50-122 ; IDMCarFollowing.cpp:61-211 ; AbstractCarFollowing.cpp:692-716
__device__ inline double law_dist(const MrgView& W, int x, double dtv, const int* lanes, int nl, double start, int L, double dist, double dn0) {
  const DevParams& P = W.P; const DevNet& n = W.N; const DevCls& c = P.cls[W.V.cls[x]]; const double v1 = W.V.vit[x];
  if (c.law == 0) {
    double trav = 0, rem = dtv;                                                            // ComputeFreeFlowDistance :70-111
    for (int k = 0; k < nl; k++) { const int ln = lanes[k]; const double sp = k == 0 ? start : 0.0; const double lex = len_ex(P, n, ln);
      double vdes = fmin(c.vx, n.link_vreg[n.lane_link[ln]]);
      if (P.accbornee) vdes = fmin(vdes, v1 + acc_max(c, v1) * dtv);
      const double tt = (n.lane_len[ln] - sp) / vdes;
      if (tt >= rem || k == nl - 1) { trav += rem * vdes; break; }
      trav += lex - sp; rem -= tt; }
    double d = trav;
    if (L >= 0) {                                                                          // ComputeCongestedDistance :204-398 (leader computed)
      const DevCls& cl = P.cls[W.V.cls[L]]; const double vL1 = W.V.vit[L]; const double vL = W.V.vit_n[L];
      const double kv = -cl.kx * cl.w / (vL1 - cl.w), kvfin = -cl.kx * cl.w / (vL - cl.w);
      const double r_self = dn0 / (-c.kx * cl.w);
      double res;
      if ((P.dt - r_self) >= -0.0001) res = fmin(1.0, (dn0 / kv + fmin(r_self / dtv * (vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      else res = fmin(1.0, (dn0 / kv + fmin((vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      const double dn1 = fmax(dn0, res);
      const double r = dn0 / (-cl.kx * cl.w);
      double pc;
      if ((P.dt - r) >= -0.0001) { pc = dist + vL * P.dt - dn1 / kvfin; if (pc <= 0.0001 && dn0 < 1) pc = dist + vL * P.dt - dn0 / kvfin; }
      else { const double alpha = -cl.w * cl.kx * P.dt / dn0;
        if (dtv < P.dt) pc = dist + vL * P.dt - dn1 / kvfin;
        else { pc = alpha * dist + vL * P.dt - alpha * dn1 / kvfin; if (pc <= 0.0001 && dn0 < 1) pc = alpha * dist + vL * P.dt - alpha * dn0 / kvfin; } }
      if (pc < d) d = pc;
    }
    return d;
  }
  const double vmaxl = fmin(c.vx, n.link_vreg[n.lane_link[lanes[0]]]);
  if (c.law == 1) {
    double vf = v1 + 2.5 * acc_max(c, v1) * dtv * (1 - v1 / vmaxl) * pow_half(0.025 + v1 / vmaxl);
    if (L >= 0) { const DevCls& cl = P.cls[W.V.cls[L]]; const double vL1 = W.V.vit[L];
      double dmax = -c.decel; if (dmax == 0) dmax = -DBL_MAX; double dlmax = -cl.decel; if (dlmax == 0) dlmax = -DBL_MAX;
      const double fac = dmax * (dtv / 2.0 + 1); const double lenL = 1 / cl.kx - cl.esp;
      const double sq = fac * fac - dmax * (2.0 * (dist - (lenL + c.esp)) - v1 * dtv - vL1 * vL1 / dlmax);
      const double vc = sq >= 0 ? fac + sqrt(sq) : 0.0; if (vc < vf) vf = vc; }
    return vf * dtv;
  }
  double relpos, relsp;
  if (L >= 0) { const DevCls& cl = P.cls[W.V.cls[L]]; relpos = dist - (1 / cl.kx - cl.esp); relsp = v1 - W.V.vit[L]; } else { relpos = DBL_MAX; relsp = v1; }
  const double amax = acc_max(c, v1); double dec = c.decel; if (dec == 0) dec = DBL_MAX;
  double z; if (relpos <= 0) z = 2; else { const double dsp = c.esp + fmax(1 * v1 + (v1 * relsp) / (2.0 * sqrt(amax * dec)), 0.0); z = dsp / relpos; }
  const double a = amax * (1.0 - pow_four(v1 / vmaxl) - pow_three_halves(z));
  return v1 * dtv + 0.5 * a * dtv * dtv;
}

// priority of an upstream lane at its point: Convergent::IsVoieAmontNonPrioritaire convergent.cpp:1230-1258
__device__ inline bool non_prioritaire(const DevMrg& M, int lane) {
  const int p = M.lane_pt[lane]; if (p < 0) return false;
  int niv = 0, nmin = INT_MAX;
  for (int a = M.pt_vam0[p]; a < M.pt_vam0[p] + M.pt_nvam[p]; a++) { if (M.vam_lane[a] == lane) niv = M.vam_niv[a]; if (nmin > M.vam_niv[a]) nmin = M.vam_niv[a]; }
  return niv > nmin && niv > 1;
}
// Convergent::CalculVitesseLeader convergent.cpp:1121-1168.  `nv` = the vehicle's sticky next lane (m_pNextVoie), already resolved.
__device__ inline double vitesse_leader(const DevParams& P, const DevNet& N, const DevMrg& M, int c, int cls, int lane1, double pos1, int nv, double inst, double pas) {
  double tl = -DBL_MIN;                                                                     // GetInstLastPassage :1170-1184 (UNDEF_INSTANT when not a point of this merge)
  for (int p = M.cvg_pt0[c]; p < M.cvg_pt0[c] + M.cvg_npt[c]; p++) if (M.pt_down[p] == nv) { tl = M.pt_last[p]; break; }
  const double tf = fmax(M.cvg_tf[c * M.n_cls + cls], 1 / debit_max(P.cls[cls]));
  const double t0 = tl >= 0 ? tl + tf - (inst - pas) : 0;
  if (t0 > 0) return (N.lane_len[lane1] - pos1) / t0;
  return P.cls[cls].vx;
}

// ---- phase A hooks (declared in micro.cuh) --------------------------------------------------------------------------------
// own: Convergent::ComputeTraffic (convergent.cpp:1368-1403, called from Vehicule::CheckNetworkInfluence vehicule.cpp:1668-1677): the head of
//      a non-priority branch is held to the speed that brings it to the merge when the follow-up time has elapsed;
// lead: NewellCarFollowing.cpp:266-298: the follower of such a head assumes it will not insert.
// Returns the draws consumed (the head's next lane is resolved once and kept: Vehicule::m_pNextVoie is never reset).
__device__ inline int mrg_phase_a(const DevParams& P, const DevNet& n, const DevVeh& V, const DevMrg* M, int i, int lane1, int L, double p1, double dtv,
                                  double inst, int* memo, MrgA* out) {
  int draws = 0;
  const int cv = M->link_cvg[n.lane_link[lane1]];
  if (cv >= 0 && V.ahead[i] < 0 && non_prioritaire(*M, lane1)) {
    int nv = M->nvoie[i];
    if (nv < 0) { int tl; nv = next_voie(n, V.route[i], V.route_pos[i], lane1, &tl); if (tl >= 0) draws += memo_touch(memo, tl); M->nvoie[i] = nv; }
    out->own = true; out->mc = vitesse_leader(P, n, *M, cv, V.cls[i], lane1, p1, nv, inst, dtv) * dtv;
  }
  if (L >= 0 && P.cls[V.cls[i]].law == 0 && V.status[L] == ST_ALIVE) {                     // the leader has a link at index 1
    const int ll = V.lane[L]; const int kl = n.lane_link[ll]; const int cl = M->link_cvg[kl];
    if (n.link_type[kl] == 'C' && cl >= 0 && M->cvg_nam[cl] >= 1 && V.ahead[L] < 0 && non_prioritaire(*M, ll)) {
      const DevCls& kc = P.cls[V.cls[L]]; const double pl = snap_pos(V.pos[L]); const double vl1 = V.vit[L];
      int nvl = M->nvoie[L]; if (nvl < 0) { int tl; nvl = next_voie(n, V.route[L], V.route_pos[L], ll, &tl); }   // the leader's own thread stores it and counts the draw
      out->lead = true; out->posl = pl; out->lenl = n.lane_len[ll];
      const double a = vl1 + acc_max(kc, V.vit[i]) * dtv, b = n.link_vreg[kl];            // sic: the leader's acceleration table at the follower's speed
      out->borne = (b < a) ? b : a;
      out->repl = (n.lane_len[ll] - pl) / dtv;
      out->vunc = vitesse_leader(P, n, *M, cl, V.cls[L], ll, pl <= SYM_UNDEF ? 0.0 : pl, nvl, inst, dtv);
    }
  }
  return draws;
}
__device__ inline void mrg_store(const DevMrg* M, int p, const MrgA& a) {
  if (!a.own && !a.lead) return;
  double* x = M->mrgx + (size_t)p * kMrgX;
  if (a.own) x[MX_OWN] = a.mc;
  if (a.lead) { x[MX_POSL] = a.posl; x[MX_LENL] = a.lenl; x[MX_BORNE] = a.borne; x[MX_REPL] = a.repl; x[MX_VUNC] = a.vunc; }
}
__device__ inline void mrg_load(const DevMrg* M, int p, int hdr, MrgA* a) {
  a->own = hdr & H_MRGOWN; a->lead = hdr & H_MRGLEAD;
  if (!a->own && !a->lead) return;
  const double* x = M->mrgx + (size_t)p * kMrgX;
  if (a->own) a->mc = x[MX_OWN];
  if (a->lead) { a->posl = x[MX_POSL]; a->lenl = x[MX_LENL]; a->borne = x[MX_BORNE]; a->repl = x[MX_REPL]; a->vunc = x[MX_VUNC]; }
}
// end of the placement loop (vehicule.cpp:1485-1502): the lanes before index kf were left during the step.  m_LstUsedLanes = the
// intermediate ones; a vehicle that crossed a whole upstream lane of a merge is the lane's candidate "from upstream" (the vehicles that
// started the step ON the lane are found in its slice, vehicule_en_attente).
__device__ inline void mrg_register(const DevNet& n, const DevMrg* M, int i, int p, const int* lanes, int kf) {
  M->nused[i] = kf > 0 ? kf - 1 : 0;
  for (int k = 1; k < kf && k <= 3; k++) M->ulanes[i * 3 + k - 1] = lanes[k];               // for the sensor counts at the end of the step (k_mrg_sensors)
  if (p < 0) return;
  for (int k = 1; k < kf; k++) if (M->link_cvg[n.lane_link[lanes[k]]] >= 0) atomicMax(&M->cand_up[lanes[k]], p);
}

// first vehicle registered at the merge (ConnectionPonctuel::m_LstInsVeh: computation order, leader first) that went through
// upstream lane `lane` and does not end the step on it: Convergent::GetVehiculeEnAttente convergent.cpp:1087-1116
__device__ inline int vehicule_en_attente(const MrgView& W, int lane) {
  const int s = W.lane_start[lane]; const double lim = W.N.lane_len[lane] - W.M.max_vit_max * W.P.dt - 1.0;   // a vehicle further upstream cannot have left the lane
  for (int q = s + W.lane_count[lane] - 1; q >= s; q--) { if (mkey(W, q) < lim) break;
    const int y = W.order[q]; if (W.V.lane_n[y] != lane) {
      // several vehicles standing at one place (the reference creates such twins when it pins vehicles at a lane end): they have the same leader,
      // so the recursion computed — and registered — them in list order, and a slot index is a place in that list
      int best = y; const double k = mkey(W, q);
      for (int q2 = q - 1; q2 >= s && mkey(W, q2) == k; q2--) { const int y2 = W.order[q2]; if (W.V.lane_n[y2] != lane && y2 < best) best = y2; }
      return best; } }
  const int k = W.M.cand_up[lane];
  if (k >= 0) { const int y = W.order[k]; if (W.V.lane_n[y] != lane) {
      // the same for vehicles that stood at one place on the lane before (side by side at a stop line) and crossed this one entirely together
      int best = y; const double key = mkey(W, k); const int s0 = W.lane_start[W.V.lane[y]];
      for (int k2 = k - 1; k2 >= s0 && mkey(W, k2) == key; k2--) { const int y2 = W.order[k2]; if (W.V.lane_n[y2] == lane || y2 > best) continue;
        const int nu = min(W.M.nused[y2], 3); for (int j = 0; j < nu; j++) if (W.M.ulanes[y2 * 3 + j] == lane) best = y2; }
      return best; } }
  return -1;
}

struct RngCursor { const int* buf; int pos; int cap; int overflow; };   // serial visits draw from the look-ahead values of k_mrg_walk

// PointDeConvergence::CalculInsertion convergent.cpp:333-990 + AbstractCarFollowing::TestConvergentInsertion (:336-664).
// mode MM_PROBE   : nothing written; A collects the vehicles the point reads at index 0 or writes, and the count-only draws
//      MM_EVAL    : processed completely unless the decision needs values of the stream -> MS_VALUES (nothing written)
//      MM_DECIDED : as MM_EVAL, the congested test's outcome is M.decide[p]
//      MM_SERIAL  : literal, drawing from the positioned stream `rng` (count-only draws are consumed from it as well)
__device__ inline int merge_point(const MrgView& W, int p, int mode, MrgAcc& A, DevRng* rng) {
  const DevMrg& M = W.M; const DevParams& P = W.P; const DevNet& N = W.N; const DevVeh& V = W.V;
  const bool probe = mode == MM_PROBE; A.probe = probe;
  const int c = M.pt_cvg[p]; const int v0 = M.pt_vam0[p], nv = M.pt_nvam[p]; const int down = M.pt_down[p];
  if (nv == 0) return MS_NONE;
  int VAmP = -1; int niv_prio = INT_MAX;
  for (int a = v0; a < v0 + nv; a++) if (M.vam_niv[a] < niv_prio) { VAmP = M.vam_lane[a]; niv_prio = M.vam_niv[a]; }      // GetVoieAmontPrioritaire :77-100
  int niv = 999, att = -1, att_lane = -1, VAmNP = -1;
  for (int a = v0; a < v0 + nv; a++) if (M.vam_niv[a] > niv_prio) {                                                        // :376-449
    const int lane = M.vam_lane[a];
    int w = vehicule_en_attente(W, lane);
    if (w >= 0) touch(A, w);
    if (!(w >= 0 && V.vit_n[w] > 0)) w = -1;
    if (w >= 0) {
      if (M.vam_niv[a] < niv) { if (att >= 0 && !probe) immobilise(W, att, att_lane, true); att = w; att_lane = lane; VAmNP = lane; niv = M.vam_niv[a]; }
      else if (!probe) immobilise(W, w, lane, true);
    }
  }
  if (att < 0) return A.ninv ? MS_DONE : MS_NONE;
  int fol = near_amont_m(W, VAmP, N.lane_len[VAmP] + 0.01, P.dt * M.max_vit_max);                                         // :463
  int TAmP = N.lane_link[VAmP]; int nVoieGir = N.lane_num[VAmP];
  auto drop = [&](int f) { if (f >= 0 && ((V.oflags[f] & O_LINK1NULL) || (V.oflags[f] & O_FEU))) return -1; return f; };  // :469-483 (every listed vehicle has a context)
  fol = drop(fol);
  if (fol < 0) for (int a = v0; a < v0 + nv; a++) if (M.vam_niv[a] < niv && M.vam_niv[a] > 1) {                           // :486-504
    fol = near_amont_m(W, M.vam_lane[a], N.lane_len[M.vam_lane[a]], -1.0);
    if (fol >= 0 && !(V.oflags[fol] & O_FEU)) { niv = M.vam_niv[a]; VAmP = M.vam_lane[a]; TAmP = N.lane_link[VAmP]; nVoieGir = N.lane_num[VAmP]; }
    else fol = -1; }
  fol = drop(fol);
  if (fol >= 0 && !is_itineraire(W, fol, TAmP)) fol = -1;                                                                  // :525-528
  touch(A, fol);
  int lead = near_aval_m(W, down, 0.0);                                                                                   // :531-533
  if (lead < 0 && !(V.oflags[att] & O_LINK1NULL)) {                                                                       // GetPrevVehicule(waiting, true) reseau.cpp:3650-3820
    lead = V.ahead[att];
    if (lead < 0) { const int k1 = N.lane_link[V.lane[att]]; const int ta = N.link_type[k1]; int voie = V.lane[att]; const int nvoie = N.lane_num[voie];
      int next_link = (ta != 'S' && ta != 'P') ? k1 : -1;
      if (ta == 'C' && M.lcav0[k1 + 1] > M.lcav0[k1]) { next_link = M.lcav[M.lcav0[k1]];
        voie = next_voie_m(W, att, voie, probe, &A.ndraw, nullptr, &A);
        if (voie < 0) voie = N.link_first[next_link] + min(nvoie, N.link_nl[next_link] - 1);
        lead = first_from_zero(W, voie, att); }
      if (lead < 0 && next_link >= 0) for (int q = M.lcav0[next_link]; q < M.lcav0[next_link + 1]; q++) { const int nn = M.lcav[q];
        if (!is_itineraire(W, att, nn)) continue;
        voie = next_voie_m(W, att, voie, probe, &A.ndraw, nullptr, &A);
        if (voie < 0) voie = N.link_first[nn] + min(nvoie, N.link_nl[nn] - 1);
        const int l2 = first_from_zero(W, voie, att); if (l2 >= 0) lead = l2; }
    }
  }
  touch(A, lead);
  bool free_flow;                                                                                                           // :538-556
  { int w = near_amont_m(W, down, M.cvg_poscpt[c], -1.0); if (w >= 0 && V.lane[w] != down) w = -1;
    free_flow = w >= 0 ? (V.oflags[w] & O_FLUIDE) != 0 : true; }
  if (!free_flow && vehicule_amont(W, A, TAmP, fmin(N.link_vreg[TAmP], M.vx0)) < 0) free_flow = true;
  double offre = 0, demande1 = 0, demande2 = 0;
  const int TNP = N.lane_link[VAmNP];
  if (!free_flow) {                                                                                                         // :560-583
    offre = min_std(M.max_debit_max, debit_moyen(W, capteur_of(W, c, M.cvg_down[c]), nVoieGir));
    const int w = vehicule_amont(W, A, TNP, fmin(N.link_vreg[TNP], M.vx0));
    const bool np_free = w >= 0 ? !(V.vit[w] < fmin(P.cls[V.cls[w]].vx, N.link_vreg[TNP])) : true;
    demande2 = np_free ? min_std(M.max_debit_max, debit_moyen(W, capteur_of(W, c, TNP), min(N.lane_num[VAmNP], nVoieGir))) : M.max_debit_max;
  } else demande1 = min_std(M.max_debit_max, debit_moyen(W, capteur_of(W, c, TAmP), nVoieGir));
  // ---- TestConvergentInsertion -----------------------------------------------------------------------------------------
  bool ins = false; double tr = 0, ta = 0; int rand_numbers = M.pt_rand[p];
  int drawn = 0;                                                                          // serial mode: count-only draws consumed from the stream so far
  RngCursor* cur = (RngCursor*)rng;                                                       // serial mode draws from the look-ahead values of k_mrg_walk
  auto draw = [&]() { const int v = cur->pos < cur->cap ? cur->buf[cur->pos] : 0; if (cur->pos >= cur->cap) cur->overflow = 1; cur->pos++; return v; };
  auto flush = [&]() { if (mode == MM_SERIAL) { for (; drawn < A.ndraw; drawn++) draw(); } };
  if (!free_flow) {                                                                                                         // :343-377
    const double q2g = M.cvg_gamma[c] * offre / (1 + M.cvg_gamma[c]);
    if (demande2 < q2g) { ins = true; rand_numbers = 0; }
    else {
      const double proba = q2g * P.dt;
      if (probe) { M.pre[p] = A.ndraw; M.kmax[p] = rand_numbers; M.proba[p] = proba; return MS_VALUES; }   // decided by k_mrg_walk from the stream's values
      if (mode == MM_EVAL) return MS_OVERFLOW;                                                                              // unreachable: the probe saw the same state
      if (mode == MM_DECIDED) { const int d = M.decide[p]; ins = d >> 16; rand_numbers -= d & 0xffff; }
      else { flush(); draw();                                                                                              // one draw, unused; the loop bound is re-read after every decrement (:362-374)
        for (int k = 0; k < rand_numbers; k++) { const double r = draw() / 32767.0; rand_numbers--; if (r < proba) { ins = true; break; } } }
    }
  } else {
    rand_numbers = 0;                                                                                                       // :381
    const DevCls& k = P.cls[V.cls[att]];
    const double tmin = fmax(cvg_tf(W, c, V.cls[att]), 1 / debit_max(k));
    int f2 = fol; bool cond = W.inst - M.pt_last[p] >= tmin;
    if (cond) {
      if (f2 >= 0) {                                                                                                        // :395-461
        if (N.link_brick[TAmP] >= 0 && M.node_type[N.link_up[TAmP]] == 'C') {                                               // get_Type_amont() == 'D'
          const int ent = M.lcam0[TAmP + 1] > M.lcam0[TAmP] ? M.lcam[M.lcam0[TAmP]] : -1;
          if (ent >= 0 && M.lcav0[ent + 1] - M.lcav0[ent] > 1) next_voie_m(W, f2, N.link_first[ent], probe, &A.ndraw, nullptr, &A);   // :432
        } else if (!is_itineraire(W, f2, M.cvg_down[c])) { f2 = -1; ins = true; }
      }
      if (N.link_nl[TAmP] > 1 && N.link_nl[M.cvg_down[c]] > 1 && nVoieGir == 0) {                                           // :472-500 (not a roundabout: one draw, then ignored)
        const int fi = near_amont_m(W, N.link_first[TAmP] + nVoieGir + 1, N.lane_len[VAmP], -1.0);
        if (fi >= 0) A.ndraw++;
      }
      const int sNP = capteur_of(W, c, TNP), sP = capteur_of(W, c, TAmP);
      double tfbar = 0;                                                                                                     // :505-513 (lane index j = 0)
      if ((sNP < 0 ? 0.0 : cpt_somme(W, sNP, 0, -1)) == 0) tfbar = cvg_tf(W, c, V.cls[att]);
      else for (int i = 0; i < M.n_cls; i++) tfbar += ratio_type(W, sNP, 0, i) * cvg_tf(W, c, i);
      double tmbar = 0;
      for (int i = 0; i < M.n_cls; i++) tmbar += ratio_type(W, sP, 0, i) * (1 / debit_max(P.cls[i]));
      const double q1mu = 1 / ((tfbar * M.cvg_mu[c]) + tmbar);
      double vit_am;                                                                                                        // :520-525
      { const int w = vehicule_amont(W, A, TAmP, fmin(N.link_vreg[TAmP], M.vx0));
        if (w >= 0 && !(V.oflags[w] & O_LINK1NULL)) vit_am = fmin(P.cls[V.cls[w]].vx, N.link_vreg[N.lane_link[V.lane[w]]]);
        else vit_am = fmin(M.max_vit_max, N.link_vreg[TAmP]); }
      if (probe) return MS_DONE;                                                                                            // every vehicle touched and every draw counted: the rest is arithmetic
      double dlag;
      if (demande1 <= q1mu) dlag = (-k.w + vit_am) / (-k.kx * k.w);
      else { const double eta = (demande1 - q1mu) / (1 / tmbar - q1mu); dlag = (-k.w + vit_am) / (-k.kx * k.w) * (1 - eta); }
      const int l1 = V.lane[att];
      if (VAmNP == l1) tr = P.dt * (N.lane_len[l1] - mpos1(W, att)) / dst_ex(W, att);                                       // :544-547
      else tr = P.dt * (N.lane_len[l1] - mpos1(W, att) + N.lane_len[VAmNP]) / dst_ex(W, att);
      const double plim = N.lane_len[VAmP] - dlag;
      if (W.inst - P.dt + tr > M.pt_last[p] + tmin) {                                                                       // :550-587
        if (f2 >= 0) { const double pf = calcul_position(W, f2, tr, VAmP); if (pf <= plim) { ins = true; ta = P.dt - tr; } else ins = false; }
        else { ins = true; ta = P.dt - tr; }
      } else {                                                                                                              // :589-628
        tr = P.dt - (W.inst - (M.pt_last[p] + tmin));
        if (f2 < 0) { ta = P.dt - tr; ins = true; }
        else { const double pf = calcul_position(W, f2, tr, VAmP); if (pf <= plim) { ta = P.dt - tr; ins = true; } else ins = false; }
      }
    } else ins = false;
  }
  if (probe) return MS_DONE;
  flush();                                                                                // count-only draws (values unused: every movement has one target lane)
  M.pt_rand[p] = rand_numbers; M.pt_free[p] = free_flow;
  if (!ins) { immobilise(W, att, VAmNP, false); M.res[p] = 2; return MS_DONE; }   // :967-987
  int la[kMrgLanes]; const int nla = ctx_lanes_of(W, att, la);
  if (free_flow) {                                                                                                          // :656-836
    int L2 = -1; double pl = 0;
    if (lead >= 0 && !(V.oflags[lead] & O_LINK1NULL)) { pl = mpos1(W, lead) + (P.dt - ta) * V.vit_n[lead]; if (down != V.lane[lead]) pl += N.lane_len[down]; L2 = lead; }
    int li[kMrgLanes]; int nli = 0; { bool add = false; for (int k = 0; k < nla; k++) { if (N.lane_link[la[k]] == N.lane_link[down]) add = true; if (add) li[nli++] = la[k]; } }
    const double pos_att = nli ? law_dist(W, att, ta, li, nli, 0.0, L2, pl, 1.0) : 0.0;
    if (pos_att > 0) {
      double dfol = 0; int lf[kMrgLanes]; int nlf = 0;
      if (fol >= 0) {                                                                                                       // :704-749
        nlf = ctx_lanes_of(W, fol, lf);
        double ptr = mpos1(W, fol) + (P.dt - ta) * V.vit_n[fol];
        int o = 0; while (nlf - o > 1 && ptr > N.link_len[N.lane_link[lf[o]]]) { ptr -= N.link_len[N.lane_link[lf[o]]]; o++; }
        double dconv = 0;
        for (int i = o; i < nlf; i++) { if (i == o) dconv = N.lane_len[lf[i]] - ptr; else dconv += N.lane_len[lf[i]]; if (N.lane_link[lf[i]] == TAmP) break; }
        dfol = law_dist(W, fol, ta, lf + o, nlf - o, ptr, att, dconv, V.dn[fol]);
        dfol = fmin(dst_ex(W, fol), (P.dt - ta) * V.vit_n[fol] + dfol);
      }
      if (fol < 0 || dfol >= 0) {
        bool dont_move = false; double trav = 0;                                                                            // :753-786
        for (int i = 0; i < nli; i++) { const int ln = li[i]; const double llen = len_ex(P, N, ln); const double ep = pos_att - trav;
          if (ln == V.lane_n[att] && ep > V.pos_n[att]) { dont_move = true; break; }
          else if (ep <= llen) { V.lane_n[att] = ln; V.pos_n[att] = ep; update_used(W, att); break; }
          trav += fmin(ep, llen); }
        if (!dont_move) { V.vit_n[att] = (pos_att + (P.dt - ta) * V.vit_n[att]) / P.dt; set_acc_m(W, att, (V.vit_n[att] - V.vit[att]) / P.dt); }
        if (fol >= 0) {                                                                                                     // :788-816
          trav = 0;
          for (int i = 0; i < nlf; i++) { const int ln = lf[i]; const double sp = i == 0 ? mpos1(W, fol) : 0.0; const double llen = len_ex(P, N, ln); const double ep = dfol - trav + sp;
            if (ep <= llen) { V.lane_n[fol] = ln; V.pos_n[fol] = ep; update_used(W, fol); break; }
            trav += fmin(ep, llen) - sp; }
          V.vit_n[fol] = (dfol + (P.dt - ta) * V.vit_n[fol]) / P.dt; set_acc_m(W, fol, (V.vit_n[fol] - V.vit[fol]) / P.dt);
        }
        M.pt_last[p] = W.inst - P.dt + tr;
        M.res[p] = 1;
        return MS_DONE;
      }
    }
    immobilise(W, att, VAmNP, false);                                                                                       // :819-830
    return MS_DONE;
  }
  if (lead < 0) return MS_DONE;                                                                                             // congested :838-966
  const double dconv = mlink1(W, att) == TNP ? N.lane_len[VAmNP] - mpos1(W, att) : N.lane_len[VAmNP] + N.lane_len[V.lane[att]] - mpos1(W, att);
  double pl = mpos1(W, lead); if (down != V.lane[lead]) pl += N.lane_len[down];
  pl += dconv;
  double d = law_dist(W, att, P.dt, la, nla, mpos1(W, att), lead, pl, 1.0);
  d = fmin(dst_ex(W, att), d);
  if (!(d >= dconv)) return MS_DONE;                                                                                        // :962-965
  { double trav = 0;
    for (int i = 0; i < nla; i++) { const int ln = la[i]; const double sp = i == 0 ? mpos1(W, att) : 0.0; const double llen = len_ex(P, N, ln); const double ep = d - trav + sp;
      if (ep <= llen) { V.lane_n[att] = ln; V.pos_n[att] = ep; update_used(W, att); break; }
      trav += fmin(ep, llen) - sp; }
    V.vit_n[att] = d / P.dt; set_acc_m(W, att, (V.vit_n[att] - V.vit[att]) / P.dt); }
  M.pt_last[p] = W.inst - P.dt;
  M.res[p] = 1;
  if (fol >= 0) {                                                                                                           // :908-958
    const double pos = mlink1(W, fol) == TAmP ? N.lane_len[VAmP] - mpos1(W, fol) : N.lane_len[VAmP] + N.lane_len[V.lane[fol]] - mpos1(W, fol);
    int lf[kMrgLanes]; const int nlf = ctx_lanes_of(W, fol, lf);
    const double dfol = fmax(0.0, law_dist(W, fol, P.dt, lf, nlf, mpos1(W, fol), att, pos, 1.0));
    if (dst_ex(W, fol) > dfol) { double trav = 0;
      for (int i = 0; i < nlf; i++) { const int ln = lf[i]; const double sp = i == 0 ? mpos1(W, fol) : 0.0; const double llen = len_ex(P, N, ln); const double ep = dfol - trav + sp;
        if (ep <= llen) { V.lane_n[fol] = ln; V.pos_n[fol] = ep; update_used(W, fol); break; }
        trav += fmin(ep, llen) - sp; }
      V.vit_n[fol] = dfol / P.dt; set_acc_m(W, fol, (V.vit_n[fol] - V.vit[fol]) / P.dt); }
  }
  return MS_DONE;
}

// The ordered visit of the value-drawing points whose number of tests depends on the values (k_mrg_walk), by one warp: entry after entry,
// the tests of one entry done by the lanes together.  rec[k] = (offset of the first tested value without the draws of the visited points
// before it, test limit, integer threshold, 0); cpre[k] = draws of the entries before k whose consumption does not depend on values;
// out[k] = (inserted << 16) | tests made.  Stops at the first entry that is not a value-drawing point.  Out of line: its registers are its own.
__device__ __noinline__ int walk_hard_run(const int4* rec, const int* cpre, const unsigned char* st, const int* hidx, int j0, int nh, int* out, int* hpre,
                                          int extra, int* H_io, const int* vals, int U) {
  const int lane = threadIdx.x & 31; int H = *H_io; int j = j0;
  for (; j < nh; j++) {
    if (lane == 0) hpre[j] = H;
    const int k = hidx[j]; if (st[k] != MS_VALUES) break;
    const int4 r = rec[k]; const int start = r.x + extra + cpre[k] + H, lim = r.y, thr = r.z;
    if (start + lim > U) break;                                                              // the values are not generated yet: the block moves the window
    int used = lim, ins = 0;
    for (int d0 = 0; d0 < lim; d0 += 32) { const int d = d0 + lane;
      const bool hit = d < lim && vals[start + d] < thr;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { used = d0 + __ffs(m); ins = 1; break; } }
    if (lane == 0) out[k] = (ins << 16) | used;
    H += used;
  }
  if (lane == 0) hpre[j] = H;
  *H_io = H;
  return j;
}
// out of line: the ordered visit of k_mrg_walk is a tight loop that must not inherit this function's register pressure
__device__ __noinline__ void merge_point_serial(const MrgView* W, int q, RngCursor* C) {
  MrgAcc A; A.ninv = 0; A.ndraw = 0; A.overflow = false;
  merge_point(*W, q, MM_SERIAL, A, (DevRng*)C);
}
// ---- kernels ----------------------------------------------------------------------------------------------------------------
// CarrefourAFeuxEx::UpdateConvergents CarrefourAFeuxEx.cpp:1386-1593.  pt_mode: 0 = re-ranked from the colours, 1 = single upstream
// lane (level 1), 2 = left at the reference levels (no controller, or behind the `break` of :1428-1436).
__global__ void k_mrg_rank(DevMrg M, const int* sl_col, const int* lane_count) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_pt) return;
  if (p == 0) *M.eval_done = 0;
  M.pt_rand[p]++;                                                                          // convergent.cpp:366, every point, every step (nothing reads it before CalculInsertion)
  const int v0 = M.pt_vam0[p], n = M.pt_nvam[p]; const int mode = M.pt_mode[p];
  for (int a = v0; a < v0 + n; a++) M.vam_niv[a] = M.vam_ref[a];
  if (mode == 1) { M.vam_niv[v0] = 1; return; }
  if (mode != 0) return;
  bool g1 = false;
  for (int m = M.pt_pm0[p]; m < M.pt_pm0[p + 1] && !g1; m++) { if (sl_col[M.pm_sl[m]] == 1) continue;
    for (int q = M.pm_occ0[m]; q < M.pm_occ0[m + 1]; q++) if (lane_count[M.pm_occ[q]] > 0) { g1 = true; break; } }
  for (int m = M.pt_pm0[p]; m < M.pt_pm0[p + 1]; m++) {
    const bool green = sl_col[M.pm_sl[m]] == 1; int grp = 2;
    if (!green) { grp = 3; for (int q = M.pm_occ0[m]; q < M.pm_occ0[m + 1]; q++) if (lane_count[M.pm_occ[q]] > 0) { grp = 1; break; } }
    const int a = M.pm_vam[m];
    if (!g1) { if (grp == 3) M.vam_niv[a] = M.vam_ref[a] + (n + 1); }
    else { if (grp == 2) M.vam_niv[a] = M.vam_ref[a] + (n + 1); else if (grp == 3) M.vam_niv[a] = M.vam_ref[a] + (2 * n + 1); }
  }
}
__global__ void k_mrg_probe(DevNet N, DevVeh V, DevMrg M, const int* order, const double* skey, const int* lane_start, const int* lane_count,
                            const int* posidx, const double* rows, double inst, int n_upd, const int* guard) {
  if (guard && *guard == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_pt) return;
  const MrgView W{cP_ref(), N, V, M, order, skey, lane_start, lane_count, posidx, rows, inst, n_upd};
  MrgAcc A; A.ninv = 0; A.ndraw = 0; A.overflow = false;
  int st = merge_point(W, p, MM_PROBE, A, nullptr);
  if (A.overflow) st = MS_OVERFLOW;
  M.st[p] = st; M.ninv[p] = A.ninv; M.ndraw[p] = 0;
  M.kf[p] = st == MS_VALUES ? (unsigned long long)(A.ndraw + 1) : st == MS_DONE ? (unsigned long long)A.ndraw : 0ULL;   // count-only draws: what the evaluation will consume (checked there)
  for (int k = 0; k < A.ninv; k++) { M.inv[p * kMrgInv + k] = A.inv[k]; atomicMin(&M.own_min[A.inv[k]], p); atomicMax(&M.own_max[A.inv[k]], p); }
}
// after every probe: points whose vehicles no other point touches are "clean"; the others go to the list of k_mrg_dirty
__global__ void k_mrg_classify(DevMrg M, const int* guard) {
  if (guard && *guard == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_pt) return;
  const int st = M.st[p];
  if (st == MS_NONE) return;
  bool clean = st != MS_OVERFLOW;
  for (int k = 0; k < M.ninv[p] && clean; k++) { const int x = M.inv[p * kMrgInv + k]; if (M.own_min[x] != p || M.own_max[x] != p) clean = false; }
  if (!clean) { M.st[p] = MS_DIRTY; M.decide[p] = st; M.kf[p] = 1ULL << 40; const int dpos = atomicAdd(M.n_dirty, 1); M.dlist[dpos] = p;   // decide[p]: what the probe found, for the first round of k_mrg_dirty
#ifdef SYM_MRG_DEBUG
    if (dpos < 6) { for (int k = 0; k < M.ninv[p]; k++) { const int x = M.inv[p * kMrgInv + k]; if (M.own_min[x] != p || M.own_max[x] != p)
      printf("dirty pt %d (down lane %d, nvam %d) inv[%d of %d] veh slot %d shared with points %d..%d\n", p, M.pt_down[p], M.pt_nvam[p], k, M.ninv[p], x, M.own_min[x], M.own_max[x]); } }
#endif
    return; }   // visited by the walk unless k_mrg_dirty settles it
  if (st == MS_VALUES) M.kf[p] |= 1ULL << 40;                                                // decided by the walk, applied by k_mrg_apply
}
__global__ void k_mrg_eval(DevNet N, DevVeh V, DevMrg M, const int* order, const double* skey, const int* lane_start, const int* lane_count,
                           const int* posidx, const double* rows, double inst, int n_upd, const int* guard) {
  if (guard && *guard == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < M.n_pt && M.st[p] == MS_DONE) {                                                    // clean points that need no values of the stream
    const MrgView W{cP_ref(), N, V, M, order, skey, lane_start, lane_count, posidx, rows, inst, n_upd};
    MrgAcc A; A.ninv = 0; A.ndraw = 0; A.overflow = false;
    const int st = merge_point(W, p, MM_EVAL, A, nullptr);
    M.st[p] = st; M.ndraw[p] = A.ndraw;
    if ((unsigned long long)A.ndraw != M.kf[p]) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_CTX_OVERFLOW], 1ULL);   // the probe's count is the offset of the points after this one: it must be what the evaluation drew
  }
  __threadfence(); __syncthreads();                                                          // this block's points are done: the walk, which runs beside this kernel, waits for all blocks before a literal visit
  if (threadIdx.x == 0) atomicAdd(M.eval_done, 1);
}
// Points that share a vehicle with another point must be processed in point order — but only relative to the points they share with.
// One block, rounds: every unfinished point stamps its vehicles with (round, point), lowest point wins; a point that holds all its
// vehicles is processed now (after a fresh probe: what it touches may have changed; anything unexpected leaves it to the walk, and so
// does a decision that needs values of the stream).  Points processed in the same round touch disjoint vehicles.
__global__ void __launch_bounds__(256) k_mrg_dirty(DevNet N, DevVeh V, DevMrg M, const int* order, const double* skey, const int* lane_start, const int* lane_count,
                                                   const int* posidx, const double* rows, double inst, int n_upd, const int* guard) {
  __shared__ int s_progress;
  if (guard && *guard == 0) return;
  const int nd = *M.n_dirty; if (nd == 0) return;
  const int tid = threadIdx.x;
  const MrgView W{cP_ref(), N, V, M, order, skey, lane_start, lane_count, posidx, rows, inst, n_upd};
  for (int round = 0; round < 6; round++) {
    const unsigned long long base = ((unsigned long long)((unsigned)n_upd * 8u + (unsigned)round + 1u)) << 32;
    for (int d = tid; d < nd; d += 256) { const int p = M.dlist[d]; if (M.st[p] != MS_DIRTY) continue;
      for (int k = 0; k < M.ninv[p]; k++) atomicMax(&M.mark[M.inv[p * kMrgInv + k]], base | (unsigned long long)(0xFFFFFFFFu - (unsigned)p)); }
    if (tid == 0) s_progress = 0;
    __threadfence(); __syncthreads();
    for (int d = tid; d < nd; d += 256) { const int p = M.dlist[d]; if (M.st[p] != MS_DIRTY) continue;
      const unsigned long long key = base | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
      bool ready = true; for (int k = 0; k < M.ninv[p] && ready; k++) ready = M.mark[M.inv[p * kMrgInv + k]] == key;
      if (!ready) continue;
      MrgAcc A; A.ninv = 0; A.ndraw = 0; A.overflow = false;
      int st; bool ok;
      if (round == 0) { st = M.decide[p]; ok = st != MS_OVERFLOW && st != MS_VALUES; }           // nothing it touches has changed since the probe of the step
      else { st = merge_point(W, p, MM_PROBE, A, nullptr); ok = !A.overflow && st != MS_VALUES; }
      for (int k = 0; k < A.ninv && ok; k++) ok = M.mark[A.inv[k]] == key;                     // everything it touches NOW is its own
#ifdef SYM_MRG_DEBUG
      if (!ok) { if (A.overflow || st == MS_VALUES) printf("upd %d round %d pt %d left: %s\n", n_upd, round, p, A.overflow ? "overflow" : "values");
        else for (int k = 0; k < A.ninv; k++) if (M.mark[A.inv[k]] != key) { const int z = A.inv[k]; const int r0 = M.own_min[z], r1 = M.own_max[z];
          printf("upd %d round %d pt %d left: veh %d (id %d) owners %d..%d st %d mark pt %d round %u\n", n_upd, round, p, z, V.id[z], r0, r1, r1 >= 0 ? M.st[r1] : -1,
                 (int)(0xFFFFFFFFu - (unsigned)(M.mark[z] & 0xFFFFFFFFu)), (unsigned)(M.mark[z] >> 32) - (unsigned)n_upd * 8u - 1u); } }
#endif
      if (!ok) continue;                                                                       // left to the walk (literal, in order); the points behind it in its cluster wait with it
      if (st == MS_DONE) { MrgAcc B; B.ninv = 0; B.ndraw = 0; B.overflow = false; merge_point(W, p, MM_EVAL, B, nullptr); M.ndraw[p] = B.ndraw; M.kf[p] = (unsigned long long)B.ndraw; }
      else { M.ndraw[p] = 0; M.kf[p] = 0ULL; }
      M.st[p] = MS_SETTLED; s_progress = 1; }
    __threadfence(); __syncthreads();
    if (!s_progress) break;
    __syncthreads();
  }
}
// ---- the random stream on the device -----------------------------------------------------------------------------------------
// RandManager::myRand = boost::mt19937 + uniform_int<>(0, 0x7FFE) by bucket rejection (RandManager.cpp:26-30).  One block of 256 threads
// MtShared / mt_twist / rng_advance / mt_load / mt_store: lanechange.cuh
__global__ void __launch_bounds__(256) k_rng_skip(DevRng* r, const int* n_ptr, int n_imm) {
  __shared__ MtShared S;
  const unsigned n = n_ptr ? (unsigned)*n_ptr : (unsigned)n_imm;
  if (n == 0) return;
  mt_load(S, r); rng_advance(S, n, nullptr); mt_store(S, r);
}
// the count-only draws of the step's car-following phase (next-lane memos, counted in *draws) advance the stream before the merge points draw
__global__ void __launch_bounds__(256) k_rng_ctx(DevRng* r, const unsigned* draws, unsigned* seen, const int* guard) {
  __shared__ MtShared S;
  if (guard && *guard == 0) return;
  const unsigned now = *draws; const unsigned n = now - *seen;
  if (n) { mt_load(S, r); rng_advance(S, n, nullptr); mt_store(S, r); }
  if (threadIdx.x == 0) *seen = now;
}
// exclusive scan of 64-bit words (block scan, scan of the block sums, add): the packed (draw count | visit flag) of every point
typedef unsigned long long u64;
__global__ void __launch_bounds__(1024) k_scan64_block(const u64* in, u64* out, u64* bsum, int n) {
  __shared__ u64 wsum[32];
  const int i = blockIdx.x * 1024 + threadIdx.x; const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u64 v = i < n ? in[i] : 0ULL; u64 incl = v;
  for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) { u64 s2 = wsum[lane]; for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, s2, o); if (lane >= o) s2 += t; } wsum[lane] = s2; }
  __syncthreads();
  const u64 off = w ? wsum[w - 1] : 0ULL;
  if (i < n) out[i] = off + incl - v;
  if (threadIdx.x == 1023) bsum[blockIdx.x] = off + incl;
}
__global__ void k_scan64_sums(u64* bsum, int nb) {                       // one warp: exclusive scan of the block sums, total in bsum[nb]
  const int lane = threadIdx.x; u64 carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 32) { const int b = b0 + lane; const u64 v = b < nb ? bsum[b] : 0ULL; u64 incl = v;
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (b < nb) bsum[b] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31); }
  if (lane == 0) bsum[nb] = carry;
}
__global__ void k_scan64_add(u64* out, const u64* bsum, int n) { const int i = blockIdx.x * 1024 + threadIdx.x; if (i < n) out[i] += bsum[blockIdx.x]; }
// the visit list in point order: rank = flag part of the scan
__global__ void k_mrg_list(DevMrg M) { const int p = blockIdx.x * blockDim.x + threadIdx.x; if (p < M.n_pt && ((M.kf[p] >> 40) & 1)) M.fsorted[(int)(M.kfo[p] >> 40)] = p; }
// The points that need the stream's VALUES (congested test) or share a vehicle with another point, in Liste_convergents order.  Every
// other point's draws are count-only and already summed (koff = exclusive scan of `known` in point order); the values the visited
// points may consume are generated ahead into rbuf, the visit itself is one thread over shared-memory records, and the stream is then
// advanced by what was really consumed.  out[0] = draws of the merge phase.
__global__ void __launch_bounds__(256) k_mrg_walk(DevNet N, DevVeh V, DevMrg M, const int* order, const double* skey, const int* lane_start,
                                                  const int* lane_count, const int* posidx, const double* rows, double inst, int n_upd, const u64* kf_total,
                                                  int* out, const int* guard, int n_eval_blocks) {
  constexpr int kBatch = 512, kRc = 4096;
  // the stream's values around the walk's position: sh_rb holds positions [sh_B, sh_B + sh_F) of the step's merge draws, S generates what follows.
  // Positions only grow, so the window slides forward; nothing is generated beyond what is consumed plus one window.
  __shared__ int sh_rb[kRc]; __shared__ int sh_B, sh_F;
  __shared__ MtShared S; __shared__ int4 e_rec[kBatch]; __shared__ int e_pt[kBatch], e_res[kBatch], e_cpre[kBatch], e_hj[kBatch], e_hpre[kBatch + 1], e_hidx[kBatch]; __shared__ unsigned char e_st[kBatch], e_done[kBatch];
  __shared__ int sh_next, sh_red2[8], sh_ctot, sh_nh, sh_H, sh_min, sh_flag;
  __shared__ int sh_red[8]; __shared__ int sh_total;
  if (guard && *guard == 0) return;
  const int tid = threadIdx.x;
#ifdef SYM_WALK_TIMING
  long long tq0 = clock64(), tq2 = 0, tq3 = 0; int n_refill = 0;
#endif
  const u64 tot = *kf_total; const int nf = (int)(tot >> 40); const int known_total = (int)(tot & ((1ULL << 40) - 1));
  if (nf == 0) {                                                                           // only count-only draws this step
    if (known_total) { mt_load(S, M.rng); rng_advance(S, (unsigned)known_total, nullptr); mt_store(S, M.rng); }
    if (tid == 0) { out[0] = known_total; M.counters[SYMGPU_CNT_MRG_DRAWS] += known_total; }
    return;
  }
  mt_load(S, M.rng);
  if (tid == 0) { sh_B = 0; sh_F = 0; }
  __syncthreads();
  // make the positions [lo, hi) available (hi - lo <= kRc, lo never below an earlier lo); whole block, uniform arguments
  auto ensure = [&](int lo, int hi) {
    const int B = sh_B, F = sh_F;
    if (hi <= B + F) return;
    __syncthreads();
    const int keep = max(0, B + F - lo); int tmp[kRc / 256]; int nk = 0;
    for (int k = tid; k < keep; k += 256) tmp[nk++] = sh_rb[lo - B + k];
    __syncthreads();
    nk = 0; for (int k = tid; k < keep; k += 256) sh_rb[k] = tmp[nk++];
    __syncthreads();
    if (lo > B + F) rng_advance(S, (unsigned)(lo - (B + F)), nullptr);
    rng_advance(S, (unsigned)(kRc - keep), sh_rb + keep);
    if (tid == 0) { sh_B = lo; sh_F = kRc; }
    __syncthreads();
#ifdef SYM_WALK_TIMING
    n_refill++;
#endif
  };
  const MrgView W{cP_ref(), N, V, M, order, skey, lane_start, lane_count, posidx, rows, inst, n_upd};
#ifdef SYM_WALK_TIMING
  tq2 = clock64();
#endif
  int extra = 0;                                                                           // draws consumed by the visited points beyond `known`
  int n_values_t = 0, n_serial = 0, overflow = 0;                                          // n_values_t: per thread, summed at the end
  for (int b0 = 0; b0 < nf; b0 += kBatch) {
    const int nb = min(kBatch, nf - b0);
    __syncthreads();
    for (int k = tid; k < nb; k += 256) { const int q = M.fsorted[b0 + k]; e_pt[k] = q; e_st[k] = (unsigned char)M.st[q]; e_done[k] = 0;
      // the test `draw / 32767.0 < proba` as an integer threshold (exactly: the quotient is monotonic in the draw), so that the ordered visit is integer work
      const double pr = M.proba[q]; int T = (int)fmin(fmax(ceil(pr * 32767.0), 0.0), 32767.0); if (!(pr == pr)) T = 0;
      while (T <= 32766 && (T / 32767.0 < pr)) T++; while (T > 0 && !((T - 1) / 32767.0 < pr)) T--;
      const int off = (int)(M.kfo[q] & ((1ULL << 40) - 1));
      // VALUES: offset of the first tested value = point offset + count-only draws + the unused draw; tests while k < m_nRandNumbers, decremented per test: ceil(n / 2) at most
      int lim = (M.kmax[q] + 1) >> 1; if (lim > kRc) { lim = kRc; overflow = 1; }              // (a point left alone for more than 2 * kRc steps)
      e_rec[k] = e_st[k] == MS_VALUES ? make_int4(off + M.pre[q] + 1, lim, T, 0) : make_int4(off + 1, 0, 0, 0); }
    __syncthreads();
    // Entries whose consumption does not depend on the values (one test at most: AbstractCarFollowing.cpp:362-374 tests while k < m_nRandNumbers,
    // decremented per test) are offsets for the others and decided in parallel once the ordered visit has passed them; only the "hard" ones
    // (two tests or more, and the points left to a literal visit) are walked in order.
    { const int i0 = 2 * tid, i1 = i0 + 1;                                                     // e_cpre = exclusive scan of the constant consumptions; hard entries compacted in order
      auto cons = [&](int k) { return k < nb && e_st[k] == MS_VALUES && e_rec[k].y <= 1 ? e_rec[k].y : 0; };
      auto hard = [&](int k) { return k < nb && !(e_st[k] == MS_VALUES && e_rec[k].y <= 1) ? 1 : 0; };
      const int a = cons(i0), b2 = cons(i1), ha = hard(i0), hb = hard(i1); int incl = a + b2, hincl = ha + hb;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o), t2 = __shfl_up_sync(0xffffffffu, hincl, o); if ((tid & 31) >= o) { incl += t; hincl += t2; } }
      if ((tid & 31) == 31) { sh_red[tid >> 5] = incl; sh_red2[tid >> 5] = hincl; }
      __syncthreads();
      int off = 0, hoff = 0; for (int w = 0; w < (tid >> 5); w++) { off += sh_red[w]; hoff += sh_red2[w]; }
      const int ex = off + incl - (a + b2), hex = hoff + hincl - (ha + hb);
      if (i0 < nb) { e_cpre[i0] = ex; e_hj[i0] = hex; if (ha) e_hidx[hex] = i0; }                // e_hj: hard entries before this one = index of the next hard entry at or after it
      if (i1 < nb) { e_cpre[i1] = ex + a; e_hj[i1] = hex + ha; if (hb) e_hidx[hex + ha] = i1; }
      if (tid == 255) { sh_ctot = off + incl; sh_nh = hoff + hincl; }
      __syncthreads(); }
    const int nh = sh_nh; const int ctot = sh_ctot;
    if (tid == 0) sh_H = 0;
    __syncthreads();
    int easy_from = 0;
    for (int j = 0;;) {                                                                        // the ordered part; e_hpre[j] = draws the hard entries before j consumed
      if (tid < 32) { int H = sh_H; const int jn = walk_hard_run(e_rec, e_cpre, e_st, e_hidx, j, nh, e_res, e_hpre, extra, &H, sh_rb - sh_B, sh_B + sh_F); if (tid == 0) { sh_H = H; sh_next = jn; } }
      __syncthreads();
      j = sh_next;
      // the easy entries the visit has passed: their positions are known; decided from the window (moved forward if they lie beyond it)
      const int klim = j < nh ? e_hidx[j] : nb;
      for (;;) {
        if (tid == 0) sh_min = INT_MAX;
        __syncthreads();
        const int B = sh_B, G = sh_B + sh_F;
        for (int k = easy_from + tid; k < klim; k += 256) if (e_st[k] == MS_VALUES && e_rec[k].y <= 1 && !e_done[k]) {
          const int pos = e_rec[k].x + extra + e_cpre[k] + e_hpre[e_hj[k]];
          if (e_rec[k].y == 1 && pos >= G) { atomicMin(&sh_min, pos); continue; }
          const bool ins = e_rec[k].y == 1 && sh_rb[pos - B] < e_rec[k].z;
          e_res[k] = ((int)ins << 16) | e_rec[k].y; e_done[k] = 1; }
        __syncthreads();
        const int mn = sh_min;
        if (mn == INT_MAX) break;
        ensure(mn, mn + 1);
      }
      easy_from = klim;
      if (j >= nh) break;
      const int k = e_hidx[j];
      if (e_st[k] == MS_VALUES) { const int start = e_rec[k].x + extra + e_cpre[k] + sh_H; ensure(start, start + e_rec[k].y); continue; }   // the window moves to this entry
      // a point that shares a vehicle with another point and was not settled by k_mrg_dirty: literally, now
      const int at = e_rec[k].x - 1 + extra + e_cpre[k] + sh_H;
      ensure(at, at + min(kRc, 64 + M.pt_rand[e_pt[k]]));                                      // its count-only draws, the unused one, at most (m_nRandNumbers + 1) / 2 tests
      if (tid == 0) { const int q = e_pt[k]; RngCursor C{sh_rb - sh_B, at, sh_B + sh_F, 0};
        { volatile int* done = M.eval_done; long long spins = 0; while (*done < n_eval_blocks && ++spins < (1LL << 26)) __nanosleep(200);   // the clean points (k_mrg_eval, beside this kernel) first, as when the kernels ran one after the other
          if (*done < n_eval_blocks) C.overflow = 1; __threadfence(); }
        merge_point_serial(&W, q, &C);
        sh_H += C.pos - at; sh_flag = C.overflow; M.st[q] = MS_DONE; }
      __syncthreads();
      overflow |= sh_flag; n_serial++; j++;
      __syncthreads();
    }
    for (int k = tid; k < nb; k += 256) if (e_st[k] == MS_VALUES) { M.decide[e_pt[k]] = e_res[k]; n_values_t++; }   // decisions out
    extra += ctot + sh_H;
    __syncthreads();
  }
  __syncthreads();
  if (n_values_t) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_MRG_VALUES], (unsigned long long)n_values_t);
  if (tid == 0) { sh_total = known_total + extra; M.counters[SYMGPU_CNT_MRG_SERIAL] += n_serial; }
  if (overflow) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_CTX_OVERFLOW], 1ULL);
  __syncthreads();
#ifdef SYM_WALK_TIMING
  tq3 = clock64();
#endif
  const int total = sh_total;
  mt_load(S, M.rng); rng_advance(S, (unsigned)total, nullptr); mt_store(S, M.rng);         // the stream itself advances by what was consumed
#ifdef SYM_WALK_TIMING
  if (tid == 0) printf("walk nf %d known %d extra %d refills %d | setup %lld loop %lld final %lld cycles\n", nf, known_total, extra, n_refill, tq2 - tq0, tq3 - tq2, clock64() - tq3);
#endif
  if (tid == 0) { out[0] = total; M.counters[SYMGPU_CNT_MRG_DRAWS] += total; }
}
__global__ void k_mrg_apply(DevNet N, DevVeh V, DevMrg M, const int* order, const double* skey, const int* lane_start, const int* lane_count,
                            const int* posidx, const double* rows, double inst, int n_upd, const int* guard) {
  if (guard && *guard == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < M.n_pt && M.st[p] == MS_VALUES) { const MrgView W{cP_ref(), N, V, M, order, skey, lane_start, lane_count, posidx, rows, inst, n_upd};
    MrgAcc A; A.ninv = 0; A.ndraw = 0; A.overflow = false; merge_point(W, p, MM_DECIDED, A, nullptr); }
}
// end of the merge phase (every point processed): ownership marks back to idle, statistics, go-ahead for the sensors of the step
__global__ void k_mrg_tally(DevMrg M, int n_upd, const int* guard) {
  if (guard && *guard == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int res = 0;
  if (p == 0) *M.sens_go = n_upd + 1;
  if (p < M.n_pt) {
    for (int k = 0; k < M.ninv[p]; k++) { const int x = M.inv[p * kMrgInv + k]; M.own_min[x] = INT_MAX; M.own_max[x] = -1; }
    res = M.res[p]; M.res[p] = 0;
  }
  const unsigned mi = __ballot_sync(0xffffffffu, res == 1), mr = __ballot_sync(0xffffffffu, res == 2);     // one atomic per warp
  if ((threadIdx.x & 31) == 0) { if (mi) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_MRG_INS], (unsigned long long)__popc(mi));
    if (mr) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_MRG_REFUS], (unsigned long long)__popc(mr)); }
}
// Reseau::UpdateConvergents reseau.cpp:2432-2600 (integer counts) after SensorsManager::UpdateTabNbVeh (:462-527): the slot the
// rotation exposes is cleared first (k_mrg_rotate), then every vehicle left on the network counts where it crossed a sensor.
__global__ void k_mrg_rotate(DevMrg M, int n_upd_after, int n_rows, const int* row_sens) {
  if (*M.sens_go != n_upd_after) return;                                                   // the step was not finished by this pass (relaxation still open)
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int s = row_sens[2 * r]; const int win = M.sens_win[s];
  if (n_upd_after > win) M.cnt[M.sens_cnt0[s] + row_sens[2 * r + 1] * win + (n_upd_after - win - 1) % win] = 0;
}
// after the counts of the step: what SensorsManager::GetDebitMoyen / GetNbVehTotal will sum during the next step (logical slots [0, min(m_nNbInst, W)))
__global__ void k_mrg_sensors_sum(DevMrg M, int n_upd_after, int n_rows, const int* row_sens) {
  if (*M.sens_go != n_upd_after) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int s = row_sens[2 * r]; const int win = M.sens_win[s]; const int kmax = min(nb_inst_of(n_upd_after, win), win); const int head = max(0, n_upd_after - win) % max(1, win);
  const int* a = M.cnt + M.sens_cnt0[s] + row_sens[2 * r + 1] * win; int sum = 0;
  for (int j = 0; j < kmax; j++) sum += a[(head + j) % win];
  M.row_tot[r] = sum;
}
__device__ inline void sens_hit(const DevNet& N, const DevMrg& M, int s, int cls, int lane, int n_upd_after) {
  const int win = M.sens_win[s]; const int nl = N.link_nl[M.sens_link[s]];
  const int last = min(n_upd_after - 1, win - 1); const int head = max(0, n_upd_after - win) % max(1, win);
  if (last >= 0) atomicAdd(&M.cnt[M.sens_cnt0[s] + (cls * nl + N.lane_num[lane]) * win + (head + last) % win], 1);
}
__global__ void k_mrg_sensors(DevNet N, DevVeh V, DevMrg M, int n, int n_upd_after) {
  if (*M.sens_go != n_upd_after) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n || V.status[x] != ST_ALIVE) return;                                           // after k_final: the vehicles still on the network; index n = end of this step
  const double p0 = V.pos_n[x];
  if (!(V.vit_n[x] > 0 && p0 >= 0)) return;
  const int l1 = V.lane[x], l0 = V.lane_n[x]; const int k1 = (V.oflags[x] & O_LINK1NULL) ? -1 : N.lane_link[l1]; const int k0 = N.lane_link[l0];
  double p1 = snap_pos(V.pos[x]);
  const int cls = V.cls[x];
  if (k1 >= 0) for (int q = M.lsens0[k1]; q < M.lsens0[k1 + 1]; q++) { const int s = M.lsens[q]; const double xs = M.sens_pos[s];
    if (p1 < xs && (p0 >= xs || k1 != k0)) sens_hit(N, M, s, cls, l1, n_upd_after); }
  if (k1 != k0) {
    for (int q = M.lsens0[k0]; q < M.lsens0[k0 + 1]; q++) { const int s = M.lsens[q]; if (p0 >= M.sens_pos[s]) sens_hit(N, M, s, cls, l0, n_upd_after); }
    const int nu = M.nused[x];                                                              // the intermediate lanes (Vehicule::m_LstUsedLanes) as they stand after the merge points
    for (int k = 0; k < nu && k < 3; k++) { const int ln = M.ulanes[x * 3 + k]; const int kk = N.lane_link[ln];
      for (int q = M.lsens0[kk]; q < M.lsens0[kk + 1]; q++) sens_hit(N, M, M.lsens[q], cls, ln, n_upd_after); }
    if (nu > 3) atomicAdd((unsigned long long*)&M.counters[SYMGPU_CNT_CTX_OVERFLOW], 1ULL);
  }
}

}  // namespace symb200
